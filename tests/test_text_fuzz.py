"""Memory-safety fuzz of the library's host text code (FASTA reader, filter, record writer) under AddressSanitizer and
UndefinedBehaviorSanitizer: tests/native/text_fuzz.cpp compiled against classeq2_b200/csrc/record_writer.cpp."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "classeq2_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_reader_and_writer_are_clean_under_asan_ubsan(tmp_path):
    exe = str(tmp_path / "text_fuzz")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-pthread",
                        os.path.join(HERE, "native", "text_fuzz.cpp"), os.path.join(CSRC, "record_writer.cpp"),
                        os.path.join(CSRC, "host_pool.cpp"), "-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip(f"cannot build with the sanitizers here: {r.stderr[-300:]}")
    r = subprocess.run([exe, "150"], env=dict(os.environ, CLS_HOST_THREADS="4"), capture_output=True, text=True, timeout=600)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and r.stdout.strip() == "bad=0", (r.returncode, r.stdout, r.stderr[-500:])
