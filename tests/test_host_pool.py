"""Stress test of the host thread pool behind ``parallel_for`` (classeq2_b200/csrc/host_pool.cpp): the loops of
``cls_place_batch`` that plan, pack and scatter a batch run on it, so a lost or doubled range is a wrong placement.
tests/native/host_pool_stress.cpp is compiled against the pool's source, plain and under ThreadSanitizer."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "native", "host_pool_stress.cpp"), os.path.join(HERE, "..", "classeq2_b200", "csrc", "host_pool.cpp")]


def _build(tmp_path, name, flags):
    exe = str(tmp_path / name)
    r = subprocess.run(["g++", "-std=c++17", "-pthread", *flags, *SRC, "-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip(f"cannot build the stress test here: {r.stderr[-300:]}")
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("threads,callers", [(8, 2), (32, 1), (64, 4)])
def test_every_index_of_every_job_is_visited_once(tmp_path, threads, callers):
    """Short jobs of changing size, more pool threads than cores (late wake-ups), several calling threads."""
    exe = _build(tmp_path, "pool_plain", ["-O2"])
    r = subprocess.run([exe, "3000" if threads == 64 else "8000", str(callers)], env=dict(os.environ, CLS_HOST_THREADS=str(threads)),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "bad=0", (r.returncode, r.stdout, r.stderr[-500:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_pool_is_clean_under_thread_sanitizer(tmp_path):
    exe = _build(tmp_path, "pool_tsan", ["-O1", "-g", "-fsanitize=thread"])
    r = subprocess.run([exe, "3000", "3"], env=dict(os.environ, CLS_HOST_THREADS="8"), capture_output=True, text=True, timeout=600)
    if "FATAL: ThreadSanitizer" in r.stderr:          # e.g. an unsupported memory layout of the box
        pytest.skip(r.stderr[-300:])
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and r.stdout.strip() == "bad=0", (r.returncode, r.stdout)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_place_sequences_loop_is_clean_under_thread_sanitizer(tmp_path):
    """cls_place_sequences places batch i + 1 while a second thread renders and writes the records of batch i (both go
    through the host pool): no data race, and the files do not depend on the batch size."""
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), "classeq2_b200", "csrc")
    exe = str(tmp_path / "place_seq_tsan")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", os.path.join(here, "native", "place_seq_tsan.cpp"),
                        os.path.join(csrc, "record_writer.cpp"), os.path.join(csrc, "host_pool.cpp"), "-o", exe], capture_output=True, text=True)
    assert "undefined reference" not in r.stderr and "error:" not in r.stderr, r.stderr[-2000:]
    if r.returncode != 0:
        pytest.skip(f"cannot build with the sanitizer here: {r.stderr[-300:]}")
    import random
    rnd = random.Random(7)
    fa = tmp_path / "q.fasta"
    with open(fa, "w") as f:
        for i in range(1500):
            f.write(f">read_{i} some description\n" + "".join(rnd.choice("ACGT") for _ in range(rnd.randrange(20, 200))) + "\n")
    outs = []
    for batch in ("64", "7", "0"):
        out = tmp_path / f"out_{batch}" / "r.x"
        r = subprocess.run([exe, str(fa), str(out)], env=dict(os.environ, CLS_HOST_THREADS="6", CLS_SEQ_BATCH=batch),
                           capture_output=True, text=True, timeout=600)
        if "FATAL: ThreadSanitizer" in r.stderr:
            pytest.skip(r.stderr[-300:])
        assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:3000]
        assert r.returncode == 0 and r.stdout.strip() == "n=1500", (r.returncode, r.stdout, r.stderr[-500:])
        outs.append(((tmp_path / f"out_{batch}" / "r.yaml").read_bytes(), (tmp_path / f"out_{batch}" / "r.error").read_bytes()))
    assert outs[0] == outs[1] == outs[2] and len(outs[0][0]) > 100_000 and len(outs[0][1]) > 0
