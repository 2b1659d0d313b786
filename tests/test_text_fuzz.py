"""Memory-safety fuzz of the library's host code under AddressSanitizer and UndefinedBehaviorSanitizer: the text code
(FASTA reader, filter, record writer: tests/native/text_fuzz.cpp against csrc/record_writer.cpp) and the model code
(index serialisation of valid and malformed model views, host model builder: tests/native/model_fuzz.cpp against
csrc/index_build.cpp and csrc/host_api.cpp)."""
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
    assert "undefined reference" not in r.stderr and "error:" not in r.stderr, r.stderr[-2000:]   # a real build error is a failure
    if r.returncode != 0:
        pytest.skip(f"cannot build with the sanitizers here: {r.stderr[-300:]}")
    r = subprocess.run([exe, "100"], env=dict(os.environ, CLS_HOST_THREADS="4"), capture_output=True, text=True, timeout=600)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and r.stdout.strip() == "bad=0", (r.returncode, r.stdout, r.stderr[-500:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_model_serialisation_refuses_malformed_views_without_crashing(tmp_path):
    exe = str(tmp_path / "model_fuzz")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-pthread",
                        os.path.join(HERE, "native", "model_fuzz.cpp"), os.path.join(CSRC, "index_build.cpp"),
                        os.path.join(CSRC, "host_api.cpp"), os.path.join(CSRC, "host_pool.cpp"), "-o", exe], capture_output=True, text=True)
    assert "undefined reference" not in r.stderr and "error:" not in r.stderr, r.stderr[-2000:]   # a real build error is a failure
    if r.returncode != 0:
        pytest.skip(f"cannot build with the sanitizers here: {r.stderr[-300:]}")
    r = subprocess.run([exe, "200"], capture_output=True, text=True, timeout=600)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and r.stdout.startswith("ok="), (r.returncode, r.stdout, r.stderr[-500:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_packer_bodies_stay_inside_their_buffers(tmp_path):
    """Every body of the host 2-bit packer on exactly-sized heap buffers (lengths 0..700): no over-read, no word written
    past ceil(len / 16), all bodies equal, invalid bytes reported."""
    exe = str(tmp_path / "pack_fuzz")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        os.path.join(HERE, "native", "pack_fuzz.cpp"), os.path.join(CSRC, "host_pack.cpp"), "-o", exe],
                       capture_output=True, text=True)
    assert "undefined reference" not in r.stderr and "error:" not in r.stderr, r.stderr[-2000:]   # a real build error is a failure
    if r.returncode != 0:
        pytest.skip(f"cannot build with the sanitizers here: {r.stderr[-300:]}")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and r.stdout.startswith("bad=0"), (r.returncode, r.stdout, r.stderr[-500:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_every_table_entry_is_found_by_the_kernels_probe_walks(tmp_path):
    """The uploaded k-mer table against the two probe walks of the kernels restated on the host (tests/native/
    table_probe.cpp): overflow chains, the 8-bit chain filter, free-slot fillers, unreachable bucket keys, shards - on
    random, clustered and tiny-valued hashes."""
    exe = str(tmp_path / "table_probe")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-pthread",
                        os.path.join(HERE, "native", "table_probe.cpp"), os.path.join(CSRC, "index_build.cpp"),
                        os.path.join(CSRC, "host_api.cpp"), os.path.join(CSRC, "host_pool.cpp"), "-o", exe], capture_output=True, text=True)
    assert "undefined reference" not in r.stderr and "error:" not in r.stderr, r.stderr[-2000:]   # a real build error is a failure
    if r.returncode != 0:
        pytest.skip(f"cannot build with the sanitizers here: {r.stderr[-300:]}")
    r = subprocess.run([exe, "120"], capture_output=True, text=True, timeout=600)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and r.stdout.startswith("bad=0 "), (r.returncode, r.stdout, r.stderr[-500:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_closed_model_records_agree_with_the_tree_and_the_sets(tmp_path):
    """What the descent kernels read for a closed model - pre-order numbering and subtree intervals of the non-leaf nodes,
    the Euler tour + sparse table behind lca_depth_node, one terminal list per node set - against the tree and the sets
    themselves on random trees (tests/native/closed_records.cpp)."""
    exe = str(tmp_path / "closed_records")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-pthread",
                        os.path.join(HERE, "native", "closed_records.cpp"), os.path.join(CSRC, "index_build.cpp"),
                        os.path.join(CSRC, "host_api.cpp"), os.path.join(CSRC, "host_pool.cpp"), "-o", exe], capture_output=True, text=True)
    assert "undefined reference" not in r.stderr and "error:" not in r.stderr, r.stderr[-2000:]   # a real build error is a failure
    if r.returncode != 0:
        pytest.skip(f"cannot build with the sanitizers here: {r.stderr[-300:]}")
    r = subprocess.run([exe, "200"], capture_output=True, text=True, timeout=600)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0 and r.stdout.startswith("bad=0 "), (r.returncode, r.stdout, r.stderr[-500:])
