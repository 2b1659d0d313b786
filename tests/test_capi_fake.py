"""The HOST side of the hot path's public call - ``cls_place_batch`` and its neighbours in classeq2_b200/csrc/capi.cu:
planning (general and just in time, with its fallback), host / device / mixed packing from pageable and pinned memory,
the chunked pipeline, scatter into the caller's arrays, resident batches, one handle over several devices, concurrent
callers, a failing launch - end to end WITHOUT a GPU.  capi.cu is compiled with g++ against a fake CUDA runtime
(tests/native/fakecuda/cuda_runtime.h: device memory is host memory, every call completes before it returns); the
placement launch unpacks the reads it is handed and gives them to the C++ oracle (tests/native/fake_kernels.cpp); the
results must equal the oracle's on the caller's ASCII batch (tests/native/capi_fake_main.cpp).  Run under
AddressSanitizer + UBSan (every scenario: a copy past a "device" buffer is a report) and under ThreadSanitizer (the
scenarios with several host threads).  The kernels themselves are what the ``-m gpu`` tests are for."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "classeq2_b200", "csrc")
FAKE = os.path.join(HERE, "native", "fakecuda")
SOURCES = [(os.path.join(CSRC, "capi.cu"), True)] + [(os.path.join(CSRC, f), False) for f in
                                                      ("index_build.cpp", "host_api.cpp", "host_pack.cpp", "host_pool.cpp", "record_writer.cpp")] + \
          [(os.path.join(HERE, "native", f), False) for f in ("capi_fake_main.cpp", "fake_kernels.cpp")]


def _build(tmp_path, name, san):
    """One object per source, compiled side by side; the oracle is compiled once, without a sanitizer (it is the checker,
    and its allocations dominate the run time under one)."""
    out = tmp_path / name
    out.mkdir()
    jobs = []
    orc = tmp_path / "orc.o"
    if not orc.exists():
        jobs.append((subprocess.Popen(["g++", "-O2", "-std=c++17", "-pthread", "-c", os.path.join(ROOT, "oracle", "classeq_oracle.cpp"), "-o", str(orc)],
                                      stderr=subprocess.PIPE, text=True), str(orc)))
    objs = []
    for src, is_cu in SOURCES:
        obj = str(out / (os.path.basename(src) + ".o"))
        objs.append(obj)
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread", *san, "-I", FAKE] + (["-x", "c++"] if is_cu else []) + ["-c", src, "-o", obj]
        jobs.append((subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True), obj))
    for p, obj in jobs:
        _, err = p.communicate()
        assert "error:" not in err, err[-3000:]                      # a real build error is a failure
        if p.returncode != 0:
            pytest.skip(f"cannot build with {san} here: {err[-300:]}")
    exe = str(out / "capi_fake")
    r = subprocess.run(["g++", "-pthread", *san, *objs, str(orc), "-o", exe], capture_output=True, text=True)
    assert "undefined reference" not in r.stderr, r.stderr[-3000:]
    if r.returncode != 0:
        pytest.skip(f"cannot link with {san} here: {r.stderr[-300:]}")
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_place_batch_host_side_equals_the_oracle_under_asan(tmp_path):
    exe = _build(tmp_path, "asan", ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and r.stdout.startswith("bad=0 "), (r.returncode, r.stdout, r.stderr[-1500:])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_place_batch_host_threads_are_clean_under_tsan(tmp_path):
    exe = _build(tmp_path, "tsan", ["-fsanitize=thread"])
    r = subprocess.run([exe, "threads"], capture_output=True, text=True, timeout=900)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and r.stdout.startswith("bad=0 "), (r.returncode, r.stdout, r.stderr[-1500:])
