"""The host-side planner of ``cls_place_batch`` / ``cls_batch_upload`` (capi.cu: plan_batch) through
``cls_debug_plan_batch``: host-decided statuses, length classes, device order and packed layout.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from classeq2_b200 import _lib

TOO_SHORT = _lib.STATUS_ERR_TOO_SHORT


def plan(lens, k):
    lens = np.asarray(lens, np.uint64)
    n = len(lens)
    offsets = np.zeros(n + 1, np.uint64)
    offsets[1:] = np.cumsum(lens)
    bases = np.zeros(1, np.uint8)            # the planner never reads the bases
    b = _lib.Batch(n, bases.ctypes.data_as(_lib.u8p), offsets.ctypes.data_as(_lib.u64p))
    pre = np.zeros(max(n, 1), np.uint8)
    perm = np.zeros(max(n, 1), np.uint32)
    woff = np.zeros(n + 1, np.uint32)
    classes = (_lib.PlanClass * 64)()
    nc, nd, nw = C.c_uint32(), C.c_uint32(), C.c_uint64()
    rc = _lib.lib.cls_debug_plan_batch(k, C.byref(b), pre.ctypes.data_as(_lib.u8p), perm.ctypes.data_as(_lib.u32p),
                                       woff.ctypes.data_as(_lib.u32p), classes, 64, C.byref(nc), C.byref(nd), C.byref(nw))
    _lib.check(rc)
    return dict(pre=pre[:n], perm=perm[:nd.value], woff=woff[:nd.value + 1], n_words=nw.value,
                classes=[(classes[c].first, classes[c].count, classes[c].max_len) for c in range(nc.value)])


def class_of(length, k):
    """Table geometry shared by a class: log2 of the de-duplication table size for that length (capi.cu)."""
    h2, c = 4 * (length - k + 1), 6
    while (1 << c) < h2:
        c += 1
    return c


def check(lens, k):
    lens = np.asarray(lens, np.int64)
    p = plan(lens, k)
    on_dev = np.flatnonzero(lens >= k)
    assert (p["pre"] == np.where(lens >= k, 0xFF, TOO_SHORT)).all()
    assert sorted(p["perm"].tolist()) == on_dev.tolist()                         # every long-enough read, once
    first = 0
    prev_class = None
    for f, cnt, mx in p["classes"]:
        assert f == first and cnt > 0
        members = p["perm"][f:f + cnt]
        assert (np.diff(members.astype(np.int64)) > 0).all()                      # input order inside a class
        cl = {class_of(int(lens[i]), k) for i in members}
        assert len(cl) == 1                                                       # one table geometry per class
        assert mx == lens[members].max()
        if prev_class is not None:
            assert cl.pop() < prev_class                                          # longest class first
            prev_class = class_of(int(lens[members[0]]), k)
        else:
            prev_class = cl.pop()
        first += cnt
    assert first == len(on_dev)
    words = (lens[p["perm"]] + 15) // 16
    assert (p["woff"] == np.concatenate([[0], np.cumsum(words)])).all() and p["n_words"] == int(words.sum())
    return p


def test_plan_uniform_short_reads():
    p = check(np.full(100_000, 150), 35)
    assert len(p["classes"]) == 1 and (p["perm"] == np.arange(100_000)).all()     # one class, caller order


def test_plan_mixed_lengths_and_short_queries():
    rng = np.random.default_rng(3)
    for k in (35, 12, 5):
        lens = np.concatenate([rng.integers(0, 60, 3000), rng.integers(100, 400, 3000), rng.integers(1000, 3000, 500),
                               [k - 1, k, k + 1, 0, 150, 151, 161, 162]])
        rng.shuffle(lens)
        p = check(lens, k)
        assert len(p["classes"]) >= 4


def test_plan_many_blocks():
    """More reads than one planner block (2^15): the per-block histograms and the scan over (class, block)."""
    rng = np.random.default_rng(4)
    lens = rng.choice([20, 150, 150, 150, 300, 1500], size=200_001)
    check(lens, 35)


def test_plan_degenerate_batches():
    assert plan([], 35)["classes"] == [] and plan([], 35)["n_words"] == 0
    p = plan([3, 0, 34], 35)
    assert p["classes"] == [] and (p["pre"] == TOO_SHORT).all()
    offsets = np.array([0, 10, 5], np.uint64)                                      # decreasing offsets are refused
    bases = np.zeros(16, np.uint8)
    b = _lib.Batch(2, bases.ctypes.data_as(_lib.u8p), offsets.ctypes.data_as(_lib.u64p))
    nc, nd, nw = C.c_uint32(), C.c_uint32(), C.c_uint64()
    rc = _lib.lib.cls_debug_plan_batch(35, C.byref(b), None, None, None, None, 0, C.byref(nc), C.byref(nd), C.byref(nw))
    assert rc == _lib.CLS_ERR_INVALID_ARGUMENT
