"""cls_place_batch with the 2-bit packing on the device (pack_kernels.cu, cls_set_pack_mode(2)): the caller's ASCII
bases cross PCIe as they are - from pageable memory through the staging ring, or straight from pinned memory - and the
results are those of the host packer (and of the oracle), field for field, whatever the lengths, cases and invalid bytes."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_rows_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cq():
    import classeq2_b200
    return classeq2_b200


@pytest.fixture()
def pack_mode():
    from classeq2_b200 import _lib
    prev = []

    def set_mode(m):
        prev.append(_lib.check(_lib.lib.cls_set_pack_mode(m)))
    yield set_mode
    _lib.lib.cls_set_pack_mode(0)


def _fields_equal(cq, a, b):
    for f, _ in cq.engine.RESULT_DTYPES:
        assert (getattr(a, f) == getattr(b, f)).all(), f


def test_device_pack_equals_host_pack_golden(cq, col_flat, col_queries, col_expected, pack_mode):
    from classeq2_b200 import _lib
    ix = cq.Index(col_flat, device=0)
    good = [s for _, s in col_queries]
    seqs = good + [good[0].lower(), good[1][:70] + "N" + good[1][71:], "", "ACGT", good[2][:34], good[2][:35], "acgtn" * 30,
                   good[3][:100] + "-" + good[3][100:]]
    pack_mode(1)
    host = ix.place_batch(seqs)
    assert ix.timing()["pack_on_device"] == 0
    pack_mode(2)
    dev = ix.place_batch(seqs)
    tm = ix.timing()
    assert tm["pack_on_device"] == 1 and tm["h2d_bytes"] >= sum(len(s) for s in seqs)
    _fields_equal(cq, host, dev)
    assert_rows_equal(ix.place_batch(good), col_expected["outcomes"]["default"])   # still in device mode
    n = len(good)
    assert dev.status[n + 1] == _lib.STATUS_ERR_INVALID_BASE and dev.status[n + 6] == _lib.STATUS_ERR_INVALID_BASE
    assert dev.status[n + 7] == _lib.STATUS_ERR_INVALID_BASE and dev.row(n) == dev.row(0)
    ix.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_device_pack_random_batches(cq, seed, pack_mode):
    """Ragged batches (lengths 0 .. 2 500: every geometry, several length classes, sequences that are not in input
    order on the device), pageable and pinned caller memory, a batch that starts in the middle of the bases array."""
    import torch
    from classeq2_b200 import synth
    rng = np.random.default_rng(seed)
    sm = synth.make_model(40, 400, 100 + seed)
    n = 3000
    lens = np.where(rng.random(n) < 0.7, rng.integers(0, 300, n), rng.integers(300, 2500, n))
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, n, lens, 7 + seed)
    bases = bases.copy()
    # lower case, and a few bytes that are not bases
    low = rng.random(len(bases)) < 0.2
    bases[low] |= 0x20
    for p in rng.integers(0, len(bases), 25):
        bases[p] = rng.choice(np.frombuffer(b"NnXx-.*\x00\xff", np.uint8))
    ix = cq.Index(sm.flat, device=0)
    pack_mode(1)
    host = cq.BatchResult(n)
    ix.place_batch_into(bases, offsets, host)
    pack_mode(2)
    dev = cq.BatchResult(n)
    ix.place_batch_into(bases, offsets, dev)
    assert ix.timing()["pack_on_device"] == 1
    _fields_equal(cq, host, dev)
    # pinned caller memory: copied straight from it
    pin = torch.empty(len(bases), dtype=torch.uint8).pin_memory()
    pb = pin.numpy()
    pb[:] = bases
    dev2 = cq.BatchResult(n)
    ix.place_batch_into(pb, offsets, dev2)
    assert ix.timing()["pack_on_device"] == 2
    _fields_equal(cq, host, dev2)
    # a batch that is a slice of a larger one (offsets[0] != 0)
    a, b = 500, 2200
    sub_h, sub_d = cq.BatchResult(b - a), cq.BatchResult(b - a)
    ix.place_batch_into(pb, offsets[a:b + 1], sub_d)
    pack_mode(1)
    ix.place_batch_into(pb, offsets[a:b + 1], sub_h)
    _fields_equal(cq, sub_h, sub_d)
    for f, _ in cq.engine.RESULT_DTYPES:
        assert (getattr(sub_d, f) == getattr(host, f)[a:b]).all(), f
    ix.close()


def test_device_pack_large_batch_many_pieces(cq, pack_mode):
    """More bases than one 32 MiB copy piece and several chunks: the pieces, the chunks and the result scatter line up."""
    from classeq2_b200 import synth
    sm = synth.make_model(200, 600, 11)
    n = 400_000
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, n, 150, 13)
    ix = cq.Index(sm.flat, device=0)
    pack_mode(1)
    host = cq.BatchResult(n)
    ix.place_batch_into(bases, offsets, host)
    pack_mode(2)
    dev = cq.BatchResult(n)
    ix.place_batch_into(bases, offsets, dev)
    tm = ix.timing()
    assert tm["pack_on_device"] == 1 and tm["h2d_bytes"] >= len(bases)
    _fields_equal(cq, host, dev)
    # mixed: some chunks packed by the host, the others on the device; a few invalid bases in both kinds of chunk
    bad = bases.copy()
    for r in (5, n // 3, n // 2, n - 9):
        bad[int(offsets[r]) + 17] = ord("N")
    pack_mode(1)
    ix.place_batch_into(bad, offsets, host)
    pack_mode(3)
    ix.place_batch_into(bad, offsets, dev)
    tm = ix.timing()
    assert tm["pack_on_device"] == 3 and 0 < tm["h2d_bytes"] < len(bases) + 20 * n
    _fields_equal(cq, host, dev)
    from classeq2_b200 import _lib
    assert (dev.status == _lib.STATUS_ERR_INVALID_BASE).sum() == 4
    ix.close()


def test_pack_mode_argument_is_checked(cq):
    from classeq2_b200 import _lib
    assert _lib.lib.cls_set_pack_mode(4) < 0 and _lib.lib.cls_set_pack_mode(-1) < 0
    assert _lib.lib.cls_set_pack_mode(0) >= 0


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_just_in_time_plan_and_its_fallback(cq, pack_mode, mode):
    """Batches whose reads all have 35 .. 162 bases are planned chunk by chunk while the GPU already works
    (plan_reads_fast); one read outside that range anywhere in the batch sends the call back to the general plan.
    Either way the results are those of the resident path, which always plans the whole batch first."""
    from classeq2_b200 import synth
    rng = np.random.default_rng(21)
    sm = synth.make_model(60, 500, 31)
    ix = cq.Index(sm.flat, device=0)
    pack_mode(mode)
    n = 150_000
    for case in ("uniform", "ragged", "late_short", "late_long", "late_invalid"):
        lens = np.full(n, 150) if case == "uniform" else rng.integers(35, 163, n)
        if case == "late_short":
            lens[n - 7] = 12
        if case == "late_long":
            lens[n - 70_000] = 700
        bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, n, lens, 5)
        if case == "late_invalid":
            bases = bases.copy()
            bases[int(offsets[n - 3]) + 5] = ord("N")
        got = cq.BatchResult(n)
        ix.place_batch_into(bases, offsets, got)
        rb = ix.upload((bases, offsets))
        rb.place()
        want = rb.fetch()
        rb.close()
        for f, _ in cq.engine.RESULT_DTYPES:
            assert (getattr(got, f) == getattr(want, f)).all(), (case, f)
    ix.close()
