"""ONE handle over several GPUs behind the C ABI (cls_index_create_devices / cls_index_create_multi): cls_place_batch
cuts the batch into one contiguous part per replica and runs them side by side; the results must equal the
single-device call field by field.  Replicas may share a GPU, so the split, the per-part threads and the error
propagation are exercised on a one-GPU box too; with several GPUs visible every device gets a replica."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched", "iterations")


@pytest.fixture(scope="module")
def cq():
    import classeq2_b200
    return classeq2_b200


@pytest.fixture(scope="module")
def workload(cq):
    from classeq2_b200 import synth
    sm = synth.make_model(80, 400, 2024)
    lens = np.concatenate([np.full(3000, 150), synth.skewed_lengths(600, 3) // 4 + 35, np.array([10, 34, 35, 0, 700, 1200])])
    rng = np.random.default_rng(1)
    rng.shuffle(lens)
    b, o, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, len(lens), lens, 2026)
    b = b.copy()
    b[int(o[7]) + 3] = ord("N")       # an invalid base: a per-query status, in whichever part the query lands
    return sm, b, o


def _same(a, b):
    for f in FIELDS:
        bad = np.flatnonzero(getattr(a, f) != getattr(b, f))
        assert bad.size == 0, (f, bad[:5], getattr(a, f)[bad[:5]], getattr(b, f)[bad[:5]])


@pytest.mark.parametrize("replicas", [1, 2, 3, 7])
def test_replicas_equal_single_device(cq, workload, replicas):
    import torch
    sm, b, o = workload
    n_gpu = torch.cuda.device_count()
    single = cq.Index(sm.flat, device=0)
    want = single.place_batch((b, o))
    devs = [d % n_gpu for d in range(replicas)]
    multi = cq.Index(sm.flat, devices=devs)
    inf = multi.info()
    assert inf["n_devices"] == replicas and inf["n_entries"] == single.info()["n_entries"]
    for kn in (dict(), dict(remove_intersection=True, min_match_coverage=0.3)):
        want = single.place_batch((b, o), cq.PlaceParams(**kn))
        _same(multi.place_batch((b, o), cq.PlaceParams(**kn)), want)
    # fewer queries than replicas, an empty batch, one query
    for n in (0, 1, 2, replicas - 1 if replicas > 1 else 1):
        sub = (b[: int(o[n])], o[: n + 1])
        _same(multi.place_batch(sub), single.place_batch(sub))
    assert multi.timing()["kernel_launches"] >= 1
    multi.close(), single.close()


def test_mask_covers_all_visible_devices(cq, workload):
    import torch
    sm, b, o = workload
    n_gpu = torch.cuda.device_count()
    single = cq.Index(sm.flat, device=0)
    multi = cq.Index(sm.flat, device_mask=0)
    assert multi.info()["n_devices"] == n_gpu
    _same(multi.place_batch((b, o)), single.place_batch((b, o)))
    with pytest.raises(cq._lib.ClsError):
        cq.Index(sm.flat, device_mask=1 << 40)
    # resident batches of a multi-device handle live on its first device
    rb = multi.upload((b, o))
    rb.place()
    _same(rb.fetch(), single.place_batch((b, o)))
    rb.close(), multi.close(), single.close()
