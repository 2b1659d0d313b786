"""Queries of any length get a per-query result (the reference has no length limit: kmers_map.rs:375-424, and isolates
failures per query: place_sequences/mod.rs:160-169).  Reads whose tables exceed the shared memory of an SM (beyond
about 4 kb at k = 35) go through the global-memory tables of csrc/giant_kernels.cuh; here 5 kb, 50 kb and 1 Mb
queries share a batch with 150-base reads and with lengths around the geometry boundary, against the C++ oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched", "iterations")


@pytest.fixture(scope="module")
def cq():
    import classeq2_b200
    return classeq2_b200


@pytest.fixture(scope="module")
def long_model(cq):
    from classeq2_b200 import synth
    sm = synth.make_model(8, 60_000, 77)
    tips = ["".join("ACGT"[c] for c in sm.ref_codes[t]) for t in range(8)]
    return sm, tips


def _check(cq, flat, qs, knobs=(dict(), dict(remove_intersection=True, min_match_coverage=0.2))):
    from oracle import cpp_oracle
    md = cpp_oracle.CppModel.from_flat(flat)
    b, o = cq.make_batch(qs)
    ix = cq.Index(flat, device=0)
    ixg = cq.Index(flat.with_general_sets(), device=0)
    for kn in knobs:
        want = md.place_batch(b, o, kn.get("max_iterations"), kn.get("min_match_coverage"), kn.get("remove_intersection"))
        rb = ix.upload((b, o))
        rb.place(cq.PlaceParams(**kn))
        for got in (ix.place_batch((b, o), cq.PlaceParams(**kn)), rb.fetch(), ixg.place_batch((b, o), cq.PlaceParams(**kn))):
            for f in FIELDS:
                bad = np.flatnonzero(getattr(got, f) != want[f])
                assert bad.size == 0, (f, kn, bad[:5], [len(qs[i]) for i in bad[:5]], getattr(got, f)[bad[:5]], want[f][bad[:5]])
        rb.close()
    md.close(), ix.close(), ixg.close()
    return want


def test_5kb_50kb_1mb_queries_in_a_mixed_batch(cq, long_model):
    sm, tips = long_model
    rng = np.random.default_rng(3)
    rnd = lambda n: "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rc = lambda s: "".join(comp[c] for c in reversed(s))
    qs = [tips[0][:5000], tips[3][100:50_100], rc(tips[5][2000:52_000]),
          "".join(tips) + rnd(1_000_000 - 8 * 60_000),               # 1 Mb: every tip once, then unrelated sequence
          tips[2][:30_000] + tips[2][:30_000] + rc(tips[2][:30_000]),  # the same k-mers three times, both strands
          rnd(7000)]                                                   # long and unrelated
    for L in (150, 151, 289, 290, 291, 1000, 3500, 4000, 4100, 4150, 4200, 4500, 6000, 9000):
        st = int(rng.integers(0, 60_000 - L))
        qs.append(tips[int(rng.integers(8))][st:st + L])
    qs += [tips[1][i:i + 150] for i in range(0, 3000, 100)]
    order = rng.permutation(len(qs))
    qs = [qs[i] for i in order]
    want = _check(cq, sm.flat, qs)
    assert (want["status"] == 6).sum() + (want["status"] == 5).sum() > len(qs) // 2
    assert want["n_matched"].max() > 400_000


def test_long_reads_other_k(cq):
    """k != 35 (generic hashing) with reads beyond the shared-memory geometry."""
    from classeq2_b200 import synth
    sm = synth.make_model(6, 20_000, 78, k_size=21, m_size=3)
    tips = ["".join("ACGT"[c] for c in sm.ref_codes[t]) for t in range(6)]
    qs = [tips[0][:12_000], tips[1][500:9_500], tips[2][:150], tips[3] + tips[4], tips[5][:3000]]
    _check(cq, sm.flat, qs, knobs=(dict(),))
