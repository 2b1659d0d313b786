"""The two independent CPU restatements (oracle/classeq_oracle.py, set-of-ints Python;
oracle/classeq_oracle.cpp, multithreaded C++) must agree - placement decisions are "parity
unpinned" against the Rust reference, so they are pinned by this agreement.  CPU only."""
import numpy as np
import pytest

from helpers import expected_row, outcome_of


@pytest.fixture(scope="module")
def cpp():
    from oracle import cpp_oracle
    return cpp_oracle


def _batch(seqs):
    bs = [s.encode() for s in seqs]
    off = np.zeros(len(bs) + 1, np.uint64)
    off[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs), np.uint8), off


def _check(out, outcomes):
    bad = []
    for i, e in enumerate(outcomes):
        want = expected_row(e)
        got = {k: out[k][i].item() for k in want}
        if got != want:
            bad.append((i, got, want))
    assert not bad, bad[:4]


def test_cpp_murmur_and_windows(cpp, oracle, pins):
    for s, h in pins["murmur3_h1"].items():
        assert cpp.murmur3_x64_128(s.encode())[0] == h
    rng = np.random.default_rng(11)
    for n in list(range(0, 50)) + [150]:
        b = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        assert cpp.murmur3_x64_128(b, 0) == oracle.murmurhash3_x64_128(b, 0)
        assert cpp.murmur3_x64_128(b, 99) == oracle.murmurhash3_x64_128(b, 99)
    s = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, 163))
    for k in (1, 4, 16, 35, 64):
        assert cpp.kmer_hashes(s.encode(), k).tolist() == [h for _, h in oracle.KmersMap(k, 0).build_kmer_from_string(s)]


@pytest.mark.parametrize("knob", ["default", "remove_intersection", "cov1", "iter2"])
def test_cpp_matches_committed_outcomes(cpp, col_flat, col_queries, col_expected, knob):
    kn = next(k for k in col_expected["knobs"] if k["name"] == knob)
    md = cpp.CppModel.from_flat(col_flat)
    bases, off = _batch([s for _, s in col_queries])
    out = md.place_batch(bases, off, kn["max_iterations"], kn["min_match_coverage"], kn["remove_intersection"])
    _check(out, col_expected["outcomes"][knob])
    one_thread = md.place_batch(bases, off, kn["max_iterations"], kn["min_match_coverage"], kn["remove_intersection"], 1)
    assert all((out[k] == one_thread[k]).all() for k in out)


@pytest.mark.parametrize("seed", range(40))
def test_cpp_matches_python_on_random_models(cpp, oracle, seed):
    import classeq2_b200 as cq
    from test_gpu_parity import _random_tree_model, tree_to_product
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.choice([35, 35, 35, 5, 11, 16, 21, 32, 40]))
    m = int(rng.choice([4, 4, 0, 1, 2, 7, k + 3 if k < 9 else 3]))
    tree, queries = _random_tree_model(oracle, rng, k, m)
    if rng.random() < 0.1:
        tree.root.children = None
    md = cpp.CppModel.from_flat(tree_to_product(cq, tree))
    bases, off = _batch([s for _, s in queries])
    for kn in [dict(), dict(remove_intersection=True), dict(min_match_coverage=1.0),
               dict(max_iterations=int(rng.integers(0, 3)), min_match_coverage=0.0)]:
        out = md.place_batch(bases, off, kn.get("max_iterations"), kn.get("min_match_coverage"), kn.get("remove_intersection"))
        _check(out, [outcome_of(oracle, h, s, tree, kn.get("max_iterations"), kn.get("min_match_coverage"),
                                kn.get("remove_intersection")) for h, s in queries])


def test_cpp_matches_python_on_synthetic(cpp, oracle):
    from classeq2_b200 import synth
    from test_gpu_parity import oracle_tree_from_flat
    sm = synth.make_model(60, 300, 4242)
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, 200, 150, 4244)
    md = cpp.CppModel.from_flat(sm.flat)
    otree = oracle_tree_from_flat(oracle, sm.flat)
    seqs = [bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode() for i in range(200)]
    for ri in (False, True):
        out = md.place_batch(bases, offsets, remove_intersection=ri)
        _check(out, [outcome_of(oracle, "r", s, otree, None, None, ri) for s in seqs])


@pytest.mark.parametrize("n_tips, l_ref, seed", [(12, 80, 3), (60, 300, 4242), (300, (200, 260), 9)])
def test_oracle_builder_equals_product_builder(cpp, n_tips, l_ref, seed):
    """The C++ oracle's own k-mer map builder (build_database/mod.rs:62, :140-168; it is what bench.py's reference arm
    builds its model with) against the library's host builder: the same (bucket key, hash) -> node ids map."""
    from classeq2_b200 import synth
    sm = synth.make_model(n_tips, l_ref, seed)
    b, o = synth.refs_to_batch(sm.ref_codes, sm.ref_lens)
    t = sm.tree
    f = cpp.build_model(35, 4, t.node_id, t.node_kind, t.child_off, t.child_idx, t.tip_node, b, o, n_threads=3)

    def as_map(F):
        so = F.set_off.astype(np.int64)
        return {(int(F.entry_bucket[e]), int(F.entry_hash[e])): tuple(sorted(F.set_node_ids[so[int(F.entry_set[e])]:so[int(F.entry_set[e]) + 1]].tolist()))
                for e in range(len(F.entry_hash))}

    assert as_map(f) == as_map(sm.flat)
    assert len(f.set_off) == len(sm.flat.set_off)   # the same number of distinct node sets
    # and placements against either model agree
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, 100, 150, seed + 2)
    m1, m2 = cpp.CppModel.from_flat(f), cpp.CppModel.from_flat(sm.flat)
    o1, o2 = m1.place_batch(bases, offsets), m2.place_batch(bases, offsets)
    for k in o1:
        assert (o1[k] == o2[k]).all(), k
