"""``cls_model_build_device`` (build-db's k-mer map on the GPU) against the host builder
``cls_model_build`` - which tests/test_abi.py holds equal to the oracle's restatement of map_kmers_to_tree
(build_database/mod.rs:26-181) - and, through an index made from the device-built model, against the committed
placement outcomes.  (The file sorts after the placement parity tests on purpose: `pytest -x` reaches those first.)"""
import os

import numpy as np
import pytest

from helpers import assert_built_equal, assert_rows_equal, random_build_case

pytestmark = pytest.mark.gpu


def _build(case, device=None):
    from classeq2_b200.model import BuiltModel
    bm = BuiltModel(*case, device=device)
    a = bm.arrays()
    bm.close()
    return a


@pytest.mark.parametrize("k,m", [(35, 4), (21, 0), (5, 7), (16, 2), (33, 4), (64, 4)])
def test_device_builder_equals_host_builder_random(k, m):
    rng = np.random.default_rng(100 * k + m)
    for it in range(6):
        case = random_build_case(rng, n_internal=int(rng.integers(0, 25)), k=k, m=m, max_len=int(rng.integers(k, 260)),
                                 dup_tips=it % 3, internal_tips=it % 2,
                                 letters=b"ACGT" if it % 2 == 0 else b"ACGTacgt")
        assert_built_equal(_build(case), _build(case, device=0))


def test_device_builder_multi_tile_sequences():
    rng = np.random.default_rng(7)
    for max_len in (2048 + 34, 2048 + 35, 5000):
        case = random_build_case(rng, n_internal=3, k=35, m=4, max_len=max_len)
        assert_built_equal(_build(case), _build(case, device=0))


def test_device_builder_degenerate_inputs():
    from classeq2_b200 import _lib
    from classeq2_b200.model import FlatModel
    tflat = FlatModel(35, 4, np.array([5, 9], np.uint64), np.array([_lib.KIND_ROOT, _lib.KIND_LEAF], np.uint8),
                      np.array([0, 1, 1], np.uint64), np.array([1], np.uint64))
    for seqs in ([], [b"ACGT"], [b"ACGTTGCATGCATGACTGACTGATCGATCGATGCA"]):
        offsets = np.zeros(len(seqs) + 1, np.uint64)
        if seqs:
            offsets[1:] = np.cumsum([len(s) for s in seqs])
        bases = np.frombuffer(b"".join(seqs) or b"\0", np.uint8).copy()
        case = (tflat, np.ones(len(seqs), np.uint64), bases, offsets)
        assert_built_equal(_build(case), _build(case, device=0))
    with pytest.raises(_lib.ClsError) as ei:
        _build((tflat, np.ones(1, np.uint64), np.frombuffer(b"ACGT", np.uint8).copy(), np.array([0, 4], np.uint64)), device=99)
    assert ei.value.code == _lib.CLS_ERR_INVALID_ARGUMENT


def test_device_built_colletotrichum_model_places_like_the_golden_one(col_queries, col_tree, col_npz, col_expected):
    """Device-built map == host-built map on the reference's Colletotrichum inputs, and an index uploaded from it
    places the 363 committed queries exactly as the committed outcomes say."""
    import classeq2_b200 as cq
    from classeq2_b200.model import BuiltModel, FlatModel
    z = col_npz
    tflat = FlatModel(35, 4, z["node_id"], z["node_kind"], z["child_off"], z["child_idx"])
    idx = {c.name: i for i, c in enumerate(col_tree.root.walk()) if c.is_leaf()}
    tips = col_queries[:171]
    tip_node = np.array([idx[h] for h, _ in tips], np.uint64)
    bases = np.frombuffer("".join(s for _, s in tips).encode(), np.uint8).copy()
    offsets = np.zeros(len(tips) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for _, s in tips])
    case = (tflat, tip_node, bases, offsets)
    assert_built_equal(_build(case), _build(case, device=0))
    bm = BuiltModel(*case, device=0)
    a = bm.arrays()
    bm.close()
    assert len(a["set_off"]) - 1 == 226 and len(a["entry_hash"]) == len(z["entry_hash"])
    flat = FlatModel(35, 4, z["node_id"], z["node_kind"], z["child_off"], z["child_idx"], a["entry_bucket"], a["entry_hash"],
                     a["entry_set"], a["set_off"], a["set_node_ids"])
    index = cq.Index(flat, device=0)
    res = index.place_batch([s for _, s in col_queries])
    by_header = dict(zip(col_expected["queries"], col_expected["outcomes"]["default"]))
    assert_rows_equal(res, [by_header[h] for h, _ in col_queries], [h for h, _ in col_queries])


def test_device_builder_synthetic_model():
    """A config-2-like synthetic model (300 tips x 1 kb): same arrays as the host builder."""
    from classeq2_b200 import synth
    tree = synth.make_tree(300, 4321)
    codes, lens = synth.make_refs(tree, 1000, 4322)
    tflat = synth.tree_only_flat(tree, 35, 4)
    bases, offs = synth.refs_to_batch(codes, lens)
    case = (tflat, tree.tip_node, bases, offs)
    a = _build(case, device=0)
    assert_built_equal(_build(case), a)
    assert len(a["entry_hash"]) > 100_000


def test_map_kmers_to_tree_on_the_device(tmp_path, col_queries):
    from classeq2_b200 import build
    nwk = "Colletotrichum_acutatum_gapdh-PhyML.nwk"
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", nwk)
    msa = tmp_path / "tips.fasta"
    msa.write_text("".join(f">{h}\n{s}\n" for h, s in col_queries[:171]))
    host = build.map_kmers_to_tree(golden, msa, pairing="own")
    dev = build.map_kmers_to_tree(golden, msa, device=0, pairing="own")
    assert host.to_obj() == dev.to_obj()


def test_device_builder_reproduces_the_reference_written_model(col_queries):
    """Fed with the reference's pairing, the device builder gives the model the reference itself wrote
    (tests/golden/reference_built_model_k12.json.gz; closed under reverse complement: it predates v0.2.3)."""
    from helpers import built_as_map, reference_built_case
    pin, tflat, tip_node, bases, offsets, want = reference_built_case(col_queries[:171])
    assert built_as_map(_build((tflat, tip_node, bases, offsets), device=0)) == want


def test_placement_against_the_reference_written_model(cq_mod, oracle, col_queries):
    """The model the reference wrote (k = 12, one bucket, forward-strand node sets) uploaded as it is: the 171 tip
    sequences, their reverse complements and the other committed queries are placed exactly as the oracle places
    them against the same model - the generic-k kernels on reference-authored node sets."""
    from helpers import load_reference_built_model, outcome_of
    cq = cq_mod
    pin = load_reference_built_model()
    otree = oracle.Tree.from_obj({"id": pin["id"], "name": pin["name"], "minBranchSupport": 70.0, "root": _with_parents(pin["root"])})
    km = oracle.KmersMap(pin["k_size"], 0)
    km.map[0] = {oracle.hash_kmer(s): set(ids) for s, ids in pin["kmers"].items()}
    otree.kmers_map = km
    flat = cq.FlatModel.from_tree(cq.Tree.from_obj(otree.to_obj()))
    queries = [(h, s) for h, s in col_queries if s][:230]
    queries += [(h + "_rc", oracle.KmersMap.reverse_complement(s)) for h, s in queries[:40]]
    for f in (flat, flat.with_general_sets()):
        ix = cq.Index(f, device=0)
        for kn in (dict(), dict(remove_intersection=True)):
            want = [outcome_of(oracle, h, s, otree, None, None, kn.get("remove_intersection")) for h, s in queries]
            assert_rows_equal(ix.place_batch([s for _, s in queries], cq.PlaceParams(**kn)), want, [h for h, _ in queries])
        ix.close()


def _with_parents(node, parent=None):
    out = dict(node, parent=parent)
    if "children" in node:
        out["children"] = [_with_parents(c, node["id"]) for c in node["children"]]
    return out


@pytest.fixture(scope="module")
def cq_mod():
    import classeq2_b200 as cq
    return cq
