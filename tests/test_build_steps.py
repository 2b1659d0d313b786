"""The index arithmetic of the device model builder (classeq2_b200/csrc/build_steps.hpp, build_prep.hpp), walked
element by element on the CPU by the test-only library tests/native/libbuild_steps_host.so, against the host
builder ``cls_model_build``.  No GPU: the kernels themselves are held equal to the host builder in
tests/test_gpu_build.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_built_equal, random_build_case

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def walk():
    from classeq2_b200 import _lib
    subprocess.run(["make", "-C", os.path.join(HERE, "native")], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(HERE, "native", "libbuild_steps_host.so"))
    lib.bsh_model_build.restype = C.c_int
    lib.bsh_model_build.argtypes = [C.POINTER(_lib.ModelView), C.c_uint64, _lib.u64p, _lib.u8p, _lib.u64p, C.POINTER(C.c_void_p)]

    def run(tflat, tip_node, bases, offsets):
        from classeq2_b200.model import BuiltModel
        h = C.c_void_p()
        rc = lib.bsh_model_build(C.byref(tflat.view), len(tip_node), tip_node.ctypes.data_as(_lib.u64p),
                                 bases.ctypes.data_as(_lib.u8p), offsets.ctypes.data_as(_lib.u64p), C.byref(h))
        if rc != 0:
            raise _lib.ClsError(rc, "bsh_model_build")
        bm = BuiltModel.__new__(BuiltModel)       # read the handle with the product's own view call
        bm.tree_only, bm._h, bm.view = tflat, h, _lib.ModelView()
        _lib.check(_lib.lib.cls_built_model_view(h, C.byref(tflat.view), C.byref(bm.view)))
        a = bm.arrays()
        bm.close()
        return a
    return run


def _host(tflat, tip_node, bases, offsets):
    from classeq2_b200.model import BuiltModel
    bm = BuiltModel(tflat, tip_node, bases, offsets)
    a = bm.arrays()
    bm.close()
    return a


@pytest.mark.parametrize("k,m", [(35, 4), (21, 0), (5, 7), (16, 2), (33, 4)])
def test_walk_equals_host_builder_random(walk, k, m):
    rng = np.random.default_rng(100 * k + m)
    for it in range(6):
        case = random_build_case(rng, n_internal=int(rng.integers(0, 25)), k=k, m=m, max_len=int(rng.integers(k, 260)),
                                 dup_tips=it % 3, internal_tips=it % 2,
                                 letters=b"ACGT" if it % 2 == 0 else b"ACGTacgt")
        assert_built_equal(_host(*case), walk(*case))
        if it % 2:   # lower-case letters are the same k-mers (both strands are upper-cased, kmers_map.rs:410)
            tflat, tip_node, bases, offsets = case
            up = np.frombuffer(bytes(bases).upper(), np.uint8).copy()
            assert_built_equal(_host(tflat, tip_node, up, offsets), _host(*case))


def test_builders_reject_non_acgt(walk):
    """Anything but A / C / G / T (either case) in a tip sequence is an error in every builder - the reference
    panics (kmers_map.rs:431-443) - and so are decreasing offsets; never a silently different model."""
    from classeq2_b200._lib import ClsError
    rng = np.random.default_rng(5)
    tflat, tip_node, bases, offsets = random_build_case(rng, n_internal=4, k=35, m=4, max_len=120)
    bad = bases.copy()
    bad[len(bad) // 2] = ord("N")
    for builder in (_host, walk):
        with pytest.raises(ClsError):
            builder(tflat, tip_node, bad, offsets)
    off2 = offsets.copy()
    if len(off2) > 2:
        off2[1], off2[2] = off2[2], off2[1]
        if off2[1] != off2[2]:
            with pytest.raises(ClsError):
                _host(tflat, tip_node, bases, off2)


def test_walk_multi_tile_sequences(walk):
    """Sequences longer than one tile of the hashing kernel (2048 windows), with a tile of one window."""
    rng = np.random.default_rng(7)
    for max_len in (2048 + 34, 2048 + 35, 5000):
        case = random_build_case(rng, n_internal=3, k=35, m=4, max_len=max_len)
        a = _host(*case)
        assert len(a["entry_hash"]) > 2048
        assert_built_equal(a, walk(*case))


def test_walk_colletotrichum(walk, col_queries, col_tree, col_npz):
    from classeq2_b200.model import FlatModel
    z = col_npz
    tflat = FlatModel(35, 4, z["node_id"], z["node_kind"], z["child_off"], z["child_idx"])
    idx = {c.name: i for i, c in enumerate(col_tree.root.walk()) if c.is_leaf()}
    tips = col_queries[:171]
    tip_node = np.array([idx[h] for h, _ in tips], np.uint64)
    bases = np.frombuffer("".join(s for _, s in tips).encode(), np.uint8).copy()
    offsets = np.zeros(len(tips) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for _, s in tips])
    a = walk(tflat, tip_node, bases, offsets)
    assert_built_equal(_host(tflat, tip_node, bases, offsets), a)
    assert len(a["set_off"]) - 1 == 226


def test_walk_degenerate_inputs(walk):
    from classeq2_b200 import _lib
    from classeq2_b200.model import FlatModel
    tflat = FlatModel(35, 4, np.array([5, 9], np.uint64), np.array([_lib.KIND_ROOT, _lib.KIND_LEAF], np.uint8),
                      np.array([0, 1, 1], np.uint64), np.array([1], np.uint64))
    # no tips; one tip shorter than k; one tip of exactly k bases
    for seqs in ([], [b"ACGT"], [b"ACGTTGCATGCATGACTGACTGATCGATCGATGCA"]):
        offsets = np.zeros(len(seqs) + 1, np.uint64)
        if seqs:
            offsets[1:] = np.cumsum([len(s) for s in seqs])
        bases = np.frombuffer(b"".join(seqs) or b"\0", np.uint8).copy()
        tip_node = np.ones(len(seqs), np.uint64)
        a, b = _host(tflat, tip_node, bases, offsets), walk(tflat, tip_node, bases, offsets)
        assert_built_equal(a, b)
        assert len(b["entry_hash"]) == (2 if seqs and len(seqs[0]) == 35 else 0)


def test_builders_reproduce_the_reference_written_model(walk, col_queries):
    """cls_model_build and the CPU walk through the device builder's steps, fed with the reference's pairing, give
    the model the reference itself wrote (closed under reverse complement: it predates v0.2.3)."""
    from helpers import built_as_map, reference_built_case
    pin, tflat, tip_node, bases, offsets, want = reference_built_case(col_queries[:171])
    host = _host(tflat, tip_node, bases, offsets)
    assert set(host["entry_bucket"].tolist()) == {0}
    assert built_as_map(host) == want
    assert built_as_map(walk(tflat, tip_node, bases, offsets)) == want


def test_map_kmers_to_tree_reference_pairing(tmp_path, oracle, col_queries):
    """build.map_kmers_to_tree(pairing="reference") == the oracle's restatement with the reference's pairing."""
    import os
    from classeq2_b200 import build
    nwk = "Colletotrichum_acutatum_gapdh-PhyML.nwk"
    golden = os.path.join(HERE, "golden", nwk)
    tips = col_queries[:171]
    msa = tmp_path / "tips.fasta"
    msa.write_text("".join(f">{h}\n{s}\n" for h, s in tips))
    got = build.map_kmers_to_tree(golden, msa, pairing="reference")
    want = oracle.tree_from_newick(open(golden).read(), nwk, 70.0)
    oracle.map_kmers_to_tree(want, tips, 35, 4, pairing="reference")
    assert {k: {h: set(n) for h, n in v.items()} for k, v in got.kmers_map.map.items()} == \
           {k: {h: set(n) for h, n in v.items()} for k, v in want.kmers_map.map.items()}
    own = build.map_kmers_to_tree(golden, msa, pairing="own")
    assert own.kmers_map.map != got.kmers_map.map


def test_toy_newick_of_the_reference(oracle):
    """core/src/tests/data/tree.nwk (six tips, supports written as fractions): the mirror's parser / sanitize and the
    oracle's agree for thresholds below, between and above the supports; above them everything collapses into the root."""
    from classeq2_b200 import build
    toy = "(((A:0.1,B:0.2)0.7:0.5,(E:0.1,F:0.2)0.95:0.3)0.98:0.1,(C:0.3,D:0.4)0.99:0.5);"
    for thr in (0.0, 0.5, 0.8, 0.97, 70.0):
        a, b = build.tree_from_newick(toy, "tree.nwk", thr), oracle.tree_from_newick(toy, "tree.nwk", thr)
        assert a.root.to_obj() == b.root.to_obj() and a.id == b.id
    star = build.tree_from_newick(toy, "tree.nwk", 70.0).root
    assert [c.name for c in star.children] == ["A", "B", "E", "F", "C", "D"] and all(c.is_leaf() for c in star.children)
    full = build.tree_from_newick(toy, "tree.nwk", 0.0).root
    assert [c.id for c in full.walk()] == list(range(11))        # phylotree numbers the nodes in pre-order


def test_build_loop_line_rules(tmp_path):
    """The MSA loop of the reference's builder (build_database/mod.rs:84-117), which is NOT the place_sequences reader:
    every '>' line sends the new header with the sequence accumulated before it, leading sequence lines go to the first
    header, the last sequence is never sent, an unknown header is an error (the reference panics, mod.rs:136-139)."""
    import os
    import pytest
    from classeq2_b200 import build
    text = "acgtn\n\n>t1\nAC-GT\nggNN\n>>t2>\r\n>t3\nTTTT\n"
    assert build.build_loop_records(text) == [("t1", "ACGT"), ("t2", "ACGTGG"), ("t3", "")]
    assert build.build_loop_records(">a\nAC\r\n>b") == [("a", ""), ("b", "AC")]
    assert build.build_loop_records("") == []
    nwk = "Colletotrichum_acutatum_gapdh-PhyML.nwk"
    golden = os.path.join(HERE, "golden", nwk)
    msa = tmp_path / "bad.fasta"
    msa.write_text(">not_a_tip\nACGT\n")
    with pytest.raises(ValueError, match="does not match any tree leaf"):
        build.map_kmers_to_tree(golden, msa)                      # the default is the reference's pairing
    assert build.map_kmers_to_tree(golden, msa, pairing="own").kmers_map.map == {}
