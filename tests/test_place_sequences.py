"""The host mirror of the reference's `place_sequences` (classeq2_b200/placement.py):
record shape and bytes against a result file the REFERENCE itself wrote
(tests/data/public/...fd7/output/result.yaml, v0.9.0, copied to tests/golden/gyrb_result_v090.yaml),
FASTA reader rules, model loading, and - on a GPU - the whole call against the oracle."""
import ctypes as C
import ctypes.util
import hashlib
import json
import os

import numpy as np
import pytest
import yaml

from conftest import GOLDEN


@pytest.fixture(scope="module")
def ps():
    from classeq2_b200 import placement
    return placement


def _load_tagged(text, ps):
    class L(yaml.SafeLoader):
        pass
    L.add_multi_constructor("!", lambda loader, suffix, node: ps.Tag(
        suffix, int(loader.construct_scalar(node)) if suffix == "Taxid" else loader.construct_scalar(node)))
    # keep floats as floats, ints as ints; YAML 1.1 would read `1e-6` as a string -> fix up below
    return list(yaml.load_all(text, Loader=L))


def _fix_floats(o):
    if isinstance(o, dict):
        return {k: (float(v) if k in ("length", "support") and v is not None else _fix_floats(v)) for k, v in o.items()}
    if isinstance(o, list):
        return [_fix_floats(v) for v in o]
    return o


def _golden_docs():
    text = open(os.path.join(GOLDEN, "gyrb_result_v090.yaml")).read()
    docs = text.split("---\n")[1:]
    return text, ["---\n" + d for d in docs]


def test_ryu_float(ps):
    cases = {72.0: "72.0", 1e-6: "1e-6", 1e-5: "0.00001", 0.000619153: "0.000619153", 1e-8: "1e-8", 100.0: "100.0",
             1e16: "1e16", 1.5e16: "1.5e16", 0.1: "0.1", 1234.5: "1234.5", 1e15: "1000000000000000.0", -2.5: "-2.5",
             0.0: "0.0", 1e-7: "1e-7", 123456.789: "123456.789"}
    for x, s in cases.items():
        assert ps.ryu_float(x) == s


def test_yaml_emitter_reproduces_reference_bytes(ps):
    """parse -> emit of every record of the reference-written result file is byte-identical."""
    text, docs = _golden_docs()
    objs = [_fix_floats(o) for o in _load_tagged(text, ps)]
    assert len(objs) == len(docs) == 14
    for o, d in zip(objs, docs):
        assert "---\n" + ps.yaml_dump(o) == d


def test_records_rebuilt_from_statuses_match_reference_bytes(ps):
    """From (status, node, one, rest) + the tree-only model + the annotations file, rebuild every
    record of the reference's result file: full Clade subtree for IdentityFound, bare id for
    MaxResolutionReached, annotations joined along the path to the root and sorted by clade."""
    from classeq2_b200 import _lib
    text, docs = _golden_docs()
    from classeq2_b200 import Clade, Tree
    # the checked-in model export is tree-only: its top level is a Clade
    root = Clade.from_obj(yaml.load(open(os.path.join(GOLDEN, "bsub-gyrb-k35.tree-only.cls.yaml")),
                                    Loader=yaml.CSafeLoader))
    tree = Tree("ce47d8bc-2885-3d2c-8247-5b8c8b28fefe", "bsub", 70.0, root)
    assert sum(1 for _ in tree.root.walk()) == 364
    tree.annotations = ps.load_annotations(os.path.join(GOLDEN, "bsub-gyrb-annotations.yaml"))
    assert len(tree.annotations) == 15
    n_ident = 0
    for o, d in zip(_load_tagged(text, ps), docs):
        if o["code"] == "IdentityFound":
            row = dict(status=_lib.STATUS_IDENTITY_FOUND, node_id=o["placement"]["clade"]["id"],
                       one=o["placement"]["one"], rest=o["placement"]["rest"], n_root_matched=0)
            n_ident += 1
        else:
            assert o["code"] == "MaxResolutionReached: LCA Accepted"
            row = dict(status=_lib.STATUS_MAX_RESOLUTION, node_id=o["placement"], one=0, rest=0, n_root_matched=0)
        obj, err = ps.placement_response(o["query"], row, tree)
        assert err is None
        assert "---\n" + ps.yaml_dump(obj) == d
        js = json.loads(ps.json_dump(obj))
        assert list(js.keys()) == ["query", "code", "annotations", "placement"]
        assert " " not in ps.json_dump({"a": [1, 2.0, None, True]}) and ps.json_dump({"a": [1, 2.0, None, True]}) == '{"a":[1,2.0,null,true]}'
    assert n_ident == 8


def test_status_strings(ps):
    from classeq2_b200 import _lib
    code, err = ps.status_code(_lib.STATUS_UNCL_NO_MATCH, 'a"b', 0)
    assert code == 'Unclassifiable: Query sequence SequenceHeader("a\\"b") may not be related to the phylogeny' and err is None
    assert ps.status_code(_lib.STATUS_UNCL_COVERAGE, "q", 17) == ("Unclassifiable: Insufficient kmers coverage: 17", None)
    assert ps.status_code(_lib.STATUS_ERR_TOO_SHORT, "q", 0) == (None, "The sequence does not contain enough kmers.")
    assert ps.status_code(_lib.STATUS_ERR_MAX_ITERATIONS, "q", 0)[1] == "The maximum number of iterations has been reached."
    # strings that need YAML quoting are single-quoted as serde_yaml does (see the golden file)
    assert ps.yaml_dump({"code": "MaxResolutionReached: LCA Accepted"}) == "code: 'MaxResolutionReached: LCA Accepted'\n"


def test_fasta_reader_rules(ps, oracle, col_queries):
    text = ">a>b\nacgtn-x\n\nACGU\r\n>empty_mid\n>c\nTT\n>trailing_empty\n"
    assert ps.read_fasta_text(text) == oracle.read_fasta_text(text) == [("ab", "ACGTACG"), ("empty_mid", ""), ("c", "TT")]
    assert ps.read_fasta_text("ACGT\n>x\nAC\n") == []
    assert ps.read_fasta_text(">\nACGT\n>x\nAC\n") == []          # header ">" is empty after removing '>'
    assert ps.read_fasta(os.path.join(GOLDEN, "colletotrichum_queries.fasta")) == col_queries


def test_load_database_zstd_and_plain(ps, tmp_path, col_tree):
    obj = col_tree.to_obj()
    obj["kmersMap"]["map"] = {k: v for k, v in list(obj["kmersMap"]["map"].items())[:3]}
    text = yaml.safe_dump(obj).encode()
    plain = tmp_path / "m.cls.yaml"
    plain.write_bytes(text)
    t1 = ps.load_database(plain)
    assert t1.kmers_map.k_size == 35 and len(t1.kmers_map.map) == 3 and t1.root.to_obj() == col_tree.root.to_obj()
    name = ctypes.util.find_library("zstd") or "libzstd.so.1"
    try:
        z = C.CDLL(name)
    except OSError:
        pytest.skip("libzstd not available")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    cap = z.ZSTD_compressBound(len(text))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, text, len(text), 0)      # level 0 as ports/cli/src/cmds/build_db.rs:73-75
    comp = tmp_path / "m.cls"
    comp.write_bytes(buf.raw[:n])
    t2 = ps.load_database(comp)
    assert t2.to_obj() == t1.to_obj()


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["yaml", "jsonl"])
def test_place_sequences_end_to_end(ps, tmp_path, col_tree, col_expected, fmt):
    """FASTA file in, result + error files out, every record equal to the oracle's response."""
    import classeq2_b200 as cq
    tree = cq.Tree.from_obj(col_tree.to_obj())
    out = tmp_path / "sub" / "result.anything"
    times = cq.place_sequences(os.path.join(GOLDEN, "colletotrichum_queries.fasta"), tree, out, output_format=fmt)
    out_path, err_path = tmp_path / "sub" / f"result.{fmt}", tmp_path / "sub" / "result.error"
    assert out_path.exists() and err_path.exists() and len(times) == len(col_expected["queries"])
    text = out_path.read_text()
    if fmt == "yaml":
        assert text.startswith("---\nquery: ")
        recs = [_fix_floats(o) for o in yaml.safe_load_all(text)]
    else:
        recs = [json.loads(ln) for ln in text.splitlines()]
    want = {h: e for h, e in zip(col_expected["queries"], col_expected["outcomes"]["default"])}
    n_err = sum(1 for e in want.values() if "error" in e)
    assert len(recs) == len(want) - n_err
    assert err_path.read_text() == "The sequence does not contain enough kmers." * n_err
    for r in recs:
        e = want[r["query"]]
        assert hashlib.sha1(json.dumps(r, sort_keys=True).encode()).hexdigest() == e["response_sha1"], r["query"]
        if r["query"] in col_expected["responses_default"]:
            assert r == col_expected["responses_default"][r["query"]]
    with pytest.raises(FileExistsError):
        cq.place_sequences(os.path.join(GOLDEN, "colletotrichum_queries.fasta"), tree, out, output_format=fmt)
    cq.place_sequences(os.path.join(GOLDEN, "colletotrichum_queries.fasta"), tree, out, output_format=fmt, overwrite=True,
                       remove_intersection=True)
    assert out_path.read_text() != "" and err_path.read_text().count("kmers.") == 2 * n_err   # error file is appended


@pytest.mark.gpu
def test_place_sequences_device_ingest_writes_the_same_files(ps, tmp_path, col_tree):
    """FASTA parsed, filtered and packed on the GPU (cls_fasta_upload): result and error files byte-identical
    to the host-reader run."""
    import classeq2_b200 as cq
    tree = cq.Tree.from_obj(col_tree.to_obj())
    fa = os.path.join(GOLDEN, "colletotrichum_queries.fasta")
    a, b = tmp_path / "host" / "r.yaml", tmp_path / "dev" / "r.yaml"
    ta = cq.place_sequences(fa, tree, a)
    tb = cq.place_sequences(fa, tree, b, ingest="device")
    assert [t.sequence for t in ta] == [t.sequence for t in tb] and len(ta) > 300
    assert a.read_bytes() == b.read_bytes() and len(a.read_bytes()) > 10000
    assert (tmp_path / "host" / "r.error").read_bytes() == (tmp_path / "dev" / "r.error").read_bytes()


def test_save_database_round_trip_and_shape(ps, tmp_path, col_tree):
    """build-db's output side (build_db.rs:67-75): serde_yaml text in a zstd frame, extension forced to .cls;
    load_database reads it back; the text has the documented shape (docs/book/02-build-db.md:139-197)."""
    import classeq2_b200 as cq
    obj = col_tree.to_obj()
    obj["kmersMap"]["map"] = {k: v for k, v in list(obj["kmersMap"]["map"].items())[:4]}
    tree = cq.Tree.from_obj(obj)
    try:
        out = cq.save_database(tree, tmp_path / "db.anything")
    except OSError:
        pytest.skip("libzstd not available")
    assert out.endswith("db.cls") and open(out, "rb").read(4) == b"\x28\xb5\x2f\xfd"      # zstd magic
    back = cq.load_database(out)
    assert back.to_obj() == tree.to_obj()
    plain = cq.save_database(tree, tmp_path / "db.yaml", compress=False)
    text = open(plain).read()
    assert cq.load_database(plain).to_obj() == tree.to_obj()
    lines = text.splitlines()
    assert lines[0].startswith("id: ") and any(ln == "kmersMap:" for ln in lines) and "  kSize: 35" in lines and "  mSize: 4" in lines
    assert "root:" in lines and "  kind: ROOT" in lines and "  children:" in lines and "  - id: 1" in lines
    key = next(iter(obj["kmersMap"]["map"]))
    assert f"    {key}:" in lines                  # u64 bucket keys are plain integers, hashes nest below them


def test_flat_model_cache_round_trip(tmp_path, col_flat):
    """The flat binary cache of a model: same arrays back, hence the same cls_model_view."""
    import classeq2_b200 as cq
    p = tmp_path / "model.npz"
    col_flat.save(p)
    back = cq.FlatModel.load(p)
    for n in cq.FlatModel._ARRAYS:
        assert (getattr(back, n) == getattr(col_flat, n)).all() and getattr(back, n).dtype == getattr(col_flat, n).dtype
    assert (back.k_size, back.m_size, back.view.n_entries, back.view.n_sets, back.view.flags) == \
           (col_flat.k_size, col_flat.m_size, col_flat.view.n_entries, col_flat.view.n_sets, col_flat.view.flags)
    g = col_flat.with_general_sets()
    g.save(p)
    assert cq.FlatModel.load(p).view.flags == g.view.flags != col_flat.view.flags


def test_build_database_mirror_equals_oracle(tmp_path, oracle, col_queries, col_tree):
    """map_kmers_to_tree (build_database/mod.rs:26-181, tree.rs:164-364) on the reference's own newick + tip
    sequences: same sanitized tree (ids, kinds, supports, lengths), same tree id, same k-mer map as the oracle's
    restatement; the result goes through save_database / load_database unchanged."""
    import shutil
    import classeq2_b200 as cq
    from classeq2_b200 import build
    nwk = "Colletotrichum_acutatum_gapdh-PhyML.nwk"
    shutil.copy(os.path.join(GOLDEN, nwk), tmp_path / nwk)
    tips = col_queries[:171]
    msa = tmp_path / "tips.fasta"
    msa.write_text("".join(f">{h}\n{s}\n" for h, s in tips))
    tree = build.map_kmers_to_tree(tmp_path / nwk, msa, pairing="own")
    want = oracle.tree_from_newick(open(os.path.join(GOLDEN, nwk)).read(), nwk, 70.0)
    oracle.map_kmers_to_tree(want, tips, 35, 4)
    assert (tree.id, tree.name, tree.min_branch_support) == (want.id, want.name, want.min_branch_support)
    assert tree.root.to_obj() == want.root.to_obj() == col_tree.root.to_obj()
    assert tree.kmers_map.k_size == 35 and tree.kmers_map.m_size == 4
    assert {k: {h: set(n) for h, n in v.items()} for k, v in tree.kmers_map.map.items()} == \
           {k: {h: set(n) for h, n in v.items()} for k, v in want.kmers_map.map.items()}
    for thr in (0.0, 95.0, 101.0):      # other collapse thresholds; a caterpillar deeper than the recursion limit
        a = build.tree_from_newick(open(os.path.join(GOLDEN, nwk)).read(), nwk, thr)
        b = oracle.tree_from_newick(open(os.path.join(GOLDEN, nwk)).read(), nwk, thr)
        assert a.root.to_obj() == b.root.to_obj()
    deep = "(" * 3000 + "t0:1" + "".join(f",t{i + 1}:1)90:0.5" for i in range(3000)) + ";"
    t = build.tree_from_newick(deep, "deep.nwk", 70.0)
    assert sum(1 for c in t.root.walk() if c.is_leaf()) == 3001
    try:
        out = cq.save_database(tree, tmp_path / "db")
    except OSError:
        pytest.skip("libzstd not available")
    assert cq.load_database(out).to_obj() == tree.to_obj()


def test_a_query_that_cannot_be_read_places_nothing_and_is_not_an_error(tmp_path, col_tree, col_flat):
    """mod.rs:111-119: both files are created, then the reader's Result is dropped (`let _ = ...`): Ok, nothing placed.
    The Python driver (both readers) and cls_sequences_open behave alike; a result file that cannot be removed under
    `overwrite` is the reference's "Could not remove file" error (:103-110)."""
    import ctypes as C
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    from test_record_writer import _OracleIndex
    tree = cq.Tree.from_obj(col_tree.to_obj())
    index = _OracleIndex(col_flat)
    for reader in ("native", "python"):
        out = tmp_path / reader / "r.out"
        times = cq.place_sequences(tmp_path / "no_such.fasta", tree, out, index=index, reader=reader)
        assert times == []
        assert (tmp_path / reader / "r.yaml").read_bytes() == b"" and (tmp_path / reader / "r.error").read_bytes() == b""
    index.close()
    h, b = C.c_void_p(), _lib.Batch()
    out = tmp_path / "c" / "r.out"
    assert _lib.lib.cls_sequences_open(str(tmp_path / "no_such.fasta").encode(), str(out).encode(), 0, 0, C.byref(h), C.byref(b)) == 0
    assert int(b.n_queries) == 0
    _lib.lib.cls_sequences_close(h)
    assert (tmp_path / "c" / "r.yaml").read_bytes() == b"" and (tmp_path / "c" / "r.error").read_bytes() == b""
    # overwrite = true, but the "file" in the way is a directory with something in it: remove() fails
    (tmp_path / "d" / "r.yaml" / "x").mkdir(parents=True)
    rc = _lib.lib.cls_sequences_open(str(tmp_path / "no_such.fasta").encode(), str(tmp_path / "d" / "r.out").encode(), 0, 1, C.byref(h), C.byref(b))
    assert rc == _lib.CLS_ERR_INVALID_ARGUMENT and _lib.last_error().startswith("Could not remove file given ")
