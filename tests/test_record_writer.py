"""The library's native record writer ``cls_records_render`` (classeq2_b200/csrc/record_writer.cpp) against the Python
mirror (placement_response + yaml_dump / json_dump), byte for byte, and against a file the reference itself wrote.
No GPU: results are made up from statuses."""
import json
import os
import re

import numpy as np
import pytest
import yaml

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ps():
    from classeq2_b200 import placement
    return placement


def _python_render(ps, headers, res, tree, fmt):
    lookup = ps._TreeLookup(tree)
    out, err = [], []
    for i, h in enumerate(headers):
        obj, e = ps.placement_response(h, res.row(i), tree, lookup)
        if e is not None:
            err.append(e)
        elif fmt == "yaml":
            out.append("---\n" + ps.yaml_dump(obj))
        else:
            out.append(ps.json_dump(obj) + "\n")
    return "".join(out).encode("utf-8"), "".join(err).encode("utf-8")


def _check(ps, headers, res, tree):
    rt = ps.RecordTree(tree)
    for fmt in ("yaml", "jsonl"):
        want = _python_render(ps, headers, res, tree, fmt)
        got = ps.render_records(headers, res, rt, fmt)
        assert got[0] == want[0], (fmt, _first_diff(got[0], want[0]))
        assert got[1] == want[1]
    return got


def _first_diff(a, b):
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return i, a[max(0, i - 60):i + 30], b[max(0, i - 60):i + 30]
    return len(a), len(b)


def test_reference_written_file_is_reproduced(ps):
    """tests/golden/gyrb_result_v090.yaml (written by the reference, v0.9.0) from (status, node, one, rest) + the
    tree-only model + the annotations file, by the native writer."""
    import classeq2_b200 as cq
    from classeq2_b200 import _lib, Clade, Tree
    root = Clade.from_obj(yaml.load(open(os.path.join(GOLDEN, "bsub-gyrb-k35.tree-only.cls.yaml")), Loader=yaml.CSafeLoader))
    tree = Tree("ce47d8bc-2885-3d2c-8247-5b8c8b28fefe", "bsub", 70.0, root)
    tree.annotations = ps.load_annotations(os.path.join(GOLDEN, "bsub-gyrb-annotations.yaml"))
    text = open(os.path.join(GOLDEN, "gyrb_result_v090.yaml")).read()
    docs = list(yaml.safe_load_all(re.sub(r"!\w+ ", "", text)))
    res, headers = cq.BatchResult(len(docs)), []
    for i, o in enumerate(docs):
        headers.append(o["query"])
        if o["code"] == "IdentityFound":
            res.status[i], res.node_id[i] = _lib.STATUS_IDENTITY_FOUND, o["placement"]["clade"]["id"]
            res.one[i], res.rest[i] = o["placement"]["one"], o["placement"]["rest"]
        else:
            res.status[i], res.node_id[i] = _lib.STATUS_MAX_RESOLUTION, o["placement"]
    out, err = _check(ps, headers, res, tree)
    assert ps.render_records(headers, res, ps.RecordTree(tree), "yaml")[0].decode() == text and err == b""
    for line in ps.render_records(headers, res, ps.RecordTree(tree), "jsonl")[0].decode().splitlines():
        assert list(json.loads(line).keys()) == ["query", "code", "annotations", "placement"]


def test_every_status_on_the_colletotrichum_outcomes(ps, col_tree, col_queries, col_expected):
    import classeq2_b200 as cq
    from helpers import expected_row
    tree = cq.Tree.from_obj(col_tree.to_obj())
    headers = [h for h, _ in col_queries]
    by_header = dict(zip(col_expected["queries"], col_expected["outcomes"]["default"]))
    res = cq.BatchResult(len(headers))
    for i, h in enumerate(headers):
        row = expected_row(by_header[h])
        for k, v in row.items():
            getattr(res, k)[i] = v
    for ann in (None, [], [{"clade": 0, "meta": [ps.Tag("Taxid", 5), ps.Tag("SciName", "x: y")]}, {"clade": 18}, {"clade": 1},
                           {"clade": 999999}, {"clade": 0, "meta": [ps.Tag("Note", "two\nlines\n")]}]):
        tree.annotations = ann
        out, err = _check(ps, headers, res, tree)
        assert out.count(b"---\n") + err.count(b".") >= 1
    # every cls_status, including the ones the fixture does not produce
    res2 = cq.BatchResult(11)
    res2.status[:] = np.arange(11)
    res2.node_id[:] = 18
    res2.n_root_matched[:] = 7
    _check(ps, [f"q{i}" for i in range(11)], res2, tree)


def test_adversarial_strings_floats_and_tree_shapes(ps):
    import classeq2_b200 as cq
    from classeq2_b200 import _lib, Clade, Tree
    names = ["plain", "with space", " lead", "trail ", "a: b", "a #b", "key:", "-dash", "- dash", "?q", ":c", "null", "Null", "~", "true",
             "YES", "no", "123", "1_000", "-1.5", ".5", "5.", "1e5", "0x1F", "0o7", "+.inf", ".nan", "inf", "nan", "it's", 'say "hi"',
             "back\\slash", "it's \"both\"", "tab\there", "ctl\x01", "del\x7f", "two\nlines", "ends\n", "\n", "é", "日本語", " nbsp",
             "nbsp ", "[list]", "{map}", "#c", "&a", "*a", "!t", "|p", ">f", "'q", "\"q", "%p", "@a", "`b", ",c", "a,b", "e5", "1e", "_", "1__0",
             "", "0", "-", "--", "-?", "?", ":", "..", "+1", "-x", ".x", "a\rb", "x y", "　wide"]
    floats = [0.0, -0.0, 1.0, 72.0, 1e-6, 1e-5, 9.999e-6, 0.000619153, 123456.789, 1e15, 1e16, 1.5e16, 1.7976931348623157e308, 5e-324, -3.25,
              0.1, 1 / 3, 100.0, 1e21, 12345678901234567.0, float("inf"), 2.5e-7]
    kids = []
    for i, nm in enumerate(names):
        sup = floats[i % len(floats)] if i % 3 else None
        kids.append(Clade(id=10 + i, parent=(1 if i % 2 else None), kind="LEAF" if i % 4 else "NODE", name=nm, support=sup,
                          length=floats[(i + 5) % len(floats)] if i % 5 else None, children=([] if i % 7 == 0 else None)))
    inner = Clade(id=1, parent=0, kind="NODE", support=99.5, length=0.25, children=kids)
    dup = Clade(id=1, parent=0, kind="NODE", name="second node with id 1", children=None)     # get_node_by_id: the first one wins
    root = Clade(id=0, parent=None, kind="ROOT", length=0.0, children=[inner, dup])
    tree = Tree("t", "t", 70.0, root)
    tree.annotations = [{"clade": 1, "meta": [ps.Tag("SciName", nm) for nm in names[:12]]}, {"clade": 0}, {"clade": 10}]
    n = len(names) + 6
    res = cq.BatchResult(n)
    headers = []
    for i in range(n):
        headers.append(names[i % len(names)])
        res.status[i] = _lib.STATUS_IDENTITY_FOUND if i % 3 == 0 else (_lib.STATUS_UNCL_NO_MATCH if i % 3 == 1 else _lib.STATUS_MAX_RESOLUTION)
        res.node_id[i] = [1, 10 + (i % len(names)), 0][i % 3] if i % 3 != 2 else [1, 4242, 10][i % 3]
        res.one[i], res.rest[i] = i * 7, i
    _check(ps, headers, res, tree)
    # an IdentityFound whose node is not in the tree is an error, not a guess
    bad = cq.BatchResult(1)
    bad.status[0], bad.node_id[0] = _lib.STATUS_IDENTITY_FOUND, 424242
    with pytest.raises(_lib.ClsError):
        ps.render_records(["q"], bad, ps.RecordTree(tree), "yaml")


def test_many_records_keep_their_order(ps, col_tree):
    """More records than one writer block (2 048): blocks are rendered on the host pool and joined in input order."""
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    tree = cq.Tree.from_obj(col_tree.to_obj())
    ids = [c.id for c in tree.root.walk() if not c.is_leaf()]
    n = 20_000
    rng = np.random.default_rng(1)
    res = cq.BatchResult(n)
    res.status[:] = rng.choice([_lib.STATUS_IDENTITY_FOUND, _lib.STATUS_MAX_RESOLUTION, _lib.STATUS_UNCL_COVERAGE, _lib.STATUS_ERR_TOO_SHORT,
                                _lib.STATUS_UNCL_NO_MATCH], n)
    res.node_id[:] = rng.choice(ids, n)
    res.one[:] = rng.integers(0, 500, n)
    res.n_root_matched[:] = rng.integers(0, 500, n)
    _check(ps, [f"read_{i}" for i in range(n)], res, tree)


def test_rendered_text_beyond_one_copy_piece_is_not_truncated(ps, col_tree, monkeypatch):
    """``ctypes.string_at`` takes a C int: a rendered batch beyond 2 GiB used to come back silently truncated (size
    modulo 2^32).  The text is copied in pieces now; with a small piece size the pieced copy must equal the whole."""
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    tree = cq.Tree.from_obj(col_tree.to_obj())
    ids = [c.id for c in tree.root.walk() if not c.is_leaf()]
    n = 300
    res = cq.BatchResult(n)
    res.status[:] = _lib.STATUS_IDENTITY_FOUND
    res.node_id[:] = np.random.default_rng(3).choice(ids, n)
    headers = [f"q{i}" for i in range(n)]
    rt = ps.RecordTree(tree)
    whole = ps.render_records(headers, res, rt, "yaml")
    assert len(whole[0]) > 100_000
    for piece in (1, 4097, 65536):
        monkeypatch.setattr(ps, "_C_TEXT_PIECE", piece if piece > 1 else 1000)
        assert ps.render_records(headers, res, rt, "yaml") == whole
    import ctypes
    buf = ctypes.create_string_buffer(b"abcdefghij", 10)
    monkeypatch.setattr(ps, "_C_TEXT_PIECE", 3)
    assert ps._c_text(ctypes.c_void_p(ctypes.addressof(buf)), 10) == b"abcdefghij"
    assert ps._c_text(ctypes.c_void_p(ctypes.addressof(buf)), 0) == b""


class _OracleIndex:
    """Stands in for ``Index`` so that ``place_sequences`` runs without a GPU: the batch is placed by the C++ oracle
    (test infrastructure).  Everything else - FASTA reader, batching, record writers, files - is the product's."""

    def __init__(self, flat):
        from oracle import cpp_oracle
        self.md = cpp_oracle.CppModel.from_flat(flat)

    def place_batch(self, seqs, params=None):
        import classeq2_b200 as cq
        bases, offsets = cq.make_batch(seqs)
        p = params or cq.PlaceParams()
        out = self.md.place_batch(bases, offsets, p.max_iterations, p.min_match_coverage, p.remove_intersection, n_threads=4)
        res = cq.BatchResult(len(offsets) - 1)
        for name, _ in cq.engine.RESULT_DTYPES:
            getattr(res, name)[:] = out[name]
        return res

    def close(self):
        self.md.close()


@pytest.mark.parametrize("fmt", ["yaml", "jsonl"])
def test_place_sequences_files_with_both_writers(ps, tmp_path, col_tree, col_flat, col_expected, fmt):
    """The whole host side of place_sequences (reader -> batches -> records -> result / error files) around an oracle
    placement: the native and the Python writer give the same bytes, and every record is the oracle's response."""
    import hashlib
    import classeq2_b200 as cq
    tree = cq.Tree.from_obj(col_tree.to_obj())
    index = _OracleIndex(col_flat)
    fa = os.path.join(GOLDEN, "colletotrichum_queries.fasta")
    a, b = tmp_path / "n" / "r", tmp_path / "p" / "r"
    cq.place_sequences(fa, tree, a, output_format=fmt, index=index, writer="native", reader="native", batch_size=100)
    cq.place_sequences(fa, tree, b, output_format=fmt, index=index, writer="python", reader="python", batch_size=1 << 20)
    index.close()
    text = (tmp_path / "n" / f"r.{fmt}").read_bytes()
    assert text == (tmp_path / "p" / f"r.{fmt}").read_bytes() and len(text) > 10000
    assert (tmp_path / "n" / "r.error").read_bytes() == (tmp_path / "p" / "r.error").read_bytes()
    recs = [json.loads(ln) for ln in text.decode().splitlines()] if fmt == "jsonl" else \
        [_floats(o) for o in yaml.safe_load_all(text.decode())]
    want = dict(zip(col_expected["queries"], col_expected["outcomes"]["default"]))
    assert len(recs) == sum(1 for e in want.values() if "error" not in e)
    for r in recs:
        assert hashlib.sha1(json.dumps(r, sort_keys=True).encode()).hexdigest() == want[r["query"]]["response_sha1"], r["query"]


def _floats(o):
    if isinstance(o, dict):
        return {k: (float(v) if k in ("length", "support") and v is not None else _floats(v)) for k, v in o.items()}
    if isinstance(o, list):
        return [_floats(v) for v in o]
    return o


@pytest.mark.parametrize("fmt", ["yaml", "jsonl"])
def test_native_sequences_open_and_write(ps, tmp_path, col_tree, col_flat, fmt):
    """cls_sequences_open / cls_sequences_write - the two host halves of the one-call cls_place_sequences - around an
    oracle placement: same files as the Python driver, same path handling (extension replaced, directory created,
    overwrite refused with the reference's message, error file appended)."""
    import ctypes as C
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    tree = cq.Tree.from_obj(col_tree.to_obj())
    tree.annotations = [{"clade": 0, "meta": [ps.Tag("Rank", "genus")]}, {"clade": 18}]
    index = _OracleIndex(col_flat)
    fa = os.path.join(GOLDEN, "colletotrichum_queries.fasta")
    cq.place_sequences(fa, tree, tmp_path / "p" / "r.whatever", output_format=fmt, index=index, writer="python", reader="python")
    rt = ps.RecordTree(tree)

    def run(out_file, overwrite, chunk):
        h, b = C.c_void_p(), _lib.Batch()
        rc = _lib.lib.cls_sequences_open(fa.encode(), str(out_file).encode(), 0 if fmt == "yaml" else 1, overwrite, C.byref(h), C.byref(b))
        if rc != 0:
            return rc, _lib.last_error()
        n = int(b.n_queries)
        offsets = np.ctypeslib.as_array(b.offsets, shape=(n + 1,)).copy()
        bases = np.ctypeslib.as_array(b.bases, shape=(int(offsets[-1]),)).copy()
        for a in range(0, n, chunk):
            off = offsets[a:a + chunk + 1]
            res = index.place_batch((bases[int(off[0]):int(off[-1])], off - off[0]))
            cr = res.to_c()
            _lib.check(_lib.lib.cls_sequences_write(h, C.byref(rt.view), len(off) - 1, C.byref(cr)))
        _lib.lib.cls_sequences_close(h)
        return 0, ""

    out = tmp_path / "n" / "r.something"
    assert run(out, 0, 97) == (0, "")
    want_o, want_e = (tmp_path / "p" / f"r.{fmt}").read_bytes(), (tmp_path / "p" / "r.error").read_bytes()
    assert (tmp_path / "n" / f"r.{fmt}").read_bytes() == want_o and len(want_o) > 10000
    assert (tmp_path / "n" / "r.error").read_bytes() == want_e
    rc, msg = run(out, 0, 97)                                   # the result file exists and overwrite is false (mod.rs:96-101)
    assert rc == _lib.CLS_ERR_INVALID_ARGUMENT
    assert msg == f'Could not overwrite existing file "{tmp_path}/n/r.{fmt}" when overwrite option is `false`.'
    assert run(out, 1, 1 << 20) == (0, "")                      # overwrite: the result file starts over, the error file is appended
    assert (tmp_path / "n" / f"r.{fmt}").read_bytes() == want_o
    assert (tmp_path / "n" / "r.error").read_bytes() == want_e + want_e
    index.close()


_ALPHABET = list("abzAZ019 _-.:,#'\"\\/|>!&*?[]{}%@`~+=()<;$^") + ["\t", "\x01", "\x7f", "é", "日", "\x85", "\xa0", "﻿", "\U0001F600", "\r", " "]


def test_quoting_agrees_with_libyaml_on_random_strings(ps):
    """Where serde_yaml leaves the scalar style to libyaml (text that does not read as a number / boolean / null), the
    emitters must choose what libyaml's emitter chooses: plain, else single-quoted, else - for text outside its printable
    set - double-quoted with its escapes.  PyYAML's C emitter IS libyaml's; 150 000 random strings."""
    import random
    dumper = getattr(yaml, "CSafeDumper", None)
    if dumper is None:
        pytest.skip("PyYAML without libyaml")
    resolver = yaml.resolver.Resolver()
    rng = random.Random(11)
    n = 0
    for _ in range(150_000):
        s = "".join(rng.choice(_ALPHABET[:-1]) for _ in range(rng.randint(1, 7)))
        if resolver.resolve(yaml.ScalarNode, s, (True, False)) != "tag:yaml.org,2002:str":
            continue                                    # PyYAML's YAML 1.1 typing would quote it for its own reasons
        if ps._serde_yaml_reads_it_as_another_type(s):
            continue                                    # serde_yaml's own rule (single quotes), not libyaml's choice
        out = yaml.dump({"k": s}, Dumper=dumper, width=10**9, allow_unicode=True, default_flow_style=False)
        want = out[3:].rstrip("\n") if out.startswith("k: ") else out[2:]
        assert ps._yaml_scalar(s, 2) == want, repr(s)
        n += 1
    assert n > 100_000


def test_native_writer_equals_python_on_random_headers(ps, col_tree):
    """The same random strings (plus number-like and multi-line ones) as query headers and clade names: C++ == Python."""
    import random
    import classeq2_b200 as cq
    from classeq2_b200 import _lib, Clade, Tree
    rng = random.Random(12)
    extra = ["1", "-1", "1e5", ".5", "0x10", "true", "NULL", "~", "yes", "007", "1_000", "+.inf", "a\nb", "x\n", "---", "...", "--- a", ".. .",
             "null", "nULL", "TRUE", "tRue", "no", "y", "+1", "++1", "-007", "0", "00", "0x1F", "0xg", "0x", "0o7", "0o8", "0b101", "0b2", "-0x1f",
             "+0x1f", "0x" + "f" * 32, "0x1" + "0" * 32, "0o3" + "7" * 42, "0o4" + "0" * 42, "0b" + "1" * 128, "0b" + "1" * 129, "0x000", "1.5",
             "1.", "+.5", "1E-5", "1e", "e5", ".", "-", "-.INF", ".NaN", ".nan", "-.nan", "inf", "nan", "infinity", "1e999", "1e-999", "12a",
             "-1x", "1 2", "\u0661\u0662", "9" * 60, "-" + "9" * 60, "+-1", "-+1", "1e+5", "1e+", ".e5", "5.e5", "-.5e-3"]
    strings = ["".join(rng.choice(_ALPHABET) for _ in range(rng.randint(1, 9))) for _ in range(6000)] + extra
    kids = [Clade(id=10 + i, parent=1, kind="LEAF", name=s, length=0.5) for i, s in enumerate(strings[:400])]
    root = Clade(id=0, parent=None, kind="ROOT", children=[Clade(id=1, parent=0, kind="NODE", children=kids)])
    tree = Tree("t", "t", 70.0, root)
    n = len(strings)
    res = cq.BatchResult(n)
    res.status[:] = [_lib.STATUS_UNCL_NO_MATCH if i % 2 else _lib.STATUS_UNCL_NO_ROOT for i in range(n)]
    res.status[0], res.node_id[0] = _lib.STATUS_IDENTITY_FOUND, 1          # one record that prints all the names
    _check(ps, strings, res, tree)


def test_read_name_fast_path_of_the_native_writer(ps, col_tree):
    """The native writer skips the quoting analysis for headers that start with a letter and hold only [A-Za-z0-9_]
    (and are not 4 or 5 letters long: null / true / false in their spellings): the same text as the general path, on
    every status, for word-like headers around that boundary."""
    import random
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    rng = random.Random(5)
    tree = cq.Tree.from_obj(col_tree.to_obj())
    tree.annotations = [{"clade": 0, "meta": [ps.Tag("Rank", "genus")]}]
    ids = [c.id for c in tree.root.walk() if not c.is_leaf()]
    words = ["null", "Null", "NULL", "nULL", "true", "True", "TRUE", "tRUE", "false", "False", "FALSE", "fALSE", "nulls", "truer", "a", "A", "z9",
             "read_1", "Read", "reads", "yes", "no", "on", "off", "y", "n", "inf", "nan", "NaN", "e5", "E5", "x0", "_a", "a_", "0a", "a0", "abcd",
             "abcde", "abcdef", "ab_d", "ab1d", "A_1", "_", "__", "a-b", "a.b", "a b", "a:b", "é", "aé", "~", "q" * 300]
    alpha = "abcxyzABCXYZ0189_"
    headers = words + ["".join(rng.choice(alpha) for _ in range(rng.randint(1, 8))) for _ in range(4000)]
    n = len(headers)
    statuses = [_lib.STATUS_IDENTITY_FOUND, _lib.STATUS_MAX_RESOLUTION, _lib.STATUS_UNCL_COVERAGE, _lib.STATUS_UNCL_NO_MATCH,
                _lib.STATUS_UNCL_NO_ROOT, _lib.STATUS_UNCL_NO_INTROSPECTION, _lib.STATUS_INCONCLUSIVE, _lib.STATUS_ERR_TOO_SHORT]
    for shift in range(2):
        res = cq.BatchResult(n)
        res.status[:] = [statuses[(i + shift * 3) % len(statuses)] for i in range(n)]
        res.node_id[:] = [ids[i % len(ids)] for i in range(n)]
        res.one[:] = np.arange(n) % 400 - 3
        res.rest[:] = np.arange(n) % 7
        res.n_root_matched[:] = np.arange(n) % 233
        _check(ps, headers, res, tree)


@pytest.mark.parametrize("fmt,batch", [("yaml", "100"), ("jsonl", "7"), ("yaml", "0")])
def test_one_call_place_sequences_around_an_oracle_placer(ps, tmp_path, col_tree, col_flat, fmt, batch):
    """cls_place_sequences itself - path handling, reader, the loop over batches with its result arrays, writer - with
    cls_place_batch supplied by the C++ oracle (tests/native/place_seq_host.cpp): same files as the Python driver around
    the same oracle.  Runs in a child process: CLS_SEQ_BATCH is read once."""
    import subprocess
    import sys
    import textwrap
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run(["make", "-C", os.path.join(here, "native"), "libplace_seq_host.so"], check=True, capture_output=True)
    import classeq2_b200 as cq
    tree = cq.Tree.from_obj(col_tree.to_obj())
    index = _OracleIndex(col_flat)
    fa = os.path.join(GOLDEN, "colletotrichum_queries.fasta")
    cq.place_sequences(fa, tree, tmp_path / "p" / "r", output_format=fmt, index=index, writer="python", reader="python", remove_intersection=True)
    index.close()
    child = textwrap.dedent(f"""
        import ctypes as C, json, os, sys
        sys.path.insert(0, {os.path.dirname(here)!r}); sys.path.insert(0, {here!r})
        import numpy as np
        import classeq2_b200 as cq
        from classeq2_b200 import _lib, placement as ps
        from oracle import cpp_oracle
        z = np.load(os.path.join({GOLDEN!r}, "colletotrichum_model.npz"))
        flat = cq.FlatModel(int(z["k_size"]), int(z["m_size"]), z["node_id"], z["node_kind"], z["child_off"], z["child_idx"],
                            z["entry_bucket"], z["entry_hash"], z["entry_set"], z["set_off"], z["set_node_ids"])
        tree = cq.Tree.from_obj({{"id": "t", "name": "t", "minBranchSupport": 70.0, "root": json.loads(bytes(z["tree_json"]).decode())["root"]}})
        md = cpp_oracle.CppModel.from_flat(flat)
        lib = C.CDLL(os.path.join({here!r}, "native", "libplace_seq_host.so"))
        lib.psh_set_placer(C.cast(cpp_oracle.lib.orc_place_batch, C.c_void_p))
        lib.psh_calls.restype = C.c_uint64
        lib.cls_last_error.restype = C.c_char_p
        lib.cls_place_sequences.argtypes = [C.c_void_p, C.POINTER(_lib.RecordTree), C.c_char_p, C.c_char_p, C.POINTER(_lib.Params), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        rt = ps.RecordTree(tree)
        p = cq.PlaceParams(remove_intersection=True).to_c()
        n = C.c_uint64()
        out = {str(tmp_path / "c" / "r.x")!r}.encode()
        fmt = {0 if fmt == "yaml" else 1}
        rc = lib.cls_place_sequences(md._h, C.byref(rt.view), {fa!r}.encode(), out, C.byref(p), fmt, 0, C.byref(n))
        assert rc == 0, lib.cls_last_error()
        rc2 = lib.cls_place_sequences(md._h, C.byref(rt.view), {fa!r}.encode(), out, C.byref(p), fmt, 0, C.byref(n))
        assert rc2 == _lib.CLS_ERR_INVALID_ARGUMENT and lib.cls_last_error().startswith(b"Could not overwrite existing file")
        rc3 = lib.cls_place_sequences(md._h, C.byref(rt.view), {fa!r}.encode(), out, C.byref(p), fmt, 1, C.byref(n))
        assert rc3 == 0
        print(json.dumps({{"n": n.value, "calls": lib.psh_calls()}}))
    """)
    r = subprocess.run([sys.executable, "-c", child], env=dict(os.environ, CLS_SEQ_BATCH=batch), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    want_o, want_e = (tmp_path / "p" / f"r.{fmt}").read_bytes(), (tmp_path / "p" / "r.error").read_bytes()
    assert (tmp_path / "c" / f"r.{fmt}").read_bytes() == want_o and len(want_o) > 10000
    assert (tmp_path / "c" / "r.error").read_bytes() == want_e + want_e          # the second successful run appends
    n_records = len(want_o.split(b"---\n")) - 1 if fmt == "yaml" else want_o.count(b"\n")
    assert info["n"] >= n_records
    b = int(batch)
    assert info["calls"] == 2 * (1 if b == 0 else -(-info["n"] // b))            # the loop really ran in batches
