"""The oracle against every known answer the reference's own docs / fixtures pin,
and against the committed golden outcomes.  CPU only."""
import hashlib
import json
import os
import uuid

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN


def test_murmur3_reference_pins(oracle, pins):
    # docs/book/02-build-db.md:181,192 print these two bucket keys; h1("") = 0 is the m = 0 key
    for s, h in pins["murmur3_h1"].items():
        assert oracle.hash_kmer(s) == h
    assert oracle.murmurhash3_x64_128(b"", 0) == (0, 0)


def test_murmur3_published_vectors(oracle):
    # Published MurmurHash3_x64_128 answers (SMHasher reference implementation, seed 0 / 123)
    assert oracle.murmurhash3_x64_128(b"hello", 0) == (0xCBD8A7B341BD9B02, 0x5B1E906A48AE1D19)
    assert oracle.murmurhash3_x64_128(b"The quick brown fox jumps over the lazy dog", 0) == (
        0xE34BBC7BBC071B6C, 0x7A433CA9C49A9347)
    assert oracle.murmurhash3_x64_128(b"hello, world", 0) == (0x342FAC623A5EBC8E, 0x4CDCBC079642414D)


def test_256_buckets_at_m4(oracle):
    # docs/book/02-build-db.md:123 - m = 4 gives 256 distinct bucket keys
    import itertools
    keys = {oracle.hash_kmer("".join(p)) for p in itertools.product("ACGT", repeat=4)}
    assert len(keys) == 256


def test_windowing_doc_example(oracle):
    # kmers_map.rs:359-372 (forward strand windows of "ATCG")
    f = lambda k: [w for w, _ in oracle.KmersMap.build_kmers_from_sequence("ATCG", k)]  # noqa: E731
    assert f(1) == ["A", "T", "C", "G"]
    assert f(2) == ["AT", "TC", "CG"]
    assert f(3) == ["ATC", "TCG"]
    assert f(4) == ["ATCG"]
    assert oracle.KmersMap(5, 0).build_kmer_from_string("ATCG") == []
    km = oracle.KmersMap(2, 0)
    assert [w for w, _ in km.build_kmer_from_string("ATCG")] == ["AT", "TC", "CG", "CG", "GA", "AT"]


def test_both_strands_all_windows_count(oracle, pins, col_queries):
    # tests/data/public/...fd7/output/result.yaml first record: `one: 3754` for a 1 911 bp query
    # = 2 * (1911 - 34): every window of both strands, counted as distinct hashes.
    g = pins["gyrb_first_query"]
    seq = dict(col_queries)[g["header"]]
    assert len(seq) == g["length"] == 1911
    kmers = oracle.KmersMap(35, 4).build_kmer_from_string(seq)
    assert len(kmers) == g["one"] == 3754
    assert len({h for _, h in kmers}) == 3754


def test_tree_id_and_preorder_ids(oracle, pins):
    # tree.rs:213-214 (uuid3 of the file name) and phylotree's pre-order node numbering, both
    # pinned by the reference's stale golden model core/src/tests/data/.../outputs/...yaml
    assert str(uuid.uuid3(uuid.NAMESPACE_DNS, pins["tree_name"])) == pins["tree_id"]
    newick = open(os.path.join(GOLDEN, pins["tree_name"])).read()
    tree = oracle.tree_from_newick(newick, pins["tree_name"], 0.0 - 1e9)  # no collapse: raw numbering
    assert tree.id == pins["tree_id"]
    leaves = {c.id: c.name for c in tree.root.walk() if c.is_leaf()}
    assert len(leaves) == 171
    for i, name in pins["stale_golden_leaf_ids"]:
        assert leaves[i] == name
    assert len(pins["stale_golden_leaf_ids"]) == 171


def test_fasta_reader_rules(oracle):
    # file_or_stdin.rs:76-116 + sequence.rs:47-56
    text = ">a>b\nacgtn-x\n\nACGU\r\n>empty_mid\n>c\nTT\n>trailing_empty\n"
    assert oracle.read_fasta_text(text) == [("ab", "ACGTACG"), ("empty_mid", ""), ("c", "TT")]
    assert oracle.read_fasta_text("ACGT\n>x\nAC\n") == []          # sequence before any header: aborted
    assert oracle.read_fasta_text(">x\nAC") == [("x", "AC")]       # no trailing newline
    assert oracle.remove_non_iupac_from_sequence("acgtNRYKM-acgt") == "ACGTACGT"


def test_result_record_shape(oracle, pins, col_expected, col_queries, col_tree):
    # wire shape of PlacementResponse pinned by ...fd7/output/result.yaml
    shape = pins["result_record_keys"]
    seqs = dict(col_queries)
    rs = {h: oracle.placement_response(h, oracle.place_sequence(h, seqs[h], col_tree), col_tree)
          for h in col_expected["responses_default"]}
    assert rs == col_expected["responses_default"]
    ident = next(r for r in rs.values() if r["code"] == "IdentityFound")
    assert list(ident.keys()) == [k for k in shape["identity"] if k != "annotations"]
    assert list(ident["placement"].keys()) == shape["identity_placement"] == ["clade", "one", "rest"]
    assert set(shape["identity_clade"]) <= set(ident["placement"]["clade"].keys()) | {"support", "length", "name", "children"}
    maxres = next(r for r in rs.values() if r["code"].startswith("MaxResolutionReached"))
    assert maxres["code"] in shape["codes"] and isinstance(maxres["placement"], int)
    uncl = next(r for r in rs.values() if r["code"].startswith("Unclassifiable"))
    assert "placement" not in uncl


def test_unclassifiable_message_uses_debug_format(oracle, col_tree):
    p = oracle.place_sequence('quote"back\\slash', "ACGT" * 20, col_tree)
    assert p.code() == ('Unclassifiable: Query sequence SequenceHeader("quote\\"back\\\\slash") may not be '
                        "related to the phylogeny")


def test_model_rebuild_matches_committed(oracle, col_queries, col_tree, col_npz, pins):
    # the committed flat model == oracle build (newick -> sanitize(70) -> k-mer map, k=35, m=4)
    newick = open(os.path.join(GOLDEN, pins["tree_name"])).read()
    tree = oracle.tree_from_newick(newick, pins["tree_name"], 70.0)
    oracle.map_kmers_to_tree(tree, col_queries[:171], 35, 4)
    assert tree.root.to_obj() == col_tree.root.to_obj()
    assert tree.kmers_map.map == col_tree.kmers_map.map
    assert tree.kmers_map.n_entries() == 9378 and len(tree.kmers_map.map) <= 256
    # builder invariant: every node set contains the root and is a union of root->tip paths
    assert all(0 in nodes for v in tree.kmers_map.map.values() for nodes in v.values())


@pytest.mark.parametrize("knob", ["default", "remove_intersection", "cov1", "iter2"])
def test_oracle_matches_committed_outcomes(oracle, col_queries, col_tree, col_expected, knob):
    from helpers import outcome_of
    kn = next(k for k in col_expected["knobs"] if k["name"] == knob)
    headers = col_expected["queries"]
    assert [h for h, _ in col_queries] == headers
    step = 1 if knob == "default" else 4  # the pure-Python oracle is slow; full pass for defaults only
    for i in range(0, len(headers), step):
        h, s = col_queries[i]
        got = outcome_of(oracle, h, s, col_tree, kn["max_iterations"], kn["min_match_coverage"], kn["remove_intersection"])
        want = {k: v for k, v in col_expected["outcomes"][knob][i].items() if k != "response_sha1"}
        assert got == want, (h, got, want)


def test_response_hashes(oracle, col_queries, col_tree, col_expected):
    for i in range(0, len(col_queries), 9):
        h, s = col_queries[i]
        e = col_expected["outcomes"]["default"][i]
        if "error" in e:
            continue
        p = oracle.place_sequence(h, s, col_tree)
        sha = hashlib.sha1(json.dumps(oracle.placement_response(h, p, col_tree), sort_keys=True).encode()).hexdigest()
        assert sha == e["response_sha1"]


def test_rust_round(oracle):
    assert [oracle.rust_round(x) for x in (0.5, 1.5, 2.5, 2.4999, 0.0, 6.3)] == [1, 2, 3, 2, 0, 6]


# ---- counter formulation (what the GPU kernel computes) == set formulation (the reference) -------
def _counter_descent(tree, hits, remove_intersection, max_iterations=1000):
    """cnt/excl/U form of place_sequence.rs:279-601 over `hits` = list of node-id sets (M_r)."""
    parent, children, it = tree.root, tree.root.children, 0
    while True:
        it += 1
        if it > max_iterations:
            return ("Err",)
        nl = [c for c in children if not c.is_leaf()]
        ids = [c.id for c in nl]
        cnt = {i: 0 for i in ids}
        excl = {i: 0 for i in ids}
        U = 0
        for s in hits:
            pres = [i for i in ids if i in s]
            for i in pres:
                cnt[i] += 1
            U += bool(pres)
            if len(pres) == 1:
                excl[pres[0]] += 1
        cand = [c for c in nl if cnt[c.id] > 0]
        props = []
        for c in cand:
            if len(cand) == 1:
                one, rest = cnt[c.id], 0
            elif remove_intersection:
                one, rest = excl[c.id], U - cnt[c.id]
            else:
                one, rest = cnt[c.id], U - excl[c.id]
            if one > rest:
                props.append((c, one, rest))
        assert len(props) <= 1  # SURVEY.md section 7 "mathematical note"
        if not props:
            return ("Unclassifiable",) if it == 1 else ("MaxResolutionReached", parent.id, it)
        c, one, rest = props[0]
        nlc = [x for x in (c.children or []) if not x.is_leaf()]
        if not nlc:
            return ("IdentityFound", c.id, one, rest, it)
        parent, children = c, nlc


@st.composite
def _random_model(draw):
    from oracle import classeq_oracle as O
    n_internal = draw(st.integers(1, 9))
    rng = np.random.Generator(np.random.PCG64(draw(st.integers(0, 2**32))))
    root = O.Clade(0, None, "ROOT", children=[])
    nodes, nid = [root], 1
    for _ in range(n_internal):
        p = nodes[int(rng.integers(len(nodes)))]
        c = O.Clade(nid * 3, p.id, "NODE", children=[] if rng.random() < 0.8 else None)  # sparse ids
        nid += 1
        p.children = (p.children or []) + [c]
        nodes.append(c)
    for p in list(nodes):
        for _ in range(int(rng.integers(0, 3))):
            leaf = O.Clade(nid * 3, p.id, "LEAF", name=f"t{nid}")
            nid += 1
            p.children = (p.children or []) + [leaf]
            nodes.append(leaf)
    ids = [n.id for n in nodes]
    n_hits = int(rng.integers(0, 40))
    hits = []
    for _ in range(n_hits):
        s = {0} | {ids[int(j)] for j in rng.integers(0, len(ids), int(rng.integers(0, 6)))}
        if rng.random() < 0.1:
            s.discard(0)
        hits.append(s)
    return O.Tree("x", "x", 70.0, root), hits


@settings(max_examples=150, deadline=None)
@given(_random_model(), st.booleans())
def test_counter_form_equals_set_form(model, ri):
    """Drive the oracle's place_sequence through a hand-made k-mer map whose hashes are unique
    per hit, so that its set algebra can be compared with the counter formulation."""
    from oracle import classeq_oracle as O
    tree, hits = model
    km = O.KmersMap(2, 0)
    # every 2-mer of ACGT-strings is a potential hash; instead of real k-mers use a query with all
    # 16 2-mers and give the index at most 16 entries keyed by those real hashes
    query = "AACAGATCCGCTGGTTA"
    hs = sorted({h for _, h in km.build_kmer_from_string(query)})
    hits = hits[: len(hs)]
    km.map[0] = {hs[i]: set(s) for i, s in enumerate(hits)}
    tree.kmers_map = km
    try:
        p = O.place_sequence("q", query, tree, None, 0.0, ri)
    except O.PlacementError:
        return
    m_r = [s for s in hits if 0 in s]
    if p.status == "Unclassifiable" and not p.message.startswith("Tree introspection"):
        assert not hits or not m_r
        return
    got = _counter_descent(tree, m_r, ri)
    if p.status == "Unclassifiable":
        assert got == ("Unclassifiable",)
    elif p.status == "MaxResolutionReached":
        assert got == ("MaxResolutionReached", p.clade, p.iterations)
    else:
        assert p.status == "IdentityFound"
        assert got == ("IdentityFound", p.clade, p.one, p.rest, p.iterations)


def test_reference_built_model_pins_tree_and_builder(oracle, col_queries):
    """The reference's OWN build output for its Colletotrichum inputs (tests/golden/reference_built_model_k12.json.gz,
    written by an early version: k = 12, string keys, forward strand only) is reproduced exactly by the oracle:
    the tree of Tree::init_from_file (tree.rs:164-364: ids, kinds, names, supports, lengths), the windows, the node
    set of every k-mer (root -> tip id paths, clade.rs:127-156) and the pairing of the MSA loop (header i with
    sequence i-1, build_database/mod.rs:93-116)."""
    import os
    from helpers import load_reference_built_model
    pin = load_reference_built_model()
    nwk = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", pin["name"])).read()
    tree = oracle.tree_from_newick(nwk, pin["name"], -1e9)       # that version collapsed nothing
    assert tree.id == pin["id"] and tree.name == pin["name"]

    def strip(n):                                                # the old schema has no `parent` field
        n = {k: v for k, v in n.items() if k != "parent"}
        if "children" in n:
            n["children"] = [strip(c) for c in n["children"]]
        return n
    assert strip(tree.root.to_obj()) == pin["root"]
    tips = col_queries[:171]                                     # the reference's gap-free FASTA, in file order
    oracle.map_kmers_to_tree(tree, tips, pin["k_size"], 0, pairing="reference", forward_only=True)
    assert list(tree.kmers_map.map) == [0]                       # m = 0: one bucket
    assert tree.kmers_map.map[0] == {oracle.hash_kmer(s): set(ids) for s, ids in pin["kmers"].items()}
    # with the corrected pairing the map differs (that is the defect), with both strands it is the closure below
    own = oracle.tree_from_newick(nwk, pin["name"], -1e9)
    oracle.map_kmers_to_tree(own, tips, pin["k_size"], 0, forward_only=True)
    assert own.kmers_map.map[0] != tree.kmers_map.map[0]
    both = oracle.tree_from_newick(nwk, pin["name"], -1e9)
    oracle.map_kmers_to_tree(both, tips, pin["k_size"], 0, pairing="reference")
    comp = str.maketrans("ACGT", "TGCA")
    want = {}
    for s, ids in pin["kmers"].items():
        rc = s[::-1].translate(comp)
        want[oracle.hash_kmer(s)] = want[oracle.hash_kmer(rc)] = set(ids) | set(pin["kmers"].get(rc, []))
    assert both.kmers_map.map[0] == want


def test_reference_written_results_obey_the_descent_rules():
    """1 111 placement records written by the reference itself (tests/golden/reference_gyrb_results_compact.json.gz:
    two result files of its bsub-gyrB model, whose k-mer map is not available) against the rules the oracle and the
    kernels implement: IdentityFound names a node WITHOUT non-leaf children (update_introspection_node.rs:32-87),
    MaxResolutionReached a node WITH some (place_sequence.rs:456-465), a proposal has one > rest (:383-418), and
    one <= 2 * (L - k + 1) - the count of distinct k-mers of both strands (kmers_map.rs:375-398) - with equality for
    the queries that match a reference sequence completely."""
    import gzip
    import yaml
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    pin = json.loads(gzip.decompress(open(os.path.join(g, "reference_gyrb_results_compact.json.gz"), "rb").read()))
    root = yaml.load(open(os.path.join(g, "bsub-gyrb-k35.tree-only.cls.yaml")), Loader=yaml.CSafeLoader)
    nodes = {}
    stack = [root]
    while stack:
        n = stack.pop()
        nodes[n["id"]] = n
        stack.extend(n.get("children") or [])
    assert len(nodes) == 364

    def has_nonleaf_child(i):
        return any(c["kind"] != "LEAF" for c in nodes[i].get("children") or [])   # is_leaf() is by kind (clade.rs:166-172)

    k = pin["k"]
    n_id = n_max = n_full = 0
    for tag, length, code, node, one, rest in pin["records"]:
        if code == "IdentityFound":
            n_id += 1
            assert nodes[node]["kind"] != "LEAF" and not has_nonleaf_child(node), (tag, node)
            assert one > rest >= 0
            assert one <= 2 * (length - k + 1)
            n_full += one == 2 * (length - k + 1)
        elif code.startswith("MaxResolutionReached"):
            n_max += 1
            assert code == "MaxResolutionReached: LCA Accepted" and has_nonleaf_child(node), (tag, node)
        else:
            assert code.startswith("Unclassifiable")          # the older layout (fd8) carries no message
    assert (n_id, n_max, len(pin["records"])) == (637, 472, 1111) and n_full == 328
