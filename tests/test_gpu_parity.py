"""Parity of the sm_100a path (through the C ABI) against the CPU oracle.  Bit-exact: hashes,
statuses, placement nodes, one/rest, all counters.  The only floating point on the path is
round(n_matched * coverage) in f64 (place_sequence.rs:231-232) - exact, no tolerance needed."""
import threading

import numpy as np
import pytest

from helpers import assert_rows_equal, outcome_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cq():
    import classeq2_b200
    return classeq2_b200


@pytest.fixture(scope="module", params=["closed", "general"])
def col_index(request, cq, col_flat):
    """The Colletotrichum model on the GPU in both node-set record formats: terminal lists + LCA
    jumps (picked automatically: the builder's sets are upward closed) and general mini-trees."""
    flat = col_flat if request.param == "closed" else col_flat.with_general_sets()
    ix = cq.Index(flat, device=0)
    assert ix.info()["closed_sets"] == (1 if request.param == "closed" else 0)
    yield ix
    ix.close()


def _rand_seq(rng, n):
    return "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))


# ---- (3) k-mer extraction + murmur3 ----------------------------------------------------------------
@pytest.mark.parametrize("k", [35, 1, 2, 4, 8, 9, 15, 16, 17, 31, 32, 33, 36, 48, 64])
def test_kmer_hashes_bit_exact(cq, oracle, k):
    rng = np.random.default_rng(k)
    for L in sorted({k, k + 1, k + 2, k + 15, k + 16, k + 17, 150, 151, 163, 1911}):
        if L < k:
            continue
        s = _rand_seq(rng, L)
        want = [h for _, h in oracle.KmersMap(k, 0).build_kmer_from_string(s)]
        got = cq.debug_kmer_hashes(s, k)
        assert got.tolist() == want, (k, L)
    assert len(cq.debug_kmer_hashes("ACGT", 35)) == 0  # shorter than k: no k-mers (kmers_map.rs:383-385)
    assert cq.debug_kmer_hashes("acgtacgtac", 4).tolist() == cq.debug_kmer_hashes("ACGTACGTAC", 4).tolist()


def test_gyrb_query_has_3754_distinct_kmers(cq, pins, col_queries):
    seq = dict(col_queries)[pins["gyrb_first_query"]["header"]]
    h = cq.debug_kmer_hashes(seq, 35)
    assert len(h) == len(set(h.tolist())) == pins["gyrb_first_query"]["one"]


# ---- whole path on the Colletotrichum fixture -------------------------------------------------------
@pytest.mark.parametrize("knob", ["default", "remove_intersection", "cov1", "iter2"])
def test_colletotrichum_golden(cq, col_index, col_queries, col_expected, knob):
    kn = next(k for k in col_expected["knobs"] if k["name"] == knob)
    params = cq.PlaceParams(kn["max_iterations"], kn["min_match_coverage"], kn["remove_intersection"])
    res = col_index.place_batch([s for _, s in col_queries], params)
    assert_rows_equal(res, col_expected["outcomes"][knob], [h for h, _ in col_queries])


def test_resident_path_equals_batch_path(cq, col_index, col_queries, col_expected):
    seqs = [s for _, s in col_queries]
    rb = col_index.upload(seqs)
    rb.place()
    res = rb.fetch()
    assert_rows_equal(res, col_expected["outcomes"]["default"])
    rb.place(cq.PlaceParams(remove_intersection=True))   # place again on the same resident batch
    assert_rows_equal(rb.fetch(), col_expected["outcomes"]["remove_intersection"])
    assert rb.nbytes() > 0
    rb.close()


def test_batch_order_and_ragged_inputs(cq, col_index, col_queries, col_expected):
    """Results come back in caller order whatever the length ordering on the device."""
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(col_queries))
    res = col_index.place_batch([col_queries[i][1] for i in perm])
    assert_rows_equal(res, [col_expected["outcomes"]["default"][i] for i in perm])
    # empty batch, all-too-short batch, single query
    assert col_index.place_batch([]).n == 0
    r = col_index.place_batch(["", "ACGT", "A" * 34])
    assert r.status.tolist() == [0, 0, 0] and r.n_query_kmers.tolist() == [0, 0, 0]
    one = col_index.place_batch([col_queries[0][1]])
    assert_rows_equal(one, [col_expected["outcomes"]["default"][0]])


def test_invalid_base_is_a_per_query_status(cq, col_index, col_queries, col_expected):
    from classeq2_b200 import _lib
    good = col_queries[0][1]
    bad = good[:50] + "N" + good[51:]
    res = col_index.place_batch([good, bad, good.lower()])
    assert res.status[1] == _lib.STATUS_ERR_INVALID_BASE
    e = col_expected["outcomes"]["default"][0]
    assert_rows_equal(res, [e], None) if False else None
    assert res.row(0) == res.row(2)
    assert res.node_id[0] == e["clade"]


def test_concurrent_calls_on_one_handle(cq, col_index, col_queries, col_expected):
    seqs = [s for _, s in col_queries]
    out, errs = {}, []

    def work(t):
        try:
            for _ in range(3):
                out[t] = col_index.place_batch(seqs)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for t in range(4):
        assert_rows_equal(out[t], col_expected["outcomes"]["default"])


# ---- adversarial random models (sparse ids, multifurcations, sets that are not upward closed,
#      sets without the root, leaf-kind nodes with children, m = 0 / m > k, small k) ---------------
def _random_tree_model(oracle, rng, k, m):
    O = oracle
    root = O.Clade(int(rng.integers(0, 5)), None, "ROOT", children=[])
    nodes, used = [root], {root.id}

    def new_id():
        while True:
            i = int(rng.integers(0, 2000)) if rng.random() < 0.9 else int(rng.integers(2**40, 2**63))
            if i not in used:
                used.add(i)
                return i

    for _ in range(int(rng.integers(1, 40))):
        p = nodes[int(rng.integers(len(nodes)))]
        if p.kind == "LEAF" and rng.random() < 0.9:
            continue
        kind = "LEAF" if rng.random() < 0.45 else "NODE"
        c = O.Clade(new_id(), p.id, kind, name="x" if kind == "LEAF" else None,
                    children=[] if (kind == "NODE" and rng.random() < 0.5) else None)
        p.children = (p.children or []) + [c]
        nodes.append(c)
    tree = O.Tree("t", "t", 70.0, root)
    # reference sequences: a few related strings so that queries hit many entries
    base = _rand_seq(rng, int(rng.integers(k + 5, 120)))
    km = O.KmersMap(k, m)
    ids = [n.id for n in nodes]
    paths = {}
    for n in nodes:  # root->node paths
        pass

    def path_to(n):
        out, cur = [n.id], n
        while cur.parent is not None:
            cur = next(x for x in nodes if x.id == cur.parent)
            out.append(cur.id)
        return set(out)

    for t in range(int(rng.integers(1, 8))):
        s = list(base)
        for i in range(len(s)):
            if rng.random() < 0.05:
                s[i] = "ACGT"[int(rng.integers(4))]
        s = "".join(s)
        n = nodes[int(rng.integers(len(nodes)))]
        mode = rng.random()
        if mode < 0.6:
            nset = path_to(n)                       # builder-like: upward closed
        elif mode < 0.8:
            nset = {ids[int(j)] for j in rng.integers(0, len(ids), int(rng.integers(1, 5)))}  # arbitrary
        elif mode < 0.9:
            nset = path_to(n) - {root.id}           # no root
        else:
            nset = path_to(n) | {int(rng.integers(3000, 4000))}  # id that is not in the tree
        for kmer, h in km.build_kmer_from_string(s):
            km.insert_or_append_kmer_hash(kmer, h, nset)
    tree.kmers_map = km
    queries = []
    for q in range(12):
        a = int(rng.integers(0, max(1, len(base) - k)))
        s = list(base[a:a + int(rng.integers(0, len(base) + 10))])
        for i in range(len(s)):
            if rng.random() < 0.03:
                s[i] = "ACGT"[int(rng.integers(4))]
        s = "".join(s)
        if rng.random() < 0.3:
            s = O.KmersMap.reverse_complement(s)
        queries.append((f"q{q}", s))
    queries.append(("rand", _rand_seq(rng, 90)))
    return tree, queries


@pytest.mark.parametrize("seed", range(60))
def test_random_models(cq, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.choice([35, 35, 35, 5, 11, 16, 21, 32, 40]))
    m = int(rng.choice([4, 4, 0, 1, 2, 7, k + 3 if k < 9 else 3]))
    tree, queries = _random_tree_model(oracle, rng, k, m)
    if rng.random() < 0.1:
        tree.root.children = None
    flat = tree_to_product(cq, tree)
    knobs = [dict(), dict(remove_intersection=True), dict(min_match_coverage=1.0),
             dict(max_iterations=int(rng.integers(0, 3)), min_match_coverage=0.0)]
    wants = [[outcome_of(oracle, h, s, tree, kn.get("max_iterations"), kn.get("min_match_coverage"),
                         kn.get("remove_intersection")) for h, s in queries] for kn in knobs]
    for f in (flat, flat.with_general_sets()):
        ix = cq.Index(f, device=0)
        for kn, want in zip(knobs, wants):
            res = ix.place_batch([s for _, s in queries], cq.PlaceParams(**kn))
            assert_rows_equal(res, want, [h for h, _ in queries])
        ix.close()


def tree_to_product(cq, otree):
    """oracle Tree -> product Tree (same serde object shape) -> FlatModel."""
    t = cq.Tree.from_obj(otree.to_obj())
    return cq.FlatModel.from_tree(t)


# ---- synthetic configs at reduced size against the Python oracle ------------------------------------
def test_synthetic_small_config(cq, oracle):
    from classeq2_b200 import synth
    sm = synth.make_model(60, 300, 4242)
    bases, offsets, truth = synth.make_reads(sm.ref_codes, sm.ref_lens, 300, 150, 4244)
    lens = synth.skewed_lengths(40, 5) // 5 + 35
    b2, o2, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, 40, lens, 4245)
    ix = cq.Index(sm.flat, device=0)
    ixg = cq.Index(sm.flat.with_general_sets(), device=0)
    otree = oracle_tree_from_flat(oracle, sm.flat)
    for bs, of in ((bases, offsets), (b2, o2)):
        seqs = [bytes(bs[int(of[i]):int(of[i + 1])]).decode() for i in range(len(of) - 1)]
        for kn in (dict(), dict(remove_intersection=True)):
            want = [outcome_of(oracle, f"r{i}", s, otree, None, None, kn.get("remove_intersection"))
                    for i, s in enumerate(seqs)]
            assert_rows_equal(ix.place_batch((bs, of), cq.PlaceParams(**kn)), want)
            assert_rows_equal(ixg.place_batch((bs, of), cq.PlaceParams(**kn)), want)
    info = ix.info()
    assert info["n_entries"] == sm.flat.n_entries and info["k_size"] == 35
    assert info["closed_sets"] == 1 and ixg.info()["closed_sets"] == 0
    ix.close()
    ixg.close()


# ---- builder-like (upward closed) random models: multifurcations, deep trees, big shared sets,
#      every knob; checked against the C++ oracle (itself held equal to the Python one on CPU) ----------
def _closed_model(rng, n_internal, max_kids, n_tips_per, k, m, ref_len):
    """Random rooted tree with multifurcations; refs evolved down the tree; node sets = unions of
    root->tip paths (the reference builder's invariant)."""
    from classeq2_b200 import _lib
    from classeq2_b200.model import BuiltModel, FlatModel
    parent, kind = [-1], [_lib.KIND_ROOT]
    internal = [0]
    for _ in range(n_internal):
        p = internal[int(rng.integers(len(internal)))] if rng.random() < 0.7 else internal[-1]
        parent.append(p), kind.append(_lib.KIND_NODE)
        internal.append(len(parent) - 1)
    tips = []
    for p in internal:
        for _ in range(int(rng.integers(0, n_tips_per + 1))):
            parent.append(p), kind.append(_lib.KIND_LEAF)
            tips.append(len(parent) - 1)
    n = len(parent)
    kids = [[] for _ in range(n)]
    for c in range(1, n):
        kids[parent[c]].append(c)
    child_off = np.zeros(n + 1, np.uint64)
    child_idx = []
    for i in range(n):
        rng.shuffle(kids[i])
        child_idx += kids[i]
        child_off[i + 1] = len(child_idx)
    node_id = (rng.permutation(4 * n)[:n] + 1).astype(np.uint64)   # sparse, shuffled ids
    if rng.random() < 0.5:
        node_id[0] = 0
    seqs = [None] * n
    seqs[0] = rng.integers(0, 4, ref_len, dtype=np.uint8)
    order = [0]
    for v in order:
        for c in kids[v]:
            s = seqs[v].copy()
            mut = np.flatnonzero(rng.random(ref_len) < 0.03)
            s[mut] = (s[mut] + rng.integers(1, 4, len(mut), dtype=np.uint8)) & 3
            seqs[c] = s
            order.append(c)
    ascii_ = np.frombuffer(b"ACGT", np.uint8)
    tflat = FlatModel(k, m, node_id, np.array(kind, np.uint8), child_off, np.array(child_idx, np.uint64))
    tip_bases = np.concatenate([ascii_[seqs[t]] for t in tips]) if tips else np.zeros(0, np.uint8)
    tip_off = np.arange(len(tips) + 1, dtype=np.uint64) * ref_len
    bm = BuiltModel(tflat, np.array(tips, np.uint64), tip_bases, tip_off)
    a = bm.arrays()
    bm.close()
    flat = FlatModel(k, m, node_id, np.array(kind, np.uint8), child_off, np.array(child_idx, np.uint64),
                     a["entry_bucket"], a["entry_hash"], a["entry_set"], a["set_off"], a["set_node_ids"])
    # queries: fragments of tip and internal sequences, both strands, some chimeric / random
    qs = []
    for _ in range(200):
        src = seqs[int(rng.integers(n))]
        ln = int(rng.integers(k - 2, ref_len + 1))
        st = int(rng.integers(0, ref_len - ln + 1))
        s = src[st:st + ln].copy()
        if rng.random() < 0.3:
            other = seqs[int(rng.integers(n))]
            cut = int(rng.integers(0, ln + 1))
            s[cut:] = other[st + cut:st + ln]
        if rng.random() < 0.5:
            s = (3 - s)[::-1]
        if rng.random() < 0.05:
            s = rng.integers(0, 4, ln, dtype=np.uint8)
        qs.append(ascii_[s].tobytes().decode())
    return flat, qs


@pytest.mark.parametrize("seed", range(24))
def test_random_closed_models(cq, seed):
    from oracle import cpp_oracle
    rng = np.random.default_rng(7000 + seed)
    k = int(rng.choice([35, 35, 35, 21, 9]))
    m = int(rng.choice([4, 4, 0, 2]))
    flat, qs = _closed_model(rng, n_internal=int(rng.integers(1, 80)), max_kids=0,
                             n_tips_per=int(rng.integers(1, 4)), k=k, m=m, ref_len=int(rng.integers(60, 260)))
    md = cpp_oracle.CppModel.from_flat(flat)
    bases, offsets = cq.make_batch(qs)
    ix, ixg = cq.Index(flat, device=0), cq.Index(flat.with_general_sets(), device=0)
    assert ix.info()["closed_sets"] == 1
    for kn in [dict(), dict(remove_intersection=True), dict(min_match_coverage=0.0, max_iterations=3),
               dict(min_match_coverage=1.0, max_iterations=1)]:
        want = md.place_batch(bases, offsets, kn.get("max_iterations"), kn.get("min_match_coverage"), kn.get("remove_intersection"))
        for index in (ix, ixg):
            got = index.place_batch((bases, offsets), cq.PlaceParams(**kn))
            for f in ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched"):
                bad = np.flatnonzero(getattr(got, f) != want[f])
                assert bad.size == 0, (f, kn, bad[:5], getattr(got, f)[bad[:5]], want[f][bad[:5]])
            ok_it = (got.iterations == want["iterations"]) | (want["status"] == 8)
            assert ok_it.all()
    ix.close(), ixg.close(), md.close()


def oracle_tree_from_flat(oracle, flat):
    O = oracle
    kinds = {0: "ROOT", 1: "NODE", 2: "LEAF"}
    n = len(flat.node_id)
    clades = [O.Clade(int(flat.node_id[i]), None, kinds[int(flat.node_kind[i])]) for i in range(n)]
    for i in range(n):
        a, b = int(flat.child_off[i]), int(flat.child_off[i + 1])
        if b > a:
            clades[i].children = [clades[int(j)] for j in flat.child_idx[a:b]]
            for c in clades[i].children:
                c.parent = clades[i].id
    km = O.KmersMap(flat.k_size, flat.m_size)
    so, sn = flat.set_off, flat.set_node_ids
    for b, h, s in zip(flat.entry_bucket.tolist(), flat.entry_hash.tolist(), flat.entry_set.tolist()):
        km.map.setdefault(b, {})[h] = set(sn[int(so[s]):int(so[s + 1])].tolist())
    return O.Tree("s", "s", 70.0, clades[0], kmers_map=km)


# ---- kb-scale reads: one CTA per read (all warps share one set of tables) --------------------------
@pytest.mark.parametrize("general", [False, True])
def test_long_reads_cta_mode(cq, general):
    from classeq2_b200 import synth
    from oracle import cpp_oracle
    sm = synth.make_model(48, 1600, 9090)
    lens = np.concatenate([synth.skewed_lengths(300, 77), np.array([35, 36, 299, 300, 301, 512, 1024, 1600, 1600])])
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, len(lens), lens, 9092)
    md = cpp_oracle.CppModel.from_flat(sm.flat)
    ix = cq.Index(sm.flat.with_general_sets() if general else sm.flat, device=0)
    for kn in (dict(), dict(remove_intersection=True), dict(min_match_coverage=0.0, max_iterations=2)):
        want = md.place_batch(bases, offsets, kn.get("max_iterations"), kn.get("min_match_coverage"), kn.get("remove_intersection"))
        got = ix.place_batch((bases, offsets), cq.PlaceParams(**kn))
        for f in ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched"):
            bad = np.flatnonzero(getattr(got, f) != want[f])
            assert bad.size == 0, (f, kn, bad[:5], getattr(got, f)[bad[:5]], want[f][bad[:5]], lens[bad[:5]])
    # low-complexity reads: every window repeats (distinct-hash semantics across warps of the CTA)
    rep = ["ACGT" * 400, "A" * 1200, ("ACGTTGCA" * 200)[:1550]]
    b2, o2 = cq.make_batch(rep)
    want = md.place_batch(b2, o2)
    got = ix.place_batch((b2, o2))
    assert got.status.tolist() == want["status"].tolist() and got.n_matched.tolist() == want["n_matched"].tolist()
    ix.close(), md.close()


def test_long_reads_foreign_bucket_keys(cq):
    """kb-scale reads against a model in which some entries sit under the bucket key of ANOTHER prefix than their own
    k-mer's: such a hit counts when any window of the query has that prefix (kmers_map.rs:55-70).  The fragment kernel
    cannot settle that from one fragment and hands the read back to the placement kernel (frag_kernels.cuh, `redo`)."""
    from classeq2_b200 import synth
    from classeq2_b200.model import FlatModel
    from oracle import cpp_oracle
    rng = np.random.default_rng(515)
    sm = synth.make_model(48, 1600, 9191)
    f0 = sm.flat
    eb = f0.entry_bucket.copy()
    idx = rng.choice(len(eb), len(eb) // 40, replace=False)
    eb[idx] = eb[rng.permutation(idx)]          # keys of other entries: all of them keys of real ACGT prefixes
    flat = FlatModel(f0.k_size, f0.m_size, f0.node_id, f0.node_kind, f0.child_off, f0.child_idx, eb, f0.entry_hash, f0.entry_set,
                     f0.set_off, f0.set_node_ids)
    lens = np.concatenate([synth.skewed_lengths(120, 78), np.array([150, 162, 163, 290, 291, 400, 1600])])
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, len(lens), lens, 9193)
    md = cpp_oracle.CppModel.from_flat(flat)
    want = md.place_batch(bases, offsets)
    base_want = cpp_oracle.CppModel.from_flat(f0).place_batch(bases, offsets)
    assert (want["n_matched"] != base_want["n_matched"]).any()   # the moved keys do change what counts
    ix = cq.Index(flat, device=0)
    got = ix.place_batch((bases, offsets))
    for f in ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched"):
        bad = np.flatnonzero(getattr(got, f) != want[f])
        assert bad.size == 0, (f, bad[:5], getattr(got, f)[bad[:5]], want[f][bad[:5]], lens[bad[:5]])
    ix.close(), md.close()


# ---- reads with MANY distinct node sets: every hand-over / fallback of the short-read path ----------
def _many_sets_model(cq, oracle, rng, n_tips, n_reads, pool):
    """A model whose k-mer index holds exactly the windows of `n_reads` random 150-mers, every window
    filed under a node set drawn from a pool of `pool` random upward-closed sets (None: one set per
    window), so that a read sees up to 232 distinct node sets."""
    from classeq2_b200 import synth
    from classeq2_b200.model import FlatModel
    tree = synth.make_tree(n_tips, int(rng.integers(1 << 30)))
    n = len(tree.node_id)
    parent = np.full(n, -1, np.int64)
    for p in range(n):
        for j in range(int(tree.child_off[p]), int(tree.child_off[p + 1])):
            parent[int(tree.child_idx[j])] = p

    def path_ids(tips):
        nodes = set()
        for t in tips:
            v = int(tree.tip_node[t])
            while v >= 0:
                nodes.add(int(tree.node_id[v]))
                v = int(parent[v])
        return sorted(nodes)

    def rand_set():
        return path_ids(rng.choice(n_tips, size=int(rng.integers(1, 5)), replace=False))

    sets = [rand_set() for _ in range(pool)] if pool else []
    reads = [_rand_seq(rng, 150) for _ in range(n_reads)]
    km = oracle.KmersMap(35, 4)
    eb, eh, es = [], [], []
    seen = set()
    for s in reads:
        for kmer, h in km.build_kmer_from_string(s):
            if h in seen:
                continue
            seen.add(h)
            if pool:
                si = int(rng.integers(pool))
            else:
                sets.append(rand_set())
                si = len(sets) - 1
            eb.append(cq.host_murmur3_h1(kmer[:4].encode())), eh.append(h), es.append(si)
    set_off = np.zeros(len(sets) + 1, np.uint64)
    set_off[1:] = np.cumsum([len(x) for x in sets])
    flat = FlatModel(35, 4, tree.node_id, tree.node_kind, tree.child_off, tree.child_idx,
                     np.array(eb, np.uint64), np.array(eh, np.uint64), np.array(es, np.uint64), set_off,
                     np.array([v for x in sets for v in x], np.uint64))
    # queries: the reads themselves, mutated copies (fewer hits) and chimeras of two reads
    qs = list(reads)
    for s in reads:
        b = bytearray(s.encode())
        for i in rng.integers(0, 150, 3):
            b[int(i)] = b"ACGT"[int(rng.integers(4))]
        qs.append(b.decode())
    for i in range(n_reads - 1):
        qs.append(reads[i][:75] + reads[i + 1][75:])
    return flat, qs


@pytest.mark.parametrize("pool", [8, 30, 50, 64, 100, None])
def test_many_distinct_node_sets_per_read(cq, oracle, pool):
    from classeq2_b200.parallel import LocalShardedPlacer
    from oracle import cpp_oracle
    rng = np.random.default_rng(4242 + (pool or 0))
    flat, qs = _many_sets_model(cq, oracle, rng, n_tips=90, n_reads=24, pool=pool)
    md = cpp_oracle.CppModel.from_flat(flat)
    bases, offsets = cq.make_batch(qs)
    ix = cq.Index(flat, device=0)
    assert ix.info()["closed_sets"] == 1
    sharded = LocalShardedPlacer(flat, 0, 4)
    for kn in [dict(), dict(remove_intersection=True), dict(min_match_coverage=0.1, max_iterations=4)]:
        want = md.place_batch(bases, offsets, kn.get("max_iterations"), kn.get("min_match_coverage"), kn.get("remove_intersection"))
        rb = ix.upload((bases, offsets))
        rb.place(cq.PlaceParams(**kn))
        for got in (ix.place_batch((bases, offsets), cq.PlaceParams(**kn)), rb.fetch(), sharded.place((bases, offsets), cq.PlaceParams(**kn))):
            for f in ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched"):
                bad = np.flatnonzero(getattr(got, f) != want[f])
                assert bad.size == 0, (f, kn, pool, bad[:5], getattr(got, f)[bad[:5]], want[f][bad[:5]])
            assert ((got.iterations == want["iterations"]) | (want["status"] == 8)).all()
        rb.close()
    md.close(), ix.close()


def test_repeated_kmers_in_short_reads(cq, oracle):
    """Distinct-hash semantics (the reference collects hashes in HashSets): tandem repeats put the same
    k-mer at several positions of a strand, hairpins put it on both strands; every copy must count once."""
    from classeq2_b200 import synth
    from classeq2_b200.model import BuiltModel, FlatModel
    from classeq2_b200.parallel import LocalShardedPlacer
    from oracle import cpp_oracle
    rng = np.random.default_rng(99)
    tree = synth.make_tree(12, 5)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rc = lambda s: "".join(comp[c] for c in reversed(s))
    tips = []
    for t in range(12):
        unit, stem = _rand_seq(rng, 41 + t), _rand_seq(rng, 70)
        tips.append(unit * 4 + stem + rc(stem))          # period-(41+t) repeat, then a perfect hairpin
    bases = np.frombuffer("".join(tips).encode(), np.uint8).copy()
    offs = np.zeros(13, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in tips])
    tflat = synth.tree_only_flat(tree)
    bm = BuiltModel(tflat, tree.tip_node, bases, offs)
    a = bm.arrays()
    bm.close()
    flat = FlatModel(35, 4, tree.node_id, tree.node_kind, tree.child_off, tree.child_idx,
                     a["entry_bucket"], a["entry_hash"], a["entry_set"], a["set_off"], a["set_node_ids"])
    qs = []
    for s in tips:
        for st in range(0, len(s) - 150, 13):
            qs.append(s[st:st + 150])
            qs.append(rc(s[st:st + 140]))
        qs.append(s[-140:])                              # the hairpin: forward and reverse strand share their k-mers
    md = cpp_oracle.CppModel.from_flat(flat)
    b, o = cq.make_batch(qs)
    want = md.place_batch(b, o)
    assert (want["n_matched"] < want["n_query_kmers"]).sum() > len(qs) // 2   # the repeats really collapse
    ix = cq.Index(flat, device=0)
    for got in (ix.place_batch((b, o)), LocalShardedPlacer(flat, 0, 3).place((b, o))):
        for f in ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched", "iterations"):
            bad = np.flatnonzero(getattr(got, f) != want[f])
            assert bad.size == 0, (f, bad[:5], getattr(got, f)[bad[:5]], want[f][bad[:5]])
    md.close(), ix.close()


# ---- per-node hit counts, level by level (cls_debug_node_counts) ------------------------------------
@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("ri", [False, True])
def test_node_counts_per_level_colletotrichum(cq, oracle, col_flat, col_tree, col_queries, general, ri):
    """|K(c)|, exclusive and union counts of every non-leaf child with votes at every level of the walk,
    against the oracle's sets (place_sequence.rs:311-428), on the reference's own data fixture
    (multifurcating tree, 35 .. 1911 bp queries: both the one-warp and the one-CTA geometry)."""
    flat = col_flat.with_general_sets() if general else col_flat
    ix = cq.Index(flat, device=0)
    params = cq.PlaceParams(remove_intersection=ri)
    want_all = ix.place_batch([s for _, s in col_queries], params)
    n_rows = 0
    for qi, (h, s) in enumerate(col_queries):
        if len(s) < 35 or (qi % 3 and len(s) > 400):
            continue
        trace = []
        try:
            oracle.place_sequence(h, s, col_tree, None, None, ri, trace=trace)
        except oracle.PlacementError:
            pass
        trace.sort(key=lambda r: (r["level"], r["child_id"]))
        rows, res = ix.debug_node_counts(s, params)
        assert rows == trace, (h, rows[:3], trace[:3])
        for f, _ in cq.engine.RESULT_DTYPES:   # the traced walk gives the production path's result
            assert getattr(res, f)[0] == getattr(want_all, f)[qi], (h, f)
        n_rows += len(rows)
    assert n_rows > 500
    ix.close()


def test_node_counts_per_level_synthetic(cq, oracle):
    from classeq2_b200 import synth
    sm = synth.make_model(60, 300, 31)
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, 40, 150, 33)
    tree = oracle_tree_from_flat(oracle, sm.flat)
    ix = cq.Index(sm.flat, device=0)
    for i in range(40):
        s = bases[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()
        trace = []
        oracle.place_sequence("q", s, tree, trace=trace)
        trace.sort(key=lambda r: (r["level"], r["child_id"]))
        rows, _ = ix.debug_node_counts(s)
        assert rows == trace, (i, rows[:3], trace[:3])
    ix.close()
