"""Edge cases of the second-generation scan kernel (csrc/scan2_kernels.cuh), through the C ABI against the C++ oracle:
read lengths around every pass / geometry boundary, the same k-mer several times inside ONE pass of 32 windows
(short-period repeats: the exact answer behind the test-and-set filter), reads with more list entries than the
per-read list holds (overflow list -> first-generation kernel), and the filler hashes of free table slots."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched", "iterations")


@pytest.fixture(scope="module")
def cq():
    import classeq2_b200
    return classeq2_b200


def _rand_seq(rng, n):
    return "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))


def _built(cq, tree, tips):
    from classeq2_b200 import synth
    from classeq2_b200.model import BuiltModel, FlatModel
    bases = np.frombuffer("".join(tips).encode(), np.uint8).copy()
    offs = np.zeros(len(tips) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in tips])
    bm = BuiltModel(synth.tree_only_flat(tree), tree.tip_node, bases, offs)
    a = bm.arrays()
    bm.close()
    return FlatModel(35, 4, tree.node_id, tree.node_kind, tree.child_off, tree.child_idx,
                     a["entry_bucket"], a["entry_hash"], a["entry_set"], a["set_off"], a["set_node_ids"])


def _check(cq, flat, qs, knobs=(dict(), dict(remove_intersection=True))):
    from oracle import cpp_oracle
    md = cpp_oracle.CppModel.from_flat(flat)
    b, o = cq.make_batch(qs)
    ix = cq.Index(flat, device=0)
    assert ix.info()["closed_sets"] == 1
    for kn in knobs:
        want = md.place_batch(b, o, kn.get("max_iterations"), kn.get("min_match_coverage"), kn.get("remove_intersection"))
        rb = ix.upload((b, o))
        rb.place(cq.PlaceParams(**kn))
        for got in (ix.place_batch((b, o), cq.PlaceParams(**kn)), rb.fetch()):
            for f in FIELDS:
                bad = np.flatnonzero(getattr(got, f) != want[f])
                assert bad.size == 0, (f, kn, bad[:5], [len(qs[i]) for i in bad[:5]], getattr(got, f)[bad[:5]], want[f][bad[:5]])
        rb.close()
    md.close(), ix.close()
    return want


def test_read_lengths_around_every_boundary(cq):
    """35 (one window) ... 162 (four passes per strand), 163 ... 290 (eight), 291+ (one CTA per read)."""
    from classeq2_b200 import synth
    sm = synth.make_model(40, 400, 777)
    rng = np.random.default_rng(5)
    lens = [35, 36, 37, 50, 65, 66, 67, 68, 97, 98, 99, 100, 129, 130, 131, 132, 149, 150, 151, 160, 161, 162,
            163, 164, 193, 194, 195, 225, 226, 227, 257, 258, 259, 288, 289, 290, 291, 292, 300, 399, 400]
    lens = np.array(lens * 6, np.int64)
    rng.shuffle(lens)
    b, o, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, len(lens), lens, 778, p_err=0.003, frac_random=0.1)
    qs = [bytes(b[int(o[i]):int(o[i + 1])]).decode() for i in range(len(lens))]
    want = _check(cq, sm.flat, qs)
    assert (want["status"] == 6).sum() > len(qs) // 2
    # the same lengths one class at a time (a launch whose longest read sits exactly on the boundary)
    for L in (35, 66, 67, 162, 163, 290, 291):
        _check(cq, sm.flat, [q for q in qs if len(q) == L], knobs=(dict(),))


@pytest.mark.parametrize("period", [1, 2, 3, 5, 7, 10, 16, 31, 32, 33])
def test_same_kmer_several_times_in_one_pass(cq, period):
    """Tandem repeats with a period below the pass width put one k-mer on several lanes of a pass; hairpins put
    it on both strands.  Every copy counts once (HashSet semantics, kmers_map.rs:273-311)."""
    from classeq2_b200 import synth
    rng = np.random.default_rng(100 + period)
    tree = synth.make_tree(8, 9)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rc = lambda s: "".join(comp[c] for c in reversed(s))
    tips = []
    for t in range(8):
        unit = _rand_seq(rng, period)
        flank = _rand_seq(rng, 60)
        tips.append(flank + unit * (200 // period + 2) + rc(flank) + unit * (80 // period + 1))
    flat = _built(cq, tree, tips)
    qs = []
    for s in tips:
        for st in range(0, len(s) - 150, 11):
            qs.append(s[st:st + 150])
            qs.append(rc(s[st:st + 120]))
            qs.append(s[st:st + 60])
    want = _check(cq, flat, qs)
    if period < 32:
        assert (want["n_matched"] * 2 < want["n_query_kmers"]).sum() > len(qs) // 4


def test_more_list_entries_than_the_list_holds(cq, oracle):
    """Every k-mer of the reads carries its own node set: 232 list entries per 150-base read against a list of
    128 -> the overflow list and the first-generation kernel; mixed with ordinary reads in one launch."""
    from classeq2_b200 import synth
    from classeq2_b200.model import FlatModel
    rng = np.random.default_rng(31)
    tree = synth.make_tree(120, 8)
    parent = tree.parent
    n_tips = tree.n_tips

    def path_ids(ts):
        nodes = set()
        for t in ts:
            v = int(tree.tip_node[t])
            while v >= 0:
                nodes.add(int(tree.node_id[v]))
                v = int(parent[v])
        return sorted(nodes)

    reads = [_rand_seq(rng, 150) for _ in range(40)]
    km = oracle.KmersMap(35, 4)
    sets, eb, eh, es, seen = [], [], [], [], set()
    for i, s in enumerate(reads):
        shared = path_ids(rng.choice(n_tips, size=2, replace=False))
        for j, (kmer, h) in enumerate(km.build_kmer_from_string(s)):
            if h in seen:
                continue
            seen.add(h)
            if i % 2 == 0:      # a set of its own for every k-mer
                sets.append(path_ids(rng.choice(n_tips, size=int(rng.integers(1, 4)), replace=False)))
            elif j % 40 == 0:   # a handful of sets per read
                sets.append(shared if j == 0 else path_ids(rng.choice(n_tips, size=2, replace=False)))
            eb.append(cq.host_murmur3_h1(kmer[:4].encode())), eh.append(h), es.append(len(sets) - 1)
    set_off = np.zeros(len(sets) + 1, np.uint64)
    set_off[1:] = np.cumsum([len(x) for x in sets])
    flat = FlatModel(35, 4, tree.node_id, tree.node_kind, tree.child_off, tree.child_idx,
                     np.array(eb, np.uint64), np.array(eh, np.uint64), np.array(es, np.uint64), set_off,
                     np.array([v for x in sets for v in x], np.uint64))
    qs = list(reads) + [reads[i][:80] + reads[i + 1][80:] for i in range(len(reads) - 1)] + [_rand_seq(rng, 150) for _ in range(10)]
    want = _check(cq, flat, qs, knobs=(dict(), dict(remove_intersection=True), dict(min_match_coverage=0.05, max_iterations=6)))
    assert (want["n_matched"] > 200).sum() >= 20


def test_tiny_tables_and_free_slot_fillers(cq):
    """Models with zero, one and a few entries: the table has at least four buckets and its free slots carry
    hashes no probe can ask for; reads whose hashes fall into empty buckets must miss."""
    from classeq2_b200 import synth
    rng = np.random.default_rng(8)
    tree = synth.make_tree(4, 3)
    for n in (35, 36, 40, 75):
        tips = [_rand_seq(rng, n) for _ in range(4)]
        flat = _built(cq, tree, tips)
        qs = tips + [_rand_seq(rng, 150) for _ in range(20)] + [t + _rand_seq(rng, 50) for t in tips]
        _check(cq, flat, qs, knobs=(dict(),))
