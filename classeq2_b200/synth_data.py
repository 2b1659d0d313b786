"""Seeded synthetic workloads of BASELINE.json's configs (SURVEY.md section 8d "Common generator").

* ``Tree(N, seed)``   random binary topology by successive random splits (Yule), internal
  ``support = 100``, ids pre-order from root = 0, tips named ``tip_%06d``.
* ``Refs(N, Lref, seed+1)``  root sequence iid uniform ACGT; evolved down the tree with a
  per-branch substitution probability of 0.01 per site (substitutions only) so that clades share
  k-mers.  When ``Lref`` is a range every tip keeps a random-length prefix of its sequence.
* ``Model``  k = 35, m = 4; built on the host by ``cls_model_build`` (every tip paired with its
  own sequence; node set = ids on the root->tip path, both ends included).
* ``Reads(R, L, seed+2)``  tip uniform, start uniform, strand uniform, 1 % substitution error;
  5 % of the reads are iid random (unrelated -> ``Unclassifiable``).

PRNG: ``numpy.random.Generator(PCG64(seed))``.  Base codes inside this module: A=0 C=1 G=2 T=3
(complement = 3 - code); everything handed out is ASCII.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np


ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
KIND_ROOT, KIND_NODE, KIND_LEAF = 0, 1, 2   # cls_node_kind (include/classeq_b200.h)


@dataclass
class SynthTree:
    n_tips: int
    node_id: np.ndarray      # pre-order ids == indices
    node_kind: np.ndarray
    parent: np.ndarray       # int64, -1 for the root
    child_off: np.ndarray
    child_idx: np.ndarray
    tip_node: np.ndarray     # node index of tip t (t in pre-order of tips)
    depth: np.ndarray

    def names(self):
        return ["tip_%06d" % t for t in range(self.n_tips)]


def make_tree(n_tips: int, seed: int) -> SynthTree:
    rng = np.random.Generator(np.random.PCG64(seed))
    n_nodes = 2 * n_tips - 1
    left = np.full(n_nodes, -1, dtype=np.int64)
    right = np.full(n_nodes, -1, dtype=np.int64)
    leaves = [0]
    nxt = 1
    picks = rng.random(max(n_tips - 1, 0))
    for i in range(n_tips - 1):
        j = int(picks[i] * len(leaves))
        node = leaves[j]
        left[node], right[node] = nxt, nxt + 1
        leaves[j] = nxt
        leaves.append(nxt + 1)
        nxt += 2
    # pre-order renumbering
    order = np.empty(n_nodes, dtype=np.int64)
    new_id = np.empty(n_nodes, dtype=np.int64)
    stack = [0]
    k = 0
    while stack:
        v = stack.pop()
        order[k] = v
        new_id[v] = k
        k += 1
        if left[v] >= 0:
            stack.append(right[v])
            stack.append(left[v])
    parent = np.full(n_nodes, -1, dtype=np.int64)
    child_off = np.zeros(n_nodes + 1, dtype=np.uint64)
    child_idx = []
    kind = np.full(n_nodes, KIND_NODE, dtype=np.uint8)
    depth = np.zeros(n_nodes, dtype=np.int32)
    tips = []
    for k in range(n_nodes):
        v = order[k]
        if left[v] >= 0:
            a, b = new_id[left[v]], new_id[right[v]]
            child_idx += [a, b]
            parent[a] = parent[b] = k
            depth[a] = depth[b] = depth[k] + 1
        else:
            kind[k] = KIND_LEAF
            tips.append(k)
        child_off[k + 1] = len(child_idx)
    kind[0] = KIND_ROOT
    return SynthTree(n_tips, np.arange(n_nodes, dtype=np.uint64), kind, parent, child_off,
                     np.array(child_idx, dtype=np.uint64), np.array(tips, dtype=np.uint64), depth)


def make_refs(tree: SynthTree, l_ref: Union[int, Tuple[int, int]], seed: int, p_sub: float = 0.01):
    """Returns (codes[n_tips, Lmax] uint8, lengths[n_tips])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo, hi = (l_ref, l_ref) if isinstance(l_ref, int) else l_ref
    n_nodes = len(tree.node_id)
    seqs = np.empty((n_nodes, hi), dtype=np.uint8)
    seqs[0] = rng.integers(0, 4, hi, dtype=np.uint8)
    for v in range(1, n_nodes):  # pre-order: parents first
        s = seqs[tree.parent[v]].copy()
        mut = np.flatnonzero(rng.random(hi) < p_sub)
        if len(mut):
            s[mut] = (s[mut] + rng.integers(1, 4, len(mut), dtype=np.uint8)) & 3
        seqs[v] = s
    tips = seqs[tree.tip_node.astype(np.int64)]
    lens = rng.integers(lo, hi + 1, tree.n_tips).astype(np.int64) if hi > lo else np.full(tree.n_tips, hi, np.int64)
    return tips, lens


def refs_to_batch(codes: np.ndarray, lens: np.ndarray):
    """Flat ASCII bases + offsets of the tip sequences."""
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    mask = np.arange(codes.shape[1])[None, :] < lens[:, None]
    return ASCII[codes[mask]], offsets


def make_reads(ref_codes: np.ndarray, ref_lens: np.ndarray, n_reads: int,
               read_len: Union[int, np.ndarray], seed: int, p_err: float = 0.01, frac_random: float = 0.05,
               chunk: int = 200_000, workers: Optional[int] = None):
    """Returns (bases ASCII uint8[], offsets uint64[n_reads+1], truth_tip int64[n_reads], -1 = random).

    Chunk c of `chunk` reads draws from its own stream ``PCG64(seed).jumped(c)``: the chunks are generated side by
    side on the host cores, and the first chunks of a long batch equal a shorter batch of whole chunks."""
    n_tips = len(ref_lens)
    lens = np.full(n_reads, read_len, dtype=np.int64) if np.isscalar(read_len) else np.asarray(read_len, np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    bases = np.empty(int(offsets[-1]), dtype=np.uint8)
    truth = np.empty(n_reads, dtype=np.int64)
    lmax_ref = ref_codes.shape[1]
    flat_refs = ref_codes.reshape(-1)

    def one(ci: int):
        a = ci * chunk
        b = min(n_reads, a + chunk)
        rng = np.random.Generator(np.random.PCG64(seed).jumped(ci))
        n = b - a
        ln = lens[a:b]
        tip = rng.integers(0, n_tips, n)
        # reads longer than the reference are clipped to it
        ln_eff = np.minimum(ln, ref_lens[tip])
        start = np.floor(rng.random(n) * (ref_lens[tip] - ln_eff + 1)).astype(np.int64)
        strand = rng.random(n) < 0.5
        is_rand = rng.random(n) < frac_random
        tot = int(ln.sum())
        rid = np.repeat(np.arange(n), ln)
        off_local = np.zeros(n + 1, dtype=np.int64)
        off_local[1:] = np.cumsum(ln)
        pos = np.arange(tot) - off_local[rid]
        within = pos < ln_eff[rid]
        src_pos = np.where(strand[rid], ln_eff[rid] - 1 - pos, pos)
        src = tip[rid] * lmax_ref + start[rid] + np.where(within, src_pos, 0)
        codes = flat_refs[src]
        codes = np.where(strand[rid], 3 - codes, codes).astype(np.uint8)
        err = rng.random(tot) < p_err
        ne = int(err.sum())
        if ne:
            codes[err] = (codes[err] + rng.integers(1, 4, ne, dtype=np.uint8)) & 3
        rnd = is_rand[rid] | ~within
        nr = int(rnd.sum())
        if nr:
            codes[rnd] = rng.integers(0, 4, nr, dtype=np.uint8)
        bases[int(offsets[a]):int(offsets[b])] = ASCII[codes]
        truth[a:b] = np.where(is_rand, -1, tip)

    n_chunks = (n_reads + chunk - 1) // chunk
    if workers is None:
        workers = min(n_chunks, os.cpu_count() or 1, 16)
    if workers <= 1:
        for ci in range(n_chunks):
            one(ci)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(one, range(n_chunks)))
    return bases, offsets, truth


# The named configurations of BASELINE.json / SURVEY.md section 8d.
CONFIGS = {
    2: dict(n_tips=1_000, l_ref=1_000, tree_seed=1001, n_reads=1_000_000, read_len=150),
    3: dict(n_tips=10_000, l_ref=(550, 650), tree_seed=1002, n_reads=10_000_000, read_len=150),
    4: dict(n_tips=5_000, l_ref=1_500, tree_seed=1003, n_reads=1_000_000, read_len="skewed", len_seed=1004),
    5: dict(n_tips=100_000, l_ref=(550, 650), tree_seed=1005, n_reads=10_000_000, read_len=150),
}


def skewed_lengths(n: int, seed: int) -> np.ndarray:
    """Config 4: 70 % U[1400,1550], 20 % U[400,1400], 10 % U[150,400]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(n)
    out = np.where(u < 0.7, rng.integers(1400, 1551, n),
                   np.where(u < 0.9, rng.integers(400, 1401, n), rng.integers(150, 401, n)))
    return out.astype(np.int64)
