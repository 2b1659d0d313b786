"""Synthetic models of BASELINE.json's configs: the seeded data generator (``synth_data``: trees, reference
sequences, reads - pure numpy, importable without the library) plus the k-mer map built by the library."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .model import BuiltModel, FlatModel
from .synth_data import (ASCII, CONFIGS, SynthTree, make_reads, make_refs, make_tree, refs_to_batch,  # noqa: F401
                         skewed_lengths)


def tree_only_flat(tree: SynthTree, k_size: int = 35, m_size: int = 4) -> FlatModel:
    return FlatModel(k_size, m_size, tree.node_id, tree.node_kind, tree.child_off, tree.child_idx)


@dataclass
class SynthModel:
    tree: SynthTree
    ref_codes: np.ndarray
    ref_lens: np.ndarray
    flat: FlatModel       # full model (tree + entries + sets), owns its arrays


def make_model(n_tips: int, l_ref, seed: int, k_size: int = 35, m_size: int = 4, device: Optional[int] = None) -> SynthModel:
    """Tree(N, seed) + Refs(N, Lref, seed+1) + the k-mer map built by ``cls_model_build`` (host) or, with
    ``device`` given, by ``cls_model_build_device`` on that GPU (the same map, an order of magnitude sooner)."""
    tree = make_tree(n_tips, seed)
    codes, lens = make_refs(tree, l_ref, seed + 1)
    tflat = tree_only_flat(tree, k_size, m_size)
    bases, offsets = refs_to_batch(codes, lens)
    bm = BuiltModel(tflat, tree.tip_node, bases, offsets, device=device)
    arr = bm.arrays()
    bm.close()
    flat = FlatModel(k_size, m_size, tree.node_id, tree.node_kind, tree.child_off, tree.child_idx,
                     arr["entry_bucket"], arr["entry_hash"], arr["entry_set"], arr["set_off"], arr["set_node_ids"])
    return SynthModel(tree, codes, lens, flat)


