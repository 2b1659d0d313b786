"""Host mirror of the reference's ``build-db`` use-case: newick tree + one sequence per tip -> a
:class:`~classeq2_b200.model.Tree` with its k-mer map, ready for ``save_database`` / ``Index``.

Reference: ``map_kmers_to_tree`` (core/src/use_cases/build_database/mod.rs:26-181), ``Tree::init_from_file``
and ``sanitize`` (core/src/domain/dtos/tree.rs:164-364).  Offline, once per model - not part of the placement
hot path; the k-mer -> node-set map itself is built by the library: on the host
(``cls_model_build``) or, with ``device=``, on the GPU (``cls_model_build_device``).  One deliberate difference,
in both builders: every tip is indexed with ITS OWN sequence (the reference pairs header i with sequence i-1
and drops the last record, build_database/mod.rs:93-116).
"""
from __future__ import annotations

import os
import re
import uuid
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from .model import BuiltModel, Clade, FlatModel, KmersMap, Tree
from .engine import filter_sequence
from .placement import read_fasta

_LABEL = re.compile(r"[^(),:;]+")


def parse_newick(text: str) -> List[dict]:
    """Nodes of a newick string, numbered in PRE-ORDER of creation as phylotree does (root 0, every '(' or tip
    takes the next id when it is opened); each node: ``parent``, ``children`` (ids), ``name`` (label text: a
    tip name or an internal support value), ``edge`` (branch length).  Iterative: no recursion limit on
    caterpillar trees."""
    text = text.strip()
    if not text.endswith(";"):
        raise ValueError("newick string does not end with ';'")
    text = text[:-1]
    nodes: List[dict] = []

    def new(parent: Optional[int]) -> int:
        nodes.append({"parent": parent, "children": [], "name": None, "edge": None})
        if parent is not None:
            nodes[parent]["children"].append(len(nodes) - 1)
        return len(nodes) - 1

    pos = 0

    def label(nid: int) -> None:
        nonlocal pos
        m = _LABEL.match(text, pos)
        if m:
            nodes[nid]["name"] = m.group(0).strip()
            pos = m.end()
        if pos < len(text) and text[pos] == ":":
            m = _LABEL.match(text, pos + 1)
            if not m:
                raise ValueError(f"newick: branch length expected at {pos}")
            nodes[nid]["edge"] = float(m.group(0))
            pos = m.end()

    stack: List[int] = []
    nid = new(None)
    while True:
        if pos < len(text) and text[pos] == "(":      # the node has children: open the first one
            pos += 1
            stack.append(nid)
            nid = new(nid)
            continue
        label(nid)
        while True:                                   # a subtree has just been closed
            if not stack:
                if pos != len(text):
                    raise ValueError(f"newick: trailing text at {pos}")
                return nodes
            if pos >= len(text):
                raise ValueError("newick: unbalanced parentheses")
            if text[pos] == ",":
                pos += 1
                nid = new(stack[-1])
                break
            if text[pos] == ")":
                pos += 1
                nid = stack.pop()
                label(nid)
                continue
            raise ValueError(f"newick: unexpected {text[pos]!r} at {pos}")


def tree_from_newick(newick: str, file_name: str, min_branch_support: float) -> Tree:
    """``Tree::init_from_file`` (tree.rs:164-246): clades from the newick nodes, internal nodes whose support is
    below the threshold collapsed into their parent (``sanitize``, :248-285), parent ids re-written, tree id =
    UUIDv3(DNS namespace, file name)."""
    nodes = parse_newick(newick)
    clades: List[Optional[Clade]] = [None] * len(nodes)
    for i in range(len(nodes) - 1, -1, -1):           # children have larger ids: build bottom-up (:292-364)
        n = nodes[i]
        if i == 0:
            clades[i] = Clade(id=0, parent=None, kind="ROOT", length=0.0, children=[clades[c] for c in n["children"]])
        elif not n["children"]:
            clades[i] = Clade(id=i, parent=n["parent"], kind="LEAF", name=n["name"] or "Unnamed", length=n["edge"])
        else:
            try:
                support = float(n["name"]) if n["name"] is not None else None
            except ValueError:
                support = None
            clades[i] = Clade(id=i, parent=n["parent"], kind="NODE", support=support, length=n["edge"],
                              children=[clades[c] for c in n["children"]])
    # sanitize bottom-up: a child with support below the threshold hands its (already sanitized) children over
    for i in range(len(nodes) - 1, -1, -1):
        cl = clades[i]
        if cl.children is None:
            continue
        kids: List[Clade] = []
        for ch in cl.children:
            if ch.support is not None and not (ch.support >= min_branch_support or ch.is_leaf()):
                kids.extend(ch.children or [])
            else:
                kids.append(ch)
        cl.children = kids or None
    root = clades[0]
    stack: List[Tuple[Clade, Optional[int]]] = [(root, None)]
    while stack:                                      # fix_parent_ids (:229-246)
        cl, parent = stack.pop()
        cl.parent = parent
        for ch in cl.children or []:
            stack.append((ch, cl.id))
    return Tree(id=str(uuid.uuid3(uuid.NAMESPACE_DNS, file_name)), name=file_name,
                min_branch_support=float(min_branch_support), root=root)


def build_loop_records(text: str) -> List[Tuple[str, str]]:
    """The (header, sequence) pairs the reference's MSA loop indexes, in order (build_database/mod.rs:84-117): empty
    lines are skipped; every line starting with '>' SENDS the new header (all '>' removed) together with the sequence
    accumulated so far - i.e. the sequence of the PREVIOUS record (the first header gets whatever preceded it,
    normally nothing) - and clears it; other lines are filtered to upper-case A/C/G/T and appended.  The sequence
    after the last header is never sent."""
    out: List[Tuple[str, str]] = []
    parts: List[str] = []
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    for n, line in enumerate(lines):
        if line.endswith("\r") and n < len(text.split("\n")) - 1:   # BufRead::lines: "\r\n" is a terminator, a bare "\r" is text
            line = line[:-1]
        if line == "":
            continue
        if line.startswith(">"):
            out.append((line.replace(">", ""), "".join(parts)))
            parts = []
        else:
            parts.append(filter_sequence(line))
    return out


def map_kmers_to_tree(tree_path: Union[str, os.PathLike], msa_path: Union[str, os.PathLike],
                      k_size: Optional[int] = None, m_size: Optional[int] = None,
                      min_branch_support: Optional[float] = None, device: Optional[int] = None,
                      pairing: str = "reference") -> Tree:
    """Same arguments and defaults as the reference (build_database/mod.rs:26-44: k = 35, m = 4, support >= 70).
    ``device``: build the k-mer map on that GPU (``cls_model_build_device``) instead of the host builder; the
    result is the same map.

    ``pairing`` decides which sequence is indexed under which tip - the library calls (``cls_model_build*``) take
    (tip, sequence) pairs and index exactly those, so the decision is made here, explicitly:

    ``"reference"`` (default: a database built here equals the one the reference builds from the same files)
        the reference's MSA loop bit for bit (:func:`build_loop_records`): header i is indexed with the sequence of
        record i - 1, the first header gets no k-mers and the last sequence is dropped (build_database/mod.rs:93-116;
        pinned by the reference's own build output, tests/golden/reference_built_model_k12.json.gz).  A header that
        names no tip of the tree raises, as the reference panics (mod.rs:136-139).
    ``"own"``
        every tip with its own sequence - what the reference presumably meant; sequences whose header names no tip
        are ignored, tips without a sequence get no k-mers.  Placements against such a database differ from
        placements against a reference-built one."""
    if pairing not in ("own", "reference"):
        raise ValueError("pairing must be 'own' or 'reference'")
    k_size = 35 if k_size is None else int(k_size)
    m_size = 4 if m_size is None else int(m_size)
    min_branch_support = 70.0 if min_branch_support is None else float(min_branch_support)
    if not os.path.exists(tree_path):
        raise FileNotFoundError("The tree file does not exist.")
    if not os.path.exists(msa_path):
        raise FileNotFoundError("The MSA file does not exist.")
    with open(tree_path, "r", encoding="utf-8") as f:
        tree = tree_from_newick(f.read(), os.path.basename(os.fspath(tree_path)), min_branch_support)
    clades, node_id, node_kind, child_off, child_idx = FlatModel.tree_arrays(tree.root)
    tip_index: Dict[str, int] = {}
    for i, cl in enumerate(clades):
        if cl.is_leaf() and cl.name is not None:
            tip_index.setdefault(cl.name, i)
    tip_node, seqs = [], []
    if pairing == "reference":
        with open(msa_path, "r", encoding="utf-8", newline="") as f:
            records = build_loop_records(f.read())
        for header, seq in records:
            i = tip_index.get(header)
            if i is None:
                raise ValueError(f"The sequence header does not match any tree leaf: {header}")
            if seq:
                tip_node.append(i)
                seqs.append(seq.encode("ascii"))
    else:
        for header, seq in read_fasta(msa_path):
            i = tip_index.get(header)
            if i is not None and seq:
                tip_node.append(i)
                seqs.append(seq.encode("ascii"))
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offsets[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(1, np.uint8)
    tree_only = FlatModel(k_size, m_size, node_id, node_kind, child_off, child_idx)
    bm = BuiltModel(tree_only, np.array(tip_node, dtype=np.uint64), bases, offsets, device=device)
    a = bm.arrays()
    bm.close()
    km = KmersMap(k_size, m_size)
    so, sn = a["set_off"], a["set_node_ids"]
    sets = [set(sn[int(so[s]):int(so[s + 1])].tolist()) for s in range(len(so) - 1)]
    for b, h, s in zip(a["entry_bucket"].tolist(), a["entry_hash"].tolist(), a["entry_set"].tolist()):
        km.map.setdefault(b, {})[h] = sets[s]
    tree.kmers_map = km
    return tree
