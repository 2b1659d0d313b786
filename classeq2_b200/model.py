"""Host-side mirror of the reference's model types - the API surface of the placement path.

Same names, fields and serde shape as the reference (paths relative to its checkout):

* :class:`Clade`     - ``core/src/domain/dtos/clade.rs:18-38`` (camelCase; ``kind`` is
  ``ROOT|NODE|LEAF``; ``is_leaf`` is by kind, ``:166-172``)
* :class:`KmersMap`  - ``core/src/domain/dtos/kmers_map.rs:77-87`` (``kSize``, ``mSize``,
  ``map``: bucket key -> k-mer hash -> node ids)
* :class:`Tree`      - ``core/src/domain/dtos/tree.rs:9-52``

plus :class:`FlatModel`, the flat arrays behind a ``cls_model_view`` (include/classeq_b200.h).
Nothing here computes on the hot path; it only marshals the model to the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Set

import numpy as np

from . import _lib

_KIND = {"ROOT": _lib.KIND_ROOT, "NODE": _lib.KIND_NODE, "LEAF": _lib.KIND_LEAF}


@dataclass
class Clade:
    id: int
    parent: Optional[int]
    kind: str
    name: Optional[str] = None
    support: Optional[float] = None
    length: Optional[float] = None
    children: Optional[List["Clade"]] = None

    def is_leaf(self) -> bool:
        return self.kind == "LEAF"

    def walk(self) -> Iterable["Clade"]:
        stack = [self]
        while stack:
            c = stack.pop()
            yield c
            if c.children:
                stack.extend(reversed(c.children))

    def get_node_by_id(self, id_: int) -> Optional["Clade"]:  # clade.rs:95-109 (first pre-order match)
        for c in self.walk():
            if c.id == id_:
                return c
        return None

    def to_obj(self) -> dict:
        o: dict = {"id": self.id, "parent": self.parent, "kind": self.kind}
        if self.name is not None:
            o["name"] = self.name
        if self.support is not None:
            o["support"] = self.support
        if self.length is not None:
            o["length"] = self.length
        if self.children is not None:
            o["children"] = [c.to_obj() for c in self.children]
        return o

    @staticmethod
    def from_obj(o: dict) -> "Clade":
        ch = o.get("children")
        # support / length are Option<f64> (clade.rs:27-34); YAML 1.1 loaders hand "1e-6" over as a string
        f64 = lambda v: None if v is None else float(v)  # noqa: E731
        return Clade(id=int(o["id"]), parent=None if o.get("parent") is None else int(o["parent"]),
                     kind=str(o["kind"]), name=None if o.get("name") is None else str(o["name"]),
                     support=f64(o.get("support")), length=f64(o.get("length")),
                     children=None if ch is None else [Clade.from_obj(c) for c in ch])


@dataclass
class KmersMap:
    k_size: int
    m_size: int
    map: Dict[int, Dict[int, Set[int]]] = field(default_factory=dict)

    def get_kmer_size(self) -> int:  # kmers_map.rs:111-113
        return self.k_size

    def get_minimizer_size(self) -> int:  # kmers_map.rs:115-117
        return self.m_size

    def get_map(self):  # kmers_map.rs:107-109
        return self.map

    def to_obj(self) -> dict:
        return {"kSize": self.k_size, "mSize": self.m_size,
                "map": {k: {h: sorted(n) for h, n in v.items()} for k, v in self.map.items()}}

    @staticmethod
    def from_obj(o: dict) -> "KmersMap":
        km = KmersMap(int(o["kSize"]), int(o["mSize"]))
        for k, v in o["map"].items():
            km.map[int(k)] = {int(h): {int(x) for x in nodes} for h, nodes in v.items()}
        return km


@dataclass
class Tree:
    id: str
    name: str
    min_branch_support: float
    root: Clade
    annotations: Optional[list] = None
    kmers_map: Optional[KmersMap] = None
    in_memory_size: Optional[str] = None

    def to_obj(self) -> dict:
        o = {"id": self.id, "name": self.name, "minBranchSupport": self.min_branch_support,
             "inMemorySize": self.in_memory_size, "root": self.root.to_obj()}
        if self.annotations is not None:
            o["annotations"] = self.annotations
        o["kmersMap"] = None if self.kmers_map is None else self.kmers_map.to_obj()
        return o

    @staticmethod
    def from_obj(o: dict) -> "Tree":
        km = o.get("kmersMap")
        return Tree(id=str(o["id"]), name=str(o["name"]), min_branch_support=float(o["minBranchSupport"]),
                    root=Clade.from_obj(o["root"]), annotations=o.get("annotations"),
                    kmers_map=None if km is None else KmersMap.from_obj(km),
                    in_memory_size=o.get("inMemorySize"))


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


class FlatModel:
    """Flat arrays of a model + the borrowed ``cls_model_view`` over them.

    Keep the object alive for as long as the view is used (the view borrows the arrays).
    """

    def __init__(self, k_size: int, m_size: int, node_id, node_kind, child_off, child_idx,
                 entry_bucket=None, entry_hash=None, entry_set=None, set_off=None, set_node_ids=None,
                 root_children_none: bool = False, keepalive=None, force_general_sets: bool = False):
        u64 = lambda a: np.ascontiguousarray(a if a is not None else [], dtype=np.uint64)  # noqa: E731
        self.k_size, self.m_size = int(k_size), int(m_size)
        self.node_id = u64(node_id)
        self.node_kind = np.ascontiguousarray(node_kind, dtype=np.uint8)
        self.child_off = u64(child_off)
        self.child_idx = u64(child_idx)
        self.entry_bucket, self.entry_hash, self.entry_set = u64(entry_bucket), u64(entry_hash), u64(entry_set)
        self.set_off = u64(set_off if set_off is not None else [0])
        self.set_node_ids = u64(set_node_ids)
        self.root_children_none = bool(root_children_none)
        self.force_general_sets = bool(force_general_sets)
        self._keepalive = keepalive
        self.view = self._make_view()

    def _make_view(self) -> _lib.ModelView:
        v = _lib.ModelView()
        v.k_size, v.m_size = self.k_size, self.m_size
        v.flags = ((_lib.MODEL_ROOT_CHILDREN_NONE if self.root_children_none else 0)
                   | (_lib.MODEL_FORCE_GENERAL_SETS if self.force_general_sets else 0))
        v.n_nodes = len(self.node_id)
        v.node_id = _ptr(self.node_id, _lib.u64p)
        v.node_kind = _ptr(self.node_kind, _lib.u8p)
        v.child_off = _ptr(self.child_off, _lib.u64p)
        v.child_idx = _ptr(self.child_idx, _lib.u64p)
        v.n_entries = len(self.entry_hash)
        v.entry_bucket = _ptr(self.entry_bucket, _lib.u64p)
        v.entry_hash = _ptr(self.entry_hash, _lib.u64p)
        v.entry_set = _ptr(self.entry_set, _lib.u64p)
        v.n_sets = len(self.set_off) - 1
        v.set_off = _ptr(self.set_off, _lib.u64p)
        v.set_node_ids = _ptr(self.set_node_ids, _lib.u64p)
        return v

    @property
    def n_entries(self) -> int:
        return len(self.entry_hash)

    # ---- flat binary cache: upload without parsing the YAML again ------------------
    _ARRAYS = ("node_id", "node_kind", "child_off", "child_idx", "entry_bucket", "entry_hash", "entry_set", "set_off",
               "set_node_ids")

    def save(self, path) -> None:
        """The arrays of the ``cls_model_view`` as one uncompressed ``.npz`` (the YAML of a 10 k-tip model is GBs;
        these arrays are what ``cls_index_create`` consumes directly)."""
        np.savez(path, k_size=np.uint32(self.k_size), m_size=np.uint32(self.m_size),
                 root_children_none=np.uint8(self.root_children_none), force_general_sets=np.uint8(self.force_general_sets),
                 **{n: getattr(self, n) for n in self._ARRAYS})

    @staticmethod
    def load(path) -> "FlatModel":
        z = np.load(path)
        return FlatModel(int(z["k_size"]), int(z["m_size"]), *[z[n] for n in FlatModel._ARRAYS],
                         root_children_none=bool(z["root_children_none"]), force_general_sets=bool(z["force_general_sets"]))

    def with_general_sets(self) -> "FlatModel":
        """The same model, forced onto the general mini-tree node-set records (testing knob)."""
        return FlatModel(self.k_size, self.m_size, self.node_id, self.node_kind, self.child_off, self.child_idx,
                         self.entry_bucket, self.entry_hash, self.entry_set, self.set_off, self.set_node_ids,
                         self.root_children_none, self._keepalive, force_general_sets=True)

    # ---- constructors -------------------------------------------------------------------------
    @staticmethod
    def tree_arrays(root: Clade):
        """Pre-order flattening of a Clade tree: (clades, node_id, node_kind, child_off, child_idx)."""
        clades = list(root.walk())
        index = {id(c): i for i, c in enumerate(clades)}
        node_id = np.array([c.id for c in clades], dtype=np.uint64)
        node_kind = np.array([_KIND[c.kind] for c in clades], dtype=np.uint8)
        child_off = np.zeros(len(clades) + 1, dtype=np.uint64)
        child_idx: List[int] = []
        for i, c in enumerate(clades):
            for ch in c.children or []:
                child_idx.append(index[id(ch)])
            child_off[i + 1] = len(child_idx)
        return clades, node_id, node_kind, child_off, np.array(child_idx, dtype=np.uint64)

    @staticmethod
    def from_tree(tree: Tree) -> "FlatModel":
        """Flatten a :class:`Tree` (with its ``kmers_map``) - what a Rust caller does right after
        ``load_database`` (ports/cli/src/cmds/place_sequences.rs:135)."""
        if tree.kmers_map is None:
            raise ValueError("The tree does not have a kmers map.")  # place_sequence.rs:77-80
        _, node_id, node_kind, child_off, child_idx = FlatModel.tree_arrays(tree.root)
        km = tree.kmers_map
        buckets, hashes, sets = [], [], []
        set_index: Dict[frozenset, int] = {}
        set_off = [0]
        set_nodes: List[int] = []
        for key, value in km.map.items():
            for h, nodes in value.items():
                fs = frozenset(nodes)
                s = set_index.get(fs)
                if s is None:
                    s = len(set_index)
                    set_index[fs] = s
                    set_nodes.extend(sorted(fs))
                    set_off.append(len(set_nodes))
                buckets.append(key)
                hashes.append(h)
                sets.append(s)
        return FlatModel(km.k_size, km.m_size, node_id, node_kind, child_off, child_idx,
                         buckets, hashes, sets, set_off, set_nodes,
                         root_children_none=tree.root.children is None)


class BuiltModel:
    """k-mer map for one sequence per tip: ``cls_model_build`` (host side, no GPU) or, with ``device`` given,
    ``cls_model_build_device`` (the same map built on that GPU; no CPU fallback)."""

    def __init__(self, tree_only: FlatModel, tip_node: np.ndarray, bases: np.ndarray, offsets: np.ndarray,
                 device: Optional[int] = None):
        self.tree_only = tree_only
        tip_node = np.ascontiguousarray(tip_node, dtype=np.uint64)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._h = C.c_void_p()
        if device is None:
            _lib.check(_lib.lib.cls_model_build(C.byref(tree_only.view), len(tip_node), _ptr(tip_node, _lib.u64p),
                                                _ptr(bases, _lib.u8p), _ptr(offsets, _lib.u64p), C.byref(self._h)))
        else:
            _lib.check(_lib.lib.cls_model_build_device(C.byref(tree_only.view), len(tip_node), _ptr(tip_node, _lib.u64p),
                                                       _ptr(bases, _lib.u8p), _ptr(offsets, _lib.u64p), int(device),
                                                       C.byref(self._h)))
        self.view = _lib.ModelView()
        _lib.check(_lib.lib.cls_built_model_view(self._h, C.byref(tree_only.view), C.byref(self.view)))

    def arrays(self) -> dict:
        """Copies of the entry/set arrays (numpy, uint64)."""
        v = self.view
        n, s = int(v.n_entries), int(v.n_sets)
        cp = lambda p, k: np.ctypeslib.as_array(p, shape=(k,)).copy() if k else np.zeros(0, np.uint64)  # noqa: E731
        set_off = cp(v.set_off, s + 1)
        return {"entry_bucket": cp(v.entry_bucket, n), "entry_hash": cp(v.entry_hash, n),
                "entry_set": cp(v.entry_set, n), "set_off": set_off,
                "set_node_ids": cp(v.set_node_ids, int(set_off[-1]))}

    def close(self):
        if self._h:
            _lib.lib.cls_built_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
