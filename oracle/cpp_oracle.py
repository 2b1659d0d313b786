"""ctypes binding of oracle/libclasseq_oracle.so (the C++ restatement).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libclasseq_oracle.so")

u8p, u32p, i32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_int32, C.c_uint64))


class ModelView(C.Structure):  # field order of orc_model_view
    _fields_ = [("k_size", C.c_uint32), ("m_size", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("n_nodes", C.c_uint64), ("node_id", u64p), ("node_kind", u8p), ("child_off", u64p), ("child_idx", u64p),
                ("n_entries", C.c_uint64), ("entry_bucket", u64p), ("entry_hash", u64p), ("entry_set", u64p),
                ("n_sets", C.c_uint64), ("set_off", u64p), ("set_node_ids", u64p)]


def _load():
    if not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", HERE], check=True)
    lib = C.CDLL(LIB_PATH)
    lib.orc_model_create.restype = C.c_void_p
    lib.orc_model_create.argtypes = [C.POINTER(ModelView)]
    lib.orc_model_destroy.argtypes = [C.c_void_p]
    lib.orc_place_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_int32, C.c_double, C.c_uint32, C.c_int,
                                    u8p, u64p, i32p, i32p, u32p, u32p, u32p, u32p]
    lib.orc_kmer_hashes.restype = C.c_uint64
    lib.orc_kmer_hashes.argtypes = [u8p, C.c_uint64, C.c_uint32, u64p, C.c_uint64]
    lib.orc_murmur3_x64_128.argtypes = [u8p, C.c_uint64, C.c_uint64, u64p]
    lib.orc_model_build.restype = C.c_void_p
    lib.orc_model_build.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, u64p, u64p, u64p, C.c_uint64, u64p, u8p, u64p, C.c_int]
    lib.orc_built_sizes.argtypes = [C.c_void_p, u64p, u64p, u64p]
    lib.orc_built_copy.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p, u64p]
    lib.orc_built_destroy.argtypes = [C.c_void_p]
    return lib


lib = _load()
FIELDS = (("status", np.uint8), ("node_id", np.uint64), ("one", np.int32), ("rest", np.int32),
          ("n_query_kmers", np.uint32), ("n_matched", np.uint32), ("n_root_matched", np.uint32), ("iterations", np.uint32))


def _p(a, t):
    return a.ctypes.data_as(t)


class CppModel:
    """Model held by the C++ oracle, built from flat arrays (cls_model_view layout)."""

    def __init__(self, k_size, m_size, node_id, node_kind, child_off, child_idx, entry_bucket, entry_hash, entry_set,
                 set_off, set_node_ids, root_children_none=False):
        u64 = lambda a: np.ascontiguousarray(a, dtype=np.uint64)  # noqa: E731
        self._a = [u64(node_id), np.ascontiguousarray(node_kind, np.uint8), u64(child_off), u64(child_idx),
                   u64(entry_bucket), u64(entry_hash), u64(entry_set), u64(set_off), u64(set_node_ids)]
        a = self._a
        v = ModelView()
        v.k_size, v.m_size, v.flags = int(k_size), int(m_size), 1 if root_children_none else 0
        v.n_nodes = len(a[0])
        v.node_id, v.node_kind, v.child_off, v.child_idx = _p(a[0], u64p), _p(a[1], u8p), _p(a[2], u64p), _p(a[3], u64p)
        v.n_entries = len(a[5])
        v.entry_bucket, v.entry_hash, v.entry_set = _p(a[4], u64p), _p(a[5], u64p), _p(a[6], u64p)
        v.n_sets = len(a[7]) - 1
        v.set_off, v.set_node_ids = _p(a[7], u64p), _p(a[8], u64p)
        self._h = C.c_void_p(lib.orc_model_create(C.byref(v)))

    @staticmethod
    def from_flat(flat) -> "CppModel":
        """From any object with the flat-array attributes (e.g. classeq2_b200.FlatModel)."""
        return CppModel(flat.k_size, flat.m_size, flat.node_id, flat.node_kind, flat.child_off, flat.child_idx,
                        flat.entry_bucket, flat.entry_hash, flat.entry_set, flat.set_off, flat.set_node_ids,
                        getattr(flat, "root_children_none", False))

    def place_batch(self, bases, offsets, max_iterations=None, min_match_coverage=None, remove_intersection=None,
                    n_threads=None):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        out = {name: np.zeros(n, dt) for name, dt in FIELDS}
        if bases.size == 0:
            bases = np.zeros(1, np.uint8)
        lib.orc_place_batch(self._h, _p(bases, u8p), _p(offsets, u64p), n,
                            1000 if max_iterations is None else int(max_iterations),
                            0.7 if min_match_coverage is None else float(min_match_coverage),
                            1 if remove_intersection else 0, int(n_threads or os.cpu_count() or 1),
                            _p(out["status"], u8p), _p(out["node_id"], u64p), _p(out["one"], i32p), _p(out["rest"], i32p),
                            _p(out["n_query_kmers"], u32p), _p(out["n_matched"], u32p), _p(out["n_root_matched"], u32p),
                            _p(out["iterations"], u32p))
        return out

    def close(self):
        if self._h:
            lib.orc_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleFlat:
    """Flat arrays of a model (the attributes CppModel.from_flat reads)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.n_entries = len(self.entry_hash)


def build_model(k_size, m_size, node_id, node_kind, child_off, child_idx, tip_node, bases, offsets, n_threads=None) -> OracleFlat:
    """The k-mer map of a tree whose tip t carries the sequence bases[offsets[t]:offsets[t+1]] (both strands, node set =
    ids on the root -> tip path), built by the oracle's own builder (build_database/mod.rs:62, :140-168)."""
    u64 = lambda a: np.ascontiguousarray(a, dtype=np.uint64)  # noqa: E731
    node_id, child_off, child_idx, tip_node, offsets = u64(node_id), u64(child_off), u64(child_idx), u64(tip_node), u64(offsets)
    bases = np.ascontiguousarray(bases, np.uint8)
    h = lib.orc_model_build(int(k_size), int(m_size), len(node_id), _p(node_id, u64p), _p(child_off, u64p), _p(child_idx, u64p),
                            len(tip_node), _p(tip_node, u64p), _p(bases, u8p), _p(offsets, u64p), int(n_threads or os.cpu_count() or 1))
    if not h:
        raise ValueError("a tip sequence holds a character other than A, C, G, T")
    h = C.c_void_p(h)
    n = (C.c_uint64 * 3)()
    lib.orc_built_sizes(h, C.cast(C.byref(n, 0), u64p), C.cast(C.byref(n, 8), u64p), C.cast(C.byref(n, 16), u64p))
    eb, eh, es = (np.empty(n[0], np.uint64) for _ in range(3))
    so, ids = np.empty(n[1] + 1, np.uint64), np.empty(max(1, n[2]), np.uint64)
    lib.orc_built_copy(h, _p(eb, u64p), _p(eh, u64p), _p(es, u64p), _p(so, u64p), _p(ids, u64p))
    lib.orc_built_destroy(h)
    return OracleFlat(k_size=int(k_size), m_size=int(m_size), node_id=node_id, node_kind=np.ascontiguousarray(node_kind, np.uint8),
                      child_off=child_off, child_idx=child_idx, entry_bucket=eb, entry_hash=eh, entry_set=es, set_off=so,
                      set_node_ids=ids[: n[2]])


def kmer_hashes(seq: bytes, k: int) -> np.ndarray:
    a = np.frombuffer(seq, np.uint8).copy() if seq else np.zeros(1, np.uint8)
    cap = max(1, 2 * (len(seq) - k + 1))
    out = np.zeros(cap, np.uint64)
    n = lib.orc_kmer_hashes(_p(a, u8p), len(seq), k, _p(out, u64p), cap)
    return out[:n]


def murmur3_x64_128(data: bytes, seed: int = 0):
    a = np.frombuffer(data, np.uint8).copy() if data else np.zeros(1, np.uint8)
    out = np.zeros(2, np.uint64)
    lib.orc_murmur3_x64_128(_p(a, u8p), len(data), seed, _p(out, u64p))
    return int(out[0]), int(out[1])
