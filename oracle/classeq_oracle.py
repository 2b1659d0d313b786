"""CPU oracle (pure Python, sets of ints) for classeq's placement hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``classeq2_b200`` + the CUDA
library) never does.

It restates, function by function, the reference's Rust code (paths relative to
the reference checkout, ``core/src/...``):

* ``murmurhash3_x64_128``  - third-party crate ``mur3 0.1.0`` (Cargo.lock:2405-2408),
  not vendored in the reference tree; this is the published MurmurHash3_x64_128
  algorithm (A. Appleby, public domain), pinned by the two bucket keys printed in
  ``docs/book/02-build-db.md:181,192`` and ``h1("") == 0``.
* ``KmersMap``             - ``domain/dtos/kmers_map.rs``
* ``Clade`` / ``Tree``     - ``domain/dtos/clade.rs``, ``domain/dtos/tree.rs``
* ``read_fasta``           - ``domain/dtos/file_or_stdin.rs:76-116`` +
  ``domain/dtos/sequence.rs:47-56``
* ``place_sequence``       - ``use_cases/place_sequences/place_sequence.rs:42-602`` +
  ``update_introspection_node.rs:13-91``
* ``placement_response``   - ``use_cases/place_sequences/mod.rs:170-239`` +
  ``domain/dtos/placement_response.rs:30-59``
* ``map_kmers_to_tree`` / ``tree_from_newick`` - model-build side
  (``use_cases/build_database/mod.rs:26-181``, ``domain/dtos/tree.rs:164-364``),
  needed only to manufacture models for the fixtures.

PARITY STATUS: the reference is Rust and cannot be compiled in this environment
(no cargo/rustc, 465 un-vendored crates).  Reference-pinned known answers exist
for the hash arithmetic, the windowing, the "both strands / all windows" count
(``one: 3754``), the result wire shape, and - through the one model the reference
itself wrote (``core/src/tests/data/.../outputs/Colletotrichum_acutatum_gapdh-PhyML.yaml``,
committed compactly as ``tests/golden/reference_built_model_k12.json.gz``) - the whole
tree of ``Tree::init_from_file`` and the MODEL CONTENT: every k-mer's node set (union
of root -> tip id paths) and the header / sequence pairing of the build loop, all 2 158
k-mers reproduced exactly (``tests/test_oracle_kats.py``).  **Placement decisions
themselves are "parity unpinned"**: the only placement goldens in the reference
depend on a missing Git-LFS model blob.  They are pinned here by agreement of two
independent restatements (this file and ``oracle/classeq_oracle.cpp``).
"""
from __future__ import annotations

import re
import uuid
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Set, Tuple

MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------
# mur3::murmurhash3_x64_128 (kmers_map.rs:1,157-159)
# --------------------------------------------------------------------------------------
def _rotl64(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & MASK64


def _fmix64(k: int) -> int:
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & MASK64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & MASK64
    k ^= k >> 33
    return k


def murmurhash3_x64_128(data: bytes, seed: int = 0) -> Tuple[int, int]:
    c1 = 0x87C37B91114253D5
    c2 = 0x4CF5AD432745937F
    h1 = seed & MASK64
    h2 = seed & MASK64
    n = len(data)
    nblocks = n // 16
    for i in range(nblocks):
        k1 = int.from_bytes(data[16 * i: 16 * i + 8], "little")
        k2 = int.from_bytes(data[16 * i + 8: 16 * i + 16], "little")
        k1 = (k1 * c1) & MASK64
        k1 = _rotl64(k1, 31)
        k1 = (k1 * c2) & MASK64
        h1 ^= k1
        h1 = _rotl64(h1, 27)
        h1 = (h1 + h2) & MASK64
        h1 = (h1 * 5 + 0x52DCE729) & MASK64
        k2 = (k2 * c2) & MASK64
        k2 = _rotl64(k2, 33)
        k2 = (k2 * c1) & MASK64
        h2 ^= k2
        h2 = _rotl64(h2, 31)
        h2 = (h2 + h1) & MASK64
        h2 = (h2 * 5 + 0x38495AB5) & MASK64
    tail = data[16 * nblocks:]
    k1 = 0
    k2 = 0
    t = len(tail)
    if t > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * c2) & MASK64
        k2 = _rotl64(k2, 33)
        k2 = (k2 * c1) & MASK64
        h2 ^= k2
    if t > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * c1) & MASK64
        k1 = _rotl64(k1, 31)
        k1 = (k1 * c2) & MASK64
        h1 ^= k1
    h1 ^= n
    h2 ^= n
    h1 = (h1 + h2) & MASK64
    h2 = (h2 + h1) & MASK64
    h1 = _fmix64(h1)
    h2 = _fmix64(h2)
    h1 = (h1 + h2) & MASK64
    h2 = (h2 + h1) & MASK64
    return h1, h2


def hash_kmer(kmer: str) -> int:
    """kmers_map.rs:157-159 - ``murmurhash3_x64_128(kmer.as_bytes(), 0).0``."""
    return murmurhash3_x64_128(kmer.encode("utf-8"), 0)[0]


# --------------------------------------------------------------------------------------
# FASTA reading (file_or_stdin.rs:76-116, sequence.rs:47-56)
# --------------------------------------------------------------------------------------
def remove_non_iupac_from_sequence(sequence: str) -> str:
    """sequence.rs:47-56 - upper-case, then keep only A, C, G, T."""
    return "".join(c for c in sequence.upper() if c in "ACGT")


def read_fasta_text(text: str) -> List[Tuple[str, str]]:
    """file_or_stdin.rs:76-116.  Returns the (header, body) records that the
    reference sends down the channel, in order.  ``BufRead::lines`` splits on
    ``\\n`` and strips one trailing ``\\r``.  A sequence seen before any header
    aborts the read (``StdinError``), which the caller ignores
    (place_sequences/mod.rs:119) - records sent so far stand."""
    out: List[Tuple[str, str]] = []
    header = ""
    sequence = ""
    lines = text.split("\n")
    terminated = [True] * len(lines)
    terminated[-1] = False  # BufRead::lines pops "\n" and then one "\r": a "\r" at end of file stays
    if lines and lines[-1] == "":
        lines.pop()
    for line, term in zip(lines, terminated):
        if term and line.endswith("\r"):
            line = line[:-1]
        if line == "":
            continue
        if line.startswith(">"):
            if header != "":
                out.append((header, sequence))
                sequence = ""
            elif sequence != "":
                return out  # Err("unexpected sequence without header")
            header = line.replace(">", "")
        else:
            sequence += remove_non_iupac_from_sequence(line)
    if header != "" and sequence != "":
        out.append((header, sequence))
    return out


# --------------------------------------------------------------------------------------
# Clade / Tree (clade.rs:5-38, tree.rs:9-52)
# --------------------------------------------------------------------------------------
@dataclass
class Clade:
    id: int
    parent: Optional[int]
    kind: str  # "ROOT" | "NODE" | "LEAF"   (clade.rs:5-16, serde UPPERCASE)
    name: Optional[str] = None
    support: Optional[float] = None
    length: Optional[float] = None
    children: Optional[List["Clade"]] = None

    def is_leaf(self) -> bool:  # clade.rs:166-172 - by kind, not by children
        return self.kind == "LEAF"

    def get_node_by_id(self, id_: int) -> Optional["Clade"]:  # clade.rs:95-109
        if self.id == id_:
            return self
        for c in self.children or []:
            n = c.get_node_by_id(id_)
            if n is not None:
                return n
        return None

    def get_path_to_root(self, root: "Clade") -> Set[int]:  # clade.rs:111-125
        path = {self.id}
        if self.parent is not None:
            path.add(self.parent)
            p = root.get_node_by_id(self.parent)
            if p is not None:
                path |= p.get_path_to_root(root)
        return path

    def get_leaves_with_paths(self, parent_ids=None):  # clade.rs:127-156
        ids = [self.id] if parent_ids is None else list(parent_ids) + [self.id]
        if self.is_leaf():
            return [(self, ids)]
        out = []
        for c in self.children or []:
            out.extend(c.get_leaves_with_paths(ids))
        return out

    def to_obj(self) -> dict:
        """serde shape: camelCase, ``None`` fields skipped except ``parent``
        (clade.rs:18-38: ``parent`` has no skip attribute -> serialised as null)."""
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
        return Clade(
            id=int(o["id"]),
            parent=None if o.get("parent") is None else int(o["parent"]),
            kind=o["kind"],
            name=o.get("name"),
            support=o.get("support"),
            length=o.get("length"),
            children=None if ch is None else [Clade.from_obj(c) for c in ch],
        )

    def walk(self) -> Iterable["Clade"]:
        yield self
        for c in self.children or []:
            yield from c.walk()


@dataclass
class KmersMap:
    """kmers_map.rs:77-87: ``map`` is bucket key -> (k-mer hash -> node-id set)."""
    k_size: int
    m_size: int
    map: Dict[int, Dict[int, Set[int]]] = field(default_factory=dict)

    # kmers_map.rs:10-13
    def _minimizer_key(self, kmer: str) -> int:
        return hash_kmer(kmer[: self.m_size])

    # kmers_map.rs:125-149 + :23-35
    def insert_or_append_kmer_hash(self, kmer: str, h: int, nodes: Set[int]) -> None:
        key = 0 if self.m_size == 0 else self._minimizer_key(kmer)
        bucket = self.map.setdefault(key, {})
        if h in bucket:
            bucket[h] |= nodes
        else:
            bucket[h] = set(nodes)

    # kmers_map.rs:405-424
    @staticmethod
    def build_kmers_from_sequence(sequence: str, size: int) -> List[Tuple[str, int]]:
        s = sequence.upper()
        return [(s[i: i + size], hash_kmer(s[i: i + size])) for i in range(len(s) - size + 1)]

    # kmers_map.rs:431-443
    @staticmethod
    def reverse_complement(sequence: str) -> str:
        comp = {"a": "T", "A": "T", "t": "A", "T": "A", "c": "G", "C": "G", "g": "C", "G": "C"}
        return "".join(comp[c] for c in reversed(sequence))  # KeyError == reference panic

    # kmers_map.rs:375-398
    def build_kmer_from_string(self, sequence: str) -> List[Tuple[str, int]]:
        if len(sequence) < self.k_size:
            return []
        return (KmersMap.build_kmers_from_sequence(sequence, self.k_size)
                + KmersMap.build_kmers_from_sequence(KmersMap.reverse_complement(sequence), self.k_size))

    # kmers_map.rs:273-311 (+ :55-70)
    def get_overlapping_hashed_kmers(self, hashed_kmers: List[Tuple[str, int]]) -> "KmersMap":
        minimizers = {self._minimizer_key(k) for k, _ in hashed_kmers}
        hashes = {h for _, h in hashed_kmers}
        out = KmersMap(self.k_size, self.m_size)
        for key, value in self.map.items():
            if key not in minimizers:
                continue
            sub = {h: set(value[h]) for h in (set(value.keys()) & hashes)}
            if sub:
                out.map[key] = sub
        return out

    # kmers_map.rs:211-229 (+ :37-53)
    def get_minimized_hashes_with_node(self, node: int) -> Optional[Dict[int, Set[int]]]:
        out = {}
        for key, value in self.map.items():
            s = {h for h, nodes in value.items() if node in nodes}
            if s:
                out[key] = s
        return out or None

    # kmers_map.rs:318-344
    def get_overlapping_minimized_hashes(self, hashed: Dict[int, Set[int]]) -> "KmersMap":
        out = KmersMap(self.k_size, self.m_size)
        for key, value in self.map.items():
            if key in hashed:
                sub = {h: set(value[h]) for h in (set(value.keys()) & hashed[key])}
                if sub:
                    out.map[key] = sub
        return out

    # kmers_map.rs:189-203
    def get_hashed_kmers_with_node(self, node: int) -> Optional[Set[int]]:
        s: Set[int] = set()
        for value in self.map.values():
            s |= {h for h, nodes in value.items() if node in nodes}
        return s or None

    def n_entries(self) -> int:
        return sum(len(v) for v in self.map.values())

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

    def to_obj(self) -> dict:  # tree.rs:9-52 (camelCase)
        o = {"id": self.id, "name": self.name, "minBranchSupport": self.min_branch_support,
             "inMemorySize": self.in_memory_size, "root": self.root.to_obj()}
        if self.annotations is not None:
            o["annotations"] = self.annotations
        o["kmersMap"] = None if self.kmers_map is None else self.kmers_map.to_obj()
        return o

    @staticmethod
    def from_obj(o: dict) -> "Tree":
        km = o.get("kmersMap")
        return Tree(id=str(o["id"]), name=o["name"], min_branch_support=float(o["minBranchSupport"]),
                    root=Clade.from_obj(o["root"]), annotations=o.get("annotations"),
                    kmers_map=None if km is None else KmersMap.from_obj(km),
                    in_memory_size=o.get("inMemorySize"))


# --------------------------------------------------------------------------------------
# Model-build side, used only to manufacture fixtures (tree.rs:164-364,
# build_database/mod.rs:26-181).
# --------------------------------------------------------------------------------------
def _parse_newick(text: str):
    """Minimal newick parser giving phylotree-0.1.2-style nodes numbered in
    PRE-ORDER of creation (root 0, each '(' opens the next id, tips get the next
    id when read) - pinned against the ids in the reference's stale golden
    ``core/src/tests/data/.../outputs/Colletotrichum_acutatum_gapdh-PhyML.yaml:5-60``."""
    text = text.strip()
    assert text.endswith(";")
    text = text[:-1]
    nodes = []  # dict(id, parent, children, name, edge)

    def new_node(parent):
        n = {"id": len(nodes), "parent": parent, "children": [], "name": None, "edge": None}
        nodes.append(n)
        if parent is not None:
            nodes[parent]["children"].append(n["id"])
        return n["id"]

    pos = 0
    tok = re.compile(r"[^(),:;]+")

    def parse_label(nid):
        nonlocal pos
        m = tok.match(text, pos)
        if m:
            nodes[nid]["name"] = m.group(0).strip()
            pos = m.end()
        if pos < len(text) and text[pos] == ":":
            pos += 1
            m = tok.match(text, pos)
            nodes[nid]["edge"] = float(m.group(0))
            pos = m.end()

    def parse_subtree(parent):
        nonlocal pos
        nid = new_node(parent)
        if text[pos] == "(":
            pos += 1
            while True:
                parse_subtree(nid)
                if text[pos] == ",":
                    pos += 1
                    continue
                assert text[pos] == ")", (pos, text[pos: pos + 20])
                pos += 1
                break
        parse_label(nid)
        return nid

    parse_subtree(None)
    return nodes


def tree_from_newick(newick: str, file_name: str, min_branch_support: float) -> Tree:
    """tree.rs:164-364: build clades from the parsed newick, collapse internal
    nodes whose support < min (children re-attached to the grand-parent,
    ``sanitize`` :248-285), then ``fix_parent_ids`` (:229-246)."""
    nodes = _parse_newick(newick)

    def children_of(nid) -> List[Clade]:  # get_children_nodes, :292-364
        out = []
        for cid in nodes[nid]["children"]:
            c = nodes[cid]
            if not c["children"]:
                out.append(Clade(id=cid, parent=nid, kind="LEAF", name=c["name"] or "Unnamed",
                                 length=c["edge"]))
            else:
                try:
                    sup = float(c["name"]) if c["name"] is not None else None
                except ValueError:
                    sup = None
                out.append(Clade(id=cid, parent=nid, kind="NODE", support=sup, length=c["edge"],
                                 children=children_of(cid)))
        return out

    def sanitize(cl: Clade) -> Clade:
        kids: List[Clade] = []
        for ch in cl.children or []:
            s = sanitize(ch)
            if s.support is not None:
                if s.support >= min_branch_support or s.is_leaf():
                    kids.append(s)
                else:
                    kids.extend(s.children or [])
            else:
                kids.append(s)
        cl.children = kids or None
        return cl

    def fix_parent_ids(cl: Clade, parent: Optional[int]) -> None:
        cl.parent = parent
        for ch in cl.children or []:
            fix_parent_ids(ch, cl.id)

    root = Clade(id=0, parent=None, kind="ROOT", length=0.0, children=children_of(0))
    root = sanitize(root)
    fix_parent_ids(root, None)
    tid = str(uuid.uuid3(uuid.NAMESPACE_DNS, file_name))  # tree.rs:213-214
    return Tree(id=tid, name=file_name, min_branch_support=min_branch_support, root=root)


def reference_pairing(records: List[Tuple[str, str]]) -> List[Tuple[str, str]]:
    """The (header, sequence) pairs the reference's MSA loop actually indexes (build_database/mod.rs:93-116): at
    every '>' line it hashes the sequence accumulated SO FAR (the previous record's) and sends it under the NEW
    header, and nothing is flushed after the last line - header i is paired with sequence i-1, the first header
    with the empty string, the last sequence is dropped.  Pinned by the reference's own build output
    (tests/golden/reference_built_model_k12.json.gz, written by the reference from its Colletotrichum inputs)."""
    return [(records[i][0], records[i - 1][1] if i else "") for i in range(len(records))]


def map_kmers_to_tree(tree: Tree, records: List[Tuple[str, str]], k_size: int = 35, m_size: int = 4,
                      pairing: str = "own", forward_only: bool = False) -> Tree:
    """build_database/mod.rs:26-181.  Node set of a k-mer = union of the root->tip id paths (both ends included,
    clade.rs:127-156) of every tip whose sequence contains it (kmers_map.rs:119-155).

    ``pairing="own"`` (default): each tip is indexed with its own sequence - the CORRECTED pairing every model
    of this repo is built with; ``pairing="reference"``: the pairing of :func:`reference_pairing` (the
    reference's build-side defect, outside the placement path - SURVEY.md section 8c note).
    ``forward_only``: windows of the forward strand only - the k-mer generation of the reference before v0.2.3
    (CHANGELOG.md:177-179 "fix the kmers generation that will not use the reverse complement"), which is the
    state that wrote the reference's golden build output; today's code indexes both strands (kmers_map.rs:387-395)."""
    if pairing not in ("own", "reference"):
        raise ValueError("pairing must be 'own' or 'reference'")
    km = KmersMap(k_size, m_size)
    leaves = {cl.name: path for cl, path in tree.root.get_leaves_with_paths(None)}
    for header, seq in (reference_pairing(records) if pairing == "reference" else records):
        path = set(leaves[header])
        kmers = (KmersMap.build_kmers_from_sequence(seq, k_size) if len(seq) >= k_size else []) if forward_only \
            else km.build_kmer_from_string(seq)
        for kmer, h in kmers:
            km.insert_or_append_kmer_hash(kmer, h, path)
    tree.kmers_map = km
    return tree


# --------------------------------------------------------------------------------------
# place_sequence (place_sequence.rs:42-602)
# --------------------------------------------------------------------------------------
class PlacementError(Exception):
    """``Err(MappedErrors)`` of place_sequence -> a line in ``<out>.error``."""


ERR_TOO_SHORT = "The sequence does not contain enough kmers."
ERR_MAX_ITER = "The maximum number of iterations has been reached."
MSG_NO_ROOT = "Query sequence has no overlapping kmers with the reference tree"
MSG_NO_INTROSPECTION = ("Tree introspection not possible. Query sequence has no overlapping kmers "
                        "with the reference tree")


def msg_no_match(header: str) -> str:
    # format!("{query:?}") of the newtype SequenceHeader(String) -> SequenceHeader("..."),
    # with Rust's Debug string escaping (place_sequence.rs:130-139).
    return f"Query sequence SequenceHeader({rust_debug_str(header)}) may not be related to the phylogeny"


def rust_debug_str(s: str) -> str:
    out = ['"']
    for ch in s:
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\0":
            out.append("\\0")
        elif ord(ch) < 0x20 or ord(ch) == 0x7F:
            out.append("\\u{%x}" % ord(ch))
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


@dataclass
class Placement:
    """Outcome of one query.  ``status`` is one of ``Unclassifiable``,
    ``IdentityFound``, ``MaxResolutionReached``, ``Inconclusive``; counters are the
    values the reference records on its tracing span (place_sequence.rs:90-182)."""
    status: str
    message: Optional[str] = None            # Unclassifiable / MaxResolutionReached / Inconclusive
    clade: Optional[int] = None              # IdentityFound (record id) / MaxResolutionReached (u64)
    one: Optional[int] = None
    rest: Optional[int] = None
    proposals: Optional[List[Tuple[int, int, int]]] = None   # Inconclusive: (clade, one, rest)
    n_query_kmers: int = 0
    n_matched: int = 0
    n_root_matched: int = 0
    iterations: int = 0

    def code(self) -> str:  # placement_response.rs:30-42
        if self.status == "IdentityFound":
            return "IdentityFound"
        return f"{self.status}: {self.message}"


def rust_round(x: float) -> float:
    """f64::round - half away from zero (place_sequence.rs:231-232)."""
    import math
    return math.floor(x + 0.5) if x >= 0 else math.ceil(x - 0.5)


def place_sequence(header: str, sequence: str, tree: Tree,
                   max_iterations: Optional[int] = None,
                   min_match_coverage: Optional[float] = None,
                   remove_intersection: Optional[bool] = None, trace: Optional[list] = None) -> Placement:
    """`trace` (test infrastructure for cls_debug_node_counts): receives one dict per evaluated level and
    non-leaf child with K(c) != None: parent_id, child_id, level, cnt = |K(c)|, excl = |K(c) - union of the
    siblings' K|, u = |union of all K|."""
    # :64-75
    remove_intersection = bool(remove_intersection) if remove_intersection is not None else False
    max_iterations = 1000 if max_iterations is None else max_iterations
    if min_match_coverage is None:
        cov = 0.7
    elif min_match_coverage > 1.0:
        cov = 1.0
    elif min_match_coverage < 0.0:
        cov = 0.0
    else:
        cov = min_match_coverage
    kmers_map = tree.kmers_map
    assert kmers_map is not None, "The tree does not have a kmers map."

    # :87-102
    query_kmers = kmers_map.build_kmer_from_string(sequence)
    if len(query_kmers) < 2:
        raise PlacementError(ERR_TOO_SHORT)
    res = Placement(status="", n_query_kmers=len(query_kmers))

    # :118-139
    query_kmers_map = kmers_map.get_overlapping_hashed_kmers(query_kmers)
    query_kmers_len = query_kmers_map.n_entries()
    res.n_matched = query_kmers_len
    if query_kmers_len == 0:
        res.status, res.message = "Unclassifiable", msg_no_match(header)
        return res

    # :156-166
    root_hashes = query_kmers_map.get_minimized_hashes_with_node(tree.root.id)
    if root_hashes is None:
        res.status, res.message = "Unclassifiable", MSG_NO_ROOT
        return res
    introspection_kmers = query_kmers_map.get_overlapping_minimized_hashes(root_hashes)

    # :199-211
    if tree.root.children is None:
        raise PlacementError("The root node does not have children. This is unexpected.")
    children = tree.root.children
    parent = tree.root
    iteration = 0

    # :231-254
    expected = rust_round(query_kmers_len * cov)
    introspection_coverage = introspection_kmers.n_entries()
    res.n_root_matched = introspection_coverage
    if introspection_coverage < int(expected):
        res.status, res.message = "Unclassifiable", f"Insufficient kmers coverage: {introspection_coverage}"
        return res

    while True:  # :279
        iteration += 1
        res.iterations = iteration
        if iteration > max_iterations:  # :295-301
            raise PlacementError(ERR_MAX_ITER)

        # PHASE 1 :311-428
        children_kmers = []
        for record in children:
            if record.is_leaf():
                continue
            k = introspection_kmers.get_hashed_kmers_with_node(record.id)
            if k is not None:
                children_kmers.append((k, record))
        children_kmers.sort(key=lambda t: -len(t[0]))  # stable, descending (:335)
        if trace is not None and children_kmers:
            union_all = set().union(*[k for k, _ in children_kmers])
            for kmers, clade in children_kmers:
                others = set().union(*[k for k, c in children_kmers if c.id != clade.id]) if len(children_kmers) > 1 else set()
                trace.append({"parent_id": parent.id, "child_id": clade.id, "level": iteration, "cnt": len(kmers),
                              "excl": len(kmers - others), "u": len(union_all)})

        proposals = []  # (clade, one, rest)
        for kmers, clade in children_kmers:
            rest = [rk for rk, nested in children_kmers if nested.id != clade.id]
            if not rest:
                one_n, rest_n = len(kmers), 0
            else:
                rest_set = set().union(*rest)
                if remove_intersection:
                    one_n, rest_n = len(kmers - rest_set), len(rest_set - kmers)
                else:
                    one_n, rest_n = len(kmers), len(rest_set)
            if one_n > rest_n:
                proposals.append((clade, one_n, rest_n))

        # PHASE 2 :436-600
        if not proposals:
            if iteration == 1:
                res.status, res.message = "Unclassifiable", MSG_NO_INTROSPECTION
                return res
            res.status, res.message, res.clade = "MaxResolutionReached", "LCA Accepted", parent.id
            return res

        if len(proposals) == 1:
            winner = proposals[0]
        else:
            by_diff: Dict[int, list] = {}
            for p in proposals:
                by_diff.setdefault(p[1] - p[2], []).append(p)
            best = by_diff[max(by_diff)]
            if len(best) != 1:
                res.status, res.message = "Inconclusive", "Multiple proposals"
                res.proposals = [(c.id, o, r) for c, o, r in proposals]
                return res
            winner = best[0]

        # update_introspection_node.rs:13-91
        clade, one_n, rest_n = winner
        non_leaf = [c for c in (clade.children or []) if not c.is_leaf()]
        if not non_leaf:
            res.status, res.clade, res.one, res.rest = "IdentityFound", clade.id, one_n, rest_n
            return res
        parent, children = clade, non_leaf


# --------------------------------------------------------------------------------------
# Result records (place_sequences/mod.rs:170-239, placement_response.rs:44-94)
# --------------------------------------------------------------------------------------
def placement_response(header: str, p: Placement, tree: Tree) -> dict:
    """The object the reference serialises per query (before YAML/JSON text).
    Field order: query, code, annotations?, placement?."""
    o: dict = {"query": header, "code": p.code()}
    placement = None
    clade_id = None
    if p.status == "IdentityFound":
        node = tree.root.get_node_by_id(p.clade)
        placement = {"clade": node.to_obj(), "one": p.one, "rest": p.rest}
        clade_id = p.clade
    elif p.status == "MaxResolutionReached":
        placement = p.clade
        clade_id = p.clade
    elif p.status == "Inconclusive":
        placement = p.code()
    if tree.annotations is not None:  # mod.rs:180-224
        ann = None
        if clade_id is not None:
            node = tree.root.get_node_by_id(clade_id)
            path = node.get_path_to_root(tree.root) if node is not None else set()
            recs = [a for a in tree.annotations if int(a["clade"]) in path]
            if recs:
                ann = sorted(recs, key=lambda a: a["clade"])
        if ann is not None:
            o["annotations"] = ann
    if placement is not None:
        o["placement"] = placement
    return o
