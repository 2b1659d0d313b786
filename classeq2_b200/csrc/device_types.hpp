// Internal layouts shared by the host-side index builder and the sm_100a kernels.
// Nothing here crosses the C ABI (include/classeq_b200.h).
#pragma once
#include <cstdint>

namespace cls {

// ---- k-mer table -------------------------------------------------------------------
// Open addressing over 32-byte buckets of two 16-byte slots (one DRAM sector per
// probe).  bucket = hash & bucket_mask; linear probing over buckets on overflow.
struct Slot {
    uint64_t hash;     // murmur3_x64_128(kmer, 0).0  (kmers_map.rs:157-159)
    uint32_t set_off;  // offset of the node-set record (SetWord units in `arena`, or u32 units in
                       // `terms`, depending on the descent mode); kEmpty = free slot
    uint32_t code;     // bits 0..23: 2-bit prefix code of the entry's bucket (MinimizerKey);
                       // bit 31 (slot 0 of a bucket only): bucket overflowed at build time
};
static_assert(sizeof(Slot) == 16, "slot must be 16 bytes");
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr uint32_t kOverflowBit = 0x80000000u;
constexpr uint32_t kCodeMask = 0x00FFFFFFu;
// bits 24..31 of the code of slot 1: an 8-bit filter over the entries whose HOME is this bucket but which were stored
// further along the chain (bit (hash >> 40) & 7).  A probe that misses in its home bucket follows the chain only if
// its own bit is set (scan2_kernel); the overflow bit alone (every kernel may use it) says "some entry went on".
constexpr uint32_t kBloomShift = 24;

// ---- node-set records, GENERAL mode ("mini-trees") -----------------------------------
// One record per DISTINCT node set: the set restricted to the non-leaf nodes of the model
// tree, closed under ancestors, in DFS pre-order.
//   arena[off]      header  {x: flags (bit0 = set contains tree.root.id), y: n_entries}
//   arena[off+1+i]  entry i {x: ordinal among the parent's NON-LEAF children | present<<31,
//                            y: size of this entry's subtree in entries (>= 1)}
// entry 0 is the root.  Children of entry e are e+1, e+1+size(e+1), ... < e+size(e).
// `present` = the node is a member of the set (ancestors added by the closure are
// structural only and never vote), which keeps arbitrary - not upward-closed - sets exact.
struct SetWord {
    uint32_t x, y;
};
constexpr uint32_t kPresentBit = 0x80000000u;
constexpr uint32_t kSetHasRoot = 1u;

// ---- node-set records, CLOSED mode ("terminal lists") ---------------------------------
// Used when every root-containing set of the model is upward closed over the non-leaf tree
// (always true for models made by the reference's builder: node sets are unions of root->tip
// paths, build_database/mod.rs:139-168).  Such a set is determined by its TERMINALS - the
// members none of whose non-leaf children are members - and with non-leaf nodes numbered in
// DFS pre-order (q ids; subtree(q) = [q, q_end[q])) membership of a node x is
// "some terminal lies in [x, q_end[x])".
//   terms[off]          header: n_terminals | (contains root) << 31
//   terms[off+1]        copy of the LAST terminal (so first and last arrive in one round trip)
//   terms[off+2 .. +n]  terminal q ids, ascending
// Sets that do not contain the root carry n = 0: they only count towards |M|.
constexpr uint32_t kTermHasRoot = 0x80000000u;

// Per non-leaf node, closed mode.  LCA queries go through an Euler tour of the non-leaf tree and
// a sparse table of (depth << 32 | q) minima: two dependent round trips, whatever the depth.
struct QInfo {
    uint32_t q_end;        // one past the last q of the subtree
    uint32_t child_count;  // non-leaf children; they tile [q+1, q_end): c1 = q+1, c2 = q_end[c1], ...
    uint32_t euler_first;  // first occurrence in the Euler tour
    uint32_t depth;
};
static_assert(sizeof(QInfo) == 16, "QInfo must be 16 bytes");

// ---- flattened tree over non-leaf nodes (dense ids q in DFS pre-order, root = 0) ------
struct QNode {
    uint32_t child_first;  // into q_child_list
    uint32_t child_count;  // number of NON-LEAF children (Clade.kind != LEAF)
};

// ---- batch ---------------------------------------------------------------------------
struct ReadDesc {
    uint32_t word_off;  // first 32-bit word of the 2-bit packed bases (16 bases / word)
    uint32_t len;       // bases
};

struct ResultRec {      // 32 bytes, one per query, written once
    uint64_t node_id;
    int32_t one;
    int32_t rest;
    uint32_t n_matched;
    uint32_t n_root_matched;
    uint32_t iterations;
    uint32_t status;
};
static_assert(sizeof(ResultRec) == 32, "result record must be 32 bytes");

struct DeviceIndex {
    const Slot *table;
    uint64_t bucket_mask;
    const SetWord *arena;         // general mode
    const uint32_t *terms;        // closed mode
    const QNode *qnodes;
    const uint32_t *q_child_list;
    const uint64_t *q_node_id;
    const QInfo *qinfo;           // closed mode
    const uint64_t *lca_table;    // closed mode: sparse table, level j at lca_table + j * euler_len
    uint32_t n_q;
    uint32_t euler_len;
    uint32_t k_size;
    uint32_t m_eff;         // min(m, k): number of prefix bases in the bucket code
    uint32_t max_fanout;    // max child_count over qnodes
    uint32_t root_children_none;
    uint32_t closed;        // 1 = terminal-list records, 0 = mini-tree records
};

struct PlaceParams {
    int32_t max_iterations;
    uint32_t remove_intersection;
    double min_match_coverage;  // already clamped
};

// 2-bit base code used everywhere: (ascii >> 1) & 3  ->  A=0 C=1 T=2 G=3 (case-insensitive);
// complement = code ^ 2.
constexpr uint32_t kAsciiLut = 0x47544341u;    // byte[code] = "ACTG"[code]

}  // namespace cls
