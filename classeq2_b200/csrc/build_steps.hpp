// Per-element steps of the device model builder (build_kernels.cu, cls_model_build_device): the pure
// index arithmetic of every kernel, written once as host+device functions so that the data flow of the
// pipeline can also be driven element by element on the CPU by the tests (tests/native/build_steps_host.cpp)
// and held equal to the host builder cls_model_build without a GPU.  Nothing in the product path calls
// these on the host.
//
// Pipeline (reference: map_kmers_to_tree, core/src/use_cases/build_database/mod.rs:26-181; the map it fills:
// kmers_map.rs:119-155 insert_or_append_kmer_hash; leaf paths: clade.rs:127-156):
//   1. tips are ranked by the pre-order position of their node; occurrences {hash, bucket key, rank} of every
//      window of both strands are generated tip by tip in rank order, one tile of windows per CTA
//   2. stable radix sorts by bucket key, then by hash: order (hash, bucket key, rank)
//   3. head flags of (hash, bucket) groups = entries; distinct ranks inside a group = the entry's tip list
//   4. tip lists are fingerprinted, entries sorted by fingerprint, neighbours compared in full: equal lists
//      share one node set; sets are numbered in order of their first entry
//   5. node set of a tip list = union of the root -> tip paths; the tips come in pre-order, so tip j adds
//      exactly the nodes from itself up to (excluding) its lowest common ancestor with tip j - 1
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CLS_HD __host__ __device__ __forceinline__
#else
#define CLS_HD inline
#endif

namespace cls {
namespace build {

constexpr uint32_t kTileWindows = 2048;  // windows of one strand per CTA of the hashing kernel

// Item x of a tile (x < 2 * nw): the tile holds forward windows [w0, w0 + nw) of a sequence with W windows per
// strand.  Its bytes f[0 .. nw + k - 1) and their reverse complement r[] (r[i] = comp(f[nw + k - 2 - i])) sit in
// shared memory; forward window w0 + i is f[i .. i + k), and r[i .. i + k) is the reverse complement of forward
// window w0 + nw - 1 - i, i.e. window W - 1 - (w0 + nw - 1 - i) of the reverse-complement strand.
struct TileItem {
    uint32_t strand;   // 0: read from f, 1: read from r
    uint32_t pos;      // byte offset of the window in f / r
    uint64_t occ;      // occurrence index relative to the sequence's first occurrence (forward windows
                       // first, then the windows of the reverse complement: kmers_map.rs:387-395)
};
CLS_HD TileItem tile_item(uint32_t x, uint32_t nw, uint64_t W, uint64_t w0) {
    TileItem t;
    t.strand = x >= nw ? 1u : 0u;
    t.pos = t.strand ? x - nw : x;
    t.occ = t.strand ? W + (W - 1 - (w0 + nw - 1 - t.pos)) : w0 + t.pos;
    return t;
}

// Reverse complement of one byte exactly as the host builder spells it (upper-cased; anything that is not
// A / T / C becomes 'C').
CLS_HD uint8_t comp_byte(uint8_t c) {
    const uint8_t u = c & 0xDFu;
    return u == 'A' ? 'T' : u == 'T' ? 'A' : u == 'C' ? 'G' : 'C';
}

// Flags of sorted occurrence i: bit 0 = first occurrence of its (hash, bucket) group (an entry starts),
// bit 32 = first occurrence of its rank inside the group (a tip-list element).  Packed so that ONE exclusive
// sum over uint64 yields the entry index (low half) and the tip-list position (high half).
CLS_HD uint64_t occ_flags(const uint64_t *hash, const uint64_t *bucket, const uint32_t *rank, uint64_t i) {
    const bool head = i == 0 || hash[i] != hash[i - 1] || bucket[i] != bucket[i - 1];
    const bool keep = head || rank[i] != rank[i - 1];
    return (head ? 1ull : 0ull) | (keep ? (1ull << 32) : 0ull);
}

CLS_HD uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// Fingerprint of a tip list (only used to bring equal lists next to each other; equality is decided by
// lists_equal, never by the fingerprint).
CLS_HD uint64_t list_fingerprint(const uint32_t *tips, uint32_t n) {
    uint64_t h = 0x13198a2e03707344ULL ^ n;
    for (uint32_t j = 0; j < n; ++j) h = mix64(h ^ tips[j]);
    return h;
}

CLS_HD bool lists_equal(const uint32_t *a, uint32_t na, const uint32_t *b, uint32_t nb) {
    if (na != nb) return false;
    for (uint32_t j = 0; j < na; ++j)
        if (a[j] != b[j]) return false;
    return true;
}

// Lowest common ancestor by climbing (parent = -1 at a root); -1 if a and b sit in different trees of a forest.
CLS_HD int32_t lca_climb(const int32_t *parent, const uint32_t *depth, int32_t a, int32_t b) {
    uint32_t da = depth[a], db = depth[b];
    while (da > db) { a = parent[a]; --da; }
    while (db > da) { b = parent[b]; --db; }
    while (a != b) {
        a = parent[a]; b = parent[b];
        if (a < 0 || b < 0) return -1;
    }
    return a;
}

// Number of nodes tip node `x` adds to a set whose previous tip node (in pre-order) is `prev` (-1: none).
CLS_HD uint32_t path_contribution(const int32_t *parent, const uint32_t *depth, int32_t prev, int32_t x) {
    if (prev < 0) return depth[x] + 1u;
    const int32_t a = lca_climb(parent, depth, prev, x);
    return a < 0 ? depth[x] + 1u : depth[x] - depth[a];
}

// Size of the node set of a tip list (ranks ascending = pre-order), and the same walk writing the ids.
CLS_HD uint64_t set_size(const int32_t *parent, const uint32_t *depth, const uint32_t *rank_node, const uint32_t *tips, uint32_t n) {
    uint64_t total = 0;
    int32_t prev = -1;
    for (uint32_t j = 0; j < n; ++j) {
        const int32_t x = (int32_t)rank_node[tips[j]];
        total += path_contribution(parent, depth, prev, x);
        prev = x;
    }
    return total;
}
CLS_HD void set_fill(const int32_t *parent, const uint32_t *depth, const uint32_t *rank_node, const uint64_t *node_id,
                     const uint32_t *tips, uint32_t n, uint64_t *out) {
    int32_t prev = -1;
    for (uint32_t j = 0; j < n; ++j) {
        const int32_t x = (int32_t)rank_node[tips[j]];
        const uint32_t c = path_contribution(parent, depth, prev, x);
        int32_t node = x;
        for (uint32_t i = 0; i < c; ++i) { *out++ = node_id[node]; node = parent[node]; }
        prev = x;
    }
}

}  // namespace build
}  // namespace cls
