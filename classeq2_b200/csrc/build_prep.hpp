// Host preparation of the device model builder (cls_model_build_device): parent / depth / pre-order arrays of
// the tree, the ranking of the tips by pre-order position, and the work list of the hashing kernel.
// O(nodes + tips) bookkeeping; all k-mer work happens in build_kernels.cu.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/classeq_b200.h"
#include "build_steps.hpp"

namespace cls {
namespace build {

struct HashItem {
    uint32_t rank;  // tip, by rank
    uint32_t tile;  // windows [tile * kTileWindows, ...) of both strands
};

struct Prep {
    std::vector<int32_t> parent;      // node index -> parent node index, -1 at a root
    std::vector<uint32_t> depth;      // node index -> distance from its root
    std::vector<uint32_t> rank_node;  // rank -> node index of the tip
    std::vector<uint32_t> rank_tip;   // rank -> caller's tip index
    std::vector<uint64_t> seq_off;    // rank -> first byte of the sequence in `bases`
    std::vector<uint32_t> seq_len;    // rank -> sequence length
    std::vector<uint64_t> occ_off;    // rank -> first occurrence; [n_tips] = number of occurrences
    std::vector<HashItem> items;
};

// Returns 0 or a negative cls_error with `err` set.
// `bases` (may be NULL: no check): sequences must hold A / C / G / T only, either case - the reference upper-cases both
// strands (kmers_map.rs:410) and panics on anything else (:431-443).
inline int prepare(const cls_model_view *tree, uint64_t n_tips, const uint64_t *tip_node, const uint64_t *offsets,
                   Prep &p, std::string &err, const uint8_t *bases = nullptr) {
    const uint64_t n_nodes = tree->n_nodes;
    const uint32_t k = tree->k_size;
    if (k == 0) { err = "k_size == 0"; return CLS_ERR_UNSUPPORTED; }
    if (n_nodes == 0 || n_nodes > 0x7FFFFFFFull || n_tips >= 0xFFFFFFFFull) { err = "bad sizes"; return CLS_ERR_INVALID_ARGUMENT; }
    // parent links exactly as the host builder reads them (the last listing of a child wins)
    p.parent.assign(n_nodes, -1);
    for (uint64_t q = 0; q < n_nodes; ++q)
        for (uint64_t j = tree->child_off[q]; j < tree->child_off[q + 1]; ++j) {
            if (tree->child_idx[j] >= n_nodes) { err = "child index out of range"; return CLS_ERR_INVALID_ARGUMENT; }
            p.parent[tree->child_idx[j]] = (int32_t)q;
        }
    // children lists derived from the parent links (a forest by construction unless the links form a cycle)
    std::vector<uint32_t> c_off(n_nodes + 1, 0), c_idx(n_nodes);
    for (uint64_t q = 0; q < n_nodes; ++q)
        if (p.parent[q] >= 0) ++c_off[p.parent[q] + 1];
    for (uint64_t q = 0; q < n_nodes; ++q) c_off[q + 1] += c_off[q];
    {
        std::vector<uint32_t> fill(c_off.begin(), c_off.end() - 1);
        for (uint64_t q = 0; q < n_nodes; ++q)
            if (p.parent[q] >= 0) c_idx[fill[p.parent[q]]++] = (uint32_t)q;
    }
    // iterative pre-order over every tree of the forest, roots in index order
    p.depth.assign(n_nodes, 0);
    std::vector<uint32_t> pre(n_nodes, 0), stack;
    uint64_t visited = 0;
    for (uint64_t root = 0; root < n_nodes; ++root) {
        if (p.parent[root] >= 0) continue;
        stack.push_back((uint32_t)root);
        while (!stack.empty()) {
            const uint32_t q = stack.back();
            stack.pop_back();
            pre[q] = (uint32_t)visited++;
            for (uint32_t j = c_off[q + 1]; j > c_off[q]; --j) {  // reversed: the first child is visited first
                const uint32_t c = c_idx[j - 1];
                p.depth[c] = p.depth[q] + 1;
                stack.push_back(c);
            }
        }
    }
    if (visited != n_nodes) { err = "the child lists contain a cycle"; return CLS_ERR_INVALID_ARGUMENT; }
    for (uint64_t t = 0; t < n_tips; ++t) {
        if (tip_node[t] >= n_nodes) { err = "tip node out of range"; return CLS_ERR_INVALID_ARGUMENT; }
        if (offsets[t + 1] < offsets[t]) { err = "offsets decrease"; return CLS_ERR_INVALID_ARGUMENT; }
        if (offsets[t + 1] - offsets[t] > 0xFFFFFFFFull) { err = "a sequence is longer than 4 GiB"; return CLS_ERR_UNSUPPORTED; }
        if (bases)
            for (uint64_t i = offsets[t]; i < offsets[t + 1]; ++i) {
                const uint8_t c = bases[i] & 0xDFu;
                if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) { err = "a tip sequence holds a character other than A, C, G, T"; return CLS_ERR_INVALID_ARGUMENT; }
            }
    }
    // rank the tips by (pre-order position of their node, tip index)
    p.rank_tip.resize(n_tips);
    for (uint64_t t = 0; t < n_tips; ++t) p.rank_tip[t] = (uint32_t)t;
    std::sort(p.rank_tip.begin(), p.rank_tip.end(), [&](uint32_t a, uint32_t b) {
        const uint32_t pa = pre[tip_node[a]], pb = pre[tip_node[b]];
        return pa != pb ? pa < pb : a < b;
    });
    p.rank_node.resize(n_tips);
    p.seq_off.resize(n_tips);
    p.seq_len.resize(n_tips);
    p.occ_off.assign(n_tips + 1, 0);
    p.items.clear();
    for (uint64_t r = 0; r < n_tips; ++r) {
        const uint32_t t = p.rank_tip[r];
        p.rank_node[r] = (uint32_t)tip_node[t];
        p.seq_off[r] = offsets[t];
        const uint64_t len = offsets[t + 1] - offsets[t];
        p.seq_len[r] = (uint32_t)len;
        const uint64_t W = len >= k ? len - k + 1 : 0;
        p.occ_off[r + 1] = p.occ_off[r] + 2 * W;
        for (uint64_t w0 = 0; w0 < W; w0 += kTileWindows) p.items.push_back(HashItem{(uint32_t)r, (uint32_t)(w0 / kTileWindows)});
    }
    return CLS_OK;
}

}  // namespace build
}  // namespace cls
