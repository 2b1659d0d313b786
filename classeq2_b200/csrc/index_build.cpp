// Model -> GPU index serialisation (host side, once per model).
//
// Reference types being flattened (paths relative to the reference checkout):
//   Tree / Clade      core/src/domain/dtos/tree.rs:9-52, clade.rs:18-38
//   KmersMap          core/src/domain/dtos/kmers_map.rs:77-87  (bucket key -> hash -> node-id set)
// Reference behaviour preserved by the layout:
//   * bucket gating of get_overlapping_hashed_kmers (kmers_map.rs:273-311): an entry can only
//     match a query whose k-mers produce the entry's bucket key h1(prefix_m); every entry
//     carries the 2-bit code of the prefix whose hash equals its bucket key, entries whose
//     bucket key is not the hash of any ACGT prefix can never match and are dropped;
//   * is_leaf() is by Clade.kind (clade.rs:166-172), never by "has children";
//   * the descent only ever consults NON-LEAF children (place_sequence.rs:319-333,
//     update_introspection_node.rs:32-45) and tree.root.id (place_sequence.rs:156-166), so
//     node sets are stored restricted to non-leaf nodes reachable from the root.
#include "index_build.hpp"

#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "murmur3_host.hpp"

namespace cls {

namespace {

struct KeptEntry {
    uint64_t hash;
    uint64_t bucket;
    uint32_t set_off;
    uint32_t code;
};

inline uint64_t mix64(uint64_t x) { return fmix64_h(x + 0x9e3779b97f4a7c15ULL); }

}  // namespace

int build_host_index(const cls_model_view *mv, HostIndex &out, std::string &err, uint32_t shard, uint32_t n_shards) {
    if (!mv) { err = "model view is NULL"; return CLS_ERR_INVALID_ARGUMENT; }
    if (mv->k_size == 0) { err = "k_size == 0 is not supported"; return CLS_ERR_UNSUPPORTED; }
    if (mv->m_size > 12) { err = "m_size > 12 is not supported (prefix code table would exceed 4^12)"; return CLS_ERR_UNSUPPORTED; }
    if (mv->n_nodes == 0 || !mv->node_id || !mv->node_kind || !mv->child_off) {
        err = "model view has no nodes"; return CLS_ERR_INVALID_ARGUMENT;
    }
    if (mv->n_nodes >= (1ull << 24)) { err = "more than 2^24 nodes is not supported"; return CLS_ERR_UNSUPPORTED; }
    if (mv->n_entries && (!mv->entry_bucket || !mv->entry_hash || !mv->entry_set || !mv->set_off)) {
        err = "model view entry arrays are NULL"; return CLS_ERR_INVALID_ARGUMENT;
    }
    const uint64_t n_nodes = mv->n_nodes;
    const uint64_t n_child = mv->child_off[n_nodes];
    if (n_child && !mv->child_idx) { err = "child_idx is NULL"; return CLS_ERR_INVALID_ARGUMENT; }
    for (uint64_t i = 0; i < n_nodes; ++i) {
        if (mv->child_off[i] > mv->child_off[i + 1]) { err = "child_off is not non-decreasing"; return CLS_ERR_INVALID_ARGUMENT; }
        if (mv->node_kind[i] > CLS_KIND_LEAF) { err = "invalid node kind"; return CLS_ERR_INVALID_ARGUMENT; }
    }
    {
        std::vector<uint8_t> seen(n_nodes, 0);
        for (uint64_t j = 0; j < n_child; ++j) {
            uint64_t c = mv->child_idx[j];
            if (c >= n_nodes || c == 0) { err = "child index out of range (or the root listed as a child)"; return CLS_ERR_INVALID_ARGUMENT; }
            if (seen[c]) { err = "a node is listed as a child twice"; return CLS_ERR_INVALID_ARGUMENT; }
            seen[c] = 1;
        }
    }
    {
        std::vector<uint64_t> ids(mv->node_id, mv->node_id + n_nodes);
        std::sort(ids.begin(), ids.end());
        if (std::adjacent_find(ids.begin(), ids.end()) != ids.end()) {
            err = "duplicated Clade ids are not supported"; return CLS_ERR_UNSUPPORTED;
        }
    }

    out = HostIndex();
    out.k_size = mv->k_size;
    out.m_size = mv->m_size;
    out.m_eff = std::min(mv->m_size, mv->k_size);
    out.root_children_none = (mv->flags & CLS_MODEL_ROOT_CHILDREN_NONE) != 0;

    // ---- flatten the tree over non-leaf nodes, DFS pre-order, root = q 0 --------------------
    constexpr uint32_t kNoQ = 0xFFFFFFFFu;
    std::vector<uint32_t> node_q(n_nodes, kNoQ);
    std::vector<uint32_t> q_node;    // q -> node index
    std::vector<uint32_t> q_parent;  // q -> parent q
    std::vector<uint32_t> q_ord;     // q -> ordinal among the parent's non-leaf children
    std::vector<uint32_t> q_end;     // q -> one past the last q of its subtree (pre-order interval)
    std::vector<uint32_t> lca_tour;  // Euler tour of the non-leaf tree
    uint32_t lca_levels = 0;
    {
        node_q[0] = 0; q_node.push_back(0); q_parent.push_back(kNoQ); q_ord.push_back(0);
        out.qnodes.push_back(QNode{0, 0});
        // a node's non-leaf children must be contiguous in q_child_list: collect them when the
        // node is first expanded, then descend into them one by one (pre-order).
        auto expand = [&](uint32_t q) {
            uint32_t node = q_node[q];
            QNode qn{(uint32_t)out.q_child_list.size(), 0};
            for (uint64_t j = mv->child_off[node]; j < mv->child_off[node + 1]; ++j) {
                uint64_t c = mv->child_idx[j];
                if (mv->node_kind[c] == CLS_KIND_LEAF) continue;
                out.q_child_list.push_back((uint32_t)c);  // node index for now; rewritten to q below
                qn.child_count++;
            }
            out.qnodes[q] = qn;
            out.max_fanout = std::max(out.max_fanout, qn.child_count);
        };
        expand(0);
        std::vector<uint32_t> cursor(1, 0);  // per open q: next child ordinal to visit
        std::vector<uint32_t> open{0};
        while (!open.empty()) {
            uint32_t q = open.back();
            QNode qn = out.qnodes[q];
            if (cursor[q] == qn.child_count) { open.pop_back(); continue; }
            uint32_t ord = cursor[q]++;
            uint32_t cnode = out.q_child_list[qn.child_first + ord];
            uint32_t cq = (uint32_t)q_node.size();
            node_q[cnode] = cq;
            q_node.push_back(cnode); q_parent.push_back(q); q_ord.push_back(ord);
            out.qnodes.push_back(QNode{0, 0}); cursor.push_back(0);
            out.q_child_list[qn.child_first + ord] = cq;
            expand(cq);
            open.push_back(cq);
        }
        const uint32_t nq = (uint32_t)q_node.size();
        q_end.assign(nq, 0);
        for (uint32_t q = nq; q-- > 0;) {
            // pre-order: the subtree of q ends where the last child's subtree ends
            QNode qn = out.qnodes[q];
            q_end[q] = qn.child_count ? q_end[out.q_child_list[qn.child_first + qn.child_count - 1]] : q + 1;
        }
        out.q_node_id.resize(nq);
        for (uint32_t q = 0; q < nq; ++q) out.q_node_id[q] = mv->node_id[q_node[q]];
        // closed-mode tree arrays: subtree intervals, depths, Euler tour + sparse table for LCA
        out.qinfo.assign(nq, QInfo{0, 0, 0, 0});
        for (uint32_t q = 0; q < nq; ++q) {
            out.qinfo[q].q_end = q_end[q];
            out.qinfo[q].child_count = out.qnodes[q].child_count;
            if (q) out.qinfo[q].depth = out.qinfo[q_parent[q]].depth + 1;  // pre-order: parents first
        }
        std::vector<uint32_t> tour;
        tour.reserve(2 * (size_t)nq);
        {
            std::vector<uint32_t> stack{0}, cur(nq, 0);
            out.qinfo[0].euler_first = 0;
            tour.push_back(0);
            while (!stack.empty()) {
                const uint32_t q = stack.back();
                const QNode qn = out.qnodes[q];
                if (cur[q] == qn.child_count) {
                    stack.pop_back();
                    if (!stack.empty()) tour.push_back(stack.back());
                    continue;
                }
                const uint32_t c = out.q_child_list[qn.child_first + cur[q]++];
                out.qinfo[c].euler_first = (uint32_t)tour.size();
                tour.push_back(c);
                stack.push_back(c);
            }
        }
        out.euler_len = (uint32_t)tour.size();
        uint32_t levels = 1;
        while ((2u << (levels - 1)) <= out.euler_len) ++levels;
        lca_levels = levels;
        lca_tour.swap(tour);
    }
    const uint32_t nq = (uint32_t)q_node.size();

    // id -> q lookup (dense table when the ids are small, hash map otherwise)
    uint64_t max_id = 0;
    for (uint32_t q = 0; q < nq; ++q) max_id = std::max(max_id, out.q_node_id[q]);
    const bool dense_ids = max_id < (1ull << 26);
    std::vector<uint32_t> id_q_dense;
    std::unordered_map<uint64_t, uint32_t> id_q_sparse;
    if (dense_ids) {
        id_q_dense.assign(max_id + 1, kNoQ);
        for (uint32_t q = 0; q < nq; ++q) id_q_dense[out.q_node_id[q]] = q;
    } else {
        id_q_sparse.reserve(nq * 2);
        for (uint32_t q = 0; q < nq; ++q) id_q_sparse.emplace(out.q_node_id[q], q);
    }
    auto id_to_q = [&](uint64_t id) -> uint32_t {
        if (dense_ids) return id <= max_id ? id_q_dense[id] : kNoQ;
        auto it = id_q_sparse.find(id);
        return it == id_q_sparse.end() ? kNoQ : it->second;
    };

    // ---- node sets -> de-duplicated mini-tree records ------------------------------------------
    const uint64_t n_sets = mv->n_sets;
    std::vector<uint32_t> set_arena_off(n_sets, kEmpty);
    if (n_sets && (!mv->set_off || (mv->set_off[n_sets] && !mv->set_node_ids))) {
        err = "set arrays are NULL"; return CLS_ERR_INVALID_ARGUMENT;
    }
    for (uint64_t s = 0; s < n_sets; ++s)
        if (mv->set_off[s] > mv->set_off[s + 1]) { err = "set_off is not non-decreasing"; return CLS_ERR_INVALID_ARGUMENT; }

    // ---- descent mode: CLOSED iff every root-containing set is upward closed over the non-leaf tree
    bool closed = (mv->flags & CLS_MODEL_FORCE_GENERAL_SETS) == 0;
    {
        std::vector<uint64_t> stamp(nq, 0);
        std::vector<uint32_t> members;
        for (uint64_t s = 0; s < n_sets && closed; ++s) {
            const uint64_t tag = s + 1;
            members.clear();
            for (uint64_t j = mv->set_off[s]; j < mv->set_off[s + 1]; ++j) {
                uint32_t q = id_to_q(mv->set_node_ids[j]);
                if (q == kNoQ || stamp[q] == tag) continue;
                stamp[q] = tag;
                members.push_back(q);
            }
            if (stamp[0] != tag) continue;  // no root: never consulted by the descent
            for (uint32_t q : members)
                if (q != 0 && stamp[q_parent[q]] != tag) { closed = false; break; }
        }
    }
    out.closed = closed;

    if (closed && (uint64_t)lca_levels * out.euler_len > (1ull << 27)) closed = out.closed = false;  // LCA table > 1 GiB
    if (closed) {
        // sparse table over the Euler tour: level j holds minima of (depth << 32 | q) over [i, i + 2^j)
        const size_t len = out.euler_len;
        out.lca_table.assign((size_t)lca_levels * len, ~0ull);
        for (size_t i = 0; i < len; ++i) out.lca_table[i] = ((uint64_t)out.qinfo[lca_tour[i]].depth << 32) | lca_tour[i];
        for (uint32_t j = 1; j < lca_levels; ++j) {
            const size_t half = (size_t)1 << (j - 1);
            const uint64_t *prev = &out.lca_table[(size_t)(j - 1) * len];
            uint64_t *cur = &out.lca_table[(size_t)j * len];
            for (size_t i = 0; i + 2 * half <= len; ++i) cur[i] = std::min(prev[i], prev[i + half]);
        }
        // terminal-list records (device_types.hpp), de-duplicated by content
        std::unordered_map<uint64_t, std::vector<uint32_t>> dedup;
        dedup.reserve(n_sets / 4 + 16);
        std::vector<uint64_t> stamp(nq, 0), has_child(nq, 0);
        std::vector<uint32_t> members, rec;
        for (uint64_t s = 0; s < n_sets; ++s) {
            const uint64_t tag = s + 1;
            members.clear();
            for (uint64_t j = mv->set_off[s]; j < mv->set_off[s + 1]; ++j) {
                uint32_t q = id_to_q(mv->set_node_ids[j]);
                if (q == kNoQ || stamp[q] == tag) continue;
                stamp[q] = tag;
                members.push_back(q);
            }
            rec.clear();
            if (stamp[0] != tag) {
                rec.push_back(0u);
                rec.push_back(0u);
            } else {
                for (uint32_t q : members) if (q != 0) has_child[q_parent[q]] = tag;
                rec.push_back(0u);
                rec.push_back(0u);
                for (uint32_t q : members) if (has_child[q] != tag) rec.push_back(q);
                std::sort(rec.begin() + 2, rec.end());
                rec[0] = (uint32_t)(rec.size() - 2) | kTermHasRoot;
                rec[1] = rec.back();
            }
            uint64_t h = 0x243f6a8885a308d3ULL;
            for (uint32_t w : rec) h = mix64(h ^ w);
            auto &cands = dedup[h];
            uint32_t found = kEmpty;
            for (uint32_t off : cands) {
                if ((out.terms[off] & ~kTermHasRoot) + 2 == rec.size() &&
                    std::memcmp(&out.terms[off], rec.data(), rec.size() * sizeof(uint32_t)) == 0) { found = off; break; }
            }
            if (found == kEmpty) {
                if (out.terms.size() + rec.size() >= 0x7FFFFFF0ull) { err = "node-set arena exceeds 2^31 words"; return CLS_ERR_UNSUPPORTED; }  // bit 31 of a record offset is free (scan2_kernel)
                found = (uint32_t)out.terms.size();
                out.terms.insert(out.terms.end(), rec.begin(), rec.end());
                cands.push_back(found);
                out.n_distinct_sets++;
            }
            set_arena_off[s] = found;
        }
        for (int pad = 0; pad < 4; ++pad) out.terms.push_back(0u);  // records may be read a little past their end
    } else {
        std::unordered_map<uint64_t, std::vector<uint32_t>> dedup;  // content hash -> arena offsets
        dedup.reserve(n_sets / 4 + 16);
        std::vector<uint32_t> stamp(nq, 0), present(nq, 0);
        std::vector<uint32_t> members;
        std::vector<SetWord> rec;
        std::vector<uint32_t> open;
        for (uint64_t s = 0; s < n_sets; ++s) {
            const uint32_t tag = (uint32_t)(s + 1);
            if (tag == 0) { std::fill(stamp.begin(), stamp.end(), 0); std::fill(present.begin(), present.end(), 0); }
            members.clear();
            // closure under ancestors; the root is always part of the record
            stamp[0] = tag; members.push_back(0);
            bool has_root = false;
            for (uint64_t j = mv->set_off[s]; j < mv->set_off[s + 1]; ++j) {
                uint32_t q = id_to_q(mv->set_node_ids[j]);
                if (q == kNoQ) continue;
                if (q == 0) has_root = true;
                present[q] = tag;
                while (stamp[q] != tag) { stamp[q] = tag; members.push_back(q); q = q_parent[q]; }
            }
            std::sort(members.begin(), members.end());
            rec.clear();
            rec.push_back(SetWord{has_root ? kSetHasRoot : 0u, (uint32_t)members.size()});
            open.clear();
            for (uint32_t i = 0; i < members.size(); ++i) {
                uint32_t q = members[i];
                while (!open.empty() && q >= q_end[members[open.back()]]) {
                    uint32_t e = open.back(); open.pop_back();
                    rec[1 + e].y = i - e;
                }
                rec.push_back(SetWord{q_ord[q] | (present[q] == tag ? kPresentBit : 0u), 0});
                open.push_back(i);
            }
            while (!open.empty()) {
                uint32_t e = open.back(); open.pop_back();
                rec[1 + e].y = (uint32_t)members.size() - e;
            }
            uint64_t h = 0x243f6a8885a308d3ULL;
            for (const SetWord &w : rec) h = mix64(h ^ (((uint64_t)w.x << 32) | w.y));
            auto &cands = dedup[h];
            uint32_t found = kEmpty;
            for (uint32_t off : cands) {
                if (out.arena[off].y + 1 == rec.size() &&
                    std::memcmp(&out.arena[off], rec.data(), rec.size() * sizeof(SetWord)) == 0) { found = off; break; }
            }
            if (found == kEmpty) {
                if (out.arena.size() + rec.size() >= 0xFFFFFFF0ull) { err = "node-set arena exceeds 2^32 words"; return CLS_ERR_UNSUPPORTED; }
                found = (uint32_t)out.arena.size();
                out.arena.insert(out.arena.end(), rec.begin(), rec.end());
                cands.push_back(found);
                out.n_distinct_sets++;
            }
            set_arena_off[s] = found;
        }
    }
    if (out.arena.empty()) out.arena.push_back(SetWord{0, 0});  // never dereferenced; keeps uploads non-empty
    if (out.terms.empty()) out.terms.push_back(0u);

    // ---- bucket key -> prefix code -------------------------------------------------------------
    std::unordered_map<uint64_t, uint32_t> key_code;
    {
        const uint32_t mm = out.m_eff;
        const uint64_t n_codes = 1ull << (2 * mm);
        key_code.reserve(n_codes * 2);
        static const char kLetters[4] = {'A', 'C', 'T', 'G'};  // code = (ascii >> 1) & 3
        uint8_t buf[16];
        for (uint64_t code = 0; code < n_codes; ++code) {
            for (uint32_t j = 0; j < mm; ++j) buf[j] = (uint8_t)kLetters[(code >> (2 * j)) & 3];
            uint64_t key = mv->m_size == 0 ? 0 : murmur3_x64_128_h1(buf, mm, 0);
            if (!key_code.emplace(key, (uint32_t)code).second) {
                err = "two distinct k-mer prefixes hash to the same bucket key"; return CLS_ERR_UNSUPPORTED;
            }
        }
    }

    // ---- entries ---------------------------------------------------------------------------------
    std::vector<KeptEntry> kept;
    kept.reserve(mv->n_entries);
    for (uint64_t i = 0; i < mv->n_entries; ++i) {
        auto it = key_code.find(mv->entry_bucket[i]);
        if (it == key_code.end()) continue;  // unreachable bucket: no ACGT query can produce this key
        if (n_shards > 1 && (uint32_t)(mv->entry_hash[i] >> 61) % n_shards != shard) continue;  // another GPU owns it
        uint64_t s = mv->entry_set[i];
        if (s >= n_sets) { err = "entry_set out of range"; return CLS_ERR_INVALID_ARGUMENT; }
        kept.push_back(KeptEntry{mv->entry_hash[i], mv->entry_bucket[i], set_arena_off[s], it->second});
    }
    {
        std::vector<uint32_t> order(kept.size());
        for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
        if (kept.size() >= 0xFFFFFFFFull) { err = "more than 2^32 index entries"; return CLS_ERR_UNSUPPORTED; }
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return kept[a].hash < kept[b].hash; });
        for (size_t i = 1; i < order.size(); ++i) {
            const KeptEntry &a = kept[order[i - 1]], &b = kept[order[i]];
            if (a.hash == b.hash) {
                if (a.bucket == b.bucket) { err = "the same (bucket key, hash) pair is listed twice"; return CLS_ERR_INVALID_ARGUMENT; }
                err = "the same k-mer hash is stored under two bucket keys (cross-bucket 64-bit collision); not supported";
                return CLS_ERR_UNSUPPORTED;
            }
        }
    }
    out.n_entries_kept = kept.size();

    // ---- open-addressed table: 32-byte buckets of two slots, load factor in (0.25, 0.5] ----------
    // A free slot of bucket j carries the "hash" (j + 1) & mask: a probe reaches bucket j only from a home bucket
    // at cyclic distance < n_buckets - 1 behind it (checked below), never from j + 1, so the hash it asks for
    // cannot equal the filler and the kernels may treat hash equality alone as a hit.
    uint64_t nb = 4;
    while (nb < kept.size()) nb <<= 1;
    for (;;) {
        out.n_buckets = nb;
        out.table.assign(2 * nb, Slot{0, kEmpty, 0});
        const uint64_t mask = nb - 1;
        uint64_t max_dist = 0;
        for (const KeptEntry &e : kept) {
            const uint64_t home = e.hash & mask;
            uint64_t b = home, dist = 0;
            for (;;) {
                Slot *s = &out.table[2 * b];
                if (s[0].set_off == kEmpty) { uint32_t ov = s[0].code & kOverflowBit; s[0] = Slot{e.hash, e.set_off, e.code | ov}; break; }
                if (s[1].set_off == kEmpty) { s[1] = Slot{e.hash, e.set_off, e.code}; break; }
                s[0].code |= kOverflowBit;
                b = (b + 1) & mask;
                ++dist;
            }
            // the home bucket is full by now: slot 1's code carries the filter of the entries that moved on
            if (dist) out.table[2 * home + 1].code |= 1u << (kBloomShift + (uint32_t)((e.hash >> 40) & 7u));
            if (dist > max_dist) max_dist = dist;
        }
        // a miss walks one bucket past the last overflowed one
        if (max_dist + 2 < nb) {
            for (uint64_t j = 0; j < nb; ++j)
                for (int q = 0; q < 2; ++q)
                    if (out.table[2 * j + q].set_off == kEmpty) out.table[2 * j + q].hash = (j + 1) & mask;
            break;
        }
        nb <<= 1;   // pathological clustering in a tiny table: spread it out
    }
    return CLS_OK;
}

}  // namespace cls
