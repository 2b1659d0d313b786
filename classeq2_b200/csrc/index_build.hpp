// Host-side serialisation of the reference's model types (Tree/Clade/KmersMap, passed as a
// cls_model_view) into the flat arrays the sm_100a kernels consume.  See device_types.hpp for
// the layouts and DESIGN.md for the rationale.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/classeq_b200.h"
#include "device_types.hpp"

namespace cls {

struct HostIndex {
    uint32_t k_size = 0, m_size = 0, m_eff = 0;
    bool root_children_none = false;
    std::vector<Slot> table;  // 2 * n_buckets slots
    uint64_t n_buckets = 0;
    uint64_t n_entries_kept = 0;
    bool closed = false;            // terminal-list records (`terms`) instead of mini-trees (`arena`)
    std::vector<SetWord> arena;
    std::vector<uint32_t> terms;
    std::vector<QInfo> qinfo;             // closed mode
    std::vector<uint64_t> lca_table;      // closed mode: n_levels x euler_len
    uint32_t euler_len = 0;
    uint64_t n_distinct_sets = 0;
    std::vector<QNode> qnodes;
    std::vector<uint32_t> q_child_list;
    std::vector<uint64_t> q_node_id;
    uint32_t max_fanout = 0;
};

// Returns CLS_OK or a negative cls_error; `err` receives the message.
// Only the entries with (hash >> 61) % n_shards == shard go into the table (hash-sharded index);
// node-set records and the tree are always complete.
int build_host_index(const cls_model_view *mv, HostIndex &out, std::string &err, uint32_t shard = 0, uint32_t n_shards = 1);

}  // namespace cls
