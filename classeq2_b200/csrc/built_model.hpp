// The handle behind cls_model_build / cls_model_build_device (include/classeq_b200.h): the entry and
// node-set arrays of a built k-mer map, in the layout of cls_model_view.
#pragma once
#include <cstdint>
#include <vector>

struct cls_built_model {
    uint32_t k_size = 0, m_size = 0;
    std::vector<uint64_t> entry_bucket, entry_hash, entry_set, set_off, set_node_ids;
};
