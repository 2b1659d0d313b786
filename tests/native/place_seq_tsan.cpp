// TEST INFRASTRUCTURE.  The loop of cls_place_sequences (classeq2_b200/csrc/record_writer.cpp) - batch i + 1 is placed while
// the records of batch i are rendered and written on a second thread - under ThreadSanitizer, without a GPU: the
// cls_place_batch supplied here fills the result arrays from the query bytes (and keeps the host pool busy, as the real
// one's packing does).  usage: place_seq_tsan queries.fasta out_path     (CLS_SEQ_BATCH sets the batch size)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/classeq_b200.h"
#include "../../classeq2_b200/csrc/host_pool.hpp"

static thread_local std::string g_err;
namespace cls {
int set_last_error(int code, const std::string &msg) { g_err = msg; return code; }
}
extern "C" const char *cls_last_error(void) { return g_err.c_str(); }

extern "C" int cls_place_batch(cls_index *, const cls_batch *b, const cls_params *, cls_result *r) {
    // a deterministic "placement" of every query from its bases: every status of the writer's table comes up
    cls::parallel_for(b->n_queries, 16, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint64_t o0 = b->offsets[i], o1 = b->offsets[i + 1];
            uint64_t h = 1469598103934665603ull;
            for (uint64_t j = o0; j < o1; ++j) h = (h ^ b->bases[j]) * 1099511628211ull;
            static const uint8_t kinds[6] = {CLS_STATUS_IDENTITY_FOUND, CLS_STATUS_MAX_RESOLUTION, CLS_STATUS_UNCL_NO_MATCH,
                                             CLS_STATUS_UNCL_COVERAGE, CLS_STATUS_ERR_TOO_SHORT, CLS_STATUS_UNCL_NO_ROOT};
            r->status[i] = kinds[h % 6];
            r->node_id[i] = r->status[i] == CLS_STATUS_IDENTITY_FOUND ? (h >> 8) % 2 ? 1 : 4 : (r->status[i] == CLS_STATUS_MAX_RESOLUTION ? 0 : 0);
            r->one[i] = (int32_t)((h >> 16) % 300); r->rest[i] = (int32_t)((h >> 24) % 50);
            r->n_query_kmers[i] = (uint32_t)(o1 - o0); r->n_matched[i] = (uint32_t)((h >> 32) % 200);
            r->n_root_matched[i] = (uint32_t)((h >> 40) % 100); r->iterations[i] = (uint32_t)((h >> 48) % 9);
        }
    });
    return CLS_OK;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    static const uint64_t node_id[7] = {0, 1, 2, 3, 4, 5, 6};
    static const uint8_t node_kind[7] = {CLS_KIND_ROOT, CLS_KIND_NODE, CLS_KIND_LEAF, CLS_KIND_LEAF, CLS_KIND_NODE, CLS_KIND_LEAF, CLS_KIND_LEAF};
    static const uint64_t child_off[8] = {0, 2, 4, 4, 4, 6, 6, 6};
    static const uint64_t child_idx[6] = {1, 4, 2, 3, 5, 6};
    static const int64_t parent_id[7] = {-1, 0, 1, 1, 0, 4, 4};
    static const uint8_t children_some[7] = {1, 1, 0, 0, 1, 0, 0}, has_name[7] = {0, 0, 1, 1, 0, 1, 1};
    double support[7], length[7];
    for (int i = 0; i < 7; ++i) { support[i] = node_kind[i] == CLS_KIND_NODE ? 100.0 : NAN; length[i] = i ? 0.01 : 0.0; }
    static const char names[] = "tip_atip_btip_ctip_d";
    static const uint64_t name_off[8] = {0, 0, 0, 5, 10, 10, 15, 20};
    cls_record_tree rt;
    std::memset(&rt, 0, sizeof rt);
    rt.n_nodes = 7; rt.node_id = node_id; rt.parent_id = parent_id; rt.node_kind = node_kind; rt.children_some = children_some;
    rt.support = support; rt.length = length; rt.has_name = has_name; rt.name_off = name_off; rt.names = names;
    rt.child_off = child_off; rt.child_idx = child_idx;
    cls_params params;
    std::memset(&params, 0, sizeof params);
    params.max_iterations = 1000; params.min_match_coverage = 0.7;
    uint64_t n = 0;
    int dummy = 0;
    const int rc = cls_place_sequences(reinterpret_cast<cls_index *>(&dummy), &rt, argv[1], argv[2], &params, 0, 1, &n);
    if (rc != CLS_OK) { std::fprintf(stderr, "rc=%d %s\n", rc, cls_last_error()); return 1; }
    std::printf("n=%llu\n", (unsigned long long)n);
    return 0;
}
