// TEST INFRASTRUCTURE.  cls_place_sequences (classeq2_b200/csrc/record_writer.cpp) without a GPU: this file supplies a
// cls_place_batch that hands the batch to the C++ oracle (a function pointer set by the test; the "index" is the oracle's
// model handle), so that the glue of the one-call use-case - batch slicing, result arrays, the loop over batches, the
// file handling - runs end to end on the CPU.  tests/test_record_writer.py builds and drives it.
#include <cstdint>
#include <string>

#include "../../include/classeq_b200.h"

typedef void (*orc_place_fn)(const void *, const uint8_t *, const uint64_t *, uint64_t, int32_t, double, uint32_t, int, uint8_t *,
                             uint64_t *, int32_t *, int32_t *, uint32_t *, uint32_t *, uint32_t *, uint32_t *);
static orc_place_fn g_place = nullptr;
static thread_local std::string g_err;
static uint64_t g_calls = 0;

namespace cls {
int set_last_error(int code, const std::string &msg) { g_err = msg; return code; }
}

extern "C" {
void psh_set_placer(void *fn) { g_place = reinterpret_cast<orc_place_fn>(fn); }
uint64_t psh_calls(void) { return g_calls; }
const char *cls_last_error(void) { return g_err.c_str(); }
int cls_place_batch(cls_index *index, const cls_batch *b, const cls_params *p, cls_result *r) {
    if (!g_place || !index || !b || !p || !r) return cls::set_last_error(CLS_ERR_INVALID_ARGUMENT, "NULL argument");
    ++g_calls;
    g_place(reinterpret_cast<const void *>(index), b->bases, b->offsets, b->n_queries, p->max_iterations, p->min_match_coverage,
            p->remove_intersection, 4, r->status, r->node_id, r->one, r->rest, r->n_query_kmers, r->n_matched, r->n_root_matched, r->iterations);
    return CLS_OK;
}
}
