// C ABI of the placement path (include/classeq_b200.h).  Host-side plumbing only: 2-bit packing
// and length ordering of the batch, pinned staging, stream/event management, and the scatter of
// the 32-byte device result records into the caller's arrays.  All arithmetic of the path runs
// in kernels.cu; there is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/classeq_b200.h"
#include "device_types.hpp"
#include "index_build.hpp"
#include "kernels.hpp"
#include "murmur3_host.hpp"

using namespace cls;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

}  // namespace
namespace cls {
int set_last_error(int code, const std::string &msg) { return fail(code, msg); }  // for host_api.cpp
}
namespace {

#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(CLS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---- device / pinned buffers that grow on demand --------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// One launch per length class: reads [first, first+count) of the length-ordered batch.
struct LengthClass {
    uint32_t first, count, max_len;
};

// Host-side view of a packed batch (buffers are owned by a Workspace or a resident batch).
struct PackedLayout {
    uint64_t n_queries = 0;      // caller's batch size
    uint32_t n_device = 0;       // queries that reach the device (valid, L >= k)
    uint64_t n_words = 0;        // 32-bit words of packed bases
    std::vector<uint32_t> perm;  // device order -> caller index
    std::vector<uint8_t> pre_status;  // caller index -> status decided on the host, or 0xFF
    std::vector<uint32_t> lens;       // caller index -> query length (for n_query_kmers)
    std::vector<LengthClass> classes;
};

struct Workspace {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    PinBuf h_words, h_descs, h_results;
    DevBuf d_words, d_descs, d_results;
    bool in_use = false;
};

}  // namespace

struct cls_index {
    int device = 0;
    int sm_count = 0;
    DevBuf d_table, d_arena, d_terms, d_qnodes, d_qchild, d_qid, d_qinfo, d_lca;
    DeviceIndex dix{};
    cls_index_info info{};
    std::mutex mu;
    std::vector<std::unique_ptr<Workspace>> pool;
    cls_timing timing{};
};

struct cls_resident_batch {
    int device = 0;
    PackedLayout lay;
    DevBuf d_words, d_descs, d_results;
    PinBuf h_results;
};

namespace {

// ---- 2-bit packing ---------------------------------------------------------------------------
// code = (ascii >> 1) & 3 -> A=0 C=1 T=2 G=3, case-insensitive.  Returns false on a non-ACGT byte.
inline bool pack8(uint64_t x, uint32_t &out16) {
    const uint64_t ones = 0x0101010101010101ULL, low7 = 0x7F7F7F7F7F7F7F7FULL;
    const uint64_t y = x & 0xDFDFDFDFDFDFDFDFULL;  // upper-case
    auto nonzero = [&](uint64_t z) { return (((z & low7) + low7) | z) & 0x8080808080808080ULL; };
    const uint64_t bad = nonzero(y ^ (ones * 0x41)) & nonzero(y ^ (ones * 0x43)) &
                         nonzero(y ^ (ones * 0x47)) & nonzero(y ^ (ones * 0x54));
    uint64_t c = (x >> 1) & 0x0303030303030303ULL;
    c = (c | (c >> 6)) & 0x000F000F000F000FULL;
    c = (c | (c >> 12)) & 0x000000FF000000FFULL;
    c = (c | (c >> 24)) & 0xFFFFULL;
    out16 = (uint32_t)c;
    return bad == 0;
}

inline bool pack_read(const uint8_t *s, uint32_t len, uint32_t *dst) {
    bool ok = true;
    uint32_t i = 0, w = 0;
    for (; i + 16 <= len; i += 16, ++w) {
        uint64_t a, b;
        std::memcpy(&a, s + i, 8);
        std::memcpy(&b, s + i + 8, 8);
        uint32_t lo, hi;
        ok &= pack8(a, lo);
        ok &= pack8(b, hi);
        dst[w] = lo | (hi << 16);
    }
    if (i < len) {
        uint32_t v = 0;
        for (uint32_t j = 0; i + j < len; ++j) {
            const uint8_t ch = s[i + j], u = ch & 0xDF;
            ok &= (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
            v |= ((uint32_t)(ch >> 1) & 3u) << (2 * j);
        }
        dst[w] = v;
    }
    return ok;
}

int host_threads() {
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    return (int)std::min(hc, 32u);
}

template <class F>
void parallel_for(uint64_t n, uint64_t grain, F f) {
    int nt = (int)std::min<uint64_t>((uint64_t)host_threads(), (n + grain - 1) / std::max<uint64_t>(grain, 1));
    if (nt <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const uint64_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        uint64_t a = t * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=] { f(a, b); });
    }
    for (auto &x : th) x.join();
}

// Decide host-side statuses, order the surviving queries by decreasing length (so that every
// launch works on one length class and long reads start first), and compute the packed layout.
int plan_batch(const cls_batch *batch, uint32_t k, uint32_t max_fanout, PackedLayout &lay,
               std::vector<uint32_t> &word_off) {
    const uint64_t n = batch->n_queries;
    if (n >= 0xFFFFFFFFull) return fail(CLS_ERR_INVALID_ARGUMENT, "more than 2^32-1 queries in one batch");
    if (n && (!batch->offsets || (batch->offsets[n] && !batch->bases)))
        return fail(CLS_ERR_INVALID_ARGUMENT, "batch arrays are NULL");
    lay = PackedLayout();
    lay.n_queries = n;
    lay.pre_status.assign(n, 0xFF);
    lay.lens.resize(n);
    uint64_t max_len = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (batch->offsets[i] > batch->offsets[i + 1])
            return fail(CLS_ERR_INVALID_ARGUMENT, "batch offsets are not non-decreasing");
        const uint64_t len = batch->offsets[i + 1] - batch->offsets[i];
        if (len >= (1ull << 31)) return fail(CLS_ERR_INVALID_ARGUMENT, "query longer than 2^31 bases");
        lay.lens[i] = (uint32_t)len;
        if (len < k) lay.pre_status[i] = CLS_STATUS_ERR_TOO_SHORT;  // kmers_map.rs:383-385, place_sequence.rs:98-102
        else max_len = std::max(max_len, len);
    }
    // counting sort by length, longest first
    std::vector<uint32_t> count(max_len + 2, 0);
    for (uint64_t i = 0; i < n; ++i)
        if (lay.pre_status[i] == 0xFF) count[batch->offsets[i + 1] - batch->offsets[i]]++;
    std::vector<uint32_t> start(max_len + 2, 0);
    uint32_t acc = 0;
    for (uint64_t len = max_len + 1; len-- > 0;) { start[len] = acc; acc += count[len]; }
    lay.n_device = acc;
    lay.perm.resize(acc);
    for (uint64_t i = 0; i < n; ++i)
        if (lay.pre_status[i] == 0xFF) lay.perm[start[batch->offsets[i + 1] - batch->offsets[i]]++] = (uint32_t)i;
    // word offsets + classes
    word_off.resize((size_t)acc + 1);
    uint64_t w = 0;
    PlaceGeom cur{};
    for (uint32_t j = 0; j < acc; ++j) {
        const uint64_t i = lay.perm[j];
        const uint32_t len = (uint32_t)(batch->offsets[i + 1] - batch->offsets[i]);
        if (w >= 0xFFFFFFFFull) return fail(CLS_ERR_INVALID_ARGUMENT, "batch exceeds 2^32 packed words; split it");
        word_off[j] = (uint32_t)w;
        w += (len + 15) / 16;
        PlaceGeom g = make_place_geom(len, k, max_fanout);
        if (lay.classes.empty() || g.t1_size != cur.t1_size) {
            // lengths are non-increasing: a new class starts when the table size drops; the class
            // geometry is that of its first (longest) read
            lay.classes.push_back(LengthClass{j, 0, len});
            cur = g;
        }
        lay.classes.back().count++;
    }
    word_off[acc] = (uint32_t)std::min<uint64_t>(w, 0xFFFFFFFFull);
    lay.n_words = w;
    return CLS_OK;
}

// Pack the planned batch into `words`/`descs` (pinned).  Invalid bases demote the query to
// CLS_STATUS_ERR_INVALID_BASE; it still occupies its device slot (the device result is ignored).
void pack_batch(const cls_batch *batch, PackedLayout &lay, const std::vector<uint32_t> &word_off,
                uint32_t *words, ReadDesc *descs) {
    parallel_for(lay.n_device, 4096, [&](uint64_t a, uint64_t b) {
        for (uint64_t j = a; j < b; ++j) {
            const uint64_t i = lay.perm[j];
            const uint32_t len = (uint32_t)(batch->offsets[i + 1] - batch->offsets[i]);
            descs[j] = ReadDesc{word_off[j], len};
            if (!pack_read(batch->bases + batch->offsets[i], len, words + word_off[j]))
                lay.pre_status[i] = CLS_STATUS_ERR_INVALID_BASE;
        }
    });
}

int launch_classes(cls_index *ix, const PackedLayout &lay, const cls_params *params, const uint32_t *d_words,
                   const ReadDesc *d_descs, ResultRec *d_results, cudaStream_t stream, uint64_t *launches) {
    PlaceParams pp;
    pp.max_iterations = params->max_iterations;
    pp.remove_intersection = params->remove_intersection ? 1u : 0u;
    double cov = params->min_match_coverage;  // place_sequence.rs:67-75
    if (std::isnan(cov)) cov = 0.0;           // NaN survives the clamp and `NaN as usize` is 0
    else if (cov > 1.0) cov = 1.0;
    else if (cov < 0.0) cov = 0.0;
    pp.min_match_coverage = cov;
    for (const LengthClass &c : lay.classes) {
        PlaceGeom g = make_place_geom(c.max_len, ix->dix.k_size, ix->dix.max_fanout);
        cudaError_t e = launch_place(ix->dix, pp, d_words, d_descs, c.first, c.count, d_results, g, ix->sm_count, stream);
        if (e == cudaErrorInvalidConfiguration)
            return fail(CLS_ERR_UNSUPPORTED, "query too long (or tree fan-out too large) for the per-warp shared-memory tables");
        if (e != cudaSuccess) return fail(CLS_ERR_CUDA, std::string("place kernel launch: ") + cudaGetErrorString(e));
        if (launches) (*launches)++;
    }
    return CLS_OK;
}

void scatter_results(const PackedLayout &lay, uint32_t k, const ResultRec *recs, cls_result *out) {
    const uint64_t n = lay.n_queries;
    // host-decided statuses first
    parallel_for(n, 65536, [&](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; ++i) {
            const uint64_t len = lay.lens[i];
            const uint8_t ps = lay.pre_status[i];
            if (out->n_query_kmers) out->n_query_kmers[i] = len >= k ? (uint32_t)(2 * (len - k + 1)) : 0u;
            if (ps != 0xFF) {
                if (out->status) out->status[i] = ps;
                if (out->node_id) out->node_id[i] = 0;
                if (out->one) out->one[i] = 0;
                if (out->rest) out->rest[i] = 0;
                if (out->n_matched) out->n_matched[i] = 0;
                if (out->n_root_matched) out->n_root_matched[i] = 0;
                if (out->iterations) out->iterations[i] = 0;
            }
        }
    });
    parallel_for(lay.n_device, 65536, [&](uint64_t a, uint64_t b) {
        for (uint64_t j = a; j < b; ++j) {
            const uint64_t i = lay.perm[j];
            if (lay.pre_status[i] != 0xFF) continue;
            const ResultRec &r = recs[j];
            if (out->status) out->status[i] = (uint8_t)r.status;
            if (out->node_id) out->node_id[i] = r.node_id;
            if (out->one) out->one[i] = r.one;
            if (out->rest) out->rest[i] = r.rest;
            if (out->n_matched) out->n_matched[i] = r.n_matched;
            if (out->n_root_matched) out->n_root_matched[i] = r.n_root_matched;
            if (out->iterations) out->iterations[i] = r.iterations;
        }
    });
}

Workspace *acquire_ws(cls_index *ix) {
    std::lock_guard<std::mutex> lk(ix->mu);
    for (auto &w : ix->pool)
        if (!w->in_use) { w->in_use = true; return w.get(); }
    auto w = std::make_unique<Workspace>();
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto &e : w->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    w->in_use = true;
    ix->pool.push_back(std::move(w));
    return ix->pool.back().get();
}

void release_ws(cls_index *ix, Workspace *w) {
    std::lock_guard<std::mutex> lk(ix->mu);
    w->in_use = false;
}

struct WsGuard {
    cls_index *ix;
    Workspace *w;
    ~WsGuard() { if (w) release_ws(ix, w); }
};

}  // namespace

extern "C" {

int cls_abi_version(void) { return CLS_ABI_VERSION; }

const char *cls_last_error(void) { return g_last_error.c_str(); }

void cls_params_default(cls_params *p) {
    if (!p) return;
    p->max_iterations = 1000;      // place_sequence.rs:65
    p->remove_intersection = 0;    // place_sequence.rs:64
    p->min_match_coverage = 0.7;   // place_sequence.rs:74
}

int cls_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(CLS_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

int cls_index_create(const cls_model_view *model, int device, cls_index **out) {
    if (!out) return fail(CLS_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    HostIndex h;
    std::string err;
    int rc = build_host_index(model, h, err);
    if (rc != CLS_OK) return fail(rc, err);
    if (h.n_buckets > (1ull << 30)) return fail(CLS_ERR_UNSUPPORTED, "k-mer table larger than 2^30 buckets");

    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(CLS_ERR_CUDA, "this library only carries sm_100a code; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));

    auto ix = std::make_unique<cls_index>();
    ix->device = device;
    ix->sm_count = prop.multiProcessorCount;
    auto up = [&](DevBuf &b, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = b.reserve(std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    CU_TRY(up(ix->d_table, h.table.data(), h.table.size() * sizeof(Slot)));
    CU_TRY(up(ix->d_arena, h.arena.data(), h.arena.size() * sizeof(SetWord)));
    CU_TRY(up(ix->d_qnodes, h.qnodes.data(), h.qnodes.size() * sizeof(QNode)));
    CU_TRY(up(ix->d_qchild, h.q_child_list.data(), h.q_child_list.size() * sizeof(uint32_t)));
    CU_TRY(up(ix->d_qid, h.q_node_id.data(), h.q_node_id.size() * sizeof(uint64_t)));
    CU_TRY(up(ix->d_terms, h.terms.data(), h.terms.size() * sizeof(uint32_t)));
    CU_TRY(up(ix->d_qinfo, h.qinfo.data(), h.qinfo.size() * sizeof(QInfo)));
    CU_TRY(up(ix->d_lca, h.lca_table.data(), h.lca_table.size() * sizeof(uint64_t)));
    ix->dix.table = (const Slot *)ix->d_table.p;
    ix->dix.bucket_mask = h.n_buckets - 1;
    ix->dix.arena = (const SetWord *)ix->d_arena.p;
    ix->dix.qnodes = (const QNode *)ix->d_qnodes.p;
    ix->dix.q_child_list = (const uint32_t *)ix->d_qchild.p;
    ix->dix.q_node_id = (const uint64_t *)ix->d_qid.p;
    ix->dix.terms = (const uint32_t *)ix->d_terms.p;
    ix->dix.qinfo = (const QInfo *)ix->d_qinfo.p;
    ix->dix.lca_table = (const uint64_t *)ix->d_lca.p;
    ix->dix.n_q = (uint32_t)h.qnodes.size();
    ix->dix.euler_len = h.euler_len;
    ix->dix.closed = h.closed ? 1u : 0u;
    ix->dix.k_size = h.k_size;
    ix->dix.m_eff = h.m_eff;
    ix->dix.max_fanout = h.max_fanout;
    ix->dix.root_children_none = h.root_children_none ? 1u : 0u;
    ix->info.k_size = h.k_size;
    ix->info.m_size = h.m_size;
    ix->info.n_entries = h.n_entries_kept;
    ix->info.n_buckets = h.n_buckets;
    ix->info.table_bytes = h.table.size() * sizeof(Slot);
    ix->info.n_distinct_sets = h.n_distinct_sets;
    ix->info.set_arena_bytes = h.closed ? h.terms.size() * sizeof(uint32_t) : h.arena.size() * sizeof(SetWord);
    ix->info.closed_sets = h.closed ? 1u : 0u;
    ix->info.n_nonleaf_nodes = h.qnodes.size();
    ix->info.max_nonleaf_fanout = h.max_fanout;
    ix->info.device = device;
    *out = ix.release();
    return CLS_OK;
}

void cls_index_destroy(cls_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    for (auto &w : ix->pool) {
        if (w->stream) { cudaStreamSynchronize(w->stream); cudaStreamDestroy(w->stream); }
        for (auto &e : w->ev) if (e) cudaEventDestroy(e);
        w->h_words.release(); w->h_descs.release(); w->h_results.release();
        w->d_words.release(); w->d_descs.release(); w->d_results.release();
    }
    ix->d_table.release(); ix->d_arena.release(); ix->d_qnodes.release(); ix->d_qchild.release(); ix->d_qid.release();
    ix->d_terms.release(); ix->d_qinfo.release(); ix->d_lca.release();
    delete ix;
}

int cls_index_get_info(const cls_index *ix, cls_index_info *info) {
    if (!ix || !info) return fail(CLS_ERR_INVALID_ARGUMENT, "NULL argument");
    *info = ix->info;
    return CLS_OK;
}

int cls_place_batch(cls_index *ix, const cls_batch *batch, const cls_params *params, cls_result *result) {
    if (!ix || !batch || !params || !result) return fail(CLS_ERR_INVALID_ARGUMENT, "NULL argument");
    const double t0 = now_ms();
    CU_TRY(cudaSetDevice(ix->device));
    PackedLayout lay;
    std::vector<uint32_t> word_off;
    int rc = plan_batch(batch, ix->dix.k_size, ix->dix.max_fanout, lay, word_off);
    if (rc != CLS_OK) return rc;
    Workspace *w = acquire_ws(ix);
    if (!w) return fail(CLS_ERR_CUDA, "could not create a stream/workspace");
    WsGuard guard{ix, w};
    cls_timing tm{};
    const size_t words_b = (size_t)lay.n_words * 4, descs_b = (size_t)lay.n_device * sizeof(ReadDesc),
                 res_b = (size_t)lay.n_device * sizeof(ResultRec);
    if (lay.n_device) {
        CU_TRY(w->h_words.reserve(words_b + 16)); CU_TRY(w->h_descs.reserve(descs_b)); CU_TRY(w->h_results.reserve(res_b));
        CU_TRY(w->d_words.reserve(words_b + 16)); CU_TRY(w->d_descs.reserve(descs_b)); CU_TRY(w->d_results.reserve(res_b));
        pack_batch(batch, lay, word_off, (uint32_t *)w->h_words.p, (ReadDesc *)w->h_descs.p);
        tm.pack_ms = now_ms() - t0;
        CU_TRY(cudaEventRecord(w->ev[0], w->stream));
        CU_TRY(cudaMemcpyAsync(w->d_words.p, w->h_words.p, words_b, cudaMemcpyHostToDevice, w->stream));
        CU_TRY(cudaMemcpyAsync(w->d_descs.p, w->h_descs.p, descs_b, cudaMemcpyHostToDevice, w->stream));
        CU_TRY(cudaEventRecord(w->ev[1], w->stream));
        rc = launch_classes(ix, lay, params, (const uint32_t *)w->d_words.p, (const ReadDesc *)w->d_descs.p,
                            (ResultRec *)w->d_results.p, w->stream, &tm.kernel_launches);
        if (rc != CLS_OK) { cudaStreamSynchronize(w->stream); return rc; }
        CU_TRY(cudaEventRecord(w->ev[2], w->stream));
        CU_TRY(cudaMemcpyAsync(w->h_results.p, w->d_results.p, res_b, cudaMemcpyDeviceToHost, w->stream));
        CU_TRY(cudaEventRecord(w->ev[3], w->stream));
        CU_TRY(cudaStreamSynchronize(w->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, w->ev[0], w->ev[1]); tm.h2d_ms = ms;
        cudaEventElapsedTime(&ms, w->ev[1], w->ev[2]); tm.kernel_ms = ms;
        cudaEventElapsedTime(&ms, w->ev[2], w->ev[3]); tm.d2h_ms = ms;
    }
    scatter_results(lay, ix->dix.k_size, (const ResultRec *)w->h_results.p, result);
    tm.total_ms = now_ms() - t0;
    { std::lock_guard<std::mutex> lk(ix->mu); ix->timing = tm; }
    return CLS_OK;
}

int cls_batch_upload(cls_index *ix, const cls_batch *batch, cls_resident_batch **out) {
    if (!ix || !batch || !out) return fail(CLS_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    CU_TRY(cudaSetDevice(ix->device));
    auto rb = std::make_unique<cls_resident_batch>();
    rb->device = ix->device;
    std::vector<uint32_t> word_off;
    int rc = plan_batch(batch, ix->dix.k_size, ix->dix.max_fanout, rb->lay, word_off);
    if (rc != CLS_OK) return rc;
    const PackedLayout &lay = rb->lay;
    const size_t words_b = (size_t)lay.n_words * 4, descs_b = (size_t)lay.n_device * sizeof(ReadDesc),
                 res_b = (size_t)lay.n_device * sizeof(ResultRec);
    std::vector<uint32_t> words(lay.n_words + 4);
    std::vector<ReadDesc> descs(lay.n_device);
    pack_batch(batch, rb->lay, word_off, words.data(), descs.data());
    CU_TRY(rb->d_words.reserve(words_b + 16)); CU_TRY(rb->d_descs.reserve(descs_b + 16)); CU_TRY(rb->d_results.reserve(res_b + 32));
    CU_TRY(rb->h_results.reserve(res_b + 32));
    CU_TRY(cudaMemcpy(rb->d_words.p, words.data(), words_b, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(rb->d_descs.p, descs.data(), descs_b, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemset(rb->d_results.p, 0xFF, res_b));
    *out = rb.release();
    return CLS_OK;
}

int cls_place_resident(cls_index *ix, cls_resident_batch *rb, const cls_params *params, void *stream) {
    if (!ix || !rb || !params) return fail(CLS_ERR_INVALID_ARGUMENT, "NULL argument");
    if (rb->device != ix->device) return fail(CLS_ERR_INVALID_ARGUMENT, "resident batch lives on another device");
    CU_TRY(cudaSetDevice(ix->device));
    uint64_t launches = 0;
    int rc = launch_classes(ix, rb->lay, params, (const uint32_t *)rb->d_words.p, (const ReadDesc *)rb->d_descs.p,
                            (ResultRec *)rb->d_results.p, (cudaStream_t)stream, &launches);
    if (rc == CLS_OK) { std::lock_guard<std::mutex> lk(ix->mu); ix->timing.kernel_launches = launches; }
    return rc;
}

int cls_resident_fetch(cls_index *ix, cls_resident_batch *rb, void *stream, cls_result *result) {
    if (!ix || !rb || !result) return fail(CLS_ERR_INVALID_ARGUMENT, "NULL argument");
    CU_TRY(cudaSetDevice(ix->device));
    const size_t res_b = (size_t)rb->lay.n_device * sizeof(ResultRec);
    if (res_b) {
        CU_TRY(cudaMemcpyAsync(rb->h_results.p, rb->d_results.p, res_b, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    }
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    scatter_results(rb->lay, ix->dix.k_size, (const ResultRec *)rb->h_results.p, result);
    return CLS_OK;
}

void cls_resident_destroy(cls_resident_batch *rb) {
    if (!rb) return;
    cudaSetDevice(rb->device);
    rb->d_words.release(); rb->d_descs.release(); rb->d_results.release(); rb->h_results.release();
    delete rb;
}

uint64_t cls_resident_bytes(const cls_resident_batch *rb) {
    if (!rb) return 0;
    return (uint64_t)rb->lay.n_words * 4 + (uint64_t)rb->lay.n_device * (sizeof(ReadDesc) + sizeof(ResultRec));
}

int cls_get_timing(const cls_index *ix, cls_timing *out) {
    if (!ix || !out) return fail(CLS_ERR_INVALID_ARGUMENT, "NULL argument");
    std::lock_guard<std::mutex> lk(const_cast<cls_index *>(ix)->mu);
    *out = ix->timing;
    return CLS_OK;
}

int cls_debug_kmer_hashes(int device, uint32_t k_size, const uint8_t *bases, uint64_t len, uint64_t *out_hashes,
                          uint64_t cap, uint64_t *n_out) {
    if (!n_out || (len && !bases) || k_size == 0) return fail(CLS_ERR_INVALID_ARGUMENT, "NULL argument or k == 0");
    *n_out = 0;
    if (len < k_size) return CLS_OK;  // kmers_map.rs:383-385
    if (len >= (1ull << 20)) return fail(CLS_ERR_UNSUPPORTED, "debug export is limited to 2^20 bases");
    CU_TRY(cudaSetDevice(device));
    const uint64_t n = 2 * (len - k_size + 1);
    std::vector<uint32_t> words((len + 15) / 16 + 2, 0);
    if (!pack_read(bases, (uint32_t)len, words.data())) return fail(CLS_ERR_INVALID_ARGUMENT, "non-ACGT byte in the sequence");
    DevBuf d_words, d_out;
    struct Free { DevBuf &a, &b; ~Free() { a.release(); b.release(); } } guard{d_words, d_out};
    CU_TRY(d_words.reserve(words.size() * 4));
    CU_TRY(d_out.reserve(n * 8));
    CU_TRY(cudaMemcpy(d_words.p, words.data(), words.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(launch_hash_only((const uint32_t *)d_words.p, (uint32_t)len, k_size, (uint64_t *)d_out.p, nullptr));
    CU_TRY(cudaDeviceSynchronize());
    std::vector<uint64_t> h(n);
    CU_TRY(cudaMemcpy(h.data(), d_out.p, n * 8, cudaMemcpyDeviceToHost));
    *n_out = n;
    if (out_hashes) std::memcpy(out_hashes, h.data(), std::min<uint64_t>(n, cap) * 8);
    return CLS_OK;
}

uint64_t cls_debug_host_murmur3_x64_128_h1(const uint8_t *data, uint64_t len, uint64_t seed) {
    return murmur3_x64_128_h1(data, len, seed);
}

// sequence.rs:47-56: `sequence.to_uppercase().chars().filter(A|C|G|T)`.  Rust upper-cases with the
// full Unicode mapping; the only non-ASCII scalars whose upper-case expansion contains an ASCII
// A/C/G/T are U+1E97 (t with diaeresis -> "T" + U+0308), U+1E9A (a with right half ring -> "A" +
// U+02BE), U+FB05 and U+FB06 (long-s-t / st ligatures -> "ST").  Every other non-ASCII byte
// sequence contributes nothing.
uint64_t cls_filter_sequence(const uint8_t *line, uint64_t len, uint8_t *out, uint64_t cap) {
    uint64_t n = 0;
    auto put = [&](uint8_t c) { if (n < cap && out) out[n] = c; ++n; };
    for (uint64_t i = 0; i < len; ++i) {
        const uint8_t c = line[i];
        if (c < 0x80) {
            const uint8_t u = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;
            if (u == 'A' || u == 'C' || u == 'G' || u == 'T') put(u);
        } else if (c == 0xE1 && i + 2 < len && line[i + 1] == 0xBA && (line[i + 2] == 0x97 || line[i + 2] == 0x9A)) {
            put(line[i + 2] == 0x97 ? 'T' : 'A');
            i += 2;
        } else if (c == 0xEF && i + 2 < len && line[i + 1] == 0xAC && (line[i + 2] == 0x85 || line[i + 2] == 0x86)) {
            put('T');
            i += 2;
        }
    }
    return n;
}

}  // extern "C"
