// TEST INFRASTRUCTURE - not part of the product.  Drives the data flow of the device model builder
// (classeq2_b200/csrc/build_kernels.cu) element by element on the CPU, with the very same per-element
// functions (build_steps.hpp) and host preparation (build_prep.hpp), std::stable_sort in place of the radix
// sorts and plain loops in place of the scans.  tests/test_build_steps.py holds its output equal to the host
// builder cls_model_build, so that the index arithmetic of the kernels is checked without a GPU; the kernels
// themselves are checked against cls_model_build on the GPU (tests/test_gpu_build.py).
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/classeq_b200.h"
#include "../../classeq2_b200/csrc/build_prep.hpp"
#include "../../classeq2_b200/csrc/build_steps.hpp"
#include "../../classeq2_b200/csrc/built_model.hpp"
#include "../../classeq2_b200/csrc/murmur3_host.hpp"

using namespace cls::build;

extern "C" int bsh_model_build(const cls_model_view *tree, uint64_t n_tips, const uint64_t *tip_node, const uint8_t *bases,
                               const uint64_t *offsets, cls_built_model **out) {
    *out = nullptr;
    Prep pr;
    std::string err;
    const int rc = prepare(tree, n_tips, tip_node, offsets, pr, err, bases);
    if (rc != CLS_OK) return rc;
    const uint32_t k = tree->k_size, m = tree->m_size;
    const uint64_t N = pr.occ_off[n_tips];
    auto bm = new cls_built_model();
    bm->k_size = k; bm->m_size = m;
    bm->set_off.push_back(0);
    *out = bm;
    if (N == 0) return CLS_OK;
    // 1. build_hash_kernel, one "CTA" per item
    std::vector<uint64_t> A(N), B(N);
    std::vector<uint32_t> R(N);
    std::vector<uint8_t> f, r;
    for (const HashItem &it : pr.items) {
        const uint64_t W = (uint64_t)pr.seq_len[it.rank] - k + 1, w0 = (uint64_t)it.tile * kTileWindows;
        const uint32_t nw = (uint32_t)std::min<uint64_t>(W - w0, kTileWindows), nb = nw + k - 1;
        const uint8_t *src = bases + pr.seq_off[it.rank] + w0;
        f.assign(nb, 0); r.assign(nb, 0);
        for (uint32_t i = 0; i < nb; ++i) { f[i] = src[i] & 0xDFu; r[nb - 1 - i] = comp_byte(src[i]); }
        const uint32_t mm = std::min(m, k);
        for (uint32_t x = 0; x < 2 * nw; ++x) {
            const TileItem ti = tile_item(x, nw, W, w0);
            const uint8_t *s = ti.strand ? r.data() : f.data();
            const uint64_t o = pr.occ_off[it.rank] + ti.occ;
            A[o] = cls::murmur3_x64_128_h1(s + ti.pos, k, 0);
            B[o] = m == 0 ? 0 : cls::murmur3_x64_128_h1(s + ti.pos, mm, 0);
            R[o] = it.rank;
        }
    }
    // 2. stable sorts by bucket key, then by hash
    std::vector<uint32_t> I0(N), I1;
    std::iota(I0.begin(), I0.end(), 0u);
    std::stable_sort(I0.begin(), I0.end(), [&](uint32_t a, uint32_t b) { return B[a] < B[b]; });
    I1 = I0;
    std::stable_sort(I1.begin(), I1.end(), [&](uint32_t a, uint32_t b) { return A[a] < A[b]; });
    std::vector<uint64_t> K(N), H1(N);
    std::vector<uint32_t> RS(N);
    for (uint64_t i = 0; i < N; ++i) { K[i] = A[I1[i]]; H1[i] = B[I1[i]]; RS[i] = R[I1[i]]; }
    // 3. flags, exclusive sum, scatter
    std::vector<uint64_t> fl(N), sc(N);
    for (uint64_t i = 0; i < N; ++i) fl[i] = occ_flags(K.data(), H1.data(), RS.data(), i);
    uint64_t run = 0;
    for (uint64_t i = 0; i < N; ++i) { sc[i] = run; run += fl[i]; }
    const uint32_t E = (uint32_t)run, T = (uint32_t)(run >> 32);
    bm->entry_hash.resize(E); bm->entry_bucket.resize(E); bm->entry_set.resize(E);
    std::vector<uint32_t> list_off((size_t)E + 1), tips(T);
    for (uint64_t i = 0; i < N; ++i) {
        const uint32_t e = (uint32_t)sc[i], t = (uint32_t)(sc[i] >> 32);
        if (fl[i] & 1ull) { bm->entry_hash[e] = K[i]; bm->entry_bucket[e] = H1[i]; list_off[e] = t; }
        if (fl[i] >> 32) tips[t] = RS[i];
    }
    list_off[E] = T;
    // 4. fingerprints, neighbours in fingerprint order, representatives, set numbers
    std::vector<uint64_t> fp(E);
    std::vector<uint32_t> eid(E);
    for (uint32_t e = 0; e < E; ++e) { fp[e] = list_fingerprint(tips.data() + list_off[e], list_off[e + 1] - list_off[e]); eid[e] = e; }
    std::stable_sort(eid.begin(), eid.end(), [&](uint32_t a, uint32_t b) { return fp[a] < fp[b]; });
    std::vector<uint32_t> head(E), rep(E), is_rep(E), set_idx(E);
    for (uint32_t p = 0; p < E; ++p) {
        bool same = false;
        if (p > 0 && fp[eid[p]] == fp[eid[p - 1]]) {
            const uint32_t a = eid[p], b = eid[p - 1];
            same = lists_equal(tips.data() + list_off[a], list_off[a + 1] - list_off[a], tips.data() + list_off[b], list_off[b + 1] - list_off[b]);
        }
        head[p] = same ? 0u : p;
    }
    for (uint32_t p = 1; p < E; ++p) head[p] = std::max(head[p], head[p - 1]);
    for (uint32_t p = 0; p < E; ++p) { const uint32_t e = eid[p], q = eid[head[p]]; rep[e] = q; is_rep[e] = e == q; }
    uint32_t S = 0;
    for (uint32_t e = 0; e < E; ++e) { set_idx[e] = S; S += is_rep[e]; }
    std::vector<uint32_t> set_rep(S);
    for (uint32_t e = 0; e < E; ++e) { bm->entry_set[e] = set_idx[rep[e]]; if (is_rep[e]) set_rep[set_idx[e]] = e; }
    // 5. node sets
    bm->set_off.assign((size_t)S + 1, 0);
    for (uint32_t s = 0; s < S; ++s) {
        const uint32_t e = set_rep[s];
        bm->set_off[s + 1] = bm->set_off[s] + set_size(pr.parent.data(), pr.depth.data(), pr.rank_node.data(), tips.data() + list_off[e], list_off[e + 1] - list_off[e]);
    }
    bm->set_node_ids.resize(bm->set_off[S]);
    for (uint32_t s = 0; s < S; ++s) {
        const uint32_t e = set_rep[s];
        set_fill(pr.parent.data(), pr.depth.data(), pr.rank_node.data(), tree->node_id, tips.data() + list_off[e], list_off[e + 1] - list_off[e],
                 bm->set_node_ids.data() + bm->set_off[s]);
    }
    return CLS_OK;
}
