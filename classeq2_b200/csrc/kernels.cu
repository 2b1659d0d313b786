// sm_100a kernels of the placement path.  One warp places one query end to end:
//
//   decode   2-bit packed bases -> forward + reverse-complement ASCII strings in shared memory
//            (reference: kmers_map.rs:375-398 build_kmer_from_string, :431-443 reverse_complement)
//   hash     murmur3_x64_128(window).0 for every window of both strands, one window per lane
//            (kmers_map.rs:405-424, :157-159)
//   probe    one 32-byte bucket (a single DRAM sector, one 256-bit load) of the open-addressed
//            table per window; bucket-key gating by 2-bit prefix code (kmers_map.rs:273-311, :55-70)
//   dedup    distinct-hash semantics of the reference's HashSets: hits de-duplicated by table
//            slot in a per-warp shared-memory set, then histogrammed by node-set record
//   descend  one-vs-rest walk from the root (place_sequence.rs:279-601,
//            update_introspection_node.rs:13-91) with per-warp shared-memory vote counters:
//            cnt(c)  = #hits whose node set contains child c
//            excl(c) = #hits whose node set contains c and no other non-leaf sibling
//            U       = #hits whose node set contains any non-leaf child of the current node
//            default mode: one = cnt(c), rest = U - excl(c); remove_intersection: one = excl(c),
//            rest = U - cnt(c); a single candidate -> (cnt(c), 0)   (place_sequence.rs:353-418)
//
// Integer/byte work bound by HBM sector rate and the integer pipes - no tensor cores.
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>

#include "device_types.hpp"
#include "kernels.hpp"
#include "murmur3_device.cuh"

namespace cls {

namespace {

constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// 4 two-bit codes (one byte) -> 4 ASCII letters via a byte permute on the "ACTG" LUT.
__device__ __forceinline__ uint32_t decode4(uint32_t b) {
    uint32_t sel = (b & 0x03u) | ((b & 0x0Cu) << 2) | ((b & 0x30u) << 4) | ((b & 0xC0u) << 6);
    return __byte_perm(kAsciiLut, 0u, sel);
}

// Reverse the order of the sixteen 2-bit bases of a word and complement them (code ^ 2).
__device__ __forceinline__ uint32_t revcomp16(uint32_t w) {
    uint32_t r = __brev(w);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return r ^ 0xAAAAAAAAu;
}

// Decode one read into two 4-byte aligned ASCII strings (forward, reverse complement).
__device__ __forceinline__ void decode_read(const uint32_t *__restrict__ packed, uint32_t len,
                                            uint32_t *str_f, uint32_t *str_r) {
    const uint32_t nw = (len + 15u) >> 4;
    const uint32_t pad = nw * 16u - len;  // unused base slots at the top of the last word
    for (uint32_t t = lane_id(); t < nw; t += 32) {
        uint32_t f = __ldg(packed + t);
        // reverse-complement word t = bases [16t, 16t+16) of the reversed string
        uint32_t a = revcomp16(__ldg(packed + (nw - 1 - t)));
        uint32_t b = (t + 1 < nw) ? revcomp16(__ldg(packed + (nw - 2 - t))) : 0u;
        uint32_t r = __funnelshift_r(a, b, 2u * pad);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            str_f[4 * t + q] = decode4((f >> (8 * q)) & 0xFFu);
            str_r[4 * t + q] = decode4((r >> (8 * q)) & 0xFFu);
        }
    }
}

__device__ __forceinline__ void ld_bucket(const Slot *table, uint64_t bucket, uint64_t &h0, uint64_t &m0,
                                          uint64_t &h1, uint64_t &m1) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(h0), "=l"(m0), "=l"(h1), "=l"(m1)
                 : "l"(table + 2 * bucket));
}

// 2-bit prefix code of the m_eff leading bases of a window (bucket gating).
__device__ __forceinline__ uint32_t prefix_code(const uint8_t *s, uint32_t pos, uint32_t m_eff) {
    uint32_t code = 0;
    for (uint32_t j = 0; j < m_eff; ++j) code |= ((uint32_t)(s[pos + j] >> 1) & 3u) << (2 * j);
    return code;
}

template <int K>
__device__ __forceinline__ uint64_t hash_window(const uint32_t *s32, uint32_t pos, uint32_t k) {
    if constexpr (K > 0) {
        return murmur_window_smem<K>(s32, pos);
    } else {
        return murmur_window_generic(reinterpret_cast<const uint8_t *>(s32), pos, k);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Debug/parity kernel: all window hashes of one read, reference order (forward then revcomp).
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void hash_only_kernel(const uint32_t *__restrict__ packed, uint32_t len, uint32_t k,
                                 uint64_t *__restrict__ out, uint32_t str_words) {
    extern __shared__ uint32_t smem[];
    uint32_t *str_f = smem;
    uint32_t *str_r = smem + str_words;
    for (uint32_t i = threadIdx.x; i < 2 * str_words; i += blockDim.x) smem[i] = 0;
    __syncthreads();
    if (threadIdx.x < 32) decode_read(packed, len, str_f, str_r);
    __syncthreads();
    const uint32_t W = len - k + 1;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * W; idx += gridDim.x * blockDim.x) {
        const bool rc = idx >= W;
        out[idx] = hash_window<K>(rc ? str_r : str_f, rc ? idx - W : idx, k);
    }
}

// ------------------------------------------------------------------------------------------
// The placement kernel: one warp per query, persistent CTAs striding over the query range.
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) place_kernel(DeviceIndex ix, PlaceParams pp,
                                                    const uint32_t *__restrict__ packed,
                                                    const ReadDesc *__restrict__ reads, uint32_t first_read,
                                                    uint32_t n_reads, ResultRec *__restrict__ results,
                                                    PlaceGeom g) {
    extern __shared__ uint32_t smem[];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t warps_per_cta = blockDim.x >> 5;
    uint32_t *wbase = smem + (size_t)warp * g.words_per_warp;
    uint32_t *str_f = wbase;
    uint32_t *str_r = str_f + g.str_words;
    uint32_t *t1 = str_r + g.str_words;  // dedup set keyed by table slot; later the (entry, weight) list
    uint32_t *t2k = t1 + g.t1_size;      // histogram keys: node-set record offsets
    uint32_t *t2c = t2k + g.t2_size;     // histogram counts
    uint32_t *cnt = t2c + g.t2_size;     // vote counters, one per non-leaf child ordinal
    uint32_t *excl = cnt + g.fan_cap;
    const uint32_t t1_shift = 32u - g.t1_log2, t2_shift = 32u - g.t2_log2;
    const uint32_t t1_mask = g.t1_size - 1u, t2_mask = g.t2_size - 1u;
    const uint32_t k = ix.k_size;
    const bool ri = pp.remove_intersection != 0;

    for (uint32_t o = lane; o < g.fan_cap; o += 32) { cnt[o] = 0; excl[o] = 0; }

    const uint32_t gwarp = blockIdx.x * warps_per_cta + warp;
    const uint32_t gstride = gridDim.x * warps_per_cta;
    for (uint32_t r = gwarp; r < n_reads; r += gstride) {
        const ReadDesc rd = reads[first_read + r];
        const uint32_t L = rd.len;
        const uint32_t W = L - k + 1;  // host guarantees L >= k
        const uint32_t nwin = 2 * W;

        // ---- reset per-read tables, decode ------------------------------------------------
        for (uint32_t i = lane; i < g.t1_size; i += 32) t1[i] = kEmpty;
        for (uint32_t i = lane; i < g.t2_size; i += 32) { t2k[i] = kEmpty; t2c[i] = 0; }
        decode_read(packed + rd.word_off, L, str_f, str_r);
        __syncwarp();

        // ---- hash + probe + dedup + histogram -------------------------------------------
        uint32_t n_matched = 0;
        for (uint32_t base = 0; base < nwin; base += 32) {
            const uint32_t idx = base + lane;
            bool fresh = false;
            if (idx < nwin) {
                const bool rc = idx >= W;
                const uint32_t pos = rc ? idx - W : idx;
                const uint32_t *s32 = rc ? str_r : str_f;
                const uint64_t h = hash_window<K>(s32, pos, k);
                uint64_t b = h & ix.bucket_mask;
                uint32_t slot_id = kEmpty, set_off = 0, code = 0;
                for (;;) {
                    uint64_t h0, m0, h1, m1;
                    ld_bucket(ix.table, b, h0, m0, h1, m1);
                    if (h0 == h && (uint32_t)m0 != kEmpty) { slot_id = (uint32_t)(2 * b); set_off = (uint32_t)m0; code = (uint32_t)(m0 >> 32); break; }
                    if (h1 == h && (uint32_t)m1 != kEmpty) { slot_id = (uint32_t)(2 * b + 1); set_off = (uint32_t)m1; code = (uint32_t)(m1 >> 32); break; }
                    if (!((uint32_t)(m0 >> 32) & kOverflowBit)) break;
                    b = (b + 1) & ix.bucket_mask;
                }
                if (slot_id != kEmpty) {
                    // bucket gating: the entry's bucket key must be among the query's prefix keys
                    const uint8_t *s8 = reinterpret_cast<const uint8_t *>(s32);
                    const uint32_t want = code & kCodeMask;
                    bool pass = prefix_code(s8, pos, ix.m_eff) == want;
                    if (!pass) {  // only possible for models whose bucket keys disagree with their k-mers
                        const uint8_t *f8 = reinterpret_cast<const uint8_t *>(str_f);
                        const uint8_t *r8 = reinterpret_cast<const uint8_t *>(str_r);
                        for (uint32_t p = 0; p < W && !pass; ++p)
                            pass = prefix_code(f8, p, ix.m_eff) == want || prefix_code(r8, p, ix.m_eff) == want;
                    }
                    if (pass) {
                        uint32_t p1 = (slot_id * 0x9E3779B1u) >> t1_shift;
                        for (;;) {
                            uint32_t old = atomicCAS(&t1[p1], kEmpty, slot_id);
                            if (old == kEmpty) { fresh = true; break; }
                            if (old == slot_id) break;
                            p1 = (p1 + 1) & t1_mask;
                        }
                        if (fresh) {
                            uint32_t p2 = (set_off * 0x9E3779B1u) >> t2_shift;
                            for (;;) {
                                uint32_t old = atomicCAS(&t2k[p2], kEmpty, set_off);
                                if (old == kEmpty || old == set_off) { atomicAdd(&t2c[p2], 1u); break; }
                                p2 = (p2 + 1) & t2_mask;
                            }
                        }
                    }
                }
            }
            n_matched += __popc(__ballot_sync(kFull, fresh));
        }
        __syncwarp();

        // ---- compact the histogram into a list of (current entry, weight) in t1 ---------
        uint32_t D = 0;
        for (uint32_t base = 0; base < g.t2_size; base += 32) {
            const uint32_t key = t2k[base + lane], c = t2c[base + lane];
            const bool occ = key != kEmpty;
            const uint32_t m = __ballot_sync(kFull, occ);
            if (occ) {
                const uint32_t j = D + __popc(m & ((1u << lane) - 1u));
                t1[2 * j] = key;
                t1[2 * j + 1] = c;
            }
            D += __popc(m);
        }
        __syncwarp();
        // restrict to sets that contain tree.root.id (M_r, place_sequence.rs:156-166)
        uint32_t n_root = 0;
        for (uint32_t j = lane; j < D; j += 32) {
            const uint32_t off = t1[2 * j];
            const SetWord hdr = ix.arena[off];
            if (hdr.x & kSetHasRoot) { n_root += t1[2 * j + 1]; t1[2 * j] = off + 1; }
            else t1[2 * j + 1] = 0;
        }
        n_root = __reduce_add_sync(kFull, n_root);
        __syncwarp();

        // ---- gates (place_sequence.rs:120-139, :156-166, :199-206, :231-254) ---------------
        ResultRec res;
        res.node_id = 0; res.one = 0; res.rest = 0;
        res.n_matched = n_matched; res.n_root_matched = n_root; res.iterations = 0;
        res.status = 0xFFFFFFFFu;
        if (n_matched == 0) res.status = CLS_DEV_UNCL_NO_MATCH;
        else if (n_root == 0) res.status = CLS_DEV_UNCL_NO_ROOT;
        else if (ix.root_children_none) res.status = CLS_DEV_ERR_ROOT_NO_CHILDREN;
        else {
            const double expected = round((double)n_matched * pp.min_match_coverage);
            if ((double)n_root < expected) res.status = CLS_DEV_UNCL_COVERAGE;
        }

        // ---- descent ---------------------------------------------------------------------------
        uint32_t p = 0;  // current parent (dense non-leaf id), root = 0
        uint32_t iteration = 0;
        while (res.status == 0xFFFFFFFFu) {
            iteration++;
            res.iterations = iteration;
            if ((int64_t)iteration > (int64_t)pp.max_iterations) { res.status = CLS_DEV_ERR_MAX_ITERATIONS; break; }
            const QNode qn = ix.qnodes[p];
            const uint32_t m = qn.child_count;

            // votes
            uint32_t u_local = 0;
            for (uint32_t j = lane; j < D; j += 32) {
                const uint32_t w = t1[2 * j + 1];
                if (w == 0) continue;
                const uint32_t cur = t1[2 * j];
                const uint32_t end = cur + ix.arena[cur].y;
                uint32_t npres = 0, last = 0;
                for (uint32_t c = cur + 1; c < end;) {
                    const SetWord e = ix.arena[c];
                    if (e.x & kPresentBit) { last = e.x & ~kPresentBit; atomicAdd(&cnt[last], w); npres++; }
                    c += e.y;
                }
                if (npres) u_local += w;
                if (npres == 1) atomicAdd(&excl[last], w);
            }
            const uint32_t U = __reduce_add_sync(kFull, u_local);
            __syncwarp();

            // one-vs-rest test over the candidates (children with a non-empty K(c))
            uint32_t ncand = 0;
            for (uint32_t o0 = 0; o0 < m; o0 += 32) {
                const uint32_t o = o0 + lane;
                ncand += __popc(__ballot_sync(kFull, o < m && cnt[o] > 0));
            }
            uint32_t nprop = 0, n_best = 0, best_ord = 0;
            int32_t best_diff = INT_MIN, best_one = 0, best_rest = 0;
            for (uint32_t o0 = 0; o0 < m; o0 += 32) {
                const uint32_t o = o0 + lane;
                const uint32_t c = o < m ? cnt[o] : 0u, x = o < m ? excl[o] : 0u;
                const int32_t one = (int32_t)((ri && ncand > 1) ? x : c);
                const int32_t rest = ncand > 1 ? (int32_t)(ri ? U - c : U - x) : 0;
                const bool prop = c > 0 && one > rest;
                const uint32_t pm = __ballot_sync(kFull, prop);
                if (pm) {
                    nprop += __popc(pm);
                    const int32_t diff = prop ? one - rest : INT_MIN;
                    const int32_t dmax = __reduce_max_sync(kFull, diff);
                    const uint32_t eq = __ballot_sync(kFull, prop && diff == dmax);
                    if (dmax > best_diff) {
                        const int src = __ffs(eq) - 1;
                        best_diff = dmax; n_best = __popc(eq);
                        best_ord = __shfl_sync(kFull, o, src);
                        best_one = __shfl_sync(kFull, one, src);
                        best_rest = __shfl_sync(kFull, rest, src);
                    } else if (dmax == best_diff) {
                        n_best += __popc(eq);
                    }
                }
            }
            __syncwarp();
            for (uint32_t o = lane; o < m; o += 32) { cnt[o] = 0; excl[o] = 0; }
            __syncwarp();

            if (nprop == 0) {
                if (iteration == 1) res.status = CLS_DEV_UNCL_NO_INTROSPECTION;
                else { res.status = CLS_DEV_MAX_RESOLUTION; res.node_id = ix.q_node_id[p]; }
                break;
            }
            if (n_best != 1) {  // several proposals tie on (one - rest): provably unreachable
                res.status = CLS_DEV_INCONCLUSIVE; res.node_id = ix.q_node_id[p];
                break;
            }
            const uint32_t cq = ix.q_child_list[qn.child_first + best_ord];
            if (ix.qnodes[cq].child_count == 0) {  // update_introspection_node.rs:32-87
                res.status = CLS_DEV_IDENTITY_FOUND; res.node_id = ix.q_node_id[cq];
                res.one = best_one; res.rest = best_rest;
                break;
            }
            p = cq;
            // every live set follows the winner (or drops out)
            for (uint32_t j = lane; j < D; j += 32) {
                if (t1[2 * j + 1] == 0) continue;
                const uint32_t cur = t1[2 * j];
                const uint32_t end = cur + ix.arena[cur].y;
                uint32_t next = 0;
                for (uint32_t c = cur + 1; c < end;) {
                    const SetWord e = ix.arena[c];
                    if ((e.x & ~kPresentBit) == best_ord) { next = c; break; }
                    c += e.y;
                }
                if (next) t1[2 * j] = next; else t1[2 * j + 1] = 0;
            }
            __syncwarp();
        }
        if (lane == 0) results[first_read + r] = res;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------
static inline uint32_t ceil_log2(uint32_t x) {
    uint32_t l = 0;
    while ((1u << l) < x) ++l;
    return l;
}

PlaceGeom make_place_geom(uint32_t max_len, uint32_t k, uint32_t max_fanout) {
    PlaceGeom g{};
    const uint32_t H = max_len >= k ? 2 * (max_len - k + 1) : 2;
    g.str_words = ((max_len + 15u) / 16u) * 4u + 4u;  // decoded in 16-base groups, + over-read pad
    g.t1_log2 = ceil_log2(H * 2 < 64 ? 64 : H * 2);
    g.t2_log2 = ceil_log2(H + 1 < 32 ? 32 : H + 1);
    g.t1_size = 1u << g.t1_log2;
    g.t2_size = 1u << g.t2_log2;
    g.fan_cap = max_fanout < 1 ? 1 : max_fanout;
    g.words_per_warp = 2 * g.str_words + g.t1_size + 2 * g.t2_size + 2 * g.fan_cap;
    g.words_per_warp = (g.words_per_warp + 3u) & ~3u;
    return g;
}

template <int K>
static cudaError_t launch_place_t(const DeviceIndex &ix, const PlaceParams &pp, const uint32_t *packed,
                                  const ReadDesc *reads, uint32_t first_read, uint32_t n_reads,
                                  ResultRec *results, const PlaceGeom &g, int sm_count, cudaStream_t stream) {
    const size_t per_warp = (size_t)g.words_per_warp * 4;
    int warps = 8;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    if (per_warp * warps > 227 * 1024) return cudaErrorInvalidConfiguration;
    const size_t smem = per_warp * warps;
    cudaError_t e = cudaFuncSetAttribute(place_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, place_kernel<K>, warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    uint32_t grid = (uint32_t)(sm_count * occ);
    const uint32_t need = (n_reads + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid == 0) return cudaSuccess;
    place_kernel<K><<<grid, warps * 32, smem, stream>>>(ix, pp, packed, reads, first_read, n_reads, results, g);
    return cudaGetLastError();
}

cudaError_t launch_place(const DeviceIndex &ix, const PlaceParams &pp, const uint32_t *packed,
                         const ReadDesc *reads, uint32_t first_read, uint32_t n_reads, ResultRec *results,
                         const PlaceGeom &g, int sm_count, cudaStream_t stream) {
    if (ix.k_size == 35) return launch_place_t<35>(ix, pp, packed, reads, first_read, n_reads, results, g, sm_count, stream);
    return launch_place_t<0>(ix, pp, packed, reads, first_read, n_reads, results, g, sm_count, stream);
}

cudaError_t launch_hash_only(const uint32_t *packed, uint32_t len, uint32_t k, uint64_t *out, cudaStream_t stream) {
    const uint32_t str_words = ((len + 15u) / 16u) * 4u + 4u + (k + 3) / 4;
    const size_t smem = (size_t)2 * str_words * 4;
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e;
    if (k == 35) {
        e = cudaFuncSetAttribute(hash_only_kernel<35>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        hash_only_kernel<35><<<1, 256, smem, stream>>>(packed, len, k, out, str_words);
    } else {
        e = cudaFuncSetAttribute(hash_only_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        hash_only_kernel<0><<<1, 256, smem, stream>>>(packed, len, k, out, str_words);
    }
    return cudaGetLastError();
}

}  // namespace cls
