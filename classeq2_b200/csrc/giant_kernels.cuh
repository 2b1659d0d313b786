// Reads whose per-read tables do not fit the shared memory of an SM (beyond about 4 kb at k = 35): the same
// algorithm as place_kernel<K, CLOSED, true> with the tables in GLOBAL memory, so that no query length makes a call
// fail (the reference has no limit: kmers_map.rs:375-424, and isolates failures per query: place_sequences/mod.rs:160-169).
// One arena of make_place_geom(...).words_per_warp words per read, laid out exactly like the shared-memory group of
// place_kernel: [t1 | t2k | t2c | lst | str_f | str_r | pk_f | pk_r | cnt | excl | n_sets, n_matched].  Three launches
// per wave of reads (as many reads as arenas fit the scratch):
//   giant_prepare_kernel   blocks (x, read): reset the tables, decode the read into its arena
//   giant_scan_kernel      blocks (x, read): every warp hashes / probes / de-duplicates a contiguous chunk range of a
//                          strand into the read's tables (atomics on global memory)
//   giant_finish_kernel    one warp per read: gates and descent (finish_read)
// Rare by construction (amplicon and read data are far shorter); built for correctness, not for the roofline.
// Included by kernels.cu inside namespace cls { namespace { ... } }.
#pragma once

struct GiantLayout {
    uint32_t *t1, *t2k, *t2c, *lst, *cnt, *excl, *n_sets, *n_matched;
    WarpMem wm;   // str_f, str_r, pk_f, pk_r (the rings stay in shared memory, per warp)
};
__device__ __forceinline__ GiantLayout giant_carve(uint32_t *arena, const PlaceGeom &g) {
    GiantLayout l;
    l.t1 = arena;
    l.t2k = l.t1 + g.t1_size;
    l.t2c = l.t2k + g.t2_size;
    l.lst = l.t2c + g.t2_size;
    l.wm.str_f = l.lst + g.t2_size;
    l.wm.str_r = l.wm.str_f + g.str_words;
    l.wm.pk_f = l.wm.str_r + g.str_words;
    l.wm.pk_r = l.wm.pk_f + g.pk_words;
    l.cnt = l.wm.pk_r + g.pk_words;
    l.excl = l.cnt + g.fan_cap;
    l.n_sets = l.excl + g.fan_cap;
    l.n_matched = l.n_sets + 1;
    l.wm.ring_a = nullptr; l.wm.ring_b = nullptr;
    return l;
}

__global__ void __launch_bounds__(256) giant_prepare_kernel(const uint32_t *__restrict__ packed, const ReadDesc *__restrict__ reads,
                                                            uint32_t first_read, PlaceGeom g, uint32_t *__restrict__ arenas) {
    const uint32_t r = blockIdx.y;
    const ReadDesc rd = reads[first_read + r];
    uint32_t *arena = arenas + (size_t)r * g.words_per_warp;
    const GiantLayout l = giant_carve(arena, g);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < g.t1_size + g.t2_size; i += nth) l.t1[i] = kEmpty;   // t1 and t2k are contiguous
    for (uint32_t i = tid; i < g.t2_size; i += nth) l.t2c[i] = 0;
    for (uint32_t i = tid; i < 2 * g.fan_cap + 2; i += nth) l.cnt[i] = 0;           // cnt, excl, n_sets, n_matched
    decode_read(packed + rd.word_off, rd.len, l.wm, g.pk_words, tid, nth);
}

template <int K>
__global__ void __launch_bounds__(256) giant_scan_kernel(DeviceIndex ix, const ReadDesc *__restrict__ reads, uint32_t first_read,
                                                         PlaceGeom g, uint32_t *__restrict__ arenas) {
    __shared__ __align__(16) uint64_t rings[8][2 * kRing];
    __shared__ uint64_t tail_lut[64];
    init_tail_lut(tail_lut);
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
    const uint32_t r = blockIdx.y;
    const ReadDesc rd = reads[first_read + r];
    const uint32_t k = ix.k_size, L = rd.len, W = L - k + 1;
    GiantLayout l = giant_carve(arenas + (size_t)r * g.words_per_warp, g);
    l.wm.ring_a = rings[warp];
    l.wm.ring_b = rings[warp] + kRing;
    const ReadTables tb{l.t1, l.t2k, l.t2c, l.lst, l.cnt, l.excl, l.n_sets, g.t1_size - 1u, g.t2_size - 1u, 32u - g.t2_log2};
    const uint32_t code_mask = ix.m_eff >= 16 ? 0xFFFFFFFFu : ((1u << (2 * ix.m_eff)) - 1u);
    WindowHasher<K> wh(l.wm, tail_lut, L, k);
    const uint32_t n_chunks = wh.n_chunks();
    // the 2 * n_chunks chunks of the two strands, dealt out to the warps of the read's blocks in contiguous ranges
    const uint32_t n_warps = gridDim.x * warps_per_cta, me = blockIdx.x * warps_per_cta + warp;
    const uint32_t per = (2 * n_chunks + n_warps - 1) / n_warps;
    const uint32_t lo = min(me * per, 2 * n_chunks), hi = min(lo + per, 2 * n_chunks);
    uint32_t n_matched = 0;
    for (uint32_t strand = 0; strand < 2; ++strand) {
        const uint32_t c_lo = strand ? (lo > n_chunks ? lo - n_chunks : 0u) : min(lo, n_chunks);
        const uint32_t c_hi = strand ? (hi > n_chunks ? hi - n_chunks : 0u) : min(hi, n_chunks);
        if (c_lo >= c_hi) continue;
        wh.begin_strand(strand, c_lo);
        for (uint32_t c = c_lo; c < c_hi; ++c) {
            uint32_t pos;
            uint64_t h = 0, h0 = 0, m0 = 0, h1 = 0, m1 = 0;
            const bool valid = wh.pass(c, pos, h);
            uint32_t slot_id = kEmpty, set_off = 0, code = 0;
            if (valid) {
                uint64_t b = h & ix.bucket_mask;
                for (;;) {
                    ld_bucket(ix.table, b, h0, m0, h1, m1);
                    if (h0 == h && (uint32_t)m0 != kEmpty) { slot_id = (uint32_t)(2 * b); set_off = (uint32_t)m0; code = (uint32_t)(m0 >> 32); break; }
                    if (h1 == h && (uint32_t)m1 != kEmpty) { slot_id = (uint32_t)(2 * b + 1); set_off = (uint32_t)m1; code = (uint32_t)(m1 >> 32); break; }
                    if (!((uint32_t)(m0 >> 32) & kOverflowBit)) break;
                    b = (b + 1) & ix.bucket_mask;
                }
            }
            bool hit = slot_id != kEmpty;
            if (hit) {  // bucket gating (kmers_map.rs:55-70)
                const uint32_t want = code & kCodeMask;
                hit = packed_bits(strand ? l.wm.pk_r : l.wm.pk_f, pos, code_mask) == want;
                if (!hit) {  // only possible for models whose bucket keys disagree with their k-mers
                    for (uint32_t p = 0; p < W && !hit; ++p)
                        hit = packed_bits(l.wm.pk_f, p, code_mask) == want || packed_bits(l.wm.pk_r, p, code_mask) == want;
                }
            }
            n_matched += insert_hits<true>(tb, hit, slot_id, set_off);
        }
    }
    if (lane == 0 && n_matched) atomicAdd(l.n_matched, n_matched);
}

template <bool CLOSED>
__global__ void __launch_bounds__(32) giant_finish_kernel(DeviceIndex ix, PlaceParams pp, uint32_t first_read, ResultRec *__restrict__ results,
                                                          PlaceGeom g, uint32_t *__restrict__ arenas) {
    const uint32_t r = blockIdx.x;
    const GiantLayout l = giant_carve(arenas + (size_t)r * g.words_per_warp, g);
    const ReadTables tb{l.t1, l.t2k, l.t2c, l.lst, l.cnt, l.excl, l.n_sets, g.t1_size - 1u, g.t2_size - 1u, 32u - g.t2_log2};
    finish_read<CLOSED>(ix, pp, tb, *l.n_sets, *l.n_matched, results + first_read + r);
}
