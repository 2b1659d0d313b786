// scanfrag_kernel: the hashing and probing of kb-scale reads (k = 35, closed models; one CTA per read in the placement
// kernel), done the way the short-read kernel does it.  A read is cut into FRAGMENTS of 128 windows per strand (162
// bases, consecutive fragments overlap by 34); a warp takes one fragment at a time from a global counter, decodes its two
// strands, and runs scan2_kernel's pass loop over it - pre-mix ring, two interleaved register sets, the bucket load of a
// pass in flight while the next pass is hashed.  What it leaves, per window of either strand of the read, is
// {table slot of the hit or kEmpty, node-set record}: 8 bytes in a global array, 256 coalesced bytes per pass.
// gather_kernel (one CTA per read) turns these into the read's {node-set record, distinct hits} pairs for the descent
// kernel.  A hit whose bucket key is not the one of its own window's prefix (models whose bucket keys disagree with
// their k-mers; kmers_map.rs:55-70 accepts the key of ANY window of the query) marks the read in `redo`, and so does a
// read that outgrows gather_kernel's tables: place_kernel<35, true, true> then places exactly those reads, as before.
//
// Included by kernels.cu inside namespace cls { namespace { ... } }, after scan2_kernels.cuh.
#pragma once

constexpr uint32_t kFragWindows = 128;                 // windows per strand and fragment (four passes)
constexpr uint32_t kFragBases = kFragWindows + 34;     // 162 = Scan2Layout<4>::kMaxLen

// The two strands of the fragment of `flen` bases that starts at base `s_f` (a multiple of 16) of the read, and of the
// same stretch of the reverse-complement strand - the reverse complement of the read's bases [s_r, s_r + flen), s_r any.
// Lanes 0-15 own the forward words, lanes 16-31 the reverse-complement words; one 16-byte store per lane.
__device__ __forceinline__ void decode_frag16(const uint32_t *__restrict__ packed, uint32_t s_f, uint32_t s_r, uint32_t flen, uint32_t wb) {
    using Ly = Scan2Layout<4>;
    const uint32_t lane = threadIdx.x & 31u, t = lane & 15u;
    const uint32_t nw = (flen + 15u) >> 4;            // at most 11
    const uint32_t pad2 = 2u * (nw * 16u - flen);     // unused bits at the top of the last virtual word
    uint32_t fw = 0, u = 0;
    if (t < nw) {
        fw = __ldg(packed + (s_f >> 4) + t);
        // word t of the stretch [s_r, s_r + flen) as if it began at a word boundary (one word past the read's last may be
        // touched: the packed buffers carry slack, and what comes from there is shifted out below)
        const uint32_t q = (s_r >> 4) + t;
        u = __funnelshift_r(__ldg(packed + q), __ldg(packed + q + 1), 2u * (s_r & 15u));
    }
    const uint32_t r = revcomp16(u);
    const uint32_t a = __shfl_sync(kFull, r, (nw - 1u - t) & 15u), b = __shfl_sync(kFull, r, (nw - 2u - t) & 15u);
    uint32_t v = fw;
    if (lane >= 16) v = t < nw ? __funnelshift_r(a, t + 1u < nw ? b : 0u, pad2) : 0u;
    uint32_t q4[4];
    decode16(v, q4);
    const uint32_t sbase = wb + (lane >= 16 ? Ly::oStrR : Ly::oStrF) + 16u * t;
    if (t < Ly::kStrWords / 4)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbase), "r"(q4[0]), "r"(q4[1]), "r"(q4[2]), "r"(q4[3]) : "memory");
    sts_u32(wb + (lane >= 16 ? Ly::oPkR : Ly::oPkF) + 4u * t, v);
}

// Consume the bucket of a pass: match, follow the overflow chain, and write {slot, node-set record} of every window.
__device__ __forceinline__ void frag_consume(Scan2Ctx &cx, Flight &f, uint2 *__restrict__ out, uint8_t *__restrict__ redo_flag) {
    const uint32_t lane = threadIdx.x & 31u;
    const bool valid = (int32_t)lane < f.lim;
    uint32_t b = (uint32_t)f.h & cx.bmask;
    bool e0 = f.q0 == f.h, e1 = f.q1 == f.h;   // free slots carry a hash no probe of their bucket asks for (index_build.cpp)
    bool chase = valid && !(e0 || e1) && (((uint32_t)(f.qm1 >> 32) >> (kBloomShift + (((uint32_t)(f.h >> 32) >> 8) & 7u))) & 1u);
    while (__any_sync(kFull, chase)) {
        if (chase) {
            b = (b + 1) & cx.bmask;
            ld_bucket_if(true, cx.table, b, f.q0, f.qm0, f.q1, f.qm1);
            e0 = f.q0 == f.h; e1 = f.q1 == f.h;
            chase = !(e0 || e1) && ((uint32_t)(f.qm0 >> 32) & kOverflowBit);
        }
    }
    const bool hit = valid && (e0 || e1);
    const uint64_t mm = e0 ? f.qm0 : f.qm1;
    const uint32_t want = (uint32_t)(mm >> 32) & kCodeMask;
    if (__any_sync(kFull, hit && f.gate != want)) {   // the placement kernel settles this read itself
        if (lane == 0) *redo_flag = 1;
    }
    if (valid) out[lane] = make_uint2(hit ? 2u * b + (e0 ? 0u : 1u) : kEmpty, (uint32_t)mm);
}

template <int HALF>
__device__ __forceinline__ void frag_step(Scan2Ctx &cx, uint32_t it, Flight &cur, Flight &prev, uint2 *__restrict__ hv, uint32_t W,
                                          uint32_t w0, uint8_t *__restrict__ redo_flag) {
    const bool more = it < cx.n_total;   // warp-uniform
    if (more) scan2_hash<4, HALF>(cx, it, cur);
    if (it > 0 && it <= cx.n_total) {
        const uint32_t pi = it - 1u, rc = pi >= cx.n_chunks ? 1u : 0u, c = pi - rc * cx.n_chunks;
        frag_consume(cx, prev, hv + (size_t)rc * W + w0 + 32u * c, redo_flag);
    }
    if (more) {
        const bool valid = (int32_t)(threadIdx.x & 31u) < cur.lim;
        ld_bucket_if(valid, cx.table, (uint32_t)cur.h & cx.bmask, cur.q0, cur.qm0, cur.q1, cur.qm1);
    }
}

// `frags_per_read` = fragments of the longest read of the launch: work unit u is fragment u % frags_per_read of read
// u / frags_per_read (units past a read's last fragment are skipped).  hits[r * hit_stride + strand * W + window].
__global__ void __launch_bounds__(256, 5)
    scanfrag_kernel(DeviceIndex ix, const uint32_t *__restrict__ packed, const ReadDesc *__restrict__ reads, uint32_t first_read,
                    uint32_t n_reads, uint32_t frags_per_read, uint2 *__restrict__ hits, uint32_t hit_stride, uint8_t *__restrict__ redo,
                    uint32_t *__restrict__ counter) {
    using Ly = Scan2Layout<4>;
    extern __shared__ __align__(1024) uint32_t smem[];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t *wbase = smem + (size_t)warp * (Ly::kBytes / 4);
    Scan2Ctx cx;
    cx.wb = smem_addr(wbase);
    for (uint32_t i = lane; i < 64; i += 32) {
        const uint64_t t = (uint64_t)((kAsciiLut >> (8 * (i & 3))) & 0xFF) | ((uint64_t)((kAsciiLut >> (8 * ((i >> 2) & 3))) & 0xFF) << 8) |
                           ((uint64_t)((kAsciiLut >> (8 * ((i >> 4) & 3))) & 0xFF) << 16);
        sts_u64(cx.wb + Ly::oLut + 8 * i, premix_k1(t));
    }
    cx.code_mask = ix.m_eff >= 16 ? 0xFFFFFFFFu : ((1u << (2 * ix.m_eff)) - 1u);
    cx.bmask = (uint32_t)ix.bucket_mask;
    cx.table = reinterpret_cast<uint64_t>(ix.table);
    cx.sh8 = (lane & 3u) * 8u; cx.sh2 = (lane & 15u) * 2u;
    cx.lt = (1u << lane) - 1u;
    cx.A0 = cx.wb + Ly::oRingA + 8u * lane;
    cx.O1 = cx.wb + Ly::oRingB + 8u * ((lane + 40u) & 63u);
    cx.O2 = cx.wb + Ly::oRingA + 8u * ((lane + 48u) & 63u);
    cx.O3 = cx.wb + Ly::oRingB + 8u * ((lane + 56u) & 63u);
    cx.str_lane = 4u * (lane >> 2); cx.pk_lane = 4u * (lane >> 4);
    cx.n_list = 0;
    __syncwarp();
    const uint64_t n_units = (uint64_t)n_reads * frags_per_read;
    constexpr uint32_t kUnitBlock = 4;   // fragments a warp takes from the counter at a time (one address: atomics on it serialise)
    uint32_t blk_base = 0, blk_used = kUnitBlock;
#pragma unroll 1
    for (;;) {
        if (blk_used == kUnitBlock) {
            uint32_t u0 = 0;
            if (lane == 0) u0 = atomicAdd(counter, kUnitBlock);
            blk_base = __shfl_sync(kFull, u0, 0);
            blk_used = 0;
        }
        const uint64_t unit = (uint64_t)blk_base + blk_used++;
        if (unit >= n_units) break;
        const uint32_t r = (uint32_t)(unit / frags_per_read), fr = (uint32_t)(unit - (uint64_t)r * frags_per_read);
        const ReadDesc rd = reads[first_read + r];
        const uint32_t L = rd.len, W = L - 34u;          // the host guarantees L >= 35
        const uint32_t w0 = fr * kFragWindows;
        if (w0 >= W) continue;
        const uint32_t flen = min(kFragBases, L - w0);
        decode_frag16(packed + rd.word_off, w0, L - w0 - flen, flen, cx.wb);
        cx.W = flen - 34u;
        cx.n_chunks = (cx.W + 31u) >> 5; cx.n_total = 2u * cx.n_chunks;
        cx.gate_next = 0;
        Flight fa, fb;
        fa.h = fa.q0 = fa.qm0 = fa.q1 = fa.qm1 = 0; fa.gate = 0; fa.lim = 0;
        cx.sb = cx.pb = 0; cx.lim = 0;
        fb = fa;
        __syncwarp();
        uint2 *hv = hits + (size_t)r * hit_stride;
#pragma unroll 1
        for (uint32_t it = 0; it <= cx.n_total; it += 2) {
            frag_step<0>(cx, it, fa, fb, hv, W, w0, redo + r);
            frag_step<1>(cx, it + 1u, fb, fa, hv, W, w0, redo + r);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// gather_kernel: de-duplication and histogram of a kb-scale read from the hits scanfrag_kernel left.  One CTA per read.
// The placement kernel's shared-memory CAS per hit (a probing loop, sixteen warps on one table) was two thirds of its time on
// these reads; here a hit costs ONE test-and-set:
//   pass A  every hit sets its bit in a 64 K-bit filter indexed by table-slot bits (ATOMS.OR); a bit that was set
//           already is recorded in a second bitmap of COLLIDED bits
//   pass B  a hit whose bit never collided is the only hit of its table slot - distinct, no table needed (the reference
//           counts distinct hashes, kmers_map.rs:273-311); the others (the hits of a collided bit, a few per cent) go
//           through a small exact set (CAS on the slot).  Distinct hits are counted per node-set record: one leader per
//           record and warp pass (match.any) adds to a 512-slot table.
// More than so.cap distinct records (the descent kernel keeps a read's records in registers): up to kGListCap of them go,
// with the counts, to the front of the read's own hits area and descend_wide_kernel walks them from shared memory
// (kRedoWide).  More than that, or a table that fills up (low-complexity reads do not: their hits share few slots):
// kRedoPlace.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kGFilterWords = 2048, kGSetSlots = 1024, kGHistSlots = 2048, kGListCap = 1024;
constexpr uint8_t kRedoPlace = 1, kRedoWide = 2;   // redo[r]: place_kernel hashes the read again / descend_wide_kernel walks its pairs

__global__ void __launch_bounds__(256)
    gather_kernel(const ReadDesc *__restrict__ reads, uint32_t first_read, uint32_t n_reads, uint2 *__restrict__ hits,
                  uint32_t hit_stride, uint8_t *__restrict__ redo, ScanOut so) {
    __shared__ __align__(16) uint32_t filt[kGFilterWords], coll[kGFilterWords], dset[kGSetSlots], hk[kGHistSlots], hc[kGHistSlots];
    __shared__ uint32_t lst[kGListCap], n_sets, n_matched, bad;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#pragma unroll 1
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        if (redo[r]) continue;   // CTA-uniform
        const uint32_t n_ent = 2u * (reads[first_read + r].len - 34u);
        uint2 *hv = hits + (size_t)r * hit_stride;
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u), e4 = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            for (uint32_t i = tid; i < kGFilterWords / 4; i += 256) { reinterpret_cast<uint4 *>(filt)[i] = z; reinterpret_cast<uint4 *>(coll)[i] = z; }
            for (uint32_t i = tid; i < kGSetSlots / 4; i += 256) reinterpret_cast<uint4 *>(dset)[i] = e4;
            for (uint32_t i = tid; i < kGHistSlots / 4; i += 256) { reinterpret_cast<uint4 *>(hk)[i] = e4; reinterpret_cast<uint4 *>(hc)[i] = z; }
            if (tid == 0) { n_sets = 0; n_matched = 0; bad = 0; }
        }
        __syncthreads();
        for (uint32_t e = tid; e < n_ent; e += 256) {
            const uint32_t slot = __ldg(&hv[e].x);
            if (slot != kEmpty) {
                const uint32_t w = (slot >> 6) & (kGFilterWords - 1u), bit = 1u << ((slot >> 1) & 31u);
                if (atomicOr(&filt[w], bit) & bit) atomicOr(&coll[w], bit);
            }
        }
        __syncthreads();
        uint32_t nm = 0;
        for (uint32_t e0 = warp * 32u; e0 < n_ent; e0 += 256u) {   // warp-uniform bounds (match.any below)
            const uint32_t e = e0 + lane;
            uint2 v = make_uint2(kEmpty, 0u);
            if (e < n_ent) v = hv[e];
            bool fresh = v.x != kEmpty;
            if (fresh) {
                const uint32_t w = (v.x >> 6) & (kGFilterWords - 1u), bit = 1u << ((v.x >> 1) & 31u);
                if (coll[w] & bit) {   // several hits share the bit: the exact answer
                    uint32_t p = ((v.x >> 1) * 0x9E3779B1u) >> 22;
                    uint32_t tries = 0;
                    for (;;) {
                        const uint32_t old = atomicCAS(&dset[p], kEmpty, v.x);
                        if (old == kEmpty) break;
                        if (old == v.x) { fresh = false; break; }
                        p = (p + 1u) & (kGSetSlots - 1u);
                        if (++tries == kGSetSlots) { bad = 1; fresh = false; break; }
                    }
                }
            }
            const uint32_t fm = __ballot_sync(kFull, fresh);
            if (fresh) {
                const uint32_t peers = __match_any_sync(fm, v.y);
                if ((uint32_t)(__ffs(peers) - 1) == lane) {
                    uint32_t p2 = (v.y * 0x9E3779B1u) >> 21;
                    uint32_t tries = 0;
                    for (;;) {
                        const uint32_t old = atomicCAS(&hk[p2], kEmpty, v.y);
                        if (old == kEmpty) {
                            const uint32_t pos = atomicAdd(&n_sets, 1u);
                            if (pos < kGListCap) lst[pos] = p2;
                        }
                        if (old == kEmpty || old == v.y) { atomicAdd(&hc[p2], (uint32_t)__popc(peers)); break; }
                        p2 = (p2 + 1u) & (kGHistSlots - 1u);
                        if (++tries == kGHistSlots) { bad = 1; break; }
                    }
                }
            }
            nm += (uint32_t)__popc(fm);
        }
        if (lane == 0 && nm) atomicAdd(&n_matched, nm);
        __syncthreads();
        const uint32_t D = n_sets;
        if (bad || D > kGListCap || D + 1u > n_ent) {
            if (tid == 0) redo[r] = kRedoPlace;
        } else if (D > so.cap) {
            // hv[0] = {n_matched, D}, hv[1 + j] = pair j: every hit of the read has been consumed (the barrier above)
            for (uint32_t j = tid; j < D; j += 256) {
                const uint32_t p2 = lst[j];
                hv[1 + j] = make_uint2(hk[p2], hc[p2]);
            }
            if (tid == 0) { hv[0] = make_uint2(n_matched, D); redo[r] = kRedoWide; so.meta[r] = make_uint2(n_matched, kDone); }
        } else {
            for (uint32_t j = tid; j < D; j += 256) {
                const uint32_t p2 = lst[j];
                so.pairs[(size_t)r * so.cap + j] = make_uint2(hk[p2], hc[p2]);
            }
            if (tid == 0) so.meta[r] = make_uint2(n_matched, D);
        }
        __syncthreads();   // the tables are free again
    }
}

// descend_wide_kernel: the reads gather_kernel marked kRedoWide (more node-set records than the register-resident descent
// holds): one warp per read walks the records from shared memory (finish_read, the walk the placement kernel itself uses).
constexpr uint32_t kWideWarps = 4;
__global__ void __launch_bounds__(32 * kWideWarps)
    descend_wide_kernel(DeviceIndex ix, PlaceParams pp, uint32_t first_read, uint32_t n_reads, const uint2 *__restrict__ hits,
                        uint32_t hit_stride, const uint8_t *__restrict__ redo, ResultRec *__restrict__ results, uint32_t fan_cap,
                        uint32_t *__restrict__ counter) {
    extern __shared__ __align__(1024) uint32_t smem[];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t *base = smem + (size_t)warp * (5u * kGListCap + 2u * fan_cap);
    uint32_t *t1 = base, *t2k = t1 + 2 * kGListCap, *t2c = t2k + kGListCap, *lst = t2c + kGListCap, *cnt = lst + kGListCap, *excl = cnt + fan_cap;
    for (uint32_t o = lane; o < fan_cap; o += 32) { cnt[o] = 0; excl[o] = 0; }
    const ReadTables tb{t1, t2k, t2c, lst, cnt, excl, nullptr, 0u, 0u, 0u};
    __syncwarp();
#pragma unroll 1
    for (;;) {
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(counter, 32u);
        const uint32_t r0 = __shfl_sync(kFull, b0, 0);
        if (r0 >= n_reads) break;
        uint32_t wide = __ballot_sync(kFull, r0 + lane < n_reads && redo[r0 + lane] == kRedoWide);
        while (wide) {
            const uint32_t r = r0 + (uint32_t)__ffs(wide) - 1u;
            wide &= wide - 1u;
            const uint2 *hv = hits + (size_t)r * hit_stride;
            const uint2 hd = hv[0];
            const uint32_t D = hd.y;
            for (uint32_t j = lane; j < D; j += 32) {
                const uint2 v = hv[1 + j];
                lst[j] = j; t2k[j] = v.x; t2c[j] = v.y;
            }
            __syncwarp();
            finish_read<true>(ix, pp, tb, D, hd.x, results + first_read + r);
            __syncwarp();
        }
    }
}
