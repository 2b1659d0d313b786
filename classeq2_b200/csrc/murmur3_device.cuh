// Device-side MurmurHash3_x64_128 (first 64-bit half, seed 0) over the ASCII bytes of one k-mer
// window: the function the reference calls for every window through
// `mur3::murmurhash3_x64_128(kmer.as_bytes(), 0).0` (core/src/domain/dtos/kmers_map.rs:157-159,
// :414-421).  The hash is NOT a rolling hash: every window is hashed independently, so the
// work per k-mer is 2 sixteen-byte blocks + a 3-byte tail + two fmix64 at k = 35
// (14 64-bit multiplies).  Integer-only; no tensor cores apply.
#pragma once
#include <cstdint>

namespace cls {

__device__ __forceinline__ uint64_t rotl64_d(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64_d(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

constexpr uint64_t kC1 = 0x87c37b91114253d5ULL;
constexpr uint64_t kC2 = 0x4cf5ad432745937fULL;

__device__ __forceinline__ void mm_block(uint64_t &h1, uint64_t &h2, uint64_t k1, uint64_t k2) {
    k1 *= kC1; k1 = rotl64_d(k1, 31); k1 *= kC2; h1 ^= k1;
    h1 = rotl64_d(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= kC2; k2 = rotl64_d(k2, 33); k2 *= kC1; h2 ^= k2;
    h2 = rotl64_d(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
}

__device__ __forceinline__ uint64_t mm_finish(uint64_t h1, uint64_t h2, uint64_t len) {
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64_d(h1); h2 = fmix64_d(h2);
    return h1 + h2;
}

// Hash the K bytes that start at byte `pos` of a 4-byte aligned ASCII string held in shared
// memory.  The string must be readable up to 4*((pos>>2) + (K+6)/4 + 1) bytes (padding is
// never mixed into the hash).  K is a compile-time constant: all block/tail indexing unrolls.
template <int K>
__device__ __forceinline__ uint64_t murmur_window_smem(const uint32_t *s32, uint32_t pos) {
    constexpr int NW = (K + 3) / 4;  // 32-bit words of window data
    const uint32_t *p = s32 + (pos >> 2);
    const uint32_t sh = (pos & 3u) * 8u;
    uint32_t raw[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; ++i) raw[i] = p[i];
    uint32_t w[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) w[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
    uint64_t h1 = 0, h2 = 0;
    constexpr int NB = K / 16;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint64_t k1 = (uint64_t)w[4 * b] | ((uint64_t)w[4 * b + 1] << 32);
        uint64_t k2 = (uint64_t)w[4 * b + 2] | ((uint64_t)w[4 * b + 3] << 32);
        mm_block(h1, h2, k1, k2);
    }
    constexpr int T = K & 15;
    if constexpr (T > 8) {
        constexpr int nb = T - 8;  // bytes of k2 (1..7)
        uint64_t k2 = (uint64_t)w[4 * NB + 2];
        if (nb > 4) k2 |= (uint64_t)w[(4 * NB + 3 < NW) ? 4 * NB + 3 : NW - 1] << 32;
        k2 &= (nb >= 8) ? ~0ULL : ((1ULL << (8 * nb)) - 1);
        k2 *= kC2; k2 = rotl64_d(k2, 33); k2 *= kC1; h2 ^= k2;
    }
    if constexpr (T > 0) {
        constexpr int nb = T > 8 ? 8 : T;  // bytes of k1 (1..8)
        uint64_t k1 = (uint64_t)w[4 * NB];
        if (nb > 4) k1 |= (uint64_t)w[(4 * NB + 1 < NW) ? 4 * NB + 1 : NW - 1] << 32;
        k1 &= (nb >= 8) ? ~0ULL : ((1ULL << (8 * nb)) - 1);
        k1 *= kC1; k1 = rotl64_d(k1, 31); k1 *= kC2; h1 ^= k1;
    }
    return mm_finish(h1, h2, (uint64_t)K);
}

// Runtime-k variant (any k >= 1): byte-wise reads, slow but exact; used for k != 35.
__device__ __forceinline__ uint64_t load_le_smem(const uint8_t *p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

__device__ inline uint64_t murmur_window_generic(const uint8_t *s, uint32_t pos, uint32_t k) {
    const uint8_t *d = s + pos;
    uint64_t h1 = 0, h2 = 0;
    const uint32_t nblocks = k / 16;
    for (uint32_t b = 0; b < nblocks; ++b)
        mm_block(h1, h2, load_le_smem(d + 16 * b, 8), load_le_smem(d + 16 * b + 8, 8));
    const uint8_t *tail = d + 16 * nblocks;
    const int t = (int)(k & 15u);
    if (t > 8) {
        uint64_t k2 = load_le_smem(tail + 8, t - 8);
        k2 *= kC2; k2 = rotl64_d(k2, 33); k2 *= kC1; h2 ^= k2;
    }
    if (t > 0) {
        uint64_t k1 = load_le_smem(tail, t > 8 ? 8 : t);
        k1 *= kC1; k1 = rotl64_d(k1, 31); k1 *= kC2; h1 ^= k1;
    }
    return mm_finish(h1, h2, (uint64_t)k);
}

}  // namespace cls
