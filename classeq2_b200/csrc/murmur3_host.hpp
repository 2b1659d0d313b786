// Host-side MurmurHash3_x64_128 (first half), the function the reference reaches through
// `mur3::murmurhash3_x64_128(kmer.as_bytes(), 0).0` (core/src/domain/dtos/kmers_map.rs:157-159;
// crate mur3 0.1.0, Cargo.lock:2405-2408 - not vendored, published algorithm by A. Appleby).
// Used while building the index (bucket keys of the 4^m prefixes) and by the host-side
// model builder.  The device twin lives in murmur3_device.cuh.
#pragma once
#include <cstdint>
#include <cstring>

namespace cls {

static inline uint64_t rotl64_h(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64_h(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

static inline uint64_t load_le(const uint8_t *p, int n) {  // n <= 8 bytes, little endian
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

static inline uint64_t murmur3_x64_128_h1(const uint8_t *data, uint64_t len, uint64_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t nblocks = len / 16;
    for (uint64_t i = 0; i < nblocks; ++i) {
        uint64_t k1 = load_le(data + 16 * i, 8), k2 = load_le(data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64_h(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64_h(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64_h(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64_h(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + 16 * nblocks;
    const int t = (int)(len & 15);
    if (t > 8) {
        uint64_t k2 = load_le(tail + 8, t - 8);
        k2 *= c2; k2 = rotl64_h(k2, 33); k2 *= c1; h2 ^= k2;
    }
    if (t > 0) {
        uint64_t k1 = load_le(tail, t > 8 ? 8 : t);
        k1 *= c1; k1 = rotl64_h(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64_h(h1); h2 = fmix64_h(h2);
    h1 += h2;
    return h1;
}

}  // namespace cls
