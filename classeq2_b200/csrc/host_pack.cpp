// Host-side 2-bit packing of query bases (A=0 C=1 T=2 G=3 = (ascii >> 1) & 3, case-insensitive,
// 16 bases per 32-bit word, base j of a word in bits 2j..2j+1).  Input side of the path: what the
// reference's FASTA reader hands to place_sequence (sequence.rs:47-56 guarantees pure ACGT).
// The body is picked at run time: AVX-512 (BW + VL: 64 bases per step, masked loads and stores so that a read has
// no scalar tail; the code bits are gathered with two multiply-adds and a 32 -> 8 bit narrowing), else AVX2 + BMI2
// (32 bases per step, PEXT gathers the two code bits of every byte), else the portable SWAR body.  Returns false on
// a non-ACGT byte.
#include "host_pack.hpp"

#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace cls {

namespace {

inline bool pack8_swar(uint64_t x, uint32_t &out16) {
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

inline bool pack_tail(const uint8_t *s, uint32_t n, uint32_t *dst) {  // n < 16 bases -> one word
    bool ok = true;
    uint32_t v = 0;
    for (uint32_t j = 0; j < n; ++j) {
        const uint8_t ch = s[j], u = ch & 0xDF;
        ok &= (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
        v |= ((uint32_t)(ch >> 1) & 3u) << (2 * j);
    }
    *dst = v;
    return ok;
}

bool pack_read_swar(const uint8_t *s, uint32_t len, uint32_t *dst) {
    bool ok = true;
    uint32_t i = 0, w = 0;
    for (; i + 16 <= len; i += 16, ++w) {
        uint64_t a, b;
        std::memcpy(&a, s + i, 8);
        std::memcpy(&b, s + i + 8, 8);
        uint32_t lo, hi;
        ok &= pack8_swar(a, lo);
        ok &= pack8_swar(b, hi);
        dst[w] = lo | (hi << 16);
    }
    if (i < len) ok &= pack_tail(s + i, len - i, dst + w);
    return ok;
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"))) bool pack_read_avx2(const uint8_t *s, uint32_t len, uint32_t *dst) {
    const __m256i up = _mm256_set1_epi8((char)0xDF);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const uint64_t sel = 0x0606060606060606ULL;
    uint32_t okmask = 0xFFFFFFFFu;
    uint32_t i = 0, w = 0;
    for (; i + 32 <= len; i += 32, w += 2) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i));
        const __m256i y = _mm256_and_si256(x, up);
        const __m256i good = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(y, cA), _mm256_cmpeq_epi8(y, cC)),
                                             _mm256_or_si256(_mm256_cmpeq_epi8(y, cG), _mm256_cmpeq_epi8(y, cT)));
        okmask &= (uint32_t)_mm256_movemask_epi8(good);
        const uint64_t q0 = (uint64_t)_mm256_extract_epi64(x, 0), q1 = (uint64_t)_mm256_extract_epi64(x, 1);
        const uint64_t q2 = (uint64_t)_mm256_extract_epi64(x, 2), q3 = (uint64_t)_mm256_extract_epi64(x, 3);
        dst[w] = (uint32_t)(_pext_u64(q0, sel) | (_pext_u64(q1, sel) << 16));
        dst[w + 1] = (uint32_t)(_pext_u64(q2, sel) | (_pext_u64(q3, sel) << 16));
    }
    bool ok = okmask == 0xFFFFFFFFu;
    if (i < len) ok &= pack_read_swar(s + i, len - i, dst + w);
    return ok;
}

// 64 bases per step.  codes c = (x >> 1) & 3 as bytes; maddubs with weights (1, 4) joins two bases into 4 bits of a
// 16-bit lane, madd with weights (1, 16) joins two of those into 8 bits of a 32-bit lane (4 bases), and vpmovdb
// narrows the sixteen lanes to sixteen bytes = four words.  The last step of a read loads and stores under a mask:
// exactly ceil(n / 16) words are written (the next read's words belong to another thread).
__attribute__((target("avx512f,avx512bw,avx512vl"))) bool pack_read_avx512(const uint8_t *s, uint32_t len, uint32_t *dst) {
    const __m512i up = _mm512_set1_epi8((char)0xDF), three = _mm512_set1_epi8(3);
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'), cT = _mm512_set1_epi8('T');
    const __m512i w1 = _mm512_set1_epi16(0x0401), w2 = _mm512_set1_epi32(0x00100001);
    __mmask64 bad = 0;
    uint8_t *out = reinterpret_cast<uint8_t *>(dst);
    uint32_t i = 0;
    for (; i + 64 <= len; i += 64, out += 16) {
        const __m512i x = _mm512_loadu_si512(s + i);
        const __m512i y = _mm512_and_si512(x, up);
        bad |= ~(_mm512_cmpeq_epi8_mask(y, cA) | _mm512_cmpeq_epi8_mask(y, cC) | _mm512_cmpeq_epi8_mask(y, cG) | _mm512_cmpeq_epi8_mask(y, cT));
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(x, 1), three);
        const __m512i q = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w1), w2);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(out), _mm512_cvtepi32_epi8(q));
    }
    if (i < len) {
        const uint32_t n = len - i;                                  // 1 .. 63 bases left
        const __mmask64 k = ((__mmask64)1 << n) - 1;
        const __m512i x = _mm512_maskz_loadu_epi8(k, s + i);
        const __m512i y = _mm512_and_si512(x, up);
        bad |= k & ~(_mm512_cmpeq_epi8_mask(y, cA) | _mm512_cmpeq_epi8_mask(y, cC) | _mm512_cmpeq_epi8_mask(y, cG) | _mm512_cmpeq_epi8_mask(y, cT));
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(x, 1), three);
        const __m512i q = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w1), w2);
        const __mmask16 kb = (__mmask16)((1u << (4 * ((n + 15) / 16))) - 1u);   // whole words: 4, 8, 12 or 16 bytes
        _mm_mask_storeu_epi8(out, kb, _mm512_cvtepi32_epi8(q));
    }
    return bad == 0;
}
#endif

using PackFn = bool (*)(const uint8_t *, uint32_t, uint32_t *);

bool cpu_has(int variant) {
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (variant == kPackAvx512) return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
    if (variant == kPackAvx2) return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
#endif
    return variant == kPackPortable;
}

PackFn pick() {
#if defined(__x86_64__)
    if (cpu_has(kPackAvx512)) return pack_read_avx512;
    if (cpu_has(kPackAvx2)) return pack_read_avx2;
#endif
    return pack_read_swar;
}

}  // namespace

bool pack_read(const uint8_t *s, uint32_t len, uint32_t *dst) {
    static const PackFn fn = pick();
    return fn(s, len, dst);
}

bool pack_read_portable(const uint8_t *s, uint32_t len, uint32_t *dst) { return pack_read_swar(s, len, dst); }

int pack_read_variant(int variant, const uint8_t *s, uint32_t len, uint32_t *dst) {
    if (variant == kPackAuto) return pack_read(s, len, dst) ? 1 : 0;
    if (!cpu_has(variant)) return -1;
#if defined(__x86_64__)
    if (variant == kPackAvx512) return pack_read_avx512(s, len, dst) ? 1 : 0;
    if (variant == kPackAvx2) return pack_read_avx2(s, len, dst) ? 1 : 0;
#endif
    return pack_read_swar(s, len, dst) ? 1 : 0;
}

}  // namespace cls
