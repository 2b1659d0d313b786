// 2-bit packing of query bases ON THE DEVICE: cls_place_batch's alternative to the host packer (host_pack.cpp) for
// hosts with few cores per GPU - the caller's ASCII bases are copied to the GPU as they are (straight from the caller's
// memory when it is pinned) and this kernel writes the packed words the placement kernels read.  Same code as the host
// packer: (ascii >> 1) & 3 (A=0 C=1 T=2 G=3, case-insensitive), 16 bases per word, base j in bits 2j..2j+1; a byte
// other than A/C/G/T in either case flags the read (the host then reports CLS_STATUS_ERR_INVALID_BASE, as it does when
// it packs itself: sequence.rs:47-56 hands place_sequence pure ACGT).
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.hpp"
#include "kernels.hpp"

namespace cls {
namespace {

// one thread per (read, packed word): reads [first, first + count) of the length-ordered batch, `words_max` words for
// the longest of them
__global__ void __launch_bounds__(256) ascii_pack_kernel(const uint8_t *__restrict__ ascii, const uint64_t *__restrict__ src_off,
                                                         const ReadDesc *__restrict__ descs, uint32_t first, uint32_t count,
                                                         uint32_t words_max, uint32_t *__restrict__ words, uint8_t *__restrict__ bad) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t j = t / words_max;
    if (j >= count) return;
    const uint32_t w = (uint32_t)(t - j * words_max);
    const ReadDesc rd = descs[first + j];
    if (16u * w >= rd.len) return;
    const uint8_t *s = ascii + src_off[first + j] + 16u * w;
    const uint32_t n = min(16u, rd.len - 16u * w);
    uint32_t v = 0;
    bool ok = true;
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) {
        if (i < n) {
            const uint32_t c = s[i], u = c & 0xDFu;
            ok &= (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
            v |= ((c >> 1) & 3u) << (2 * i);
        }
    }
    words[rd.word_off + w] = v;
    if (!ok) bad[first + j] = 1;   // every writer stores the same value
}

}  // namespace

cudaError_t launch_ascii_pack(const uint8_t *ascii, const uint64_t *src_off, const ReadDesc *descs, uint32_t first, uint32_t count,
                              uint32_t max_len, uint32_t *words, uint8_t *bad, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const uint32_t words_max = (max_len + 15u) / 16u;
    const uint64_t threads = (uint64_t)count * words_max;
    const uint64_t grid = (threads + 255) / 256;
    if (grid > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    ascii_pack_kernel<<<(uint32_t)grid, 256, 0, stream>>>(ascii, src_off, descs, first, count, words_max, words, bad);
    return cudaGetLastError();
}

}  // namespace cls
