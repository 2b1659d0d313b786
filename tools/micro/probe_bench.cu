// Micro-benchmark: the rate of random 32-byte (and 64-byte) probes of a table in global memory on one B200, by
// table size - the ceiling of the k-mer table probe.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_bench probe_bench.cu
//   ./probe_bench [mib ...]            (under ncu: --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// BYTES = 32: one 256-bit load per probe; 64: two (one 64-byte aligned line); 16: one 128-bit load
template <int BYTES, int UNROLL>
__global__ void __launch_bounds__(256) probe_kernel(const uint8_t *__restrict__ table, uint64_t n_units, uint32_t iters, uint64_t *out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0, s = tid * 0x9E3779B97F4A7C15ULL + 1;
    for (uint32_t i = 0; i < iters; i += UNROLL) {
        uint64_t v[UNROLL][BYTES / 8];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            s = mix(s + u + i);
            const uint8_t *p = table + __umul64hi(s, n_units) * BYTES;
            if constexpr (BYTES == 16) {
                asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v[u][0]), "=l"(v[u][1]) : "l"(p));
            } else {
#pragma unroll
                for (int q = 0; q < BYTES / 32; ++q)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                                 : "=l"(v[u][4 * q]), "=l"(v[u][4 * q + 1]), "=l"(v[u][4 * q + 2]), "=l"(v[u][4 * q + 3]) : "l"(p + 32 * q));
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int q = 0; q < BYTES / 8; ++q) acc ^= v[u][q];
    }
    if (acc == 0x1234567) out[0] = acc;
}

template <int BYTES>
static void run(const uint8_t *table, uint64_t bytes, uint64_t *out, int sms) {
    const uint64_t n_units = bytes / BYTES;
    const uint32_t iters = 2048;
    const int grid = sms * 8;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(a));
        probe_kernel<BYTES, 4><<<grid, 256>>>(table, n_units, iters, out);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep && ms < best) best = ms;
    }
    const double probes = (double)grid * 256 * iters;
    printf("table %5llu MiB  %2d-byte probes: %7.2f G probes/s  %7.1f GB/s algorithmic  (%.3f ms)\n", (unsigned long long)(bytes >> 20), BYTES,
           probes / best / 1e6, probes * BYTES / best / 1e6, best);
    fflush(stdout);
}

int main(int argc, char **argv) {
    std::vector<uint64_t> sizes;
    for (int i = 1; i < argc; ++i) sizes.push_back((uint64_t)atoll(argv[i]) << 20);
    if (sizes.empty()) sizes = {32ull << 20, 64ull << 20, 96ull << 20, 128ull << 20, 192ull << 20, 256ull << 20, 512ull << 20, 1024ull << 20, 4096ull << 20};
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    uint64_t mx = 0; for (auto s : sizes) mx = s > mx ? s : mx;
    uint8_t *table; uint64_t *out;
    CK(cudaMalloc(&table, mx)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(table, 0x5A, mx));
    for (auto s : sizes) { run<32>(table, s, out, prop.multiProcessorCount); run<64>(table, s, out, prop.multiProcessorCount); run<16>(table, s, out, prop.multiProcessorCount); }
    return 0;
}
