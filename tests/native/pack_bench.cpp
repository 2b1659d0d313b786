// TEST / MEASUREMENT INFRASTRUCTURE: single-thread throughput of the host 2-bit packer bodies
// (classeq2_b200/csrc/host_pack.cpp) on reads of a given length.  usage: pack_bench [read_len=150] [n_reads=1000000]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../classeq2_b200/csrc/host_pack.hpp"

int main(int argc, char **argv) {
    const uint32_t L = argc > 1 ? atoi(argv[1]) : 150;
    const uint32_t n = argc > 2 ? atoi(argv[2]) : 1000000;
    std::vector<uint8_t> bases((size_t)L * n);
    uint64_t x = 88172645463325252ull;
    for (auto &b : bases) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = "ACGT"[x & 3]; }
    const uint32_t wpr = (L + 15) / 16;
    std::vector<uint32_t> words((size_t)wpr * n + 16);
    const char *names[] = {"auto", "portable", "avx2+bmi2", "avx512"};
    for (int v = 0; v < 4; ++v) {
        double best = 1e30;
        int ok = 1;
        for (int rep = 0; rep < 5; ++rep) {
            const auto t0 = std::chrono::steady_clock::now();
            for (uint32_t r = 0; r < n; ++r) ok &= cls::pack_read_variant(v, bases.data() + (size_t)r * L, L, words.data() + (size_t)r * wpr);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt < best) best = dt;
        }
        uint64_t sum = 0;
        for (uint32_t w : words) sum += w;
        printf("%-10s ok=%d  %7.2f ns/read  %6.2f GB/s  checksum %llu\n", names[v], ok, best / n * 1e9, (double)L * n / best / 1e9, (unsigned long long)sum);
    }
    return 0;
}
