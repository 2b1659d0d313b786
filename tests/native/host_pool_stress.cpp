// TEST INFRASTRUCTURE.  Stress test of the host thread pool (classeq2_b200/csrc/host_pool.cpp): many short jobs of
// changing size from several caller threads; every index of every job must be visited exactly once; every eleventh job
// throws from one of its ranges (on whichever thread runs it) - the caller gets that exception back, no worker is left
// holding the job, and the pool goes on.  Built by tests/test_host_pool.py with and without -fsanitize=thread.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../classeq2_b200/csrc/host_pool.hpp"

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 20000;
    const int callers = argc > 2 ? atoi(argv[2]) : 2;
    std::atomic<long> bad{0};
    std::vector<std::thread> th;
    for (int c = 0; c < callers; ++c)
        th.emplace_back([&, c] {
            std::vector<uint8_t> seen;
            uint64_t x = 88172645463325252ull + (uint64_t)c;
            for (int r = 0; r < rounds; ++r) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                const uint64_t n = 1 + x % 5000, grain = 1 + (x >> 20) % 64;
                seen.assign(n, 0);
                cls::parallel_for(n, grain, [&](uint64_t a, uint64_t b) {
                    if (b > n || a >= b) { bad++; return; }
                    for (uint64_t i = a; i < b; ++i) seen[i]++;
                });
                for (uint64_t i = 0; i < n; ++i)
                    if (seen[i] != 1) { bad++; break; }
                if (r % 11 == 5) {                                   // a job whose function throws (ADVICE: Pool::run was not exception-safe)
                    const uint64_t poison = x % n;
                    bool caught = false;
                    std::vector<uint8_t> local(n, 0);              // on this stack: a worker still inside the job after the throw would scribble on a dead frame
                    try {
                        cls::parallel_for(n, grain, [&](uint64_t a, uint64_t b) {
                            for (uint64_t i = a; i < b; ++i) local[i]++;
                            if (a <= poison && poison < b) throw std::runtime_error("poisoned range");
                        });
                    } catch (const std::runtime_error &) {
                        caught = true;
                    }
                    if (!caught) bad++;
                    for (uint64_t i = 0; i < n; ++i)
                        if (local[i] > 1) { bad++; break; }        // ranges may be skipped after the throw, never run twice
                }
            }
        });
    for (auto &t : th) t.join();
    printf("bad=%ld\n", bad.load());
    return bad.load() ? 1 : 0;
}
