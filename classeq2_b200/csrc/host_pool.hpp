// A small persistent host thread pool for the packing / scattering loops around the device calls
// (spawning std::threads per loop costs more than the loops themselves at 1 M reads per batch).
#pragma once
#include <cstdint>
#include <functional>

namespace cls {

int host_threads();

// Runs f(begin, end) over disjoint sub-ranges of [0, n) on the pool (the caller participates) and
// returns when all of it is done.  Concurrent callers are serialised.  `grain` = smallest range
// worth a task.
void parallel_for_impl(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)> &f);

template <class F>
inline void parallel_for(uint64_t n, uint64_t grain, F f) {
    if (n == 0) return;
    if (n <= grain || host_threads() <= 1) { f(0, n); return; }
    parallel_for_impl(n, grain, std::function<void(uint64_t, uint64_t)>(f));
}

}  // namespace cls
