#include "host_pool.hpp"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace cls {

int host_threads() {
    // CLS_HOST_THREADS overrides the default (all cores, at most 32).  Under torchrun every rank keeps the full
    // pool: the ranks pack at different times, and a static split would idle cores while a rank waits for its GPU
    static const int n = [] {
        if (const char *e = std::getenv("CLS_HOST_THREADS")) {
            const int v = std::atoi(e);
            if (v >= 1) return std::min(v, 64);
        }
        unsigned hc = std::thread::hardware_concurrency();
        if (hc == 0) hc = 4;
        return (int)std::min(hc, 32u);
    }();
    return n;
}

namespace {

struct Pool {
    std::mutex job_mu;  // one job at a time
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    // current job
    const std::function<void(uint64_t, uint64_t)> *fn = nullptr;
    uint64_t n = 0, step = 0;
    std::atomic<uint64_t> next{0};
    uint64_t generation = 0;
    int active = 0;
    bool stop = false;

    Pool() {
        const int nw = host_threads() - 1;
        for (int i = 0; i < nw; ++i) workers.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
    }
    void run_chunks() {
        for (;;) {
            const uint64_t st = step;  // one value for the claim and for the range it covers
            const uint64_t a = next.fetch_add(st);
            if (a >= n) break;
            (*fn)(a, std::min(n, a + st));
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                ++active;
            }
            run_chunks();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    void run(uint64_t n_, uint64_t grain, const std::function<void(uint64_t, uint64_t)> &f) {
        std::lock_guard<std::mutex> job(job_mu);
        const uint64_t parts = (uint64_t)host_threads() * 4;
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = &f; n = n_;
            step = std::max<uint64_t>(grain, (n_ + parts - 1) / parts);
            next.store(0);
            ++generation;
        }
        cv_work.notify_all();
        run_chunks();
        std::unique_lock<std::mutex> lk(mu);
        // workers that woke up for this generation must have left run_chunks before fn goes away;
        // workers that have not woken yet will find next >= n and do nothing
        cv_done.wait(lk, [&] { return active == 0; });
        fn = nullptr;
    }
};

Pool &pool() {
    static Pool p;
    return p;
}

}  // namespace

void parallel_for_impl(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)> &f) { pool().run(n, grain, f); }

}  // namespace cls
