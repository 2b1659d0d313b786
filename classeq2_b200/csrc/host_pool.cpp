#include "host_pool.hpp"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

namespace cls {

int host_threads() {
    // CLS_HOST_THREADS overrides the default: all cores, at most 32 - and under a launcher that runs one process per
    // GPU (LOCAL_WORLD_SIZE = N > 1, torchrun) this rank's share of the cores.  N pools of 32 threads on a 32-core box
    // spent their time in the scheduler (pack_ms 3.2 -> 10.9 ms from 1 to 8 ranks, SCALE_r01); twice the share (round 2's
    // first choice, good for host packing alone) still left every parallel loop of a rank waiting for a worker that had
    // to get a time slice on a core another rank was using: fifty short loops per call took 15 ms of a 20 ms call with
    // device packing (profiles/r2b/bench_cfg3_n8_*.json).
    static const int n = [] {
        if (const char *e = std::getenv("CLS_HOST_THREADS")) {
            const int v = std::atoi(e);
            if (v >= 1) return std::min(v, 64);
        }
        unsigned hc = std::thread::hardware_concurrency();
        if (hc == 0) hc = 4;
        unsigned want = std::min(hc, 32u);
        if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) {
            const int lws = std::atoi(e);
            if (lws > 1) want = std::max(2u, std::min(want, hc / (unsigned)lws));
        }
        return (int)want;
    }();
    return n;
}

namespace {

// One parallel_for call.  It lives on the caller's stack for the duration of the call; workers only ever see it
// through Pool::cur, under the pool mutex, and are counted in Pool::active for as long as they hold it - so every
// field a worker reads belongs to the job it joined, never to a later one.
struct Job {
    const std::function<void(uint64_t, uint64_t)> &fn;
    const uint64_t n, step;
    std::atomic<uint64_t> next{0};
    std::mutex err_mu;
    std::exception_ptr err;   // the first exception thrown by fn on any thread; rethrown by Pool::run on the caller's
    Job(const std::function<void(uint64_t, uint64_t)> &f, uint64_t n_, uint64_t step_) : fn(f), n(n_), step(step_) {}
    void work() {
        for (;;) {
            const uint64_t a = next.fetch_add(step);
            if (a >= n) break;
            try {
                fn(a, std::min(n, a + step));
            } catch (...) {
                {
                    std::lock_guard<std::mutex> lk(err_mu);
                    if (!err) err = std::current_exception();
                }
                next.store(n);   // nobody starts another range
            }
        }
    }
};

struct Pool {
    std::mutex job_mu;  // one job at a time
    std::mutex mu;      // guards cur, generation, active, stop
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    Job *cur = nullptr;       // the job workers may join (null between jobs and while a job drains)
    uint64_t generation = 0;  // bumped once per job: a worker joins a job at most once
    int active = 0;           // workers currently inside Job::work
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
    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || (cur && generation != seen); });
            if (stop) return;
            seen = generation;
            Job *j = cur;
            ++active;
            lk.unlock();
            j->work();
            lk.lock();
            if (--active == 0) cv_done.notify_all();
        }
    }
    void run(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)> &f) {
        std::lock_guard<std::mutex> one(job_mu);
        const uint64_t parts = (uint64_t)host_threads() * 4;
        Job job(f, n, std::max<uint64_t>(grain, (n + parts - 1) / parts));
        {
            std::lock_guard<std::mutex> lk(mu);
            cur = &job;
            ++generation;
        }
        cv_work.notify_all();
        job.work();   // never throws: exceptions of fn are parked in the job
        {
            std::unique_lock<std::mutex> lk(mu);
            cur = nullptr;  // a worker that wakes up from here on finds no job; those inside are counted in `active`
            cv_done.wait(lk, [&] { return active == 0; });
        }
        if (job.err) std::rethrow_exception(job.err);   // the job (on this stack) is no longer referenced by any worker
    }
};

Pool &pool() {
    static Pool p;
    return p;
}

}  // namespace

void parallel_for_impl(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)> &f) { pool().run(n, grain, f); }

}  // namespace cls
