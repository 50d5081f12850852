// host_par.h -- host threads for the set-up code (factorisation, schedules, packing).
//
// Every use hands DISJOINT output ranges to the threads and computes each output element exactly as
// the serial code does, so results never depend on the number of threads (LSSPG_HOST_THREADS, default:
// the cores of the process' affinity mask divided by LOCAL_WORLD_SIZE, at most 32; 1 = run inline).
#pragma once
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

namespace lsspg {

// std::vector whose resize(n) leaves new elements uninitialised: multi-GB arrays are then first touched by the
// threads that fill them instead of being zero-filled by one thread.
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = default_init_allocator<U>;
    };
    default_init_allocator() = default;
    template <class U>
    default_init_allocator(const default_init_allocator<U> &) {}
    template <class U, class... Args>
    void construct(U *p, Args &&...args)
    {
        if constexpr (sizeof...(Args) == 0) ::new ((void *)p) U;
        else ::new ((void *)p) U(std::forward<Args>(args)...);
    }
};
using IVec = std::vector<int, default_init_allocator<int>>;
using DVec = std::vector<double, default_init_allocator<double>>;

inline int host_threads()
{
    static const int nt = [] {
        int v = 0;
        if (const char *e = getenv("LSSPG_HOST_THREADS")) v = atoi(e);
        if (v <= 0) {
            // the cores this process may run on, shared with the other ranks of a torchrun launch on this node
            cpu_set_t set;
            v = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            if (const char *w = getenv("LOCAL_WORLD_SIZE"))
                if (atoi(w) > 1) v /= atoi(w);
        }
        return std::max(1, std::min(v, 32));
    }();
    return nt;
}

// fn(begin, end, piece) over [0, n) cut into `pieces` contiguous ranges of (almost) equal length;
// pieces = 0: one per thread.  Small ranges run inline.
template <class F>
void parallel_ranges(long long n, F fn, int pieces = 0, long long min_per_piece = 1 << 14)
{
    int np = pieces > 0 ? pieces : host_threads();
    if (pieces <= 0 && n / np < min_per_piece) np = (int)std::max<long long>(1, n / min_per_piece);
    if (np <= 1) {
        fn((long long)0, n, 0);
        return;
    }
    const int nt = std::min(np, host_threads());
    auto piece = [&](int p) { fn(n * p / np, n * (p + 1) / np, p); };
    if (nt <= 1) {
        for (int p = 0; p < np; p++) piece(p);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            for (int p = t; p < np; p += nt) piece(p);
        });
    for (auto &x : th) x.join();
}

// memcpy of a large array over the threads (first touch of the destination pages is the cost)
inline void parallel_copy(void *dst, const void *src, size_t bytes)
{
    parallel_ranges((long long)bytes, [&](long long b, long long e, int) {
        memcpy((char *)dst + b, (const char *)src + b, (size_t)(e - b));
    }, 0, 1 << 22);
}

// exclusive prefix sum in place over counts[0..n) -> returns the total; counts[i] becomes the offset of i.
// Two passes over per-piece partial sums; identical to the serial scan (integer arithmetic).
template <class T>
long long parallel_exclusive_scan(T *counts, long long n)
{
    const int np = host_threads();
    if (np <= 1 || n < (1 << 18)) {
        long long run = 0;
        for (long long i = 0; i < n; i++) { const long long c = counts[i]; counts[i] = (T)run; run += c; }
        return run;
    }
    std::vector<long long> part(np + 1, 0);
    parallel_ranges(n, [&](long long b, long long e, int p) {
        long long s = 0;
        for (long long i = b; i < e; i++) s += counts[i];
        part[p + 1] = s;
    }, np);
    for (int p = 0; p < np; p++) part[p + 1] += part[p];
    parallel_ranges(n, [&](long long b, long long e, int p) {
        long long run = part[p];
        for (long long i = b; i < e; i++) { const long long c = counts[i]; counts[i] = (T)run; run += c; }
    }, np);
    return part[np];
}

// Rows whose dependencies are only discovered while they are computed (ILU(k) symbolic phase, ILUT): a row may
// need any FINISHED row of smaller index.  Chunks of consecutive rows are handed out in natural order; a thread
// runs the rows of its chunk in order and, before reading row r of an earlier chunk, waits until that chunk's
// progress counter has passed r.  Waits only ever target chunks handed out earlier, and the lowest chunk in
// flight never waits, so the scheme cannot deadlock; every row sees exactly the finished rows the serial loop
// would show it, so results do not depend on the number of threads or on timing.  With the chunk length equal to
// the innermost grid dimension of a stencil matrix, a thread trails the owner of the previous grid line by one row.
class RowPipeline {
public:
    RowPipeline(int n, int chunk) : n_(n), chunk_(std::max(1, chunk)), nchunks_((n + chunk_ - 1) / chunk_),
                                    progress_(new Counter[(size_t)std::max(nchunks_, 1)]), ticket_(0)
    {
        for (int c = 0; c < nchunks_; c++) progress_[c].v.store(0, std::memory_order_relaxed);
    }
    // row r (smaller than the caller's current row) is finished on return
    void wait(int r, int my_chunk_begin) const
    {
        if (r >= my_chunk_begin) return;   // the caller finished it itself
        const int c = r / chunk_, need = r - c * chunk_ + 1;
        int spins = 0;
        while (progress_[c].v.load(std::memory_order_acquire) < need)
            if (++spins > 4000) { std::this_thread::yield(); spins = 0; }
    }
    // body(thread, row, chunk_begin)
    template <class F>
    void run(int threads, F body)
    {
        auto worker = [&](int t) {
            for (;;) {
                const int c = ticket_.fetch_add(1, std::memory_order_relaxed);
                if (c >= nchunks_) return;
                const int b = c * chunk_, e = std::min(n_, b + chunk_);
                for (int r = b; r < e; r++) {
                    body(t, r, b);
                    // published every few rows (and at the end): the followers' polling costs the owner a cache miss
                    if (((r - b) & 3) == 3 || r == e - 1) progress_[c].v.store(r - b + 1, std::memory_order_release);
                }
            }
        };
        if (threads <= 1) {
            worker(0);
            return;
        }
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(worker, t);
        worker(0);
        for (auto &x : th) x.join();
    }

private:
    struct alignas(64) Counter {
        std::atomic<int> v;
    };
    int n_, chunk_, nchunks_;
    std::unique_ptr<Counter[]> progress_;
    std::atomic<int> ticket_;
};

// Blocks of ints/doubles that never move once handed out (rows publish pointers into them).
template <class T>
class BlockArena {
public:
    T *take(size_t count)
    {
        if (count > left_) {
            const size_t cap = std::max(count, (size_t)1 << 18);
            blocks_.emplace_back(new T[cap]);
            cur_ = blocks_.back().get();
            left_ = cap;
        }
        T *p = cur_;
        cur_ += count;
        left_ -= count;
        return p;
    }

private:
    std::vector<std::unique_ptr<T[]>> blocks_;
    T *cur_ = nullptr;
    size_t left_ = 0;
};

// Rows grouped by dependency level, run level after level by a team of host threads (spinning barrier between
// levels; the rows of one level are cut into contiguous pieces, one per thread).  fn(row) may read everything
// rows of EARLIER levels wrote and must write only what belongs to its own row -- results are then independent
// of the number of threads.
class LevelTeam {
public:
    template <class F>
    static void run(int nlev, const int *start, const int *order, F fn)
    {
        const int nt = host_threads();
        long long rows = nlev > 0 ? (long long)start[nlev] - start[0] : 0;
        if (nt <= 1 || rows < (1 << 15) || rows / std::max(nlev, 1) < 4 * nt) {
            for (int l = 0; l < nlev; l++)
                for (int q = start[l]; q < start[l + 1]; q++) fn(order[q]);
            return;
        }
        std::atomic<int> arrived(0), sense(0);
        auto worker = [&](int t) {
            int local = 0;
            for (int l = 0; l < nlev; l++) {
                const long long b = start[l], cnt = start[l + 1] - b;
                const long long q0 = b + cnt * t / nt, q1 = b + cnt * (t + 1) / nt;
                for (long long q = q0; q < q1; q++) fn(order[q]);
                // sense-reversing barrier
                local ^= 1;
                if (arrived.fetch_add(1, std::memory_order_acq_rel) == nt - 1) {
                    arrived.store(0, std::memory_order_relaxed);
                    sense.store(local, std::memory_order_release);
                }
                else {
                    int spins = 0;
                    while (sense.load(std::memory_order_acquire) != local)
                        if (++spins > 2000) std::this_thread::yield();
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(worker, t);
        worker(0);
        for (auto &x : th) x.join();
    }
};

}  // namespace lsspg
