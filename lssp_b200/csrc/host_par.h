// host_par.h -- host threads for the set-up code (factorisation, schedules, packing).
//
// Every use hands DISJOINT output ranges to the threads and computes each output element exactly as
// the serial code does, so results never depend on the number of threads (LSSPG_HOST_THREADS, default:
// the hardware concurrency, at most 32; 1 = run inline).
#pragma once
#include <stdlib.h>
#include <string.h>
#include <algorithm>
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
        if (v <= 0) v = (int)std::thread::hardware_concurrency();
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

}  // namespace lsspg
