// host_parallel.h — the two tools of the host analysis of systems with 10^5..10^6 variables:
//
//   parallel_ranges   runs fn(begin, end, part) over [0, n) cut into contiguous parts on host threads; one part (the caller's
//                     thread, no thread is started) below `grain` items per part, so small sketches pay nothing;
//   uvec<T>           a std::vector that does not zero what resize() adds.  On a fresh mapping the zero-fill of a
//                     std::vector is a SERIAL walk over pages nobody has touched yet (about 0.5 ms per MB of page faults on
//                     the boxes measured: the 177 MB of analysed constraints of a 1M-variable sketch cost 90 ms before a
//                     single slot was computed); a uvec leaves the first touch to the threads that fill it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>

namespace ezs {

template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = default_init_allocator<U>;
    };
    using std::allocator<T>::allocator;
    template <class U>
    void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
        ::new (static_cast<void*>(p)) U;
    }
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
};
template <class T>
using uvec = std::vector<T, default_init_allocator<T>>;

constexpr uint32_t kHostGrain = 1u << 15;  // items per host thread before a phase is worth splitting

inline uint32_t host_threads(uint32_t n, uint32_t grain) {
    static const uint32_t cores = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    uint32_t cap = cores;  // EZPZ_B200_HOST_THREADS lowers it (read per call: tests compare 1 thread against many)
    if (const char* e = std::getenv("EZPZ_B200_HOST_THREADS")) cap = std::max(1u, std::min(cap, (uint32_t)std::strtoul(e, nullptr, 10)));
    return std::max(1u, std::min(cap, n / std::max(1u, grain)));
}

template <class F>
inline void parallel_ranges(uint32_t n, uint32_t grain, F&& fn, uint32_t* parts_out = nullptr) {
    const uint32_t nt = host_threads(n, grain);
    if (parts_out) *parts_out = nt;
    if (nt <= 1) {
        fn(0u, n, 0u);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (uint32_t t = 1; t < nt; ++t)
        th.emplace_back([&, t] { fn((uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt), t); });
    fn(0u, (uint32_t)((uint64_t)n / nt), 0u);
    for (auto& x : th) x.join();
}

// counter[0]++ returning the old value: a relaxed atomic when several host threads share the counters, a plain increment on
// the single-thread path of small systems (a locked instruction per pair is most of their pattern pass).
inline uint32_t bump(uint32_t* counter, bool shared) {
    if (shared) return __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED);
    return (*counter)++;
}

// Exclusive prefix sum in place over counts stored at v[1..n] (v[0] = 0 on entry): v[i + 1] += v[i].
template <class V>
inline void prefix_sum(V& v) {
    for (size_t i = 1; i < v.size(); ++i) v[i] += v[i - 1];
}

// dst[0..count) = src[0..count) on host threads (first touch of dst spread over the threads).
template <class T>
inline void parallel_copy(T* dst, const T* src, size_t count) {
    if (count >= (1ull << 32)) {
        std::copy(src, src + count, dst);
        return;
    }
    parallel_ranges((uint32_t)count, (uint32_t)std::max<size_t>(1, (4u << 20) / sizeof(T)),
                    [&](uint32_t b, uint32_t e, uint32_t) { std::copy(src + b, src + e, dst + b); });
}
template <class T>
inline void parallel_fill(T* dst, size_t count, T value) {
    if (count >= (1ull << 32)) {
        std::fill(dst, dst + count, value);
        return;
    }
    parallel_ranges((uint32_t)count, (uint32_t)std::max<size_t>(1, (4u << 20) / sizeof(T)),
                    [&](uint32_t b, uint32_t e, uint32_t) { std::fill(dst + b, dst + e, value); });
}

}  // namespace ezs
