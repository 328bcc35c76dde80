// host_parallel.h — the two tools of the host analysis of systems with 10^5..10^6 variables:
//
//   parallel_ranges   runs fn(begin, end, part) over [0, n) cut into contiguous parts on the threads of a persistent pool
//                     (HostPool); one part (the caller's thread, the pool is not touched) below `grain` items per part, so
//                     small sketches pay nothing;
//   uvec<T>           a std::vector that does not zero what resize() adds.  On a fresh mapping the zero-fill of a
//                     std::vector is a SERIAL walk over pages nobody has touched yet (about 0.5 ms per MB of page faults on
//                     the boxes measured: the 177 MB of analysed constraints of a 1M-variable sketch cost 90 ms before a
//                     single slot was computed); a uvec leaves the first touch to the threads that fill it.
#pragma once
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
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

// The threads behind parallel_ranges: a process-wide pool created on first use (min(cores, 16) - 1 workers; the caller is
// the last one).  Starting and joining 15 threads costs ~300 us — more than most phases of the analysis of a mid-size sketch and
// as much as copying 9 MB — so the workers persist: they spin for a moment after a job (the phases of an analysis follow each
// other within microseconds) and then sleep on a condition variable.  One job at a time, in lockstep: every worker checks in
// for every job, so none can be late into the next one; a second caller (or a worker) that finds the pool busy runs its
// parts on its own thread.  A forked child starts a pool of its own.
class HostPool {
   public:
    static HostPool& get() {
        static std::atomic<HostPool*> inst{nullptr};
        static std::mutex make;
        HostPool* p = inst.load(std::memory_order_acquire);
        if (p && p->pid_ == getpid()) return *p;
        std::lock_guard<std::mutex> lock(make);
        p = inst.load(std::memory_order_acquire);
        if (!p || p->pid_ != getpid()) {  // (a forked child inherits the object but not its threads: the old one is abandoned)
            p = new HostPool();
            inst.store(p, std::memory_order_release);
        }
        return *p;
    }
    uint32_t workers() const { return (uint32_t)threads_.size(); }
    // fn(part) for part in [0, parts): on the pool and the caller; returns when every part is done.
    template <class F>
    void run(uint32_t parts, F&& fn) {
        if (parts == 0) return;
        if (parts == 1 || threads_.empty() || in_worker() || !busy_.try_lock()) {
            for (uint32_t i = 0; i < parts; ++i) fn(i);
            return;
        }
        using Fn = typename std::remove_reference<F>::type;
        call_ = [](void* a, uint32_t i) { (*static_cast<Fn*>(a))(i); };
        arg_ = const_cast<void*>(static_cast<const void*>(&fn));
        parts_ = parts;
        next_.store(0, std::memory_order_relaxed);
        arrived_.store(0, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lock(mu_);
            generation_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (uint32_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < parts;) fn(i);
        const uint32_t n = (uint32_t)threads_.size();
        for (uint32_t spins = 0; arrived_.load(std::memory_order_acquire) < n; ++spins)
            if (spins > 2000) std::this_thread::yield();
        busy_.unlock();
    }

   private:
    static bool& in_worker() {
        static thread_local bool flag = false;
        return flag;
    }
    HostPool() : pid_(getpid()) {
        if (const char* e = std::getenv("EZPZ_B200_POOL_SPIN_US")) spin_us_ = std::strtol(e, nullptr, 10);
        const uint32_t cores = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
        for (uint32_t t = 1; t < cores; ++t) {
            threads_.emplace_back([this] { loop(); });
            threads_.back().detach();
        }
    }
    void loop() {
        in_worker() = true;
        uint64_t last = 0;
        for (;;) {
            // a short spin (the next phase usually follows at once), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            while (generation_.load(std::memory_order_acquire) == last) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) {
                    std::unique_lock<std::mutex> lock(mu_);
                    cv_.wait(lock, [&] { return generation_.load(std::memory_order_acquire) != last; });
                    break;
                }
            }
            last = generation_.load(std::memory_order_acquire);
            const uint32_t parts = parts_;
            for (uint32_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < parts;) call_(arg_, i);
            arrived_.fetch_add(1, std::memory_order_release);
        }
    }
    pid_t pid_;
    long spin_us_ = 200;  // how long a worker spins for the next job before it sleeps (EZPZ_B200_POOL_SPIN_US)
    std::vector<std::thread> threads_;
    std::mutex mu_, busy_;
    std::condition_variable cv_;
    std::atomic<uint64_t> generation_{0};
    std::atomic<uint32_t> next_{0}, arrived_{0};
    void (*call_)(void*, uint32_t) = nullptr;
    void* arg_ = nullptr;
    uint32_t parts_ = 0;
};

template <class F>
inline void parallel_ranges(uint32_t n, uint32_t grain, F&& fn, uint32_t* parts_out = nullptr) {
    const uint32_t nt = host_threads(n, grain);
    if (parts_out) *parts_out = nt;
    if (nt <= 1) {
        fn(0u, n, 0u);
        return;
    }
    // (the partition depends on n, the grain and the thread cap only — not on who executes the parts)
    HostPool::get().run(nt, [&](uint32_t t) { fn((uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt), t); });
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
