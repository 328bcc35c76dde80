// Host-side costs behind the staging decisions (run on the GPU box): spawning + joining k threads, and copying 8.4 MB of
// pageable memory with k threads that already exist (a barrier-released pool).  g++ -O2 -pthread host_copy_bench.cpp
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t N = 8388608 + 589824;
    std::vector<char> src(N, 1), dst(N, 2);
    for (int k : {1, 2, 4, 8, 15}) {
        double best = 1e18;
        for (int rep = 0; rep < 20; ++rep) {
            const double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < k; ++t) th.emplace_back([] {});
            for (auto& x : th) x.join();
            best = std::min(best, now() - t0);
        }
        std::printf("spawn + join %2d threads: %7.1f us\n", k, best);
    }
    for (int k : {1, 2, 4, 8, 16}) {
        std::atomic<int> go{0}, done{0};
        std::atomic<bool> quit{false};
        std::vector<std::thread> th;
        for (int t = 1; t < k; ++t)
            th.emplace_back([&, t] {
                int seen = 0;
                while (true) {
                    while (go.load(std::memory_order_acquire) == seen && !quit.load()) {}
                    if (quit.load()) return;
                    ++seen;
                    const size_t b = N * t / k, e = N * (t + 1) / k;
                    std::memcpy(dst.data() + b, src.data() + b, e - b);
                    done.fetch_add(1, std::memory_order_release);
                }
            });
        double best = 1e18;
        for (int rep = 0; rep < 30; ++rep) {
            const double t0 = now();
            done.store(0);
            go.fetch_add(1, std::memory_order_release);
            std::memcpy(dst.data(), src.data(), N / k);
            while (done.load(std::memory_order_acquire) < k - 1) {}
            best = std::min(best, now() - t0);
        }
        quit.store(true);
        for (auto& x : th) x.join();
        std::printf("copy %.1f MB with %2d spinning threads: %7.1f us = %5.1f GB/s\n", N / 1e6, k, best, N / best / 1e3);
    }
}
