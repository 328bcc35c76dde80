// multi.cpp — one call, several GPUs: ezpz_b200_solve_batch_multi and friends.
//
// The reference's seam is one synchronous in-process call (ezpz/src/lib.rs:80-87); a drop-in caller cannot start one
// process per GPU.  A multi-context therefore owns one worker thread per device, each with its own ezpz_context (stream,
// workspace); a batch call cuts the batch into contiguous shards (ezpz_b200_shard_range: problems are independent, there is
// no exchange step and hence no collective — SURVEY.md §8e), hands shard k to worker k and returns when all are done.
// Every worker runs the same single-device entry point on its slice of the CALLER's buffers, so a problem's result does
// not depend on how many devices took part.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ezpz_b200.h"
#include "structure.h"

namespace {

struct Job {
    const ezpz_structure_t* s = nullptr;
    const ezpz_config_t* config = nullptr;
    uint64_t count = 0;
    ezpz_batch_io_t io{};
    int32_t* status = nullptr;  // where the job's own status goes (ezpz_b200_solve_jobs_multi)
};

// A slice [b0, b1) of a batch call as a job of its own.
Job slice_job(const ezpz_structure_t* s, const ezpz_config_t* config, const ezpz_batch_io_t* io, uint64_t b0, uint64_t b1) {
    uint32_t m = 0, n = 0;
    uint64_t nnz = 0;
    ezpz_b200_structure_dims(s, &m, &n, &nnz, nullptr, nullptr, nullptr);
    const size_t nc = s->n_cons, uw = (s->n_cons + 31) / 32, vw = (n + 31) / 32;
    Job job;
    job.s = s;
    job.config = config;
    job.count = b1 - b0;
    job.io.guesses = io->guesses + b0 * n;
    job.io.params = io->params ? io->params + b0 * nc : nullptr;
    job.io.final_values = io->final_values + b0 * n;
    job.io.iterations = io->iterations + b0;
    job.io.status = io->status + b0;
    job.io.unsat_mask = io->unsat_mask ? io->unsat_mask + b0 * uw : nullptr;
    job.io.degen_count = io->degen_count ? io->degen_count + b0 * nc : nullptr;
    job.io.jacobian = io->jacobian ? io->jacobian + b0 * nnz : nullptr;
    job.io.under_mask = io->under_mask ? io->under_mask + b0 * vw : nullptr;
    return job;
}

struct Worker {
    int device = 0;
    ezpz_context_t* ctx = nullptr;
    std::thread thread;
    std::mutex m;
    std::condition_variable cv;
    bool has_job = false, done = true, quit = false;
    std::atomic<uint32_t> posted{0};     // jobs handed to this worker so far (what a spinning worker watches)
    std::atomic<uint32_t> completed{0};  // jobs it has finished (what the caller watches)
    std::vector<Job> jobs;  // what the worker runs next, in order
    int32_t rc = EZPZ_OK;
    ezpz_error_detail_t detail{};
};

}  // namespace

struct ezpz_multi {
    std::vector<Worker*> workers;
    std::mutex call_mutex;  // one batch call at a time per multi-context
};

namespace {

// A wake-up through a condition variable costs tens of microseconds, as much as a shard's kernel.  A worker that has just
// finished a job therefore watches its counter for a short while (callers solve batch after batch) before it goes to sleep,
// and the caller watches the completion counters instead of sleeping on the workers' condition variables.
constexpr int kSpinIterations = 20000;  // ~100-200 us

void worker_main(Worker* w) {
    cudaSetDevice(w->device);
    uint32_t seen = 0;
    for (;;) {
        for (int k = 0; k < kSpinIterations && w->posted.load(std::memory_order_acquire) == seen; ++k) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        std::unique_lock<std::mutex> lock(w->m);
        w->cv.wait(lock, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        w->has_job = false;
        std::vector<Job> jobs;
        jobs.swap(w->jobs);
        lock.unlock();
        int32_t rc = EZPZ_OK;
        for (Job& job : jobs) {
            ezpz_error_detail_t det;
            const int32_t r = ezpz_b200_solve_batch(w->ctx, job.s, job.config, job.count, &job.io, &det);
            if (job.status) *job.status = r;
            if (r != EZPZ_OK && rc == EZPZ_OK) {
                rc = r;
                w->detail = det;
            }
        }
        lock.lock();
        w->rc = rc;
        w->done = true;
        seen = w->posted.load(std::memory_order_acquire);
        lock.unlock();
        w->completed.fetch_add(1, std::memory_order_release);
    }
}

}  // namespace

extern "C" {

int32_t ezpz_b200_multi_create(const int32_t* devices, int32_t n_devices, ezpz_multi_t** out, ezpz_error_detail_t* detail) {
    if (!out) return EZPZ_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (detail) std::memset(detail, 0, sizeof *detail);
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        if (detail) std::snprintf(detail->message, sizeof detail->message, "no CUDA device (%s)", cudaGetErrorString(e));
        cudaGetLastError();
        return EZPZ_ERR_NO_DEVICE;
    }
    std::vector<int> devs;
    if (!devices || n_devices <= 0) {
        const int want = n_devices > 0 ? std::min(n_devices, count) : count;
        for (int d = 0; d < want; ++d) devs.push_back(d);
    } else {
        for (int k = 0; k < n_devices; ++k) devs.push_back(devices[k]);
    }
    ezpz_multi* mg = new (std::nothrow) ezpz_multi();
    if (!mg) return EZPZ_ERR_INVALID_ARGUMENT;
    for (int d : devs) {
        Worker* w = new (std::nothrow) Worker();
        if (!w) {
            ezpz_b200_multi_destroy(mg);
            return EZPZ_ERR_INVALID_ARGUMENT;
        }
        w->device = d;
        const int32_t rc = ezpz_b200_context_create(d, &w->ctx, detail);
        if (rc != EZPZ_OK) {
            delete w;
            ezpz_b200_multi_destroy(mg);
            return rc;
        }
        w->thread = std::thread(worker_main, w);
        mg->workers.push_back(w);
    }
    *out = mg;
    return EZPZ_OK;
}

void ezpz_b200_multi_destroy(ezpz_multi_t* mg) {
    if (!mg) return;
    for (Worker* w : mg->workers) {
        {
            std::lock_guard<std::mutex> lock(w->m);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->thread.joinable()) w->thread.join();
        ezpz_b200_context_destroy(w->ctx);
        delete w;
    }
    delete mg;
}

int32_t ezpz_b200_multi_device_count(const ezpz_multi_t* mg) { return mg ? (int32_t)mg->workers.size() : 0; }

ezpz_context_t* ezpz_b200_multi_context(ezpz_multi_t* mg, int32_t index) {
    if (!mg || index < 0 || index >= (int32_t)mg->workers.size()) return nullptr;
    return mg->workers[(size_t)index]->ctx;
}

uint64_t ezpz_b200_multi_launches(const ezpz_multi_t* mg) {
    uint64_t total = 0;
    if (mg)
        for (const Worker* w : mg->workers) total += ezpz_b200_context_launches(w->ctx);
    return total;
}

// Hands list k to worker k and waits for all of them.  The calling thread runs the first non-empty list itself, on that
// worker's context (a context belongs to whoever solves with it, one at a time; the worker's own thread stays asleep): a
// one-worker multi-context costs no hand-off at all, and the other workers' wake-up latency hides behind the caller's own work.
static int32_t run_lists(ezpz_multi_t* mg, std::vector<std::vector<Job>>& lists, ezpz_error_detail_t* detail) {
    std::vector<Worker*> used;
    std::vector<uint32_t> target;
    Worker* own_worker = nullptr;
    std::vector<Job> own;
    for (size_t k = 0; k < lists.size(); ++k) {
        if (lists[k].empty()) continue;
        Worker* w = mg->workers[k];
        if (!own_worker) {
            own_worker = w;
            own.swap(lists[k]);
            continue;
        }
        {
            std::lock_guard<std::mutex> lock(w->m);
            w->jobs.swap(lists[k]);
            w->has_job = true;
            w->done = false;
            target.push_back(w->completed.load(std::memory_order_relaxed) + 1);
            w->posted.fetch_add(1, std::memory_order_release);
        }
        w->cv.notify_all();
        used.push_back(w);
    }
    int32_t rc = EZPZ_OK;
    for (Job& job : own) {
        ezpz_error_detail_t det;
        const int32_t r = ezpz_b200_solve_batch(own_worker->ctx, job.s, job.config, job.count, &job.io, &det);
        if (job.status) *job.status = r;
        if (r != EZPZ_OK && rc == EZPZ_OK) {
            rc = r;
            if (detail) *detail = det;
        }
    }
    for (size_t k = 0; k < used.size(); ++k) {
        Worker* w = used[k];
        for (uint64_t spins = 0; w->completed.load(std::memory_order_acquire) != target[k]; ++spins) {
            if (spins > 2000) std::this_thread::yield();
#if defined(__x86_64__)
            else __builtin_ia32_pause();
#endif
        }
        std::lock_guard<std::mutex> lock(w->m);
        if (w->rc != EZPZ_OK && rc == EZPZ_OK) {
            rc = w->rc;
            if (detail) *detail = w->detail;
        }
    }
    return rc;
}

int32_t ezpz_b200_solve_batch_multi(ezpz_multi_t* mg, const ezpz_structure_t* s, const ezpz_config_t* config, uint64_t batch,
                                    const ezpz_batch_io_t* io, ezpz_error_detail_t* detail) {
    if (!mg || !s || !config || !io || mg->workers.empty()) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (batch == 0) return EZPZ_OK;
    if (!io->guesses || !io->final_values || !io->iterations || !io->status) return EZPZ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> call(mg->call_mutex);
    const uint32_t world = (uint32_t)mg->workers.size();
    // Shards are whole groups of 32 problems (the batched kernel's unit) except the last.
    const uint64_t groups = (batch + 31) / 32;
    std::vector<std::vector<Job>> lists(world);
    for (uint32_t k = 0; k < world; ++k) {
        uint64_t g0 = 0, g1 = 0;
        ezpz_b200_shard_range(groups, k, world, &g0, &g1);
        const uint64_t b0 = std::min<uint64_t>(batch, g0 * 32), b1 = std::min<uint64_t>(batch, g1 * 32);
        if (b1 > b0) lists[k].push_back(slice_job(s, config, io, b0, b1));
    }
    return run_lists(mg, lists, detail);
}

// Several batches at once — the structure-homogeneous sub-batches of a mixed workload (BASELINE.json configs[4]).  Large jobs
// are cut over the workers like a single batch call; the others go whole to the least loaded worker, largest first, and every
// worker runs its list in order.  A multi-context may hold SEVERAL workers per device (ezpz_b200_multi_create with a device
// listed more than once): their streams overlap on the GPU, which is what keeps it busy when the sub-batches are small.
int32_t ezpz_b200_solve_jobs_multi(ezpz_multi_t* mg, const ezpz_config_t* config, ezpz_batch_job_t* jobs, uint32_t n_jobs,
                                   ezpz_error_detail_t* detail) {
    if (!mg || !config || (n_jobs && !jobs) || mg->workers.empty()) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    for (uint32_t j = 0; j < n_jobs; ++j) {
        jobs[j].status = EZPZ_OK;
        if (!jobs[j].structure) return EZPZ_ERR_INVALID_ARGUMENT;
        if (jobs[j].batch && (!jobs[j].io.guesses || !jobs[j].io.final_values || !jobs[j].io.iterations || !jobs[j].io.status))
            return EZPZ_ERR_INVALID_ARGUMENT;
    }
    std::lock_guard<std::mutex> call(mg->call_mutex);
    const uint32_t world = (uint32_t)mg->workers.size();
    std::vector<std::vector<Job>> lists(world);
    std::vector<uint64_t> load(world, 0);  // modelled cost per worker: problems x constraints
    std::vector<uint32_t> order;
    for (uint32_t j = 0; j < n_jobs; ++j)
        if (jobs[j].batch) order.push_back(j);
    auto cost = [&](uint32_t j) { return jobs[j].batch * (uint64_t)std::max<uint32_t>(1, jobs[j].structure->n_cons); };
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cost(x) > cost(y); });
    constexpr uint64_t kSplitProblems = 65536;  // jobs of this size and more are cut over all workers
    for (uint32_t j : order) {
        const ezpz_batch_job_t& jb = jobs[j];
        if (world > 1 && jb.batch >= kSplitProblems) {
            const uint64_t groups = (jb.batch + 31) / 32;
            for (uint32_t k = 0; k < world; ++k) {
                uint64_t g0 = 0, g1 = 0;
                ezpz_b200_shard_range(groups, k, world, &g0, &g1);
                const uint64_t b0 = std::min<uint64_t>(jb.batch, g0 * 32), b1 = std::min<uint64_t>(jb.batch, g1 * 32);
                if (b1 <= b0) continue;
                Job job = slice_job(jb.structure, config, &jb.io, b0, b1);
                job.status = &jobs[j].status;  // (any failing slice marks the job)
                lists[k].push_back(job);
                load[k] += (b1 - b0) * (uint64_t)std::max<uint32_t>(1, jb.structure->n_cons);
            }
        } else {
            const uint32_t k = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
            Job job = slice_job(jb.structure, config, &jb.io, 0, jb.batch);
            job.status = &jobs[j].status;
            lists[k].push_back(job);
            load[k] += cost(j);
        }
    }
    return run_lists(mg, lists, detail);
}

int32_t ezpz_b200_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return EZPZ_ERR_INVALID_ARGUMENT;
    const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return EZPZ_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? EZPZ_ERR_NO_DEVICE : EZPZ_ERR_CUDA;
    }
    return EZPZ_OK;
}

int32_t ezpz_b200_host_unregister(void* ptr) {
    if (!ptr) return EZPZ_ERR_INVALID_ARGUMENT;
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return EZPZ_ERR_CUDA;
    }
    return EZPZ_OK;
}

int32_t ezpz_b200_host_alloc(uint64_t bytes, void** out) {
    if (!out) return EZPZ_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? EZPZ_ERR_NO_DEVICE : EZPZ_ERR_CUDA;
    }
    return EZPZ_OK;
}

void ezpz_b200_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
