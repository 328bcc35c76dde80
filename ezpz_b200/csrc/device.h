// device.h — declarations shared by the CUDA translation units of libezpz_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

#include "structure.h"

namespace ezs {

// The small program compiled for one (roles, shared-memory stride) pair (structure.h: RoleBlob) with its device copy.
struct RoleTables {
    RoleBlob blob;
    uint32_t* dev = nullptr;
};

// Device-resident copy of an analysed structure (one per CUDA device, created lazily).
struct DeviceCopy {
    int device = -1;
    DevCons* cons = nullptr;         // [n_cons]
    std::vector<RoleTables*> roles;  // one per (roles, stride) used so far
    uint32_t* csc_to_csr = nullptr;  // [nnz] position in CSR order of each CSC entry
    uint32_t* csc_col_ptr = nullptr;  // [n + 1], [nnz]: the pattern of J for the freedom analysis (freedom.cu), on first use
    uint32_t* csc_row_idx = nullptr;
    // LargeDevice (large.cu: the structure's tables AND the work buffers of a solve), one per CONTEXT that solved with this
    // structure on this device: two contexts (threads) may solve the same structure on one device at the same time
    std::vector<std::pair<const void*, void*>> large;
};

int32_t cuda_fail(cudaError_t e, ezpz_error_detail_t* detail, const char* what);
int32_t get_device_copy(ezpz_context* ctx, const ezpz_structure* s, DeviceCopy** out, ezpz_error_detail_t* detail);
int32_t ensure_ws(ezpz_context* ctx, size_t bytes, ezpz_error_detail_t* detail);
int32_t ensure_pin(ezpz_context* ctx, size_t bytes, ezpz_error_detail_t* detail);
int32_t ensure_fa(ezpz_context* ctx, size_t bytes, ezpz_error_detail_t* detail);
// Freedom analysis on device-resident Jacobians, enqueued on st, no host round trip (freedom.cu).
int32_t freedom_device(ezpz_context* ctx, const ezpz_structure* s, uint64_t batch, const double* d_jac, uint32_t* d_mask,
                       cudaStream_t st, ezpz_error_detail_t* detail);
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
void release_large(DeviceCopy* d);  // large.cu
// Single system that does not fit the thread-per-problem kernel (large.cu).
int32_t solve_large(ezpz_context* ctx, const ezpz_structure* s, const ezpz_config_t* config, const ezpz_one_io_t* io,
                    ezpz_error_detail_t* detail);

// A batch of such systems, one CTA per problem; device pointers in io, enqueued on st (large.cu).
int32_t solve_large_batch(ezpz_context* ctx, const ezpz_structure* s, const ezpz_config_t* config, uint64_t batch,
                          const ezpz_batch_io_t* io, cudaStream_t st, ezpz_error_detail_t* detail);

// ezpz_b200_eval for structures with a large programme: the large path's own assembly kernel (large.cu).
int32_t eval_large(ezpz_context* ctx, const ezpz_structure* s, const double* x, double* r, double* jac_csc, double* jac_csr,
                   bool* have_csr, ezpz_error_detail_t* detail);

}  // namespace ezs

struct ezpz_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};  // copy/compute pipeline of the host-buffer batch call
    cudaEvent_t pipe_done[3] = {nullptr, nullptr, nullptr};
    uint64_t shape_batch = 0;  // nonzero: launch_small sizes its CTAs for a batch of this many problems instead of the call's (lane pipeline)
    std::vector<cudaEvent_t> lane_ev;  // events of the lane pipeline (copy-in lane -> kernel lanes -> copy-out lane), made on demand
    int sm_count = 0;
    size_t smem_optin = 0;
    uint64_t launches = 0;
    // grow-only device workspace for the host-buffer entry points
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // grow-only pinned host staging buffer (small single-system solves: one DMA each way instead of pageable copies)
    void* pin = nullptr;
    size_t pin_bytes = 0;
    // freedom analysis (freedom.cu): grow-only scratch of the dense factorisations, the Jacobians of a fused solve + analysis
    // call, and the event that orders launches of different streams on them
    void* fa_ws = nullptr;
    size_t fa_bytes = 0;
    void* fa_jac = nullptr;
    size_t fa_jac_bytes = 0;
    cudaEvent_t fa_done = nullptr;
    bool fa_busy = false;
    cudaStream_t fa_last_stream = nullptr;
    // topology cache of ezpz_b200_solve (host_api.cpp): analysed structures keyed by their constraint list
    void* structure_cache = nullptr;
};

namespace ezs {
void release_structure_cache(ezpz_context* ctx);  // host_api.cpp
}

#define EZ_CUDA(call, what)                                                \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return ezs::cuda_fail(e__, detail, what);  \
    } while (0)
