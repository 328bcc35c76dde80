// freedom.cu — the "underconstrained" verdict on the device: batched freedom analysis.
//
// Follows ezpz/src/solver/find_dof.rs:15-104: densify J (values cached at the last accepted point),
// column-pivoted QR, rank = number of leading |R_ii| > 1e-8 * max|R_ii| (take_while), a basis of
// null(J P^T) by back substitution, un-permute, orthonormalise, and flag variable j when the squared
// norm of row j of the orthonormal basis exceeds (1e-3 * max)^2.  faer's ColPivQr / thin-Q are not in
// tree; here: Householder QR with the largest-remaining-column-norm pivot rule and modified
// Gram-Schmidt (twice).  The participation norms are the diagonal of the orthogonal projector onto
// null(J), so they do not depend on which orthonormal basis is produced.
//
// One thread per problem; each thread's dense work arrays live in a global scratch buffer interleaved
// [element][thread] so that a warp touching element e of its 32 problems reads one contiguous 256-byte
// row.  This is a verdict pass run once per structure change (lib.rs:86-90 says so), not the inner loop.
#include <algorithm>
#include <cstring>

#include "device.h"
#include "dmath.cuh"

namespace {

struct FreedomArgs {
    const uint32_t* csc_col_ptr;
    const uint32_t* csc_row_idx;
    const double* jac;   // [count * nnz]
    uint32_t* mask;      // [count * words]
    double* scratch;     // [per_thread * threads]
    uint32_t m, n, nnz, words, count, threads;
};

__global__ void __launch_bounds__(128) freedom_kernel(const FreedomArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const uint32_t m = a.m, n = a.n, TT = a.threads;
    double* A = a.scratch + t;                        // A(i,j) = A[(j*m + i) * TT]
    double* N = A + (size_t)m * n * TT;               // N(i,f) = N[(f*n + i) * TT]
    double* perm = N + (size_t)n * n * TT;            // [n]
    double* rdiag = perm + (size_t)n * TT;            // [min(m,n)]
    double* part = rdiag + (size_t)(m < n ? m : n) * TT;  // [n]
#define AE(i, j) A[((size_t)(j) * m + (i)) * TT]
#define NE(i, f) N[((size_t)(f) * n + (i)) * TT]
    for (uint32_t j = 0; j < n; ++j) {
        for (uint32_t i = 0; i < m; ++i) AE(i, j) = 0.0;
        for (uint32_t e = a.csc_col_ptr[j]; e < a.csc_col_ptr[j + 1]; ++e)
            AE(a.csc_row_idx[e], j) = a.jac[(size_t)t * a.nnz + e];
        perm[(size_t)j * TT] = (double)j;
        part[(size_t)j * TT] = 0.0;
    }
    const uint32_t ndiag = m < n ? m : n;
    for (uint32_t k = 0; k < ndiag; ++k) {
        uint32_t best = k;
        double bestn = -1.0;
        for (uint32_t j = k; j < n; ++j) {
            double s = 0.0;
            for (uint32_t i = k; i < m; ++i) s += AE(i, j) * AE(i, j);
            if (s > bestn) {
                bestn = s;
                best = j;
            }
        }
        if (best != k) {
            for (uint32_t i = 0; i < m; ++i) {
                const double tmp = AE(i, k);
                AE(i, k) = AE(i, best);
                AE(i, best) = tmp;
            }
            const double tp = perm[(size_t)k * TT];
            perm[(size_t)k * TT] = perm[(size_t)best * TT];
            perm[(size_t)best * TT] = tp;
        }
        const double norm = sqrt(bestn);
        if (norm == 0.0) {
            rdiag[(size_t)k * TT] = 0.0;
            continue;
        }
        const double akk = AE(k, k);
        const double alpha = akk > 0 ? -norm : norm;
        const double vk = akk - alpha;
        for (uint32_t i = k + 1; i < m; ++i) AE(i, k) = AE(i, k) / vk;
        const double tau = -vk / alpha;
        AE(k, k) = alpha;
        rdiag[(size_t)k * TT] = alpha;
        for (uint32_t j = k + 1; j < n; ++j) {
            double s = AE(k, j);
            for (uint32_t i = k + 1; i < m; ++i) s += AE(i, k) * AE(i, j);
            s *= tau;
            AE(k, j) -= s;
            for (uint32_t i = k + 1; i < m; ++i) AE(i, j) -= s * AE(i, k);
        }
    }
    uint32_t* mask = a.mask + (size_t)t * a.words;
    for (uint32_t w = 0; w < a.words; ++w) mask[w] = 0;
    double largest = ezm::ez_abs(rdiag[0]);
    for (uint32_t i = 1; i < ndiag; ++i) largest = ezm::ez_fmax(largest, ezm::ez_abs(rdiag[(size_t)i * TT]));
    const double tolerance = 1e-8 * largest;
    uint32_t rank = 0;
    while (rank < ndiag && ezm::ez_abs(rdiag[(size_t)rank * TT]) > tolerance) ++rank;
    const uint32_t nullity = n - rank;
    if (nullity == 0) return;
    for (uint32_t f = 0; f < nullity; ++f) {
        const uint32_t free_var = rank + f;
        for (uint32_t i = 0; i < n; ++i) NE(i, f) = 0.0;
        NE(free_var, f) = 1.0;
        for (uint32_t ii = rank; ii-- > 0;) {
            double rhs = AE(ii, free_var);
            for (uint32_t j = ii + 1; j < rank; ++j) rhs += AE(ii, j) * NE(j, f);
            NE(ii, f) = -rhs / AE(ii, ii);
        }
    }
    for (uint32_t f = 0; f < nullity; ++f) {
        for (int pass = 0; pass < 2; ++pass)
            for (uint32_t g = 0; g < f; ++g) {
                double s = 0.0;
                for (uint32_t i = 0; i < n; ++i) s += NE(i, g) * NE(i, f);
                for (uint32_t i = 0; i < n; ++i) NE(i, f) -= s * NE(i, g);
            }
        double s = 0.0;
        for (uint32_t i = 0; i < n; ++i) s += NE(i, f) * NE(i, f);
        s = sqrt(s);
        for (uint32_t i = 0; i < n; ++i) NE(i, f) = NE(i, f) / s;
    }
    for (uint32_t f = 0; f < nullity; ++f)
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t var = (uint32_t)perm[(size_t)i * TT];
            part[(size_t)var * TT] += NE(i, f) * NE(i, f);
        }
    double max_p = 0.0;
    for (uint32_t j = 0; j < n; ++j) max_p = ezm::ez_fmax(max_p, part[(size_t)j * TT]);
    const double var_tol = 1e-3 * max_p;
    const double squared_tol = var_tol * var_tol;
    for (uint32_t j = 0; j < n; ++j)
        if (part[(size_t)j * TT] > squared_tol) mask[j >> 5] |= 1u << (j & 31u);
#undef AE
#undef NE
}

}  // namespace

extern "C" int32_t ezpz_b200_freedom_analysis(ezpz_context_t* ctx, const ezpz_structure_t* s, uint64_t batch,
                                              const double* jacobian, uint32_t* under_mask,
                                              ezpz_error_detail_t* detail) {
    if (!ctx || !s || !under_mask || (batch && !jacobian)) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    const uint32_t m = s->m, n = s->n;
    if (std::min(m, n) == 0) return EZPZ_ERR_EMPTY_SYSTEM;  // find_dof.rs:41-44
    if (n > 256) {
        if (detail) std::snprintf(detail->message, sizeof detail->message, "freedom analysis is dense O(m n^2); n = %u > 256", n);
        return EZPZ_ERR_TOO_LARGE;
    }
    if (batch == 0) return EZPZ_OK;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t nnz = s->csc_row_idx.size();
    const uint32_t words = (n + 31) / 32;
    const size_t per_thread = (size_t)m * n + (size_t)n * n + 2 * (size_t)n + std::min(m, n);
    // chunk so that the scratch stays below 256 MiB
    size_t chunk = std::max<size_t>(128, ((size_t)256 << 20) / (per_thread * 8) / 128 * 128);
    chunk = std::min<size_t>(chunk, (batch + 127) / 128 * 128);
    const size_t b_ptr = ezs::align_up((n + 1) * 4, 256), b_idx = ezs::align_up(nnz * 4, 256);
    const size_t b_jac = ezs::align_up(chunk * nnz * 8, 256), b_mask = ezs::align_up(chunk * words * 4, 256);
    const size_t b_scr = ezs::align_up(per_thread * chunk * 8, 256);
    int32_t rc = ezs::ensure_ws(ctx, b_ptr + b_idx + b_jac + b_mask + b_scr, detail);
    if (rc != EZPZ_OK) return rc;
    char* w = (char*)ctx->ws;
    uint32_t* d_ptr = (uint32_t*)w; w += b_ptr;
    uint32_t* d_idx = (uint32_t*)w; w += b_idx;
    double* d_jac = (double*)w; w += b_jac;
    uint32_t* d_mask = (uint32_t*)w; w += b_mask;
    double* d_scr = (double*)w;
    cudaStream_t st = ctx->stream;
    EZ_CUDA(cudaMemcpyAsync(d_ptr, s->csc_col_ptr.data(), (n + 1) * 4, cudaMemcpyHostToDevice, st), "H2D col_ptr");
    if (nnz) EZ_CUDA(cudaMemcpyAsync(d_idx, s->csc_row_idx.data(), nnz * 4, cudaMemcpyHostToDevice, st), "H2D row_idx");
    for (uint64_t first = 0; first < batch; first += chunk) {
        const uint32_t count = (uint32_t)std::min<uint64_t>(chunk, batch - first);
        if (nnz) EZ_CUDA(cudaMemcpyAsync(d_jac, jacobian + first * nnz, (size_t)count * nnz * 8, cudaMemcpyHostToDevice, st), "H2D jacobian");
        FreedomArgs a;
        a.csc_col_ptr = d_ptr;
        a.csc_row_idx = d_idx;
        a.jac = d_jac;
        a.mask = d_mask;
        a.scratch = d_scr;
        a.m = m;
        a.n = n;
        a.nnz = (uint32_t)nnz;
        a.words = words;
        a.count = count;
        a.threads = (uint32_t)chunk;
        freedom_kernel<<<(count + 127) / 128, 128, 0, st>>>(a);
        ctx->launches += 1;
        EZ_CUDA(cudaGetLastError(), "freedom_kernel launch");
        EZ_CUDA(cudaMemcpyAsync(under_mask + first * words, d_mask, (size_t)count * words * 4, cudaMemcpyDeviceToHost, st), "D2H mask");
        EZ_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    }
    return EZPZ_OK;
}
