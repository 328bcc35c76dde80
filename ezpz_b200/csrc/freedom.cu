// freedom.cu — the "underconstrained" verdict on the device: freedom analysis for any number of variables.
//
// Follows ezpz/src/solver/find_dof.rs:15-104: densify J (values cached at the last accepted point),
// column-pivoted QR, rank = number of leading |R_ii| > 1e-8 * max|R_ii| (take_while), a basis of
// null(J P^T) by back substitution, un-permute, orthonormalise, and flag variable j when the squared
// norm of row j of the orthonormal basis exceeds (1e-3 * max)^2.  faer's ColPivQr / thin-Q are not in
// tree; here: Householder QR with the largest-remaining-column-norm pivot rule (first maximum wins) and
// Gram-Schmidt with re-orthogonalisation.  The participation norms are the diagonal of the orthogonal
// projector onto null(J), so they do not depend on which orthonormal basis is produced.
//
// One kernel, freedom_team_kernel, run by a TEAM of threads per problem; three deployments, by size:
//   * a warp / CTA per problem with the dense matrix in SHARED memory (n up to ~64: batches of small sketches,
//     config 5 of BASELINE.json) — tens of CTAs per SM, no global scratch traffic;
//   * a CTA per problem with the matrix in a global (L2-resident) scratch slot (n up to 1,024, batches);
//   * the whole grid on ONE problem (cooperative launch, grid barrier; single systems of any size — the
//     reference benches its analysis at 200 variables and solves massive_parallel_system at 2,400).
// The matrix is stored ROW-major and a thread owns a COLUMN: the sums of a column run over its rows in ascending
// order in one thread — exactly the oracle's order (oracle/ezpz_oracle.cpp freedom_analysis), so the QR, its
// pivot choices (ties included) and the rank decision are bit-identical to the oracle — while the 32 threads of
// a warp read 32 consecutive doubles of a row.  A Householder step is two barriers: (swap the pivot column in and
// form v) | (per column: dot with v, update, and the squared norm of what remains, fused in one sweep).
// Orthonormalisation uses parallel dot products (not bit-identical to the oracle's sequential ones; the
// verdict's thresholds are relative 1e-6).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <utility>

#include "device.h"
#include "dmath.cuh"

namespace cg = cooperative_groups;

namespace {

struct FreedomArgs {
    const uint32_t* csc_col_ptr;
    const uint32_t* csc_row_idx;
    const double* jac;  // [count * nnz]
    uint32_t* mask;     // [count * words]
    double* scratch;    // [slots * per_problem] (global deployments)
    uint64_t per_problem;  // doubles of scratch per problem
    uint32_t m, n, nnz, words, count;
    uint32_t in_smem;   // matrix and work vectors in dynamic shared memory
    uint32_t v_smem;    // doubles of shared memory available for staging v in the global deployments (0 = none)
};

// Work layout of one problem (doubles): A[m*n] row-major | N[n*n] row-major n x nullity | v[m] | norms[n] | dk[n] | perm[n] |
// rdiag[n] | part[n] | s[n] | red[one per CTA of the team]
constexpr uint32_t kMaxTeamCtas = 256;
__host__ __device__ inline uint64_t freedom_doubles(uint64_t m, uint64_t n, bool grid) {
    return m * n + n * n + m + 6 * n + 8 + (grid ? kMaxTeamCtas : 8);
}

template <bool GRID>
struct Team {
    uint32_t tid, size;
    __device__ Team() {
        if (GRID) {
            tid = blockIdx.x * blockDim.x + threadIdx.x;
            size = gridDim.x * blockDim.x;
        } else {
            tid = threadIdx.x;
            size = blockDim.x;
        }
    }
    __device__ void sync() const {
        if (GRID) cg::this_grid().sync();
        else __syncthreads();
    }
};

// Loads of data other CTAs of the team may have written go to L2 in the grid deployment.
template <bool GRID>
__device__ __forceinline__ double ld(const double* p) {
    if (GRID) return __ldcg(p);
    return *p;
}

template <bool GRID>
__device__ void freedom_problem(const FreedomArgs& a, const Team<GRID>& team, double* W, const double* jac, uint32_t* mask,
                                double* vstage, uint32_t* sh_u, double* sh_d) {
    const uint32_t m = a.m, n = a.n;
    const uint32_t ndiag = m < n ? m : n;
    double* A = W;
    double* N = A + (size_t)m * n;
    double* v = N + (size_t)n * n;
    double* norms = v + m;
    double* dk = norms + n + 8;  // (norms holds the pivot candidates: 2 * (n / 32 + 1) <= n + 8 values)
    double* perm = dk + n;
    double* rdiag = perm + n;
    double* part = rdiag + n;
    double* sdot = part + n;
    double* red = sdot + n;  // [256] per-CTA partial results of team-wide reductions
    // ---- densify (thread per column; find_dof.rs:16-18)
    for (uint32_t j = team.tid; j < n; j += team.size) {
        for (uint32_t i = 0; i < m; ++i) A[(size_t)i * n + j] = 0.0;
        for (uint32_t e = a.csc_col_ptr[j]; e < a.csc_col_ptr[j + 1]; ++e) A[(size_t)a.csc_row_idx[e] * n + j] = jac[e];
        perm[j] = (double)j;
    }
    team.sync();
    // ---- column-pivoted Householder QR
    // Pivot candidates: every warp that produces the squared norms of 32 consecutive remaining columns leaves their maximum
    // (first one among equals) in cand[]; a pivot search then reads (n - k) / 32 pairs instead of n - k norms.
    const uint32_t lane = threadIdx.x & 31u, warp_first = team.tid - lane;
    double* cand_v = norms;
    double* cand_i = norms + (n + 31) / 32 + 1;
    auto leave_candidate = [&](bool act, double nrm, uint32_t j, uint32_t slot) {
        double bn = (act && nrm == nrm) ? nrm : -1.0;  // (a NaN norm never wins, as in `s > bestn`)
        uint32_t bj = (act && nrm == nrm) ? j : 0xffffffffu;
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, bn, off);
            const uint32_t oj = __shfl_down_sync(0xffffffffu, bj, off);
            if (ob > bn || (ob == bn && oj < bj)) {
                bn = ob;
                bj = oj;
            }
        }
        if (lane == 0) {
            cand_v[slot] = bn;
            cand_i[slot] = (double)bj;
        }
    };
    constexpr int U = GRID ? 16 : 8;  // loads in flight per thread in the column sweeps
    bool norms_valid = false;
    for (uint32_t k = 0; k < ndiag; ++k) {
        if (!norms_valid) {  // squared norms of the remaining columns over rows k.. (first step, or after a zero pivot)
            for (uint32_t base = k + warp_first; base < n; base += team.size) {
                const uint32_t j = base + lane;
                const bool act = j < n;
                double s = 0.0;
                if (act) {
                    const double* col = A + j;
                    uint32_t i = k;
                    for (; i + U <= m; i += U) {
                        double x[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) x[u] = ld<GRID>(&col[(size_t)(i + u) * n]);
#pragma unroll
                        for (int u = 0; u < U; ++u) s += x[u] * x[u];
                    }
                    for (; i < m; ++i) {
                        const double x = ld<GRID>(&col[(size_t)i * n]);
                        s += x * x;
                    }
                    dk[j] = ld<GRID>(&col[(size_t)k * n]);
                }
                leave_candidate(act, s, j, (base - k) / 32);
            }
            team.sync();
        }
        // pivot: the largest remaining norm, the first one among equals (every CTA computes it for itself)
        double bestn = -1.0;
        uint32_t best = k;
        const uint32_t n_cand = (n - k + 31) / 32;
        for (uint32_t c = threadIdx.x; c < n_cand; c += blockDim.x) {
            const double s = ld<GRID>(&cand_v[c]);
            const uint32_t j = (uint32_t)ld<GRID>(&cand_i[c]);
            if (s > bestn || (s == bestn && j < best)) {
                bestn = s;
                best = j;
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, bestn, off);
            const uint32_t oj = __shfl_down_sync(0xffffffffu, best, off);
            if (ob > bestn || (ob == bestn && oj < best)) {
                bestn = ob;
                best = oj;
            }
        }
        if (blockDim.x > 32) {
            __syncthreads();
            if (lane == 0) {
                sh_d[threadIdx.x >> 5] = bestn;
                sh_u[threadIdx.x >> 5] = best;
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                const uint32_t nw = blockDim.x >> 5;
                bestn = threadIdx.x < nw ? sh_d[threadIdx.x] : -2.0;
                best = threadIdx.x < nw ? sh_u[threadIdx.x] : 0xffffffffu;
                for (int off = 16; off > 0; off >>= 1) {
                    const double ob = __shfl_down_sync(0xffffffffu, bestn, off);
                    const uint32_t oj = __shfl_down_sync(0xffffffffu, best, off);
                    if (ob > bestn || (ob == bestn && oj < best)) {
                        bestn = ob;
                        best = oj;
                    }
                }
                if (threadIdx.x == 0) {
                    sh_d[32] = bestn;
                    sh_u[32] = best;
                }
            }
            __syncthreads();
            bestn = sh_d[32];
            best = sh_u[32];
        } else {
            bestn = __shfl_sync(0xffffffffu, bestn, 0);
            best = __shfl_sync(0xffffffffu, best, 0);
        }
        const double norm = sqrt(bestn);
        const bool zero = norm == 0.0;
        const double akk = ld<GRID>(&dk[best]);  // A(k, best) before the swap
        const double alpha = akk > 0 ? -norm : norm;
        const double vk = akk - alpha;
        const double tau = -vk / alpha;
        // swap columns k and best in every row; form v below the diagonal (thread per row)
        for (uint32_t i = team.tid; i < m; i += team.size) {
            const double ab = ld<GRID>(&A[(size_t)i * n + best]);
            if (best != k) A[(size_t)i * n + best] = ld<GRID>(&A[(size_t)i * n + k]);
            if (zero || i < k) A[(size_t)i * n + k] = ab;
            else if (i == k) A[(size_t)i * n + k] = alpha;
            else v[i] = ab / vk;
        }
        if (team.tid == 0) {
            const double tp = perm[k];
            perm[k] = perm[best];
            perm[best] = tp;
            rdiag[k] = zero ? 0.0 : alpha;
        }
        team.sync();
        if (zero) {
            norms_valid = false;
            continue;
        }
        // per column j > k: s = (A(k,j) + sum_i v_i A(i,j)) tau;  A(k,j) -= s;  A(i,j) -= s v_i;  next norm over rows k+1..
        const double* vv = v;
        if (vstage) {  // global deployments: the CTA stages v in shared memory once per step
            const uint32_t len = m - (k + 1);
            if (len <= a.v_smem) {
                for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) vstage[i] = ld<GRID>(&v[k + 1 + i]);
                __syncthreads();
                vv = vstage - (k + 1);
            }
        }
        const bool v_l2 = GRID && vv == v;
        for (uint32_t base = k + 1 + warp_first; base < n; base += team.size) {
            const uint32_t j = base + lane;
            const bool act = j < n;
            double nrm = 0.0;
            if (act) {
                double* col = A + j;
                double s = ld<GRID>(&col[(size_t)k * n]);
                const double akj = s;
                uint32_t i = k + 1;
                for (; i + U <= m; i += U) {
                    double x[U], w[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        x[u] = ld<GRID>(&col[(size_t)(i + u) * n]);
                        w[u] = v_l2 ? __ldcg(&vv[i + u]) : vv[i + u];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) s += w[u] * x[u];
                }
                for (; i < m; ++i) s += (v_l2 ? __ldcg(&vv[i]) : vv[i]) * ld<GRID>(&col[(size_t)i * n]);
                s *= tau;
                col[(size_t)k * n] = akj - s;
                double first = 0.0;
                i = k + 1;
                for (; i + U <= m; i += U) {
                    double x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u] = ld<GRID>(&col[(size_t)(i + u) * n]) - s * (v_l2 ? __ldcg(&vv[i + u]) : vv[i + u]);
                    if (i == k + 1) first = x[0];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        col[(size_t)(i + u) * n] = x[u];
                        nrm += x[u] * x[u];
                    }
                }
                for (; i < m; ++i) {
                    const double x = ld<GRID>(&col[(size_t)i * n]) - s * (v_l2 ? __ldcg(&vv[i]) : vv[i]);
                    col[(size_t)i * n] = x;
                    nrm += x * x;
                    if (i == k + 1) first = x;
                }
                dk[j] = first;
            }
            leave_candidate(act, nrm, j, (base - (k + 1)) / 32);
        }
        norms_valid = true;
        team.sync();
    }
    // ---- rank (find_dof.rs:40-52)
    double largest = ezm::ez_abs(ld<GRID>(&rdiag[0]));
    for (uint32_t i = 1; i < ndiag; ++i) largest = ezm::ez_fmax(largest, ezm::ez_abs(ld<GRID>(&rdiag[i])));
    const double tolerance = 1e-8 * largest;
    uint32_t rank = 0;
    while (rank < ndiag && ezm::ez_abs(ld<GRID>(&rdiag[rank])) > tolerance) ++rank;
    const uint32_t F = n - rank;
    if (F == 0) {
        for (uint32_t w = team.tid; w < a.words; w += team.size) mask[w] = 0;
        team.sync();  // (the work area is reused by the next problem of this team)
        return;
    }
    // ---- basis of null(J P^T): thread per free column f, back substitution over R11 (find_dof.rs:56-72)
    for (uint32_t f = team.tid; f < F; f += team.size) {
        for (uint32_t i = rank; i < n; ++i) N[(size_t)i * F + f] = (i == rank + f) ? 1.0 : 0.0;
        for (uint32_t ii = rank; ii-- > 0;) {
            const double* Rrow = A + (size_t)ii * n;
            double rhs = ld<GRID>(&Rrow[rank + f]);
            for (uint32_t j = ii + 1; j < rank; ++j) rhs += ld<GRID>(&Rrow[j]) * ld<GRID>(&N[(size_t)j * F + f]);
            N[(size_t)ii * F + f] = -rhs / ld<GRID>(&Rrow[ii]);
        }
    }
    team.sync();
    // ---- orthonormalise the columns of N: Gram-Schmidt against all earlier columns at once, twice, then normalise
    const uint32_t cta = GRID ? blockIdx.x : 0, n_cta = GRID ? gridDim.x : 1;
    for (uint32_t f = 0; f < F; ++f) {
        for (int pass = 0; pass < 2 && f > 0; ++pass) {
            for (uint32_t g = team.tid; g < f; g += team.size) {  // thread per earlier column
                double s = 0.0;
                for (uint32_t i = 0; i < n; ++i) s += ld<GRID>(&N[(size_t)i * F + g]) * ld<GRID>(&N[(size_t)i * F + f]);
                sdot[g] = s;
            }
            team.sync();
            for (uint32_t i = team.tid; i < n; i += team.size) {  // thread per row
                double* row = N + (size_t)i * F;
                double z = ld<GRID>(&row[f]);
                for (uint32_t g = 0; g < f; ++g) z -= ld<GRID>(&sdot[g]) * ld<GRID>(&row[g]);
                row[f] = z;
            }
            team.sync();
        }
        double s = 0.0;
        for (uint32_t i = team.tid; i < n; i += team.size) {
            const double z = ld<GRID>(&N[(size_t)i * F + f]);
            s += z * z;
        }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        __syncthreads();
        if ((threadIdx.x & 31u) == 0) sh_d[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) t += sh_d[w];
            red[cta] = t;
        }
        team.sync();
        double total = 0.0;
        for (uint32_t c = 0; c < n_cta; ++c) total += ld<GRID>(&red[c]);
        const double nrm = sqrt(total);
        team.sync();  // (red is rewritten for the next column)
        for (uint32_t i = team.tid; i < n; i += team.size) N[(size_t)i * F + f] = ld<GRID>(&N[(size_t)i * F + f]) / nrm;
        team.sync();
    }
    // ---- participation of every ORIGINAL variable, threshold, mask (find_dof.rs:82-104)
    double pmax = 0.0;
    for (uint32_t i = team.tid; i < n; i += team.size) {
        const double* row = N + (size_t)i * F;
        double p = 0.0;
        for (uint32_t f = 0; f < F; ++f) {
            const double z = ld<GRID>(&row[f]);
            p += z * z;
        }
        part[(uint32_t)ld<GRID>(&perm[i])] = p;
        pmax = ezm::ez_fmax(pmax, p);
    }
    for (int off = 16; off > 0; off >>= 1) pmax = ezm::ez_fmax(pmax, __shfl_down_sync(0xffffffffu, pmax, off));
    __syncthreads();
    if ((threadIdx.x & 31u) == 0) sh_d[threadIdx.x >> 5] = pmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) t = ezm::ez_fmax(t, sh_d[w]);
        red[cta] = t;
    }
    team.sync();
    double max_p = 0.0;
    for (uint32_t c = 0; c < n_cta; ++c) max_p = ezm::ez_fmax(max_p, ld<GRID>(&red[c]));
    const double var_tol = 1e-3 * max_p;
    const double squared_tol = var_tol * var_tol;
    for (uint32_t w = team.tid; w < a.words; w += team.size) {
        uint32_t bits = 0;
        for (uint32_t b = 0; b < 32 && w * 32 + b < n; ++b)
            if (ld<GRID>(&part[w * 32 + b]) > squared_tol) bits |= 1u << b;
        mask[w] = bits;
    }
    team.sync();
}

template <bool GRID>
__global__ void __launch_bounds__(GRID ? 256 : 1024) freedom_team_kernel(const FreedomArgs a) {
    extern __shared__ double fsm[];
    __shared__ double sh_d[33];
    __shared__ uint32_t sh_u[33];
    const Team<GRID> team;
    if constexpr (GRID) {
        for (uint32_t p = 0; p < a.count; ++p)
            freedom_problem<true>(a, team, a.scratch, a.jac + (size_t)p * a.nnz, a.mask + (size_t)p * a.words,
                                  a.v_smem ? fsm : nullptr, sh_u, sh_d);
    } else {
        for (uint32_t p = blockIdx.x; p < a.count; p += gridDim.x) {
            double* W = a.in_smem ? fsm : a.scratch + (size_t)blockIdx.x * a.per_problem;
            freedom_problem<false>(a, team, W, a.jac + (size_t)p * a.nnz, a.mask + (size_t)p * a.words,
                                   (!a.in_smem && a.v_smem) ? fsm : nullptr, sh_u, sh_d);
        }
    }
}

// One Householder step of the register-resident QR (freedom_warp_kernel); K is a template parameter so that every index
// into the per-lane arrays is a compile-time constant (a runtime k would push the arrays into local memory).
// A warp holds 32 / MAXD problems side by side: SUB-WARPS of MAXD lanes (lane `sl` of the sub-warp = column sl and, in the
// division exchange, row sl).  Every shuffle, ballot and barrier is confined to the sub-warp (`seg` = its lane mask, `shift`
// = its first lane), so sub-warps may take different data-dependent paths (zero pivots, ranks) without waiting for each other.
struct SubWarp {
    unsigned seg;    // lane mask of this sub-warp
    uint32_t shift;  // its first lane
    uint32_t sl;     // this lane's index inside it
};
template <int MAXD>
__device__ __forceinline__ uint32_t sub_ballot(const SubWarp& w, bool pred) {
    return (__ballot_sync(w.seg, pred) >> w.shift) & ((MAXD >= 32) ? 0xffffffffu : ((1u << MAXD) - 1u));
}
template <int K, int MAXD>
__device__ __forceinline__ void qr_step(double (&A)[MAXD], double (&rd)[MAXD], uint32_t& pos, double& nrm, const SubWarp& w, bool active,
                                        uint32_t m, uint32_t ndiag, double* tile) {
    const uint32_t lane = w.sl;
    rd[K] = 0.0;
    if ((uint32_t)K >= ndiag) return;
    // pivot among the columns at positions >= K: largest norm, first position among equals; a NaN norm never wins
    const bool cand = active && pos >= (uint32_t)K && nrm == nrm;
    double bn = cand ? nrm : -1.0;
    uint32_t bp = cand ? pos : 0xffffffffu;
#pragma unroll
    for (int off = MAXD / 2; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(w.seg, bn, off, MAXD);
        const uint32_t op = __shfl_xor_sync(w.seg, bp, off, MAXD);
        if (ob > bn || (ob == bn && op < bp)) {
            bn = ob;
            bp = op;
        }
    }
    if (bp == 0xffffffffu) bp = (uint32_t)K;  // (nothing but NaNs: the oracle keeps column K)
    const uint32_t best_lane = __ffs(sub_ballot<MAXD>(w, active && pos == bp)) - 1u;
    const uint32_t lane_k = __ffs(sub_ballot<MAXD>(w, active && pos == (uint32_t)K)) - 1u;
    if (lane == best_lane) pos = (uint32_t)K;
    else if (lane == lane_k) pos = bp;
    const double norm = sqrt(bn);
    const double akk = __shfl_sync(w.seg, A[K], best_lane, MAXD);
    if (norm == 0.0) {  // zero pivot: R_kk = 0, the remaining norms restart one row lower (find_dof: `continue`)
        nrm = 0.0;
#pragma unroll
        for (int i = K + 1; i < MAXD; ++i)
            if ((uint32_t)i < m) nrm += A[i] * A[i];
        return;
    }
    const double alpha = akk > 0 ? -norm : norm;
    const double vk = akk - alpha;
    const double tau = -vk / alpha;
    rd[K] = alpha;
    // v_i = A(i, pivot column) / vk for the rows below K.  A division is some thirty instructions for the whole warp whether
    // one lane or all of them divide: the pivot column goes through the sub-warp's row of the shared-memory tile so that LANE i divides
    // row i — one division per step instead of one per row — and every lane reads the quotients back (broadcast loads).
    double v[MAXD];
    if (lane == best_lane) {
#pragma unroll
        for (int i = K + 1; i < MAXD; ++i) tile[i] = A[i];
    }
    __syncwarp(w.seg);
    const double vi = (lane > (uint32_t)K && lane < m) ? tile[lane] / vk : 0.0;
    __syncwarp(w.seg);
    tile[lane] = vi;
    __syncwarp(w.seg);
#pragma unroll
    for (int i = K + 1; i < MAXD; ++i) v[i] = tile[i];
    __syncwarp(w.seg);
    if (lane == best_lane) {
        A[K] = alpha;
    } else if (active && pos > (uint32_t)K) {
        double s = A[K];
#pragma unroll
        for (int i = K + 1; i < MAXD; ++i)
            if ((uint32_t)i < m) s += v[i] * A[i];
        s *= tau;
        A[K] -= s;
        nrm = 0.0;
#pragma unroll
        for (int i = K + 1; i < MAXD; ++i)
            if ((uint32_t)i < m) {
                A[i] -= s * v[i];
                nrm += A[i] * A[i];
            }
    }
}
template <int MAXD, int... Ks>
__device__ __forceinline__ void qr_steps(std::integer_sequence<int, Ks...>, double (&A)[MAXD], double (&rd)[MAXD], uint32_t& pos, double& nrm,
                                         const SubWarp& w, bool active, uint32_t m, uint32_t ndiag, double* tile) {
    (qr_step<Ks, MAXD>(A, rd, pos, nrm, w, active, m, ndiag, tile), ...);
}

// Back substitution row II of the null-space basis (freedom_warp_kernel), II a compile-time constant for the same reason.
template <int II, int MAXD>
__device__ __forceinline__ void back_row(const double (&A)[MAXD], double (&z)[MAXD], const uint32_t (&lop)[MAXD], uint32_t rank,
                                         const SubWarp& w) {
    if ((uint32_t)II >= rank) return;
    double rhs = A[II];
#pragma unroll
    for (int j = II + 1; j < MAXD; ++j)
        if ((uint32_t)j < rank) rhs += __shfl_sync(w.seg, A[II], lop[j], MAXD) * z[j];
    const double diagonal = __shfl_sync(w.seg, A[II], lop[II], MAXD);
    z[II] = -rhs / diagonal;
}
template <int MAXD, int... Is>
__device__ __forceinline__ void back_rows(std::integer_sequence<int, Is...>, const double (&A)[MAXD], double (&z)[MAXD],
                                          const uint32_t (&lop)[MAXD], uint32_t rank, const SubWarp& w) {
    (back_row<MAXD - 1 - Is, MAXD>(A, z, lop, rank, w), ...);  // rows descending
}

// ---- Small sketches: a WARP per problem, the matrix in REGISTERS --------------------------------------------------------
// For m, n <= 16 (the sketches of BASELINE.json's config 5) lane j holds column j of the Jacobian in registers; a
// Householder step is a handful of shuffles (the pivot column's entries broadcast as v) and each lane's own fma-free
// chains — no shared memory, no barriers, ~10x fewer instructions than the team kernel, which spent 73 % of its issue slots
// on index arithmetic and barriers for 16 x 16 matrices.  Columns are never moved: a lane tracks the POSITION of its column
// in the oracle's swapped order, so ties between equal norms break exactly as the oracle's "first maximum" does.  Every
// sum runs in the oracle's order (rows ascending, free columns ascending, modified Gram-Schmidt g ascending), so the whole
// analysis — not only the QR — is bit-identical to oracle/ezpz_oracle.cpp freedom_analysis.
// (16 x 16: six blocks per SM = 24 warps at 80 registers and 144 bytes of spills measured 353 us against 387 us for four blocks at
// 114 registers; the 8 x 8 instance needs 79 registers anyway)
template <int MAXD>
__global__ void __launch_bounds__(128, MAXD >= 16 ? 6 : 1) freedom_warp_kernel(const FreedomArgs a) {
    extern __shared__ double fsm[];  // per warp: MAXD x 32 doubles, only to densify the sparse columns
    constexpr uint32_t kPerWarp = 32u / MAXD;  // problems per warp: sub-warps of MAXD lanes
    const uint32_t wlane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t sub = wlane / MAXD;
    const SubWarp w{MAXD >= 32 ? 0xffffffffu : (((1u << MAXD) - 1u) << (sub * MAXD)), sub * MAXD, wlane % MAXD};
    const uint32_t lane = w.sl;
    const uint32_t p = (blockIdx.x * (blockDim.x >> 5) + warp) * kPerWarp + sub;
    if (p >= a.count) return;
    const uint32_t m = a.m, n = a.n, ndiag = m < n ? m : n;
    const bool active = lane < n;
    // the sub-warp's MAXD columns of the warp's [MAXD][32] block; its row 0 doubles as the exchange row of the QR steps
    double* tile = fsm + (size_t)warp * MAXD * 32 + w.shift;
    const double* jac = a.jac + (size_t)p * a.nnz;
#pragma unroll
    for (int i = 0; i < MAXD; ++i) tile[i * 32 + lane] = 0.0;
    __syncwarp(w.seg);
    if (active)
        for (uint32_t e = a.csc_col_ptr[lane]; e < a.csc_col_ptr[lane + 1]; ++e) tile[a.csc_row_idx[e] * 32 + lane] = jac[e];
    __syncwarp(w.seg);
    double A[MAXD];
#pragma unroll
    for (int i = 0; i < MAXD; ++i) A[i] = tile[i * 32 + lane];
    uint32_t pos = lane;  // position of this lane's column in the oracle's (swapped) column order
    double rd[MAXD];
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
        if ((uint32_t)i < m) nrm += A[i] * A[i];
    __syncwarp(w.seg);
    qr_steps<MAXD>(std::make_integer_sequence<int, MAXD>{}, A, rd, pos, nrm, w, active, m, ndiag, tile);
    // ---- rank (find_dof.rs:40-52)
    double largest = ezm::ez_abs(rd[0]);
#pragma unroll
    for (int i = 1; i < MAXD; ++i)
        if ((uint32_t)i < ndiag) largest = ezm::ez_fmax(largest, ezm::ez_abs(rd[i]));
    const double tolerance = 1e-8 * largest;
    uint32_t rank = 0;
    {
        bool run = true;
#pragma unroll
        for (int i = 0; i < MAXD; ++i) {
            run = run && (uint32_t)i < ndiag && ezm::ez_abs(rd[i]) > tolerance;
            if (run) rank = (uint32_t)i + 1u;
        }
    }
    uint32_t* mask = a.mask + (size_t)p * a.words;
    if (rank == n) {
        if (lane == 0) mask[0] = 0;
        return;
    }
    uint32_t lop[MAXD];  // lane that holds the column at each position
#pragma unroll
    for (int q = 0; q < MAXD; ++q) lop[q] = (uint32_t)q < n ? __ffs(sub_ballot<MAXD>(w, active && pos == (uint32_t)q)) - 1u : 0u;
    // ---- basis of null(J P^T): back substitution over R11 for this lane's column (meaningful in the free lanes)
    double z[MAXD];
#pragma unroll
    for (int i = 0; i < MAXD; ++i) z[i] = 0.0;
    back_rows<MAXD>(std::make_integer_sequence<int, MAXD>{}, A, z, lop, rank, w);
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
        if ((uint32_t)i >= rank) z[i] = ((uint32_t)i == pos) ? 1.0 : 0.0;
    // ---- modified Gram-Schmidt (twice) over the free columns in position order, then normalise
#pragma unroll 1
    for (uint32_t pf = rank; pf < n; ++pf) {
        uint32_t cur = 0;
#pragma unroll
        for (int q = 0; q < MAXD; ++q)
            if ((uint32_t)q == pf) cur = lop[q];
        double zc[MAXD];  // the column being orthonormalised, replicated in every lane
#pragma unroll
        for (int i = 0; i < MAXD; ++i) zc[i] = (uint32_t)i < n ? __shfl_sync(w.seg, z[i], cur, MAXD) : 0.0;
        for (int pass = 0; pass < 2; ++pass)
#pragma unroll 1
            for (uint32_t pg = rank; pg < pf; ++pg) {
                uint32_t gl = 0;
#pragma unroll
                for (int q = 0; q < MAXD; ++q)
                    if ((uint32_t)q == pg) gl = lop[q];
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < MAXD; ++i)
                    if ((uint32_t)i < n) s += z[i] * zc[i];  // (the value of lane gl is the one used)
                s = __shfl_sync(w.seg, s, gl, MAXD);
#pragma unroll
                for (int i = 0; i < MAXD; ++i)
                    if ((uint32_t)i < n) zc[i] -= s * __shfl_sync(w.seg, z[i], gl, MAXD);
            }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < MAXD; ++i)
            if ((uint32_t)i < n) s += zc[i] * zc[i];
        s = sqrt(s);
        if (lane == cur) {
#pragma unroll
            for (int i = 0; i < MAXD; ++i)
                if ((uint32_t)i < n) z[i] = zc[i] / s;
        }
    }
    // ---- participation of the variable at every position: sum over the free columns, ascending (find_dof.rs:82-104)
    double my_part = 0.0;
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
        if ((uint32_t)i < n) {
            const double sq = z[i] * z[i];
            double total = 0.0;
#pragma unroll 1
            for (uint32_t pf = rank; pf < n; ++pf) {
                uint32_t fl = 0;
#pragma unroll
                for (int q = 0; q < MAXD; ++q)
                    if ((uint32_t)q == pf) fl = lop[q];
                total += __shfl_sync(w.seg, sq, fl, MAXD);
            }
            if (pos == (uint32_t)i) my_part = total;
        }
    double max_p = active ? my_part : 0.0;
#pragma unroll
    for (int off = MAXD / 2; off > 0; off >>= 1) max_p = ezm::ez_fmax(max_p, __shfl_xor_sync(w.seg, max_p, off, MAXD));
    max_p = ezm::ez_fmax(0.0, max_p);
    const double var_tol = 1e-3 * max_p;
    const double squared_tol = var_tol * var_tol;
    const uint32_t bits = sub_ballot<MAXD>(w, active && my_part > squared_tol);
    if (lane == 0) mask[0] = bits;
}

std::mutex g_attr_mutex;

}  // namespace

namespace ezs {

// Every analysis ends here: the event that later users of the context's Jacobian buffer and scratch (other streams of the copy /
// compute pipeline) wait for.  (The warp-per-problem path once returned without it: on a loaded 8-GPU box the next chunk's
// solve kernel then overwrote Jacobians an analysis was still reading — 273 wrong masks in 1,048,577 problems.)
static int32_t freedom_mark_done(ezpz_context* ctx, cudaStream_t st, ezpz_error_detail_t* detail) {
    if (!ctx->fa_done) EZ_CUDA(cudaEventCreateWithFlags(&ctx->fa_done, cudaEventDisableTiming), "cudaEventCreate");
    EZ_CUDA(cudaEventRecord(ctx->fa_done, st), "cudaEventRecord");
    ctx->fa_busy = true;
    ctx->fa_last_stream = st;
    return EZPZ_OK;
}

// Device-resident form: `d_jac` [batch * nnz] and `d_mask` [batch * ceil(n/32)] are device pointers on the context's device; the
// kernels are enqueued on `st`, nothing is copied to or from the host and the call does not synchronise.
int32_t freedom_device(ezpz_context* ctx, const ezpz_structure* s, uint64_t batch, const double* d_jac, uint32_t* d_mask,
                       cudaStream_t st, ezpz_error_detail_t* detail) {
    const uint32_t m = s->m, n = s->n;
    if (std::min(m, n) == 0) return EZPZ_ERR_EMPTY_SYSTEM;  // find_dof.rs:41-44
    if (batch == 0) return EZPZ_OK;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    DeviceCopy* dc = nullptr;
    int32_t rc = get_device_copy(ctx, s, &dc, detail);
    if (rc != EZPZ_OK) return rc;
    const size_t nnz = s->csc_row_idx.size();
    {
        std::lock_guard<std::mutex> lock(const_cast<ezpz_structure*>(s)->dev_mutex);
        if (!dc->csc_col_ptr) {
            EZ_CUDA(cudaMalloc(&dc->csc_col_ptr, sizeof(uint32_t) * (n + 1)), "cudaMalloc(col_ptr)");
            EZ_CUDA(cudaMemcpy(dc->csc_col_ptr, s->csc_col_ptr.data(), sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice), "cudaMemcpy(col_ptr)");
            EZ_CUDA(cudaMalloc(&dc->csc_row_idx, sizeof(uint32_t) * std::max<size_t>(1, nnz)), "cudaMalloc(row_idx)");
            if (nnz) EZ_CUDA(cudaMemcpy(dc->csc_row_idx, s->csc_row_idx.data(), sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice), "cudaMemcpy(row_idx)");
        }
    }
    const uint64_t per_cta = freedom_doubles(m, n, false), per_grid = freedom_doubles(m, n, true);
    uint64_t per = per_cta;
    FreedomArgs a;
    a.csc_col_ptr = dc->csc_col_ptr;
    a.csc_row_idx = dc->csc_row_idx;
    a.m = m;
    a.n = n;
    a.nnz = (uint32_t)nnz;
    a.words = (n + 31) / 32;
    a.scratch = nullptr;
    // the work area of a team is reused launch after launch: launches of different streams take turns
    if (ctx->fa_busy && ctx->fa_last_stream != st) EZ_CUDA(cudaStreamWaitEvent(st, ctx->fa_done, 0), "cudaStreamWaitEvent");
    // m, n <= 16: a warp per problem, matrix in registers
    if (m <= 16 && n <= 16 && !std::getenv("EZPZ_B200_FREEDOM_TEAM")) {
        a.per_problem = 0;
        a.in_smem = a.v_smem = 0;
        uint64_t done = 0;
        while (done < batch) {
            const uint64_t count = std::min<uint64_t>(batch - done, (uint64_t)1 << 30);
            a.jac = d_jac + done * nnz;
            a.mask = d_mask + done * a.words;
            a.count = (uint32_t)count;
            // four warps per block, 32 / MAXD problems per warp (sub-warps of MAXD lanes)
            if (m <= 8 && n <= 8) freedom_warp_kernel<8><<<(uint32_t)((count + 15) / 16), 128, 4 * 8 * 32 * sizeof(double), st>>>(a);
            else freedom_warp_kernel<16><<<(uint32_t)((count + 7) / 8), 128, 4 * 16 * 32 * sizeof(double), st>>>(a);
            ctx->launches += 1;
            EZ_CUDA(cudaGetLastError(), "freedom_warp_kernel launch");
            done += count;
        }
        return freedom_mark_done(ctx, st, detail);
    }
    const size_t smem_cap = ctx->smem_optin - 1024;
    const bool small = per * 8 <= (size_t)64 << 10;                       // matrix in shared memory
    const bool cta_team = small || (n <= 1024 && batch >= (uint64_t)ctx->sm_count / 2);
    if (!cta_team) per = per_grid;
    a.per_problem = per;
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        static bool attr_set[64] = {};
        if (ctx->device < 64 && !attr_set[ctx->device]) {
            EZ_CUDA(cudaFuncSetAttribute((const void*)freedom_team_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap), "cudaFuncSetAttribute");
            EZ_CUDA(cudaFuncSetAttribute((const void*)freedom_team_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap), "cudaFuncSetAttribute");
            attr_set[ctx->device] = true;
        }
    }
    if (cta_team) {
        const uint32_t threads = std::min<uint32_t>(1024, std::max<uint32_t>(32, (std::max(n, small ? 0u : std::min(m, 256u)) + 31) / 32 * 32));
        uint64_t done = 0;
        while (done < batch) {
            uint64_t count = batch - done;
            size_t smem = 0;
            uint32_t grid;
            if (small) {
                a.in_smem = 1;
                a.v_smem = 0;
                smem = per * 8;
                count = std::min<uint64_t>(count, 1u << 30);
                grid = (uint32_t)count;
            } else {
                a.in_smem = 0;
                a.v_smem = (uint32_t)std::min<size_t>(m, ((size_t)48 << 10) / 8);
                smem = (size_t)a.v_smem * 8;
                const uint64_t slots = std::max<uint64_t>(1, std::min<uint64_t>(count, ((uint64_t)1 << 30) / (per * 8)));
                rc = ensure_fa(ctx, slots * per * 8, detail);
                if (rc != EZPZ_OK) return rc;
                a.scratch = (double*)ctx->fa_ws;
                grid = (uint32_t)slots;
            }
            a.jac = d_jac + done * nnz;
            a.mask = d_mask + done * a.words;
            a.count = (uint32_t)std::min<uint64_t>(count, 0xffffffffu);
            freedom_team_kernel<false><<<grid, threads, smem, st>>>(a);
            ctx->launches += 1;
            EZ_CUDA(cudaGetLastError(), "freedom_team_kernel launch");
            done += a.count;
        }
    } else {
        // the whole grid on one problem after the other: one warp per SM while the columns fit (every column's chain of sums is
        // the critical path, and an SM's L2 bandwidth is shared by its warps), more warps per CTA beyond
        rc = ensure_fa(ctx, per * 8, detail);
        if (rc != EZPZ_OK) return rc;
        a.scratch = (double*)ctx->fa_ws;
        a.in_smem = 0;
        uint32_t threads = (uint32_t)std::min<uint64_t>(256, ((uint64_t)(n + ctx->sm_count - 1) / ctx->sm_count + 31) / 32 * 32);
        threads = std::max(32u, threads);
        a.v_smem = (uint32_t)std::min<size_t>(m, ((size_t)64 << 10) / 8);
        const size_t smem = (size_t)a.v_smem * 8;
        int per_sm = 0;
        EZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, freedom_team_kernel<true>, (int)threads, smem), "occupancy");
        const uint32_t grid = std::max(1u, std::min<uint32_t>(std::min<uint32_t>((uint32_t)std::max(1, per_sm) * ctx->sm_count, kMaxTeamCtas),
                                                              (n + threads - 1) / threads));
        uint64_t done = 0;
        while (done < batch) {
            a.jac = d_jac + done * nnz;
            a.mask = d_mask + done * a.words;
            a.count = (uint32_t)std::min<uint64_t>(batch - done, 1u << 20);
            void* params[] = {(void*)&a};
            EZ_CUDA(cudaLaunchCooperativeKernel((void*)freedom_team_kernel<true>, dim3(grid), dim3(threads), params, smem, st),
                    "cudaLaunchCooperativeKernel(freedom_team_kernel)");
            ctx->launches += 1;
            done += a.count;
        }
    }
    return freedom_mark_done(ctx, st, detail);
}

}  // namespace ezs

extern "C" int32_t ezpz_b200_freedom_analysis_device(ezpz_context_t* ctx, const ezpz_structure_t* s, uint64_t batch,
                                                     const double* jacobian, uint32_t* under_mask, void* cuda_stream,
                                                     ezpz_error_detail_t* detail) {
    if (!ctx || !s || !under_mask || (batch && !jacobian)) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    return ezs::freedom_device(ctx, s, batch, jacobian, under_mask, cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream, detail);
}

extern "C" int32_t ezpz_b200_freedom_analysis(ezpz_context_t* ctx, const ezpz_structure_t* s, uint64_t batch,
                                              const double* jacobian, uint32_t* under_mask,
                                              ezpz_error_detail_t* detail) {
    if (!ctx || !s || !under_mask || (batch && !jacobian)) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (std::min(s->m, s->n) == 0) return EZPZ_ERR_EMPTY_SYSTEM;
    if (batch == 0) return EZPZ_OK;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t nnz = s->csc_row_idx.size(), words = (s->n + 31) / 32;
    // host buffers: staged through the context's workspace in chunks of at most 256 MiB of Jacobian values
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(batch, ((uint64_t)256 << 20) / std::max<size_t>(8, nnz * 8)));
    const size_t b_jac = ezs::align_up(chunk * nnz * 8, 256), b_mask = ezs::align_up(chunk * words * 4, 256);
    int32_t rc = ezs::ensure_ws(ctx, b_jac + b_mask, detail);
    if (rc != EZPZ_OK) return rc;
    double* d_jac = (double*)ctx->ws;
    uint32_t* d_mask = (uint32_t*)((char*)ctx->ws + b_jac);
    cudaStream_t st = ctx->stream;
    for (uint64_t first = 0; first < batch; first += chunk) {
        const uint64_t count = std::min<uint64_t>(chunk, batch - first);
        if (nnz) EZ_CUDA(cudaMemcpyAsync(d_jac, jacobian + first * nnz, count * nnz * 8, cudaMemcpyHostToDevice, st), "H2D jacobian");
        rc = ezs::freedom_device(ctx, s, count, d_jac, d_mask, st, detail);
        if (rc == EZPZ_OK) {
            cudaError_t e = cudaMemcpyAsync(under_mask + first * words, d_mask, count * words * 4, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = ezs::cuda_fail(e, detail, "D2H mask");
        }
        cudaError_t e = cudaStreamSynchronize(st);  // (also on failure: nothing of this call is still in flight on return)
        if (rc != EZPZ_OK) return rc;
        if (e != cudaSuccess) return ezs::cuda_fail(e, detail, "cudaStreamSynchronize");
    }
    return EZPZ_OK;
}
