// device.cu — CUDA side of libezpz_b200.so for sm_100a: contexts, device copies of an analysed
// structure, the batched small-system Levenberg–Marquardt kernel and the assembly kernel.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -DEZPZ_NO_FMAD=1 (see
// __graft_entry__.build()).  -fmad=false because the constraint formulas must round exactly as the
// reference's Rust does (no contraction); the linear-algebra phases ask for fused multiply-adds
// explicitly through __fma_rn, following the arithmetic-order spec in DESIGN.md §3.
//
// Kernels
//   lm_small_kernel   one THREAD per problem; the whole LM loop of ezpz/src/solver/newton.rs:29-145 plus
//                     the post-solve check of ezpz/src/lib.rs:305-327 runs on the device.  Per-problem
//                     state (x, r, r_next, J values, A/L values, step) lives in shared memory laid out
//                     [slot][thread] so that a warp's access to one slot is one conflict-free 256-byte
//                     row.  All threads of a launch share one structure, so the constraint loop and the
//                     linear-algebra tape (structure.cpp) are warp-uniform: no divergence except in the
//                     degenerate-geometry branches and in the iteration count.
//   assemble_kernel   one thread per constraint over a single system in global memory: residuals and
//                     Jacobian values scattered through precomputed slots (Model::residual +
//                     Model::refresh_jacobian, ezpz/src/solver.rs:318-440).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "eval.cuh"
#include "structure.h"

using namespace ezs;

#include "device.h"
#include "host_parallel.h"

namespace {

__constant__ uint8_t c_rows[EZPZ_K_COUNT];
__constant__ uint8_t c_emit_len[EZPZ_K_COUNT][2];

struct SmallArgs {
    const uint32_t* tables;  // RoleBlob words (structure.h) on the device
    const double* guesses;
    const double* params;
    double* finals;
    uint32_t* iterations;
    uint8_t* status;
    uint32_t* unsat;
    uint32_t* degen;
    double* jac;
    uint64_t batch;
    double residual_tolerance, step_tolerance, initial_lambda;
    uint32_t max_iterations;
    uint32_t n_cons, n, m, W, X0, R0, RN0, J0, L0, D0, S0, F0, unsat_words, nnz;
    uint32_t table_words;  // length of the blob in 32-bit words
    uint32_t cons_word;    // word offset of the DevCons array inside the blob
    uint32_t T, R;         // problems per CTA (a multiple of 32) and roles (warps per 32 problems)
    uint32_t stride;       // doubles between consecutive slots of V: T + 1 (see small_stride)
    uint32_t weights_one;  // 1: every weight is exactly 1.0 (r == unweighted residuals)
};

struct GlobalX {
    const double* p;
    __device__ __forceinline__ double operator()(uint32_t id) const { return __ldg(p + id); }
};

// Per-thread view of the shared-memory state: column `problem` of V[slot][problem].  Slots are addressed
// by BYTE offsets that are uniform across the warp (slot * stride * 8), so an access is base + uniform.
struct VView {
    char* base;   // &V[0][problem]
    uint32_t sb;  // stride in bytes between consecutive slots (= problems per CTA * 8)
    __device__ __forceinline__ double ld(uint32_t off) const { return *reinterpret_cast<const double*>(base + off); }
    __device__ __forceinline__ void stp(bool pred, uint32_t off, double v) const {
        if (pred) *reinterpret_cast<double*>(base + off) = v;
    }
    __device__ __forceinline__ double lds(uint32_t slot) const { return ld(slot * sb); }
};
struct SmemX {
    VView v;
    uint32_t x0;  // byte offset of x[0]
    __device__ __forceinline__ double operator()(uint32_t id) const { return v.ld(x0 + id * v.sb); }
};

// The R warps that share 32 problems meet at a named barrier (ids 1..15, one per problem group of the CTA).
__device__ __forceinline__ void group_sync(uint32_t bar_id, uint32_t bar_threads) {
    __syncwarp();
    asm volatile("barrier.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
}

// One pass over the constraints of this role's list for this thread's problem.  Lanes with `act` false run along
// (the warp is uniform) but store nothing.
//   RES: write weight*residual to V[rdst + row] and count Warning::Degenerate of Model::residual
//   JAC: write the Jacobian values to V[jdst + slot]; degenerate rows are reported through jac_degen_any and
//        counted by the caller only if the point is accepted (Model::refresh_jacobian runs only then)
template <bool RES, bool JAC>
__device__ __forceinline__ void eval_list(const SmallArgs& a, const DevCons* __restrict__ cons, const uint32_t* __restrict__ list,
                                          uint32_t n_list, const VView& V, uint32_t rdst, uint32_t jdst,
                                          const double* __restrict__ prow, uint32_t* __restrict__ degen_row, bool act,
                                          bool& res_degen_any, bool& jac_degen_any) {
    const SmemX X{V, a.X0 * V.sb};
#pragma unroll 1
    for (uint32_t q = 0; q < n_list; ++q) {
        // list entry: constraint index | rows << 16 | partials of row 0 << 20 | partials of row 1 << 24 (structure.cpp)
        const uint32_t entry = list[q];
        const uint32_t c = entry & 0xffffu, rows = (entry >> 16) & 3u;
        const DevCons& dc = cons[c];
        const uint32_t kind = dc.kind;
        uint32_t side = dc.flags;
        if (dc.side_slot != 0xffffffffu) side = (uint32_t)V.lds(a.S0 + dc.side_slot);
        const double p0 = prow ? prow[c] : dc.p0;
        ezd::EvalOut o;
        ezd::eval_constraint<JAC>(kind, side, dc.ids, p0, dc.p1, X, o);
        const double w = dc.weight;
        if (RES) {
            V.stp(act, (rdst + dc.row0) * V.sb, w * o.res[0]);
            if (rows == 2) V.stp(act, (rdst + dc.row0 + 1) * V.sb, w * o.res[1]);
            if (o.res_degen && act) {
                res_degen_any = true;
                if (degen_row) degen_row[c] += 1;
            }
        }
        if (JAC) {
            if (o.jac_degen && act) jac_degen_any = true;
#pragma unroll
            for (int row = 0; row < 2; ++row) {
                if (row < (int)rows) {
                    const uint32_t len = (entry >> (20 + 4 * row)) & 15u;
                    // (leaves at the row's last partial — a uniform branch — instead of predicating eight stores off)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (k >= (int)len) break;
                        const uint32_t s = dc.slot[row][k];
                        const uint32_t off = (jdst + (s & ~kAccumulate)) * V.sb;
                        if (s & kAccumulate) {
                            if (o.emit[row]) V.stp(act, off, V.ld(off) + w * o.pd[row][k]);
                        } else {
                            V.stp(act, off, o.emit[row] ? 0.0 + w * o.pd[row][k] : 0.0);
                        }
                    }
                }
            }
        }
    }
}

// Warning::Degenerate bookkeeping of Model::refresh_jacobian (solver.rs:385-391) for an accepted point: only
// run when some constraint of this role's list was degenerate there and the caller asked for per-constraint counts.
__device__ __noinline__ void count_jacobian_degenerates(const SmallArgs& a, const DevCons* __restrict__ cons,
                                                        const uint32_t* __restrict__ list, uint32_t n_list, VView V,
                                                        const double* __restrict__ prow, uint32_t* __restrict__ degen_row) {
    const SmemX X{V, a.X0 * V.sb};
    for (uint32_t q = 0; q < n_list; ++q) {
        const uint32_t c = list[q] & 0xffffu;
        const DevCons& dc = cons[c];
        uint32_t side = dc.flags;
        if (dc.side_slot != 0xffffffffu) side = (uint32_t)V.lds(a.S0 + dc.side_slot);
        ezd::EvalOut o;
        ezd::eval_constraint<true>(dc.kind, side, dc.ids, prow ? prow[c] : dc.p0, dc.p1, X, o);
        if (o.jac_degen) degen_row[c] += 1;
    }
}

// A run of multiply-adds of one tape op: acc = fma(+-V[a], V[b], acc) over the {a, b} byte-offset pairs in tape words
// [w, we).  One 64-bit load fetches a pair; the NEXT pair's offsets are requested before this pair's operands, so the
// loop's dependent chain is operand load -> multiply-add, not offsets -> operands -> multiply-add (whatever follows the
// last pair — pad words or the next header — is loaded and ignored; the tape is padded at the end).
template <bool NEG>
__device__ __forceinline__ double tape_run(const uint32_t* __restrict__ tape, uint32_t& w, uint32_t we, const VView& V, double acc) {
    if (w < we) {
        uint2 p = *reinterpret_cast<const uint2*>(tape + w);
#pragma unroll 1
        do {
            w += 2;
            const uint2 q = *reinterpret_cast<const uint2*>(tape + w);
            const double x = V.ld(p.x), y = V.ld(p.y);
            acc = __fma_rn(NEG ? -x : x, y, acc);
            p = q;
        } while (w < we);
    }
    return acc;
}

// This role's share of the linear-algebra tape of one LM iteration: A = JtJ + lambda*I, b = -Jt r, A = L Lt, L y = b,
// Lt d = y.  Per op a 4-word header {dst byte offset, fin byte offset, added pairs | subtracted pairs << 16, shape |
// words to the next header << 8} followed by one {a, b} byte-offset pair per multiply-add, the added ones first
// (padded to 16 bytes).  The shape names one of the few forms an op takes (structure.cpp), so that each is straight-line
// code around its loops instead of a chain of flag tests; TAPE_BARRIER is the meeting point of the roles in front of an
// op that depends on another role's work.  Offsets are uniform across the warp: an operand access is LDS [thread base +
// uniform].  Returns true when a pivot of this role was not positive and finite ("LltError::Numeric", newton.rs:96-99).
__device__ __forceinline__ bool run_role_tape(const uint32_t* __restrict__ tape, uint32_t n_ops, const VView& V, double lambda,
                                              bool act, uint32_t bar_id, uint32_t bar_threads) {
    bool fail = false;
    uint32_t w = 0;  // 32-bit word index: keeps the loop control out of 64-bit pointer arithmetic
    // The warp issues in order and an op is a chain of dependent latencies, so whatever can be requested early is: the
    // NEXT header while this op computes (the tape ends with a pad header), the finalisation operand with the first pairs.
    uint4 h = *reinterpret_cast<const uint4*>(tape);
#pragma unroll 1
    for (uint32_t op = 0; op < n_ops; ++op) {
        const uint32_t dst = h.x, fin = h.y, counts = h.z, shape = h.w & 0xffu;
        const uint32_t w_next = w + (h.w >> 8);
        h = *reinterpret_cast<const uint4*>(tape + w_next);
        w += 4;
        if (shape == TAPE_BARRIER) {
            group_sync(bar_id, bar_threads);
            w = w_next;
            continue;
        }
        const uint32_t wp = w + 2 * (counts & 0xffffu), we = wp + 2 * (counts >> 16);
        double acc;
        if (shape == TAPE_ENTRY) {  // L[i][j] = (A[i][j] - sum L[i][k] L[j][k]) / L[j][j]; y[i] = (b[i] - sum L[i][k] y[k]) / L[i][i]
            const double fin_v = V.ld(fin);
            acc = tape_run<false>(tape, w, wp, V, 0.0);
            acc = tape_run<true>(tape, w, we, V, acc);
            acc = __dmul_rn(acc, fin_v);
        } else if (shape == TAPE_PIVOT) {  // 1 / sqrt(A[j][j] + lambda - sum L[j][k]^2)
            acc = tape_run<false>(tape, w, wp, V, 0.0);
            acc = __dadd_rn(acc, lambda);
            acc = tape_run<true>(tape, w, we, V, acc);
            if (!(acc > 0.0) || !ezm::ez_isfinite(acc)) fail = true;
            acc = __ddiv_rn(1.0, __dsqrt_rn(acc));
        } else {  // TAPE_BACKWARD: d[j] = (y[j] - sum L[i][j] d[i]) / L[j][j], in place
            const double fin_v = V.ld(fin);
            acc = tape_run<true>(tape, w, we, V, V.ld(dst));
            acc = __dmul_rn(acc, fin_v);
        }
        w = w_next;
        V.stp(act, dst, acc);
    }
    return fail;
}

// Moves the rows of a problem group between global memory (row-major, `width` doubles per problem, the group's problems
// consecutive) and shared memory ([slot][problem]) with the group's R warps side by side: flat element i of the
// 32 x width block belongs to problem i / width, slot i % width, and consecutive lanes take consecutive elements, so a warp
// instruction moves 256 contiguous bytes of global memory — full lines whether that memory is HBM or, for the
// host-buffer entry point, pinned host memory read and written across PCIe by the kernel itself.
template <bool LOAD>
__device__ __forceinline__ void group_rows(char* gbase, uint32_t off0, uint32_t sb, double* __restrict__ rows, uint32_t width,
                                           uint32_t n_valid, uint32_t role, uint32_t R, uint32_t lane) {
    const uint32_t total = n_valid * width, stride = 32u * R;
    const uint32_t dp = stride / width, dj = stride - dp * width;
    uint32_t i = role * 32u + lane;
    uint32_t p = i / width, j = i - p * width;
    for (; i < total; i += stride) {
        double* s = reinterpret_cast<double*>(gbase + off0 + j * sb + p * 8u);
        if (LOAD) *s = __ldcg(rows + i);  // (read once; and, streamed, written by the copy engine while the kernel runs)
        else rows[i] = *s;
        p += dp;
        j += dj;
        if (j >= width) {
            j -= width;
            ++p;
        }
    }
}

// The whole solve of one problem — newton.rs:29-145 + lib.rs:305-327 — by the R threads (one per role warp) that share
// its shared-memory column.  Every role runs the same control flow on the same values (the sums of squares, maxima and
// verdicts are recomputed by each role from shared memory, which is cheaper than broadcasting them), so the roles take
// every branch together; the work that scales — constraint evaluation, the tape, the copies — is split.  A lane whose
// problem has finished (or failed a factorisation: newton.rs:96-99 `continue`) keeps running with its stores switched
// off until every lane of the warp is done, which is what SIMT divergence would cost anyway and keeps the named
// barriers aligned.
__device__ __forceinline__ void lm_roles_body(const SmallArgs& a, const uint32_t* __restrict__ tb, const VView V, uint64_t b,
                                              bool valid, uint32_t role, uint32_t bar_id) {
    const uint32_t R = a.R, bar_threads = 32u * R;
    const uint32_t* __restrict__ hdr = tb + role * kRoleHdrWords;
    const DevCons* __restrict__ cons = reinterpret_cast<const DevCons*>(tb + a.cons_word);
    const uint32_t* __restrict__ list = tb + hdr[0];
    const uint32_t n_list = hdr[1];
    const uint32_t* __restrict__ tape = tb + hdr[2];
    const uint32_t n_ops = hdr[3];
    const uint32_t x_lo = hdr[4], x_hi = hdr[5], r_lo = hdr[6], r_hi = hdr[7], j_lo = hdr[8], j_hi = hdr[9];
    const double* __restrict__ prow = (valid && a.params) ? a.params + b * a.n_cons : nullptr;
    uint32_t* __restrict__ degen_row = (valid && a.degen) ? a.degen + b * a.n_cons : nullptr;
    const uint32_t sb = V.sb;
    const uint32_t oX = a.X0 * sb, oR = a.R0 * sb, oRN = a.RN0 * sb, oJ = a.J0 * sb, oL = a.L0 * sb, oD = a.D0 * sb;
    uint32_t* const flag = reinterpret_cast<uint32_t*>(V.base + a.F0 * sb);  // bit 0 degenerate, bits 1-2 failed pivot (by parity)
    bool any_degen = false;

    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t b0 = b - lane;  // first problem of the group
    const uint32_t n_valid = b0 < a.batch ? (uint32_t)(a.batch - b0 < 32u ? a.batch - b0 : 32u) : 0u;
    char* const gbase = V.base - lane * 8u;  // column 0 of the group
    // initial guesses (lib.rs:275: values are positional == by id)
    group_rows<true>(gbase, oX, sb, const_cast<double*>(a.guesses) + b0 * a.n, a.n, n_valid, role, R, lane);
    if (degen_row)
        for (uint32_t q = 0; q < n_list; ++q) degen_row[list[q] & 0xffffu] = 0;
    if (role == 0) *flag = 0u;
    if (R > 1) group_sync(bar_id, bar_threads);
    else __syncwarp();
    {  // Constraint::set_from_initial_values (lib.rs:183-186)
        const SmemX X{V, oX};
        for (uint32_t q = 0; q < n_list; ++q) {
            const DevCons& dc = cons[list[q] & 0xffffu];
            if (dc.side_slot != 0xffffffffu)
                V.stp(valid, (a.S0 + dc.side_slot) * sb, (double)ezd::resolve_side(dc.kind, dc.flags, dc.ids, X));
        }
    }

    double lambda = a.initial_lambda;
    {
        bool rd = false, jd = false;
        eval_list<true, true>(a, cons, list, n_list, V, a.R0, a.J0, prow, degen_row, valid, rd, jd);
        if (jd && degen_row) count_jacobian_degenerates(a, cons, list, n_list, V, prow, degen_row);
        any_degen = rd || jd;
    }
    if (R > 1) group_sync(bar_id, bar_threads);
    // S = sum r^2 is a strictly sequential fold (Rust's iterator sum); max |r_i| with libm::fmax semantics
    // (NaN-ignoring: NaN < x is false, so a NaN candidate never wins and a NaN incumbent is replaced by any number) is
    // order-independent and rides along on the same loads.  Unrolled by four so that the loads and squares of four
    // elements are in flight while the adds chain.
    double S = 0.0, largest = ezm::ez_abs(V.ld(oR));
#pragma unroll 4
    for (uint32_t i = 0; i < a.m; ++i) {
        const double r = V.ld(oR + i * sb);
        S = S + r * r;
        const double v = ezm::ez_abs(r);
        largest = (largest < v || largest != largest) ? v : largest;
    }
    uint32_t iterations = a.max_iterations;
    bool converged = false;
    bool x_dirty = false;  // x was last modified by a rejected step: r no longer belongs to the bits of x
    bool alive = valid;
#pragma unroll 1
    for (uint32_t it = 0; it < a.max_iterations; ++it) {
        if (alive && largest <= a.residual_tolerance) {  // `largest` = max |r_i| of the current r, kept up to date below

            iterations = it;
            converged = true;
            alive = false;
        }
        if (!__any_sync(0xffffffffu, alive)) break;  // (the same lanes in every role's warp: the roles leave together)
        bool failed = run_role_tape(tape, n_ops, V, lambda, alive, bar_id, bar_threads);
        if (R > 1) {  // a failed pivot of any role fails the factorisation for all of them
            const uint32_t fbit = 2u << (it & 1u);
            if (failed && alive) atomicOr(flag, fbit);
            group_sync(bar_id, bar_threads);
            const uint32_t f = *reinterpret_cast<volatile uint32_t*>(flag);
            failed = (f & fbit) != 0u;
            if (role == 0 && (f & (fbit ^ 6u))) atomicAnd(flag, ~(fbit ^ 6u));  // the other parity's bit: read an iteration ago
        }
        const bool stepping = alive && !failed;
        if (alive && failed) lambda *= 10.0;  // newton.rs:96-99: the iteration is spent, no step test
        double step = ezm::ez_abs(V.ld(oD));
#pragma unroll 4
        for (uint32_t j = 1; j < a.n; ++j) {
            const double v = ezm::ez_abs(V.ld(oD + j * sb));
            step = (step < v || step != step) ? v : step;
        }
        for (uint32_t j = x_lo; j < x_hi; ++j) V.stp(stepping, oX + j * sb, V.ld(oX + j * sb) + V.ld(oD + j * sb));
        if (R > 1) group_sync(bar_id, bar_threads);
        // One fused pass at the trial point: residuals into r_next and, speculatively, the Jacobian into the
        // (now dead) A/L region; an accepted step copies both over, a rejected one leaves r and J untouched,
        // exactly as Model::residual + refresh_jacobian do (newton.rs:115-131).
        bool rd = false, jd = false;
        eval_list<true, true>(a, cons, list, n_list, V, a.RN0, a.L0, prow, degen_row, stepping, rd, jd);
        any_degen = any_degen || rd;
        if (R > 1) group_sync(bar_id, bar_threads);
        double S2 = 0.0, largest2 = ezm::ez_abs(V.ld(oRN));
#pragma unroll 4
        for (uint32_t i = 0; i < a.m; ++i) {
            const double r = V.ld(oRN + i * sb);
            S2 = S2 + r * r;
            const double v = ezm::ez_abs(r);
            largest2 = (largest2 < v || largest2 != largest2) ? v : largest2;
        }
        const bool accept = stepping && (S2 < S);
        const bool reject = stepping && !accept;
        for (uint32_t i = r_lo; i < r_hi; ++i) V.stp(accept, oR + i * sb, V.ld(oRN + i * sb));
        for (uint32_t k = j_lo; k < j_hi; ++k) V.stp(accept, oJ + k * sb, V.ld(oL + k * sb));
        for (uint32_t j = x_lo; j < x_hi; ++j) V.stp(reject, oX + j * sb, V.ld(oX + j * sb) - V.ld(oD + j * sb));
        if (accept) {
            if (jd) {
                any_degen = true;
                if (degen_row) count_jacobian_degenerates(a, cons, list, n_list, V, prow, degen_row);
            }
            S = S2;
            largest = largest2;  // r_next becomes r
            lambda *= 0.1;
            x_dirty = false;
        } else if (reject) {
            lambda *= 10.0;
            x_dirty = true;
        }
        if (R > 1) group_sync(bar_id, bar_threads);
        if (stepping && step <= a.step_tolerance) {
            iterations = it;
            converged = true;
            alive = false;
        }
    }

    if (R == 1) __syncwarp();  // (R > 1: every iteration ends at the group's barrier)
    group_rows<false>(gbase, oX, sb, a.finals + b0 * a.n, a.n, n_valid, role, R, lane);
    if (a.jac)  // Jacobian cached at the last accepted point (what freedom_analysis reads)
        group_rows<false>(gbase, oJ, sb, a.jac + b0 * a.nnz, a.nnz, n_valid, role, R, lane);
    if (R > 1) {
        if (any_degen && valid) atomicOr(flag, 1u);
        group_sync(bar_id, bar_threads);
        if (role != 0) return;
        any_degen = (*reinterpret_cast<volatile uint32_t*>(flag) & 1u) != 0u;
    }
    if (!valid) return;
    // lib.rs:305-327: unweighted residuals at the final point, |r| < 1e-4 per component.  When every weight
    // is 1.0 and x still holds the bits r was evaluated at, r IS that residual vector and is reused.
    bool any_unsat = false;
    {
        const bool reuse = a.weights_one && !x_dirty;
        const SmemX X{V, oX};
        uint32_t* __restrict__ urow = a.unsat ? a.unsat + b * a.unsat_words : nullptr;
        uint32_t word = 0;
        for (uint32_t c = 0; c < a.n_cons; ++c) {
            const DevCons& dc = cons[c];
            double r0, r1;
            if (reuse) {
                r0 = V.ld(oR + dc.row0 * sb);
                r1 = (c_rows[dc.kind] == 2) ? V.ld(oR + (dc.row0 + 1) * sb) : 0.0;
            } else {
                uint32_t side = dc.flags;
                if (dc.side_slot != 0xffffffffu) side = (uint32_t)V.lds(a.S0 + dc.side_slot);
                const double p0 = prow ? prow[c] : dc.p0;
                ezd::EvalOut o;
                ezd::eval_constraint<false>(dc.kind, side, dc.ids, p0, dc.p1, X, o);
                r0 = o.res[0];
                r1 = (c_rows[dc.kind] == 2) ? o.res[1] : 0.0;
            }
            const bool sat = (ezm::ez_abs(r0) < ezd::kEps) && (ezm::ez_abs(r1) < ezd::kEps);
            if (!sat) {
                any_unsat = true;
                word |= 1u << (c & 31u);
            }
            if ((c & 31u) == 31u || c + 1 == a.n_cons) {
                if (urow) urow[c >> 5] = word;
                word = 0;
            }
        }
    }
    a.iterations[b] = iterations;
    a.status[b] = (uint8_t)((converged ? EZPZ_ST_CONVERGED : 0u) | (any_unsat ? EZPZ_ST_UNSATISFIED : 0u) |
                            (any_degen ? EZPZ_ST_DEGENERATE : 0u));
}

// Thread layout: warp w of the CTA = role (w % R) of problem group (w / R); lane l of that warp works on the problem in
// shared-memory column group * 32 + l.  The tables (RoleBlob) sit in shared memory behind the per-problem state when they
// fit (STAGED), else they are read from global memory through L1.  MAXT bounds the registers ptxas may use.
template <bool STAGED, int MAXT>
__global__ void __launch_bounds__(MAXT) lm_small_kernel(const SmallArgs a) {
    extern __shared__ double smem[];
    const uint32_t* tb = a.tables;
    if (STAGED) {
        uint32_t* s_tb = reinterpret_cast<uint32_t*>(smem + (((size_t)a.W * a.stride + 1u) & ~(size_t)1));  // 16-byte aligned: the tape is read in 128-bit words
        for (uint32_t i = threadIdx.x; i < a.table_words; i += blockDim.x) s_tb[i] = a.tables[i];
        __syncthreads();
        tb = s_tb;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t group = warp / a.R, role = warp - group * a.R;
    const uint32_t col = group * 32u + lane;
    const uint64_t b = (uint64_t)blockIdx.x * a.T + col;
    const VView V{reinterpret_cast<char*>(smem + col), a.stride * 8u};
    lm_roles_body(a, tb, V, b, b < a.batch, role, 1u + group);
}

// ---- assembly over one system in global memory ----------------------------------------------------
struct AsmArgs {
    const DevCons* cons;
    const double* x;
    double* r;        // [m] weighted residuals, or nullptr
    double* jvals;    // [nnz] CSC order, or nullptr
    uint8_t* degen;   // [n_cons] bit0 residual, bit1 Jacobian, or nullptr
    uint32_t n_cons;
};

template <bool JAC>
__global__ void __launch_bounds__(256) assemble_kernel(const AsmArgs a) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cons) return;
    const DevCons dc = a.cons[c];
    const GlobalX X{a.x};
    const uint32_t side = ezd::resolve_side(dc.kind, dc.flags, dc.ids, X);
    ezd::EvalOut o;
    ezd::eval_constraint<JAC>(dc.kind, side, dc.ids, dc.p0, dc.p1, X, o);
    const uint32_t rows = c_rows[dc.kind];
    const double w = dc.weight;
    if (a.r) {
        a.r[dc.row0] = w * o.res[0];
        if (rows == 2) a.r[dc.row0 + 1] = w * o.res[1];
    }
    if (JAC && a.jvals) {
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            if (row < (int)rows) {
                const uint32_t len = c_emit_len[dc.kind][row];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < (int)len) {
                        const uint32_t s = dc.slot[row][k];
                        double* dst = a.jvals + (s & ~kAccumulate);
                        if (s & kAccumulate) {
                            if (o.emit[row]) *dst = *dst + w * o.pd[row][k];
                        } else {
                            *dst = o.emit[row] ? 0.0 + w * o.pd[row][k] : 0.0;
                        }
                    }
                }
            }
        }
    }
    if (a.degen) a.degen[c] = (uint8_t)((o.res_degen ? 1 : 0) | ((JAC && o.jac_degen) ? 2 : 0));
}

__global__ void permute_kernel(const double* __restrict__ src, const uint32_t* __restrict__ perm, double* __restrict__ dst,
                               uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[perm[i]] = src[i];
}

int32_t upload_tables(ezpz_error_detail_t* detail) {
    uint8_t rows[EZPZ_K_COUNT], emit_len[EZPZ_K_COUNT][2];
    for (int k = 0; k < EZPZ_K_COUNT; ++k) {
        rows[k] = ezk::kKinds[k].rows;
        emit_len[k][0] = ezk::kKinds[k].emit_len[0];
        emit_len[k][1] = ezk::kKinds[k].emit_len[1];
    }
    EZ_CUDA(cudaMemcpyToSymbol(c_rows, rows, sizeof rows), "cudaMemcpyToSymbol(c_rows)");
    EZ_CUDA(cudaMemcpyToSymbol(c_emit_len, emit_len, sizeof emit_len), "cudaMemcpyToSymbol(c_emit_len)");
    return EZPZ_OK;
}

}  // namespace

namespace ezs {

int32_t cuda_fail(cudaError_t e, ezpz_error_detail_t* detail, const char* what) {
    if (detail) {
        detail->a = (uint64_t)e;
        std::snprintf(detail->message, sizeof detail->message, "%s: %s", what, cudaGetErrorString(e));
    }
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return EZPZ_ERR_NO_DEVICE;
    return EZPZ_ERR_CUDA;
}

int32_t get_device_copy(ezpz_context* ctx, const ezpz_structure* cs, DeviceCopy** out, ezpz_error_detail_t* detail) {
    ezpz_structure* s = const_cast<ezpz_structure*>(cs);
    std::lock_guard<std::mutex> lock(s->dev_mutex);
    for (DeviceCopy* d : s->dev)
        if (d->device == ctx->device) {
            *out = d;
            return EZPZ_OK;
        }
    DeviceCopy* d = new (std::nothrow) DeviceCopy();
    if (!d) return EZPZ_ERR_INVALID_ARGUMENT;
    d->device = ctx->device;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    // (the analysed constraints and the CSC -> CSR permutation are only read by ezpz_b200_eval: uploaded on its first call, so
    // that the first SOLVE of a new topology does not pay for them)
    s->dev.push_back(d);
    *out = d;
    return EZPZ_OK;
}

int32_t ensure_ws(ezpz_context* ctx, size_t bytes, ezpz_error_detail_t* detail) {
    if (bytes <= ctx->ws_bytes) return EZPZ_OK;
    if (ctx->ws) cudaFree(ctx->ws);
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    EZ_CUDA(cudaMalloc(&ctx->ws, want), "cudaMalloc(workspace)");
    ctx->ws_bytes = want;
    return EZPZ_OK;
}

int32_t ensure_fa(ezpz_context* ctx, size_t bytes, ezpz_error_detail_t* detail) {
    if (bytes <= ctx->fa_bytes) return EZPZ_OK;
    if (ctx->fa_ws) cudaFree(ctx->fa_ws);  // (cudaFree waits for whatever still uses the old block)
    ctx->fa_ws = nullptr;
    ctx->fa_bytes = 0;
    EZ_CUDA(cudaMalloc(&ctx->fa_ws, bytes), "cudaMalloc(freedom analysis scratch)");
    ctx->fa_bytes = bytes;
    return EZPZ_OK;
}

int32_t ensure_pin(ezpz_context* ctx, size_t bytes, ezpz_error_detail_t* detail) {
    if (bytes <= ctx->pin_bytes) return EZPZ_OK;
    if (ctx->pin) cudaFreeHost(ctx->pin);
    ctx->pin = nullptr;
    ctx->pin_bytes = 0;
    size_t want = std::max(bytes, (size_t)256 << 10);
    EZ_CUDA(cudaHostAlloc(&ctx->pin, want, cudaHostAllocDefault), "cudaHostAlloc(staging)");
    ctx->pin_bytes = want;
    return EZPZ_OK;
}

// The small program compiled for (roles, stride) — see build_role_blob in structure.cpp — with its device copy.  Built
// once per pair and kept with the structure's device copy.
int32_t get_role_tables(ezpz_context* ctx, const ezpz_structure* cs, DeviceCopy* d, uint32_t roles, uint32_t stride,
                        RoleTables** out, ezpz_error_detail_t* detail) {
    ezpz_structure* s = const_cast<ezpz_structure*>(cs);
    std::lock_guard<std::mutex> lock(s->dev_mutex);
    for (RoleTables* t : d->roles)
        if (t->blob.stride == stride && t->blob.R == roles) {
            *out = t;
            return EZPZ_OK;
        }
    RoleTables* t = new (std::nothrow) RoleTables();
    if (!t) return EZPZ_ERR_INVALID_ARGUMENT;
    build_role_blob(*s, roles, stride, t->blob);
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    EZ_CUDA(cudaMalloc(&t->dev, sizeof(uint32_t) * std::max<size_t>(1, t->blob.words.size())), "cudaMalloc(role tables)");
    EZ_CUDA(cudaMemcpy(t->dev, t->blob.words.data(), sizeof(uint32_t) * t->blob.words.size(), cudaMemcpyHostToDevice),
            "cudaMemcpy(role tables)");
    if (const char* dbg = std::getenv("EZPZ_B200_DEBUG"); dbg && dbg[0] == '1')
        std::fprintf(stderr, "[small] roles %u stride %u: %zu table bytes, %u barriers per tape, busiest role %.0f %% of the tape\n", roles,
                     stride, t->blob.words.size() * 4, t->blob.tape_barriers, 100.0 * t->blob.busiest_share);
    d->roles.push_back(t);
    *out = t;
    return EZPZ_OK;
}

constexpr uint64_t kPipelineMinBatch = 32768;  // below: the kernel reads and writes page-locked caller buffers across PCIe itself

// Slot stride of the shared-memory state V[slot][problem], in doubles: T + 1.  With a stride of T (a multiple of 32) the
// 16 lanes that move the 16 variables of ONE problem between global memory and V (group_rows: consecutive lanes take
// consecutive doubles of a row) all hit the same bank — a 16-way conflict on 13 % of the kernel's wavefronts (ncu, r02f);
// an odd stride spreads them over 16 different 8-byte banks, and the arithmetic phases (a warp on one slot of 32 consecutive
// problems) are conflict-free either way.
inline uint32_t small_stride(uint32_t T) { return T + 1u; }

struct SmallShape {
    uint32_t T = 0, R = 0;
    size_t smem = 0;
    bool stage = false;
};

// Launch shape of the batched kernel for `batch` problems of a structure (pure host arithmetic; the device's SM count
// and shared memory per block are parameters so that tests can ask without a device).
//   R  roles = warps that cooperate on 32 problems (EZPZ_B200_ROLES overrides): the per-problem state fills the SM's
//      shared memory at ~5 warps' worth of problems, far too few independent instruction streams to cover the latency of
//      the dependent LDS -> DFMA chains; R warps per problem group multiply the streams without multiplying the state.
//      Candidates R = 1..4 are scored by problem groups resident per SM over the modelled time of the busiest role
//      (RoleBlob::critical_cost): more roles shorten the critical path when the work splits, but every group then
//      occupies R of the SM's 16 warps (128 registers per thread), so a structure with little state per problem — or a
//      batch that fills the SMs anyway — is better served by more groups, and a batch too small to fill them by more roles.
//   T  problems per CTA: as many groups of 32 as fit in shared memory next to the tables (at most 15: one named barrier per
//      group), but no more than an even split of the batch over the SMs needs — a small batch spreads over all SMs
//      instead of filling a few.
int32_t small_shape_host(const ezpz_structure* cs, uint64_t batch, uint32_t sm_count, size_t smem_optin, SmallShape* out) {
    static const uint32_t env_roles = [] {
        const char* e = std::getenv("EZPZ_B200_ROLES");
        return e ? (uint32_t)std::min(8l, std::max(0l, std::strtol(e, nullptr, 10))) : 0u;
    }();
    static const uint32_t max_threads = [] {
        const char* e = std::getenv("EZPZ_B200_SMALL_THREADS");
        return e ? (uint32_t)std::min(768l, std::max(32l, std::strtol(e, nullptr, 10))) : 512u;
    }();
    if (!cs->small.valid || sm_count == 0) return EZPZ_ERR_UNSUPPORTED;
    ezpz_structure* s = const_cast<ezpz_structure*>(cs);
    const size_t per_group = (size_t)cs->small.W * sizeof(double) * 32u;
    const uint64_t n_groups = batch == ~0ull ? ~0ull : (batch + 31) / 32;
    const uint64_t need = n_groups == ~0ull ? ~0ull : std::max<uint64_t>(1, (n_groups + sm_count - 1) / sm_count);  // groups per SM, one wave
    uint32_t R = 0;
    uint64_t g_max = 0;
    size_t tables = 0;
    bool stage = false;
    double best = -1.0;
    for (uint32_t cand = (env_roles ? env_roles : 1u); cand <= (env_roles ? env_roles : 4u); ++cand) {
        // The size of the tables does not depend on the stride: a stride-1 blob, built once per R, tells it (and the model's cost).
        const RoleBlob* probe = nullptr;
        {
            std::lock_guard<std::mutex> lock(s->dev_mutex);
            for (const RoleBlob* p : s->role_probes)
                if (p->R == cand) probe = p;
            if (!probe) {
                RoleBlob* p = new RoleBlob();
                build_role_blob(*s, cand, 1u, *p);
                s->role_probes.push_back(p);
                probe = p;
            }
        }
        const size_t tb = probe->words.size() * sizeof(uint32_t);
        bool stg = tb <= 64 * 1024 && per_group + (size_t)cs->small.W * sizeof(double) + 8 + tb <= smem_optin;
        if (const char* e = std::getenv("EZPZ_B200_STAGE"); e && e[0] == '0') stg = false;
        if (per_group + (size_t)cs->small.W * sizeof(double) + 8 > smem_optin) continue;
        const size_t pad = (size_t)cs->small.W * sizeof(double) + 8;  // the extra column of small_stride (+ alignment of the tables)
        const size_t avail = smem_optin - (stg ? tb : 0) - std::min<size_t>(pad, smem_optin - (stg ? tb : 0));
        uint64_t g = std::min<uint64_t>({avail / per_group, (uint64_t)15, (uint64_t)(max_threads / (32u * cand))});
        if (const char* e = std::getenv("EZPZ_B200_GROUPS")) g = std::min<uint64_t>(g, std::max<uint64_t>(1, std::strtoull(e, nullptr, 10)));
        if (g < 1) continue;
        // modelled time: critical path x waves the batch takes x slowdown per resident warp (2.5 % each: the warps of an SM
        // share its issue slots and shared-memory pipe); an open-ended batch (chunk sizing) is scored by throughput
        const double contention = 1.0 + 0.025 * (double)(std::min(g, need) * cand);
        const double waves = n_groups == ~0ull ? 1.0 / (double)g : (double)((n_groups + (uint64_t)sm_count * g - 1) / ((uint64_t)sm_count * g));
        const double score = 1.0 / (probe->critical_cost * waves * contention);
        if (score > best) {
            best = score;
            R = cand;
            g_max = g;
            tables = tb;
            stage = stg;
        }
    }
    if (R == 0) return EZPZ_ERR_TOO_LARGE;
    uint64_t g = g_max;
    if (n_groups != ~0ull) {
        const uint64_t per_wave = (uint64_t)sm_count * g_max;
        const uint64_t waves = std::max<uint64_t>(1, (n_groups + per_wave - 1) / per_wave);
        const uint64_t slots = (uint64_t)sm_count * waves;
        g = std::min<uint64_t>(g_max, std::max<uint64_t>(1, (n_groups + slots - 1) / slots));
    }
    out->T = (uint32_t)(32 * g);
    out->R = R;
    out->stage = stage;
    out->smem = per_group * g + (size_t)cs->small.W * sizeof(double) + 8 + (stage ? tables : 0);
    return EZPZ_OK;
}

int32_t small_shape(ezpz_context* ctx, const ezpz_structure* cs, uint64_t batch, SmallShape* out) {
    return small_shape_host(cs, batch, (uint32_t)ctx->sm_count, ctx->smem_optin, out);
}

void release_device_copies(ezpz_structure* s) {
    std::lock_guard<std::mutex> lock(s->dev_mutex);
    for (DeviceCopy* d : s->dev) {
        if (cudaSetDevice(d->device) == cudaSuccess) {
            if (d->cons) cudaFree(d->cons);
            for (RoleTables* t : d->roles) {
                if (t->dev) cudaFree(t->dev);
                delete t;
            }
            if (d->csc_to_csr) cudaFree(d->csc_to_csr);
            if (d->csc_col_ptr) cudaFree(d->csc_col_ptr);
            if (d->csc_row_idx) cudaFree(d->csc_row_idx);
            if (!d->large.empty()) release_large(d);
        }
        delete d;
    }
    s->dev.clear();
}
}  // namespace ezs

extern "C" {

int32_t ezpz_b200_context_create(int32_t device, ezpz_context_t** out, ezpz_error_detail_t* detail) {
    if (!out) return EZPZ_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (detail) std::memset(detail, 0, sizeof *detail);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return cuda_fail(e, detail, "cudaGetDeviceCount");
    if (count == 0 || device < 0 || device >= count) {
        if (detail) std::snprintf(detail->message, sizeof detail->message, "no CUDA device %d (found %d)", device, count);
        return EZPZ_ERR_NO_DEVICE;
    }
    EZ_CUDA(cudaSetDevice(device), "cudaSetDevice");
    ezpz_context* ctx = new (std::nothrow) ezpz_context();
    if (!ctx) return EZPZ_ERR_INVALID_ARGUMENT;
    ctx->device = device;
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    ctx->sm_count = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    ctx->smem_optin = (size_t)v;
    {  // freed structure tables stay in the device's pool for the next structure (large.cu: DevicePlan)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
        }
    }
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return cuda_fail(e, detail, "cudaStreamCreate");
    }
    for (int k = 0; k < 3 && e == cudaSuccess; ++k) {
        e = cudaStreamCreateWithFlags(&ctx->pipe[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->pipe_done[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        delete ctx;
        return cuda_fail(e, detail, "cudaStreamCreate(pipeline)");
    }
    int32_t rc = upload_tables(detail);
    if (rc != EZPZ_OK) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return rc;
    }
    e = cudaSuccess;
    for (const void* fn : {(const void*)lm_small_kernel<true, 384>, (const void*)lm_small_kernel<false, 384>,
                           (const void*)lm_small_kernel<true, 512>, (const void*)lm_small_kernel<false, 512>,
                           (const void*)lm_small_kernel<true, 768>, (const void*)lm_small_kernel<false, 768>})
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    if (e != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return cuda_fail(e, detail, "cudaFuncSetAttribute(lm_small_kernel)");
    }
    *out = ctx;
    return EZPZ_OK;
}

void ezpz_b200_context_destroy(ezpz_context_t* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    for (int k = 0; k < 3; ++k) {
        if (ctx->pipe[k]) cudaStreamDestroy(ctx->pipe[k]);
        if (ctx->pipe_done[k]) cudaEventDestroy(ctx->pipe_done[k]);
    }
    for (cudaEvent_t e : ctx->lane_ev) cudaEventDestroy(e);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->fa_ws) cudaFree(ctx->fa_ws);
    if (ctx->fa_jac) cudaFree(ctx->fa_jac);
    if (ctx->fa_done) cudaEventDestroy(ctx->fa_done);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    ezs::release_structure_cache(ctx);
    delete ctx;
}

uint64_t ezpz_b200_context_launches(const ezpz_context_t* ctx) { return ctx ? ctx->launches : 0; }

int32_t ezpz_b200_context_synchronize(ezpz_context_t* ctx) {
    if (!ctx) return EZPZ_ERR_INVALID_ARGUMENT;
    ezpz_error_detail_t* detail = nullptr;
    EZ_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    return EZPZ_OK;
}

static int32_t launch_small(ezpz_context_t* ctx, const ezpz_structure_t* s, const ezpz_config_t* config, uint64_t batch,
                            const ezpz_batch_io_t* io, cudaStream_t st, ezpz_error_detail_t* detail);

// The solve + freedom analysis form of the batch call (io->under_mask set): the solve kernel leaves its cached Jacobians in a
// device buffer of the context, freedom_device reads them there.  Batches whose Jacobians exceed 1 GiB go piece by piece.
static int32_t solve_batch_with_analysis(ezpz_context_t* ctx, const ezpz_structure_t* s, const ezpz_config_t* config, uint64_t batch,
                                         const ezpz_batch_io_t* io, cudaStream_t st, ezpz_error_detail_t* detail) {
    const size_t nnz = s->csc_row_idx.size(), n = s->n, nc = s->n_cons, uw = (s->n_cons + 31) / 32, vw = (s->n + 31) / 32;
    const uint64_t piece = std::max<uint64_t>(1, std::min<uint64_t>(batch, ((uint64_t)1 << 30) / std::max<size_t>(8, nnz * 8)));
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t need = std::max<size_t>(256, piece * nnz * 8);
    if (need > ctx->fa_jac_bytes) {
        if (ctx->fa_jac) cudaFree(ctx->fa_jac);
        ctx->fa_jac = nullptr;
        ctx->fa_jac_bytes = 0;
        EZ_CUDA(cudaMalloc(&ctx->fa_jac, need), "cudaMalloc(jacobians for the freedom analysis)");
        ctx->fa_jac_bytes = need;
    }
    // the buffer is shared by every stream that solves with analysis on this context: they take turns
    if (ctx->fa_busy && ctx->fa_last_stream != st) EZ_CUDA(cudaStreamWaitEvent(st, ctx->fa_done, 0), "cudaStreamWaitEvent");
    for (uint64_t b0 = 0; b0 < batch; b0 += piece) {
        const uint64_t cnt = std::min(piece, batch - b0);
        ezpz_batch_io_t sub = *io;
        sub.guesses = io->guesses + b0 * n;
        sub.params = io->params ? io->params + b0 * nc : nullptr;
        sub.final_values = io->final_values + b0 * n;
        sub.iterations = io->iterations + b0;
        sub.status = io->status + b0;
        sub.unsat_mask = io->unsat_mask ? io->unsat_mask + b0 * uw : nullptr;
        sub.degen_count = io->degen_count ? io->degen_count + b0 * nc : nullptr;
        sub.jacobian = (double*)ctx->fa_jac;
        sub.under_mask = nullptr;
        int32_t rc = ezpz_b200_solve_batch_device(ctx, s, config, cnt, &sub, st, detail);
        if (rc != EZPZ_OK) return rc;
        if (io->jacobian && nnz)
            EZ_CUDA(cudaMemcpyAsync(io->jacobian + b0 * nnz, ctx->fa_jac, cnt * nnz * 8, cudaMemcpyDefault, st), "copy of the jacobians");
        rc = ezs::freedom_device(ctx, s, cnt, (const double*)ctx->fa_jac, io->under_mask + b0 * vw, st, detail);
        if (rc != EZPZ_OK) return rc;
    }
    return EZPZ_OK;
}

int32_t ezpz_b200_solve_batch_device(ezpz_context_t* ctx, const ezpz_structure_t* s, const ezpz_config_t* config,
                                     uint64_t batch, const ezpz_batch_io_t* io, void* cuda_stream,
                                     ezpz_error_detail_t* detail) {
    if (!ctx || !s || !config || !io) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (batch == 0) return EZPZ_OK;
    if (!io->guesses || !io->final_values || !io->iterations || !io->status) return EZPZ_ERR_INVALID_ARGUMENT;
    if (s->n_cons == 0 || s->m == 0) return EZPZ_ERR_EMPTY_SYSTEM;
    if (io->under_mask)
        return solve_batch_with_analysis(ctx, s, config, batch, io, cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream, detail);
    if (!s->small.valid)  // beyond the thread-per-problem kernel: the persistent LM kernel, one CTA per problem
        return ezs::solve_large_batch(ctx, s, config, batch, io, cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream, detail);
    return launch_small(ctx, s, config, batch, io, cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream, detail);
}

}  // extern "C"

static int32_t launch_small(ezpz_context_t* ctx, const ezpz_structure_t* s, const ezpz_config_t* config, uint64_t batch,
                            const ezpz_batch_io_t* io, cudaStream_t st, ezpz_error_detail_t* detail) {
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    DeviceCopy* dc = nullptr;
    int32_t rc = get_device_copy(ctx, s, &dc, detail);
    if (rc != EZPZ_OK) return rc;
    const SmallProgram& P = s->small;
    SmallShape shape;
    rc = small_shape(ctx, s, ctx->shape_batch ? ctx->shape_batch : batch, &shape);
    if (rc != EZPZ_OK) return rc;
    RoleTables* tables = nullptr;
    rc = get_role_tables(ctx, s, dc, shape.R, small_stride(shape.T), &tables, detail);
    if (rc != EZPZ_OK) return rc;
    SmallArgs a;
    a.tables = tables->dev;
    a.table_words = (uint32_t)tables->blob.words.size();
    a.cons_word = tables->blob.cons_word;
    a.guesses = io->guesses;
    a.params = io->params;
    a.finals = io->final_values;
    a.iterations = io->iterations;
    a.status = io->status;
    a.unsat = io->unsat_mask;
    a.degen = io->degen_count;
    a.jac = io->jacobian;
    a.nnz = (uint32_t)s->csc_row_idx.size();
    a.batch = batch;
    a.residual_tolerance = config->residual_tolerance;
    a.step_tolerance = config->step_tolerance;
    a.initial_lambda = config->initial_lambda;
    a.max_iterations = (uint32_t)std::min<uint64_t>(config->max_iterations, 0x7fffffffu);
    a.n_cons = s->n_cons;
    a.n = s->n;
    a.m = s->m;
    a.W = P.W;
    a.X0 = P.X0;
    a.R0 = P.R0;
    a.RN0 = P.RN0;
    a.J0 = P.J0;
    a.L0 = P.L0;
    a.D0 = P.D0;
    a.S0 = P.S0;
    a.F0 = P.F0;
    a.unsat_words = (s->n_cons + 31) / 32;
    a.T = shape.T;
    a.stride = small_stride(shape.T);
    a.R = shape.R;
    a.weights_one = s->all_weights_one ? 1u : 0u;
    const uint64_t grid = (batch + shape.T - 1) / shape.T;
    if (grid > 0x7fffffffull) return EZPZ_ERR_TOO_LARGE;
    const unsigned threads = shape.T * shape.R;
#define EZ_LAUNCH_SMALL(MAXT)                                                                            \
    do {                                                                                                 \
        if (shape.stage) lm_small_kernel<true, MAXT><<<(unsigned)grid, threads, shape.smem, st>>>(a);    \
        else lm_small_kernel<false, MAXT><<<(unsigned)grid, threads, shape.smem, st>>>(a);               \
    } while (0)
    if (threads <= 384) EZ_LAUNCH_SMALL(384);
    else if (threads <= 512) EZ_LAUNCH_SMALL(512);
    else EZ_LAUNCH_SMALL(768);
#undef EZ_LAUNCH_SMALL
    ctx->launches += 1;
    EZ_CUDA(cudaGetLastError(), "lm_small_kernel launch");
    return EZPZ_OK;
}

extern "C" {

// Problems per full wave of the batched kernel on this device (the chunks of the copy/compute pipeline are whole waves).
static uint64_t small_wave_problems(ezpz_context* ctx, const ezpz_structure* s) {
    if (!s->small.valid) return 0;
    SmallShape shape;
    if (small_shape(ctx, s, ~0ull, &shape) != EZPZ_OK) return 0;
    const size_t by_smem = std::max<size_t>(1, ctx->smem_optin / std::max<size_t>(1, shape.smem));
    const size_t by_threads = std::max<size_t>(1, 2048 / (shape.T * shape.R));
    return (uint64_t)ctx->sm_count * shape.T * std::min(by_smem, by_threads);
}

int32_t ezpz_b200_solve_batch(ezpz_context_t* ctx, const ezpz_structure_t* s, const ezpz_config_t* config,
                              uint64_t batch, const ezpz_batch_io_t* io, ezpz_error_detail_t* detail) {
    if (!ctx || !s || !config || !io) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (batch == 0) return EZPZ_OK;
    if (!io->guesses || !io->final_values || !io->iterations || !io->status) return EZPZ_ERR_INVALID_ARGUMENT;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    // Zero-copy form: when every buffer of the call is page-locked host memory the device can address (cudaHostAlloc /
    // cudaHostRegister / ezpz_b200_host_register), the kernel reads the guesses and writes the results across PCIe
    // itself, in coalesced 256-byte warp accesses (group_rows): no staging buffers, no copy calls, and the transfers of one
    // CTA overlap the arithmetic of the others instead of being pipelined by hand.  EZPZ_B200_ZERO_COPY=0 disables it.
    bool all_pinned = true;  // every buffer of the call is page-locked host memory
    if (s->small.valid) {
        const char* zc = std::getenv("EZPZ_B200_ZERO_COPY");  // (read per call: bench.py times both forms in one process)
        const bool zero_copy = !(zc && zc[0] == '0');
        ezpz_batch_io_t dio;
        bool ok = zero_copy;
        auto map = [&](const void* host, bool required) -> void* {
            if (!host) {
                if (required) ok = false;
                return nullptr;
            }
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
                cudaGetLastError();
                ok = all_pinned = false;
                return nullptr;
            }
            if ((at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeManaged) || !at.devicePointer) ok = all_pinned = false;
            return at.devicePointer;
        };
        dio.guesses = (const double*)map(io->guesses, true);
        dio.params = (const double*)map(io->params, false);
        dio.final_values = (double*)map(io->final_values, true);
        dio.iterations = (uint32_t*)map(io->iterations, true);
        dio.status = (uint8_t*)map(io->status, true);
        dio.unsat_mask = (uint32_t*)map(io->unsat_mask, false);
        dio.degen_count = (uint32_t*)map(io->degen_count, false);
        dio.jacobian = (double*)map(io->jacobian, false);
        dio.under_mask = (uint32_t*)map(io->under_mask, false);
        // Larger batches go through the copy/compute pipeline below instead: the copy engines move page-locked memory at link
        // rate (54-56 GB/s each way), kernel-issued reads across PCIe reach about half of that and stall a wave of CTAs at a
        // time (tools/time_e2e_split.py, tools/time_e2e_modes.py: 65,536 problems 433 us zero-copy, 347 us pipelined; 8,192
        // problems 108 us against 147 us).  EZPZ_B200_HOST_MODE=zerocopy|pipeline forces a form.
        const char* hm = std::getenv("EZPZ_B200_HOST_MODE");
        if ((hm && (hm[0] == 'p' || hm[0] == 'z')) ? hm[0] == 'p' : batch >= kPipelineMinBatch) ok = false;
        if (ok) {
            const int32_t rc = ezpz_b200_solve_batch_device(ctx, s, config, batch, &dio, ctx->stream, detail);
            if (rc != EZPZ_OK) return rc;
            EZ_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
            return EZPZ_OK;
        }
    }
    const size_t n = s->n, nc = s->n_cons, uw = (s->n_cons + 31) / 32;
    // Small calls on ordinary (pageable) memory — a single sketch through ezpz_b200_solve_one above all: every pageable
    // cudaMemcpyAsync is a staged, synchronising copy of ~10 us, and the call needs five of them.  Instead the buffers are
    // copied by the CPU into the context's page-locked block, the kernel reads and writes that block across PCIe, and the CPU
    // copies the results out: one launch and one synchronisation per call.
    if (s->small.valid) {
        const size_t nnz1 = s->csc_row_idx.size(), vw1 = (s->n + 31) / 32;
        const size_t o_g = 0, o_p = o_g + align_up(batch * n * 8, 64), o_f = o_p + (io->params ? align_up(batch * nc * 8, 64) : 0);
        const size_t o_it = o_f + align_up(batch * n * 8, 64), o_st = o_it + align_up(batch * 4, 64), o_un = o_st + align_up(batch, 64);
        const size_t o_dg = o_un + (io->unsat_mask ? align_up(batch * uw * 4, 64) : 0);
        const size_t o_jc = o_dg + (io->degen_count ? align_up(batch * nc * 4, 64) : 0);
        const size_t o_uc = o_jc + (io->jacobian ? align_up(batch * nnz1 * 8, 64) : 0);
        const size_t total = o_uc + (io->under_mask ? align_up(batch * vw1 * 4, 64) : 0);
        const char* hm = std::getenv("EZPZ_B200_HOST_MODE");
        if (total <= ((size_t)256 << 10) && !(hm && hm[0] == 'p')) {
            int32_t rc = ensure_pin(ctx, total, detail);
            if (rc != EZPZ_OK) return rc;
            char* h = (char*)ctx->pin;
            void* dv = nullptr;
            EZ_CUDA(cudaHostGetDevicePointer(&dv, ctx->pin, 0), "cudaHostGetDevicePointer");
            char* d = (char*)dv;
            std::memcpy(h + o_g, io->guesses, batch * n * 8);
            if (io->params) std::memcpy(h + o_p, io->params, batch * nc * 8);
            ezpz_batch_io_t dio;
            dio.guesses = (const double*)(d + o_g);
            dio.params = io->params ? (const double*)(d + o_p) : nullptr;
            dio.final_values = (double*)(d + o_f);
            dio.iterations = (uint32_t*)(d + o_it);
            dio.status = (uint8_t*)(d + o_st);
            dio.unsat_mask = io->unsat_mask ? (uint32_t*)(d + o_un) : nullptr;
            dio.degen_count = io->degen_count ? (uint32_t*)(d + o_dg) : nullptr;
            dio.jacobian = io->jacobian ? (double*)(d + o_jc) : nullptr;
            dio.under_mask = io->under_mask ? (uint32_t*)(d + o_uc) : nullptr;
            rc = ezpz_b200_solve_batch_device(ctx, s, config, batch, &dio, ctx->stream, detail);
            const cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (rc != EZPZ_OK) return rc;
            if (e != cudaSuccess) return cuda_fail(e, detail, "cudaStreamSynchronize");
            std::memcpy(io->final_values, h + o_f, batch * n * 8);
            std::memcpy(io->iterations, h + o_it, batch * 4);
            std::memcpy(io->status, h + o_st, batch);
            if (io->unsat_mask) std::memcpy(io->unsat_mask, h + o_un, batch * uw * 4);
            if (io->degen_count) std::memcpy(io->degen_count, h + o_dg, batch * nc * 4);
            if (io->jacobian) std::memcpy(io->jacobian, h + o_jc, batch * nnz1 * 8);
            if (io->under_mask) std::memcpy(io->under_mask, h + o_uc, batch * vw1 * 4);
            return EZPZ_OK;
        }
        // Mid-size calls on ordinary (pageable) memory — what a Rust Vec<f64> is: the copy engines cannot read it, and a pageable
        // cudaMemcpyAsync is a synchronous copy staged by the driver on one thread.  While the call fits the host's caches
        // (up to 8 MB: 16,384 two_rectangles problems) the host pool's threads (host_parallel.h) copy the inputs into the
        // context's page-locked block, the call runs there in its page-locked form, and the threads copy the results out:
        // 16,384 problems 196 us against 450 us, 4,096 problems 124 against 202 us (tools/time_pageable.py).  Larger calls
        // stay with the driver: behind sixteen copy threads the DMA of the freshly touched block crawls (the same pipeline
        // 1.5 ms instead of 0.34 ms at 65,536 problems; 1.1 ms in all with two copy threads against 1.6 ms for the driver, but
        // slower again at 32,768 and 131,072 — profiles/r02v_pageable_buffers.log), so only the clear case is taken.
        // EZPZ_B200_HOST_MODE=direct | staged forces a form, EZPZ_B200_STAGE_THREADS the copy threads of the forced form.
        const bool stage_here = hm && hm[0] == 's' ? true : (hm && hm[0] == 'd' ? false : total <= ((size_t)8 << 20));
        if (!all_pinned && total <= ((size_t)1 << 31) && stage_here) {
            int32_t rc = ensure_pin(ctx, total, detail);
            if (rc != EZPZ_OK) return rc;
            char* h = (char*)ctx->pin;
            static const uint32_t stage_threads = [] {  // (0 = as many as the pool has)
                const char* e = std::getenv("EZPZ_B200_STAGE_THREADS");
                return e ? (uint32_t)std::strtoul(e, nullptr, 10) : 0u;
            }();
            auto copy = [&](void* dst, const void* src, size_t bytes) {
                const uint32_t units = (uint32_t)((bytes + 63) / 64);
                const uint32_t grain = std::max<uint32_t>((1u << 18) / 64, stage_threads ? (units + stage_threads - 1) / stage_threads : 0u);
                if (bytes) ezs::parallel_ranges(units, grain, [&](uint32_t b, uint32_t e, uint32_t) {
                    std::memcpy((char*)dst + (size_t)b * 64, (const char*)src + (size_t)b * 64, std::min((size_t)e * 64, bytes) - (size_t)b * 64);
                });
            };
            static const bool trace4 = [] { const char* e = std::getenv("EZPZ_B200_DEBUG"); return e && e[0] == '4'; }();
            const auto t_a = std::chrono::steady_clock::now();
            copy(h + o_g, io->guesses, batch * n * 8);
            if (io->params) copy(h + o_p, io->params, batch * nc * 8);
            const auto t_b = std::chrono::steady_clock::now();
            ezpz_batch_io_t pio;
            pio.guesses = (const double*)(h + o_g);
            pio.params = io->params ? (const double*)(h + o_p) : nullptr;
            pio.final_values = (double*)(h + o_f);
            pio.iterations = (uint32_t*)(h + o_it);
            pio.status = (uint8_t*)(h + o_st);
            pio.unsat_mask = io->unsat_mask ? (uint32_t*)(h + o_un) : nullptr;
            pio.degen_count = io->degen_count ? (uint32_t*)(h + o_dg) : nullptr;
            pio.jacobian = io->jacobian ? (double*)(h + o_jc) : nullptr;
            pio.under_mask = io->under_mask ? (uint32_t*)(h + o_uc) : nullptr;
            rc = ezpz_b200_solve_batch(ctx, s, config, batch, &pio, detail);  // (every buffer page-locked now: no second staging)
            if (rc != EZPZ_OK) return rc;
            const auto t_c = std::chrono::steady_clock::now();
            copy(io->final_values, h + o_f, batch * n * 8);
            copy(io->iterations, h + o_it, batch * 4);
            copy(io->status, h + o_st, batch);
            if (io->unsat_mask) copy(io->unsat_mask, h + o_un, batch * uw * 4);
            if (io->degen_count) copy(io->degen_count, h + o_dg, batch * nc * 4);
            if (io->jacobian) copy(io->jacobian, h + o_jc, batch * nnz1 * 8);
            if (io->under_mask) copy(io->under_mask, h + o_uc, batch * vw1 * 4);
            if (trace4) {
                const auto t_d = std::chrono::steady_clock::now();
                auto us = [](auto x, auto y) { return std::chrono::duration<double, std::micro>(y - x).count(); };
                std::fprintf(stderr, "[solve_batch staged] copy in %.0f us, solve %.0f us, copy out %.0f us\n", us(t_a, t_b), us(t_b, t_c), us(t_c, t_d));
            }
            return EZPZ_OK;
        }
    }
    const size_t b_x = align_up(batch * n * sizeof(double), 256);
    const size_t b_p = io->params ? align_up(batch * nc * sizeof(double), 256) : 0;
    const size_t b_it = align_up(batch * sizeof(uint32_t), 256);
    const size_t b_st = align_up(batch, 256);
    const size_t b_un = io->unsat_mask ? align_up(batch * uw * sizeof(uint32_t), 256) : 0;
    const size_t b_dg = io->degen_count ? align_up(batch * nc * sizeof(uint32_t), 256) : 0;
    const size_t nnz = s->csc_row_idx.size();
    const size_t b_jc = io->jacobian ? align_up(batch * nnz * sizeof(double), 256) : 0;
    const size_t vw = (s->n + 31) / 32;
    const size_t b_uc = io->under_mask ? align_up(batch * vw * sizeof(uint32_t), 256) : 0;
    int32_t rc = ensure_ws(ctx, 2 * b_x + b_p + b_it + b_st + b_un + b_dg + b_jc + b_uc, detail);
    if (rc != EZPZ_OK) return rc;
    char* w = (char*)ctx->ws;
    double* d_g = (double*)w; w += b_x;
    double* d_f = (double*)w; w += b_x;
    double* d_p = io->params ? (double*)w : nullptr; w += b_p;
    uint32_t* d_it = (uint32_t*)w; w += b_it;
    uint8_t* d_st = (uint8_t*)w; w += b_st;
    uint32_t* d_un = io->unsat_mask ? (uint32_t*)w : nullptr; w += b_un;
    uint32_t* d_dg = io->degen_count ? (uint32_t*)w : nullptr; w += b_dg;
    double* d_jc = io->jacobian ? (double*)w : nullptr; w += b_jc;
    uint32_t* d_uc = io->under_mask ? (uint32_t*)w : nullptr; w += b_uc;
    // whatever fails from here on, nothing of this call may still be reading or writing the caller's buffers on return
    std::vector<cudaEvent_t> tev;  // (EZPZ_B200_DEBUG=3 trace events)
    struct Drain {
        ezpz_context* c;
        std::vector<cudaEvent_t>* ev;
        ~Drain() {
            for (int k = 0; k < 3; ++k) cudaStreamSynchronize(c->pipe[k]);
            cudaStreamSynchronize(c->stream);
            for (cudaEvent_t e : *ev) cudaEventDestroy(e);
        }
    } drain{ctx, &tev};
    // Copy/compute pipeline: the batch is cut into chunks that rotate over three streams, so the H2D copy of
    // chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap (with pinned host buffers the
    // copies are true DMA on the two copy engines; pageable buffers still work, just without the overlap).
    // A chunk is a whole number of kernel waves (a partial wave costs as much kernel time as a full one), at most
    // eight chunks per call.
    uint64_t chunk = batch;
    if (const uint64_t wave = small_wave_problems(ctx, s); wave > 0 && batch > wave) {
        const uint64_t waves = (batch + wave - 1) / wave;
        chunk = wave * ((waves + 7) / 8);
    }
    if (const char* env = std::getenv("EZPZ_B200_CHUNKS"); env && s->small.valid) {  // (the one-CTA-per-problem path keeps its
                                                                                      // per-problem state in ONE buffer: one chunk)
        const uint64_t k = std::max<uint64_t>(1, std::strtoull(env, nullptr, 10));
        chunk = (batch + k - 1) / k;
    }
    const uint64_t n_chunks = (batch + chunk - 1) / chunk;
    EZ_CUDA(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    // EZPZ_B200_DEBUG=3: CUDA events after every copy and kernel of the pipeline, printed per call (where a slow call loses time)
    static const bool trace = [] {
        const char* e = std::getenv("EZPZ_B200_DEBUG");
        return e && e[0] == '3';
    }();
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    // Lane form of the pipeline (batched-kernel structures without the fused analysis): ONE stream carries every copy-in back
    // to back, one every copy-out, two carry the kernels alternately, events link a chunk's three steps.  Each copy engine
    // then streams without gaps (with the steps of a chunk on one stream, copy-in k + 3 queues behind copy-out k), and chunks
    // smaller than a kernel wave overlap on the SMs.  EZPZ_B200_PIPE_LANES=<chunks> (0 = the three-stream form below).
    uint64_t lane_chunks = 0;
    // Measured (profiles/r02s_e2e_lane_pipeline.log): 65,536 problems 329 us with six lane chunks against 350 us for three
    // whole-wave chunks; no gain at 32,768 and at 262,144, so the lane form is the default in between only.
    if (s->small.valid && !d_jc && !d_uc && batch >= 16384) {
        lane_chunks = batch >= 49152 && batch <= 131072 ? 6 : 0;
        if (const char* env = std::getenv("EZPZ_B200_PIPE_LANES")) lane_chunks = std::strtoull(env, nullptr, 10);
    }
    if (lane_chunks > 0) {
        uint64_t per = (batch + lane_chunks - 1) / lane_chunks;
        per = (per + 31) / 32 * 32;  // whole 32-problem groups
        const uint64_t nl = (batch + per - 1) / per;
        while (ctx->lane_ev.size() < 2 * nl) {
            cudaEvent_t e;
            EZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
            ctx->lane_ev.push_back(e);
        }
        // chunks keep the CTA shape of the whole batch (full CTAs on a part of the SMs: the kernels of two chunks run side by side)
        struct Shape {
            ezpz_context* c;
            ~Shape() { c->shape_batch = 0; }
        } shape_guard{ctx};
        if (const char* env = std::getenv("EZPZ_B200_LANE_PACK"); !env || env[0] != '0') ctx->shape_batch = batch;
        cudaStream_t s_in = ctx->pipe[0], s_out = ctx->pipe[1], s_k[2] = {ctx->pipe[2], ctx->stream};
        if (trace) mark(s_in);
        for (uint64_t c = 0; c < nl; ++c) {
            const uint64_t b0 = c * per, cnt = std::min(per, batch - b0);
            EZ_CUDA(cudaMemcpyAsync(d_g + b0 * n, io->guesses + b0 * n, cnt * n * sizeof(double), cudaMemcpyHostToDevice, s_in), "H2D guesses");
            if (d_p) EZ_CUDA(cudaMemcpyAsync(d_p + b0 * nc, io->params + b0 * nc, cnt * nc * sizeof(double), cudaMemcpyHostToDevice, s_in), "H2D params");
            EZ_CUDA(cudaEventRecord(ctx->lane_ev[2 * c], s_in), "cudaEventRecord");
            cudaStream_t sk = s_k[c & 1];
            EZ_CUDA(cudaStreamWaitEvent(sk, ctx->lane_ev[2 * c], 0), "cudaStreamWaitEvent");
            ezpz_batch_io_t dio;
            dio.guesses = d_g + b0 * n;
            dio.params = d_p ? d_p + b0 * nc : nullptr;
            dio.final_values = d_f + b0 * n;
            dio.iterations = d_it + b0;
            dio.status = d_st + b0;
            dio.unsat_mask = d_un ? d_un + b0 * uw : nullptr;
            dio.degen_count = d_dg ? d_dg + b0 * nc : nullptr;
            dio.jacobian = nullptr;
            dio.under_mask = nullptr;
            rc = ezpz_b200_solve_batch_device(ctx, s, config, cnt, &dio, sk, detail);
            if (rc != EZPZ_OK) return rc;
            EZ_CUDA(cudaEventRecord(ctx->lane_ev[2 * c + 1], sk), "cudaEventRecord");
            EZ_CUDA(cudaStreamWaitEvent(s_out, ctx->lane_ev[2 * c + 1], 0), "cudaStreamWaitEvent");
            EZ_CUDA(cudaMemcpyAsync(io->final_values + b0 * n, d_f + b0 * n, cnt * n * sizeof(double), cudaMemcpyDeviceToHost, s_out), "D2H finals");
            if (d_dg) EZ_CUDA(cudaMemcpyAsync(io->degen_count + b0 * nc, d_dg + b0 * nc, cnt * nc * sizeof(uint32_t), cudaMemcpyDeviceToHost, s_out), "D2H degen");
            if (trace) mark(s_out);
        }
        // the small per-problem outputs in one copy each behind the last chunk's finals (the copy-out lane has waited for
        // every kernel); the copy-in lane is idle by now and takes two of them once the last kernel is done
        EZ_CUDA(cudaStreamWaitEvent(s_in, ctx->lane_ev[2 * (nl - 1) + 1], 0), "cudaStreamWaitEvent");
        if (nl > 1) EZ_CUDA(cudaStreamWaitEvent(s_in, ctx->lane_ev[2 * (nl - 2) + 1], 0), "cudaStreamWaitEvent");
        EZ_CUDA(cudaMemcpyAsync(io->iterations, d_it, batch * sizeof(uint32_t), cudaMemcpyDeviceToHost, s_in), "D2H iterations");
        EZ_CUDA(cudaMemcpyAsync(io->status, d_st, batch, cudaMemcpyDeviceToHost, s_in), "D2H status");
        if (d_un) EZ_CUDA(cudaMemcpyAsync(io->unsat_mask, d_un, batch * uw * sizeof(uint32_t), cudaMemcpyDeviceToHost, s_in), "D2H unsat");
        EZ_CUDA(cudaStreamSynchronize(s_out), "cudaStreamSynchronize");
        EZ_CUDA(cudaStreamSynchronize(s_in), "cudaStreamSynchronize");
        if (trace && tev.size() >= 2) {
            std::fprintf(stderr, "[solve_batch lanes] us since call start, copy-out of every chunk done:");
            for (size_t k = 1; k < tev.size(); ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, tev[0], tev[k]);
                std::fprintf(stderr, " %.0f", ms * 1e3);
            }
            std::fprintf(stderr, "\n");
        }
        return EZPZ_OK;
    }
    if (trace) mark(ctx->pipe[0]);
    for (uint64_t c = 0; c < n_chunks; ++c) {
        const uint64_t b0 = c * chunk;
        if (b0 >= batch) break;
        const uint64_t cnt = std::min(chunk, batch - b0);
        cudaStream_t st = ctx->pipe[c % 3];
        EZ_CUDA(cudaMemcpyAsync(d_g + b0 * n, io->guesses + b0 * n, cnt * n * sizeof(double), cudaMemcpyHostToDevice, st), "H2D guesses");
        if (d_p) EZ_CUDA(cudaMemcpyAsync(d_p + b0 * nc, io->params + b0 * nc, cnt * nc * sizeof(double), cudaMemcpyHostToDevice, st), "H2D params");
        mark(st);
        ezpz_batch_io_t dio;
        dio.guesses = d_g + b0 * n;
        dio.params = d_p ? d_p + b0 * nc : nullptr;
        dio.final_values = d_f + b0 * n;
        dio.iterations = d_it + b0;
        dio.status = d_st + b0;
        dio.unsat_mask = d_un ? d_un + b0 * uw : nullptr;
        dio.degen_count = d_dg ? d_dg + b0 * nc : nullptr;
        dio.jacobian = d_jc ? d_jc + b0 * nnz : nullptr;
        dio.under_mask = d_uc ? d_uc + b0 * vw : nullptr;
        rc = ezpz_b200_solve_batch_device(ctx, s, config, cnt, &dio, st, detail);
        if (rc != EZPZ_OK) break;
        mark(st);
        EZ_CUDA(cudaEventRecord(ctx->pipe_done[c % 3], st), "cudaEventRecord");  // this stream's kernels so far are done
        EZ_CUDA(cudaMemcpyAsync(io->final_values + b0 * n, d_f + b0 * n, cnt * n * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H finals");
        if (d_dg) EZ_CUDA(cudaMemcpyAsync(io->degen_count + b0 * nc, d_dg + b0 * nc, cnt * nc * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H degen");
        if (d_jc) EZ_CUDA(cudaMemcpyAsync(io->jacobian + b0 * nnz, d_jc + b0 * nnz, cnt * nnz * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H jacobian");
        if (d_uc) EZ_CUDA(cudaMemcpyAsync(io->under_mask + b0 * vw, d_uc + b0 * vw, cnt * vw * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H underconstrained");
        mark(st);
    }
    if (rc != EZPZ_OK) return rc;
    // The small per-problem outputs (9 bytes a problem: iterations, status, unsatisfied mask) leave in ONE copy each for the
    // whole batch once every chunk is done: per chunk they cost three more calls and copy-engine round trips than bytes.
    // They only wait for the KERNELS of all chunks (not for the last chunk's finals) and go out on the three streams side by side.
    {
        const int used = (int)std::min<uint64_t>(3, n_chunks);
        for (int q = 0; q < 3; ++q)
            for (int k = 0; k < used; ++k)
                if (k != q) EZ_CUDA(cudaStreamWaitEvent(ctx->pipe[q], ctx->pipe_done[k], 0), "cudaStreamWaitEvent");
        EZ_CUDA(cudaMemcpyAsync(io->iterations, d_it, batch * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->pipe[0]), "D2H iterations");
        EZ_CUDA(cudaMemcpyAsync(io->status, d_st, batch, cudaMemcpyDeviceToHost, ctx->pipe[1]), "D2H status");
        if (d_un) EZ_CUDA(cudaMemcpyAsync(io->unsat_mask, d_un, batch * uw * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->pipe[2]), "D2H unsat");
    }
    if (trace) mark(ctx->pipe[0]);
    for (int k = 0; k < 3; ++k) EZ_CUDA(cudaStreamSynchronize(ctx->pipe[k]), "cudaStreamSynchronize");
    if (trace && tev.size() >= 2) {
        std::fprintf(stderr, "[solve_batch] us since call start, per chunk (H2D done, kernel done, D2H done):");
        for (size_t k = 1; k < tev.size(); ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tev[0], tev[k]);
            std::fprintf(stderr, "%s%.0f", (k - 1) % 3 == 0 ? " | " : " ", ms * 1e3);
        }
        std::fprintf(stderr, "\n");
    }
    return EZPZ_OK;
}

int32_t ezpz_b200_solve_one(ezpz_context_t* ctx, const ezpz_structure_t* s, const ezpz_config_t* config,
                            const ezpz_one_io_t* io, ezpz_error_detail_t* detail) {
    if (!ctx || !s || !config || !io) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (!io->guesses || !io->final_values || !io->iterations || !io->status) return EZPZ_ERR_INVALID_ARGUMENT;
    if (io->lin_iters) *io->lin_iters = 0;
    if (s->small.valid) {
        ezpz_batch_io_t b;
        b.guesses = io->guesses;
        b.params = nullptr;
        b.final_values = io->final_values;
        b.iterations = io->iterations;
        b.status = io->status;
        b.unsat_mask = io->unsat_mask;
        b.degen_count = io->degen_count;
        b.jacobian = io->jacobian;
        b.under_mask = nullptr;
        if (io->path_used) *io->path_used = 0;
        return ezpz_b200_solve_batch(ctx, s, config, 1, &b, detail);
    }
    return ezs::solve_large(ctx, s, config, io, detail);
}

int32_t ezpz_b200_eval(ezpz_context_t* ctx, const ezpz_structure_t* s, const double* x, double* r, double* jac_csc,
                       double* jac_csr, uint8_t* degen, ezpz_error_detail_t* detail) {
    if (!ctx || !s || !x) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (s->n_cons == 0) return EZPZ_OK;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    DeviceCopy* dc = nullptr;
    int32_t rc = get_device_copy(ctx, s, &dc, detail);
    if (rc != EZPZ_OK) return rc;
    {
        std::lock_guard<std::mutex> lock(const_cast<ezpz_structure*>(s)->dev_mutex);
        if (!dc->cons) {
            EZ_CUDA(cudaMalloc(&dc->cons, sizeof(DevCons) * s->n_cons), "cudaMalloc(cons)");
            EZ_CUDA(cudaMemcpy(dc->cons, s->dev_cons.data(), sizeof(DevCons) * s->n_cons, cudaMemcpyHostToDevice), "cudaMemcpy(cons)");
        }
        if (!dc->csc_to_csr && !s->csc_to_csr.empty()) {
            EZ_CUDA(cudaMalloc(&dc->csc_to_csr, sizeof(uint32_t) * s->csc_to_csr.size()), "cudaMalloc(perm)");
            EZ_CUDA(cudaMemcpy(dc->csc_to_csr, s->csc_to_csr.data(), sizeof(uint32_t) * s->csc_to_csr.size(), cudaMemcpyHostToDevice),
                    "cudaMemcpy(perm)");
        }
    }
    const size_t n = s->n, m = s->m, nnz = s->csc_row_idx.size(), nc = s->n_cons;
    const size_t b_x = align_up(n * 8, 256), b_r = align_up(m * 8, 256), b_j = align_up(nnz * 8, 256), b_d = align_up(nc, 256);
    rc = ensure_ws(ctx, b_x + b_r + 2 * b_j + b_d, detail);
    if (rc != EZPZ_OK) return rc;
    char* w = (char*)ctx->ws;
    double* d_x = (double*)w; w += b_x;
    double* d_r = (double*)w; w += b_r;
    double* d_j = (double*)w; w += b_j;
    double* d_j2 = (double*)w; w += b_j;
    uint8_t* d_d = (uint8_t*)w;
    cudaStream_t st = ctx->stream;
    EZ_CUDA(cudaMemcpyAsync(d_x, x, n * 8, cudaMemcpyHostToDevice, st), "H2D x");
    // Structures that take the large path are evaluated by that path's own assembly kernel; the generic kernel
    // below then only supplies what that kernel does not keep (the CSR permutation, the two degenerate bits).
    bool large_done = false, large_csr = false;
    if (s->large.built) {
        rc = ezs::eval_large(ctx, s, x, r, jac_csc, jac_csr, &large_csr, detail);
        if (rc != EZPZ_OK) return rc;
        large_done = true;
        if (!degen && (!jac_csr || large_csr)) return EZPZ_OK;
    }
    AsmArgs a;
    a.cons = dc->cons;
    a.x = d_x;
    a.r = d_r;
    a.jvals = d_j;
    a.degen = d_d;
    a.n_cons = s->n_cons;
    const unsigned T = 128, grid = (unsigned)((nc + T - 1) / T);
    assemble_kernel<true><<<grid, T, 0, st>>>(a);
    ctx->launches += 1;
    EZ_CUDA(cudaGetLastError(), "assemble_kernel launch");
    if (r && !large_done) EZ_CUDA(cudaMemcpyAsync(r, d_r, m * 8, cudaMemcpyDeviceToHost, st), "D2H r");
    if (jac_csc && !large_done) EZ_CUDA(cudaMemcpyAsync(jac_csc, d_j, nnz * 8, cudaMemcpyDeviceToHost, st), "D2H jac");
    if (jac_csr && nnz && !(large_done && large_csr)) {
        permute_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(d_j, dc->csc_to_csr, d_j2, (uint32_t)nnz);
        ctx->launches += 1;
        EZ_CUDA(cudaGetLastError(), "permute_kernel launch");
        EZ_CUDA(cudaMemcpyAsync(jac_csr, d_j2, nnz * 8, cudaMemcpyDeviceToHost, st), "D2H jac csr");
    }
    if (degen) EZ_CUDA(cudaMemcpyAsync(degen, d_d, nc, cudaMemcpyDeviceToHost, st), "D2H degen");
    EZ_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return EZPZ_OK;
}

int32_t ezpz_b200_structure_batch_shape(const ezpz_structure_t* s, uint64_t batch, uint32_t sm_count, uint64_t smem_per_block,
                                        uint32_t* roles, uint32_t* problems_per_cta) {
    if (!s) return EZPZ_ERR_INVALID_ARGUMENT;
    SmallShape shape;
    const int32_t rc = ezs::small_shape_host(s, batch, sm_count, (size_t)smem_per_block, &shape);
    if (rc != EZPZ_OK) return rc;
    if (roles) *roles = shape.R;
    if (problems_per_cta) *problems_per_cta = shape.T;
    return EZPZ_OK;
}

void ezpz_b200_angle_sincos(double radians, double* sin_out, double* cos_out) {
    double s, c;
    ezm::ez_sincos(radians, s, c);
    if (sin_out) *sin_out = s;
    if (cos_out) *cos_out = c;
}

double ezpz_b200_hypot(double x, double y) { return ezm::ez_hypot(x, y); }

}  // extern "C"
