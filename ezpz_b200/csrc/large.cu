// large.cu — one large system on the device: the whole Levenberg–Marquardt loop of
// ezpz/src/solver/newton.rs:29-145 for systems that do not fit the thread-per-problem kernel (configs 3 and
// 4 of BASELINE.json: massive_parallel_system at 2,000 x 2,000 and the synthetic 1M-variable sketch).
//
// One persistent cooperative kernel (lm_large_kernel) runs assembly, the damped step solve, the tentative
// step, accept/reject, lambda adaptation and both convergence tests; the host launches it once and reads the
// result.  Phases are separated by barriers whose kind follows the launch shape: __syncthreads for one CTA (tiny
// systems), the hardware cluster barrier for one thread-block cluster of 8 CTAs (mid-size systems such as the
// 2,000-variable massive_parallel_system), the cooperative grid barrier for the whole GPU.
//   * assembly      one thread per constraint; record tiles of 32 constraints of one kind (tile-local kind sort), so
//                   a warp runs one kind and neighbouring tiles touch neighbouring rows, slots and variables;
//                   residuals to r, partials through precomputed slots into J (CSC order) and, on the PCG path, a
//                   CSR-ordered copy for the row-wise SpMV (Model::residual / refresh_jacobian, solver.rs:318-440);
//   * step solve    (a) supernodal sparse Cholesky (sparse_direct.cpp: natural order for shallow elimination trees
//                   such as the block-diagonal massive_parallel_system, nested dissection otherwise; relaxed
//                   etree-chain supernodes stored as dense panels): one phase per stage of the supernode tree for
//                   factor + forward substitution, one per stage for the backward substitution; a panel is
//                   factorised left-looking by one thread, one warp or one CTA depending on its size and on how
//                   many panels the stage has.  Same arithmetic-order spec as the small path applied to P A Pt, so
//                   results match the oracle (given the same order) bit for bit;
//                   (b) when the factor would be too large, Jacobi-preconditioned conjugate gradients on
//                   (JtJ + lambda I) d = -Jt r, applied matrix-free as Jt (J p) + lambda p (two SpMVs/iteration);
//   * sum r^2       single-CTA systems: a strictly sequential left fold, as Rust's `.map(|x| x * x).sum()` is
//                   (newton.rs:45,116) — the accept test `S' < S` is a floating-point tie-breaker, so the
//                   summation order is part of the reference semantics.  Multi-CTA systems (more than 4,096
//                   values): the same fold inside chunks of rows, then a sequential fold of the chunk sums (a 1M-row
//                   sequential chain would cost 5 ms per evaluation); the chunk is the power of two next to sqrt(m) within
//                   [64, 1024] (ezs::sum_chunk_for) and the oracle reproduces the fold with that sum_chunk.  max|r| and
//                   max|d| are order-independent.
// spmv_csr_kernel / assemble_large_kernel are also exported as stand-alone launches (ezpz_b200_large_bench)
// so that their HBM throughput can be timed with CUDA events and profiled with ncu.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "device.h"
#include "eval.cuh"

namespace cg = cooperative_groups;
#ifndef EZPZ_PAIR_TEAM_MAX
#define EZPZ_PAIR_TEAM_MAX 32
#endif
using namespace ezs;

namespace {

__constant__ uint8_t cl_rows[EZPZ_K_COUNT];
__constant__ uint8_t cl_emit_len[EZPZ_K_COUNT][2];
__constant__ uint8_t cl_nids[EZPZ_K_COUNT];

struct LargeCtrl {
    double lambda, S, S2, largest, step;
    double rz, bb;
    uint32_t iterations, converged, fail, done;
    uint32_t any_unsat, any_degen, lin_iters, pad;
    // nanoseconds per phase, accumulated by thread 0 (printed by solve_large when EZPZ_B200_DEBUG=1):
    // 0 assembly + residual sweeps, 1 sums and maxima, 2 A/rhs assembly, 3 factor levels (grid), 4 top of the tree
    // (one CTA: factor + backward), 5 backward levels (grid), 6 PCG, 7 everything
    unsigned long long t[8];
};

// (record tiles: see the comment above assemble_slot)
struct KindLayout {  // word offsets inside a record; 0xff = absent
    uint8_t n_words, ids, p0, p1, weight, slot0, slot1, jr;
};
struct TileDesc {
    uint32_t off16;  // tile start in the record array, in 16-byte units
    uint32_t meta;   // kind | n_valid << 8
};
constexpr uint32_t kMaxRecWords = 1 + 8 + 2 + 2 + 2 + 16 + 3;  // row0, ids, p0, p1, weight, slots, CSR base + 2 offset words

struct LargeArgs {
    const uint32_t* recs;      // record tiles, back to back
    const TileDesc* tiles;     // [n_tiles]
    const uint32_t* slot_orig; // processing slot -> index in the caller's constraint list (rare paths)
    KindLayout layout[EZPZ_K_COUNT];
    const uint32_t *csr_row_ptr, *csr_col_idx, *csc_col_ptr, *csc_row_idx, *csc_to_csr;
    // sparse direct path (structure.h: LargeProgram)
    const uint32_t *perm, *sn_rows, *upd_rel, *upd_rec, *stage_ptr, *stage_rec, *aent_slot, *aprod_ptr, *aprod_a, *aprod_b, *diag_slot;
    double* vg;          // [x | r | rn | J csc | L panels | 1/pivot | y | d]
    double* jr;          // J values in CSR order (PCG path)
    double* cgv;         // PCG vectors: p, res, ap, dinv (n each), q (m)
    double* partials;    // 3 * gridDim.x
    double* sumsq;       // chunk sums of r^2 (multi-CTA grids)
    unsigned long long* lvl_ns;  // debug (EZPZ_B200_DEBUG=1): per level, ns spent in factor / backward phases, else NULL
    uint8_t* side;       // resolved side per processing slot (tangent kinds only)
    const uint8_t* side_flags;  // the caller's side per processing slot (0 = Undefined: resolved from the guesses)
    uint32_t* degen;     // per-constraint Warning::Degenerate counters
    uint32_t* unsat;     // bit mask
    LargeCtrl* ctrl;
    double residual_tolerance, step_tolerance, initial_lambda, cg_rtol;
    uint32_t max_iterations, cg_max_iters;
    uint32_t n_cons, n_slots, n_tiles, tile_bytes_max, n, m, nnz, n_levels, nnz_l, n_aent;
    uint32_t unit_weights;
    uint32_t cluster;  // launched as one thread-block cluster (mid-size systems)
    const uint32_t* jmap;   // direct path: position in the J region of each CSC entry (tile order); null = CSC order
    uint32_t sum_chunk;     // sum of squares folded in chunks of this many rows (ezs::sum_chunk_for); 0 = one sequential fold
    // CTAs cooperating on ONE system and this CTA's rank among them (set by the kernel: the whole grid, or 1 / 0 in
    // batch mode, where every CTA solves its own problem with CTA-level barriers)
    uint32_t vgrid;  // set by the host: the launch's grid size, or 1 in batch mode (this CTA's rank: vblock_of)
    // batch mode (ezpz_b200_solve_batch on structures beyond the thread-per-problem kernel): problem b = blockIdx.x uses
    // vg + b * vg_stride, jr + b * jr_stride, ... ; 0 = one system per launch
    uint32_t batch;
    size_t vg_stride, jr_stride, cgv_stride, sumsq_stride, side_stride, degen_stride, unsat_stride;
    uint32_t X0, R0, RN0, J0, L0, RV0, Y0, D0;
    uint32_t direct;
    uint32_t pipe_asm;  // in-kernel assembly through the bulk-copy pipeline (assemble_phase_pipe)
    uint32_t* sn_flag;  // [n_sn] per supernode in stage order: the factorisation whose diagonal block is final (sn_factor_slice)
    uint32_t n_sn;
};

// Rank of this CTA among the CTAs that cooperate on one system: the block index, or 0 in batch mode (one CTA per problem).
__device__ __forceinline__ uint32_t vblock_of(const LargeArgs& a) { return a.batch ? 0u : blockIdx.x; }

struct GX {
    const double* p;
    __device__ __forceinline__ double operator()(uint32_t id) const { return p[id]; }
};
// Variables already gathered into registers; "ids" are then the positions 0..7.
struct RegX {
    const double* v;
    __device__ __forceinline__ double operator()(uint32_t pos) const { return v[pos]; }
};

// Record tiles — the analysed constraints as the large path reads them.  Processing order (structure.cpp): tiles
// of kAssemblyTile consecutive input constraints, stably sorted by kind inside the tile, every kind group padded
// to whole warps.  One RECORD TILE = 32 consecutive processing slots = ONE kind; it stores only the words that
// kind needs (KindLayout), word-transposed: word w of slot l at tile[w * 32 + l], so a tile is n_words * 128
// contiguous bytes.  A warp therefore runs one kind (no divergence in the per-kind switch), reads its tile as one
// contiguous block — the stand-alone assembly kernel pulls it into shared memory with ONE bulk async copy
// (cp.async.bulk + mbarrier, three to eight tiles in flight per warp) — and the rows, Jacobian slots and variables a group
// of neighbouring tiles touches stay within a few hundred KB, so partial-sector writes merge in L2.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ double rec_double(const uint32_t* rec, uint32_t word) {
    return __hiloint2double((int)rec[(word + 1) * 32], (int)rec[word * 32]);
}

// One constraint: `rec` points at this lane's column of the tile (global or shared memory), k = processing slot.
//   RES: weighted residuals into vg[rdst + row];  JAC: Jacobian values into J (CSC order, through the precomputed
//   scatter slots) and, when write_jr, into the CSR-ordered copy used by the row-wise SpMV (the entries of a
//   constraint's rows are contiguous there: one base + 4-bit offsets).
// The (at most 8) variables of one constraint, gathered BEFORE the per-kind switch so that the warp issues its
// gathers together.  Ids mostly come as the (x, y) pair of a point at (2i, 2i + 1): such a pair is one aligned
// 16-byte load, which halves the sectors the L1 has to look up.
__device__ __forceinline__ void gather_x(const LargeArgs& a, const KindLayout ly, uint32_t kind, const uint32_t* rec, double (&xv)[8]) {
    const double* __restrict__ x = a.vg + a.X0;  // X0 == 0: 16-byte aligned
    const uint32_t nids = cl_nids[kind];
#pragma unroll
    for (uint32_t q = 0; q < 8; q += 2) {
        xv[q] = 0.0;
        xv[q + 1] = 0.0;
        if (q < nids) {
            const uint32_t i0 = rec[(ly.ids + q) * 32];
            if (q + 1 < nids) {
                const uint32_t i1 = rec[(ly.ids + q + 1) * 32];
                if (i1 == i0 + 1 && !(i0 & 1u)) {
                    const double2 v = *reinterpret_cast<const double2*>(x + i0);
                    xv[q] = v.x;
                    xv[q + 1] = v.y;
                } else {
                    xv[q] = x[i0];
                    xv[q + 1] = x[i1];
                }
            } else {
                xv[q] = x[i0];
            }
        }
    }
}

template <bool RES, bool JAC, bool NO_JR = false>
__device__ __forceinline__ void assemble_slot(const LargeArgs& a, const KindLayout ly, uint32_t kind, const uint32_t* rec, uint32_t k,
                                              uint32_t rdst, bool write_jr, const double (&xv)[8], uint32_t side) {
    const double p0 = ly.p0 != 0xff ? rec_double(rec, ly.p0) : 0.0;
    const double p1 = ly.p1 != 0xff ? rec_double(rec, ly.p1) : 0.0;
    const double w = ly.weight != 0xff ? rec_double(rec, ly.weight) : 1.0;
    const uint32_t row0 = rec[0];
    // (argument fields are read once: inside the LM kernel the block lives in local memory and every store below would
    // otherwise be followed by reloads of a.vg / a.J0 / a.jr)
    double* const rout = a.vg + rdst + row0;
    double* const jbase = a.vg + a.J0;
    double* const jrbase = a.jr;
    const uint32_t ident[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    const RegX XR{xv};
    ezd::EvalOut o;
    ezd::eval_constraint<JAC>(kind, side, ident, p0, p1, XR, o);
    const uint32_t rows = cl_rows[kind];
    uint32_t ndeg = 0;
    if (RES) {
        rout[0] = w * o.res[0];
        if (rows == 2) rout[1] = w * o.res[1];
        if (o.res_degen) ++ndeg;
    }
    if (JAC) {
        if (o.jac_degen) ++ndeg;
        const bool jr_on = !NO_JR && write_jr && ly.jr != 0xff;
        const uint32_t jr0 = jr_on ? rec[ly.jr * 32] : 0u;
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            if (row < (int)rows) {
                const uint32_t len = cl_emit_len[kind][row];
                const uint32_t jrc = jr_on ? rec[(ly.jr + 1 + row) * 32] : 0u;
                const uint32_t sb = row == 0 ? ly.slot0 : ly.slot1;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q < (int)len) {
                        const uint32_t s = rec[(sb + q) * 32];
                        double* dst = jbase + (s & ~kAccumulate);
                        double v;
                        if (s & kAccumulate) v = o.emit[row] ? *dst + w * o.pd[row][q] : *dst;
                        else v = o.emit[row] ? 0.0 + w * o.pd[row][q] : 0.0;
                        *dst = v;
                        if (jr_on) jrbase[jr0 + ((jrc >> (4 * q)) & 15u)] = v;
                    }
                }
            }
        }
    }
    if (ndeg) {
        a.degen[a.slot_orig[k]] += ndeg;
        a.ctrl->any_degen = 1;
    }
}

// All tiles by the calling warps (gw = global warp index of nw), records read straight from global memory.
template <bool RES, bool JAC>
__device__ __forceinline__ void assemble_phase_inl(const LargeArgs& a, uint32_t rdst, uint32_t tid, uint32_t nth, bool write_jr) {
    const uint32_t lane = threadIdx.x & 31u, nw = nth >> 5;
    // (argument fields read once, see assemble_slot)  The next tile's descriptor is fetched one iteration ahead and its
    // record lines are pulled into L2 while the current tile is evaluated.
    const TileDesc* const tiles = a.tiles;
    const uint32_t* const recs = a.recs;
    const uint8_t* const sides = a.side;
    const uint32_t n_tiles = a.n_tiles;
    uint32_t t = tid >> 5;
    TileDesc td = t < n_tiles ? tiles[t] : TileDesc{0, 0};
    for (; t < n_tiles; t += nw) {
        const TileDesc tn = t + nw < n_tiles ? tiles[t + nw] : TileDesc{0, 0};
        const uint32_t kind = td.meta & 0xffu, n_valid = td.meta >> 8;
        const KindLayout ly = a.layout[kind];
        if (lane < n_valid) {
            const uint32_t* rec = recs + (size_t)td.off16 * 4 + lane;
            double xv[8];
            gather_x(a, ly, kind, rec, xv);
            const uint32_t side = (kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE || kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE) ? sides[t * 32 + lane] : 0u;
            if (t + nw < n_tiles) {
                const uint32_t lines = a.layout[tn.meta & 0xffu].n_words;  // one 128-byte line per record word
                for (uint32_t q = lane; q < lines; q += 32) prefetch_l2(recs + (size_t)tn.off16 * 4 + 32 * q);
            }
            assemble_slot<RES, JAC>(a, ly, kind, rec, t * 32 + lane, rdst, write_jr, xv, side);
        }
        td = tn;
    }
}
// Out-of-line copy for the persistent kernel (keeps its register allocation apart from the solver phases).
template <bool RES, bool JAC>
__device__ __noinline__ void assemble_phase(const LargeArgs& a, uint32_t rdst, uint32_t tid, uint32_t nth, bool write_jr) {
    assemble_phase_inl<RES, JAC>(a, rdst, tid, nth, write_jr);
}

// ---- bulk-copy staging (stand-alone assembly kernel) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

constexpr uint32_t kAsmStagesMax = 8;   // tiles in flight per warp: as many as fit its 13 KB of shared memory (3 for the
                                        // widest record layout, 4-8 for structures whose kinds need fewer words)
constexpr uint32_t kTileBytesMax = kMaxRecWords * 128;    // 4352

// NaN-ignoring max |v[i]| over the grid: per-block partial to partials[blockIdx.x]; caller syncs, then every
// thread folds the partials (same order everywhere).
// The assembly phase INSIDE the persistent LM kernel, pipelined like the stand-alone kernel below: every warp keeps two to
// three record tiles in flight with cp.async.bulk + one mbarrier per stage (its slice of the CTA's dynamic shared memory, the
// stage the factorisation teams use in their phases), its tile descriptors come through a 64-entry ring one block ahead, and
// the variable gathers of tile k + 1 are issued before tile k is evaluated.  Out of line on purpose (its own register
// allocation; the solver phases of the kernel are at the register cap).  Falls back to assemble_phase when a warp's slice does
// not hold two tiles.
template <bool RES, bool JAC>
__device__ __noinline__ void assemble_phase_pipe(const LargeArgs& a, uint32_t rdst, uint32_t tid, uint32_t nth, bool write_jr,
                                                 double* warp_stage, uint32_t warp_stage_bytes) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = tid >> 5, nw = nth >> 5;
    const uint32_t kStageBytes = a.tile_bytes_max;
    constexpr uint32_t kTail = kAsmStagesMax * 8u + 64u * 8u;  // mbarriers + descriptor ring
    const uint32_t kAsmStages = min(kAsmStagesMax, (warp_stage_bytes - kTail) / max(kStageBytes, 128u));
    unsigned char* stage_base = reinterpret_cast<unsigned char*>(warp_stage);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + warp_stage_bytes - kTail);
    uint2* ring = reinterpret_cast<uint2*>(bars + kAsmStagesMax);
    // The stage was last written with ordinary stores by the factorisation teams (any warp of the CTA, CTA teams span the
    // warps' slices) and is about to be written by the async proxy: every thread orders its own stores in front of the
    // async proxy, then the CTA meets, then the copies are issued.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (lane == 0) {
        for (uint32_t st = 0; st < kAsmStages; ++st) mbar_init(&bars[st], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const uint32_t count = gw < a.n_tiles ? (a.n_tiles - gw + nw - 1) / nw : 0u;
    const TileDesc* const tiles = a.tiles;
    const uint32_t* const recs = a.recs;
    const uint8_t* const sides = a.side;
    auto load_block = [&](uint32_t b) {
        const uint32_t k = b * 32 + lane;
        return k < count ? *reinterpret_cast<const uint2*>(&tiles[gw + k * nw]) : make_uint2(0u, 0u);
    };
    auto desc = [&](uint32_t k) {
        const uint2 v = ring[k & 63u];
        return TileDesc{v.x, v.y};
    };
    ring[lane] = load_block(0);
    ring[32 + lane] = load_block(1);
    uint2 pf = load_block(2);
    __syncwarp();
    auto issue = [&](uint32_t st, const TileDesc td) {  // lane 0: start the copy of a tile into stage st
        const uint32_t bytes = (uint32_t)a.layout[td.meta & 0xffu].n_words * 128u;
        mbar_expect_tx(&bars[st], bytes);
        bulk_copy_g2s(stage_base + (size_t)st * kStageBytes, recs + (size_t)td.off16 * 4, bytes, &bars[st]);
    };
    if (lane == 0)
        for (uint32_t k = 0; k < kAsmStages && k < count; ++k) issue(k, desc(k));
    auto is_tangent = [](uint32_t kind) { return kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE || kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE; };
    double xv[8], xn[8];
    uint32_t side = 0, side_n = 0;
    TileDesc td = desc(0);
    if (count) {
        mbar_wait(&bars[0], 0);
        if (lane < (td.meta >> 8)) {
            gather_x(a, a.layout[td.meta & 0xffu], td.meta & 0xffu, reinterpret_cast<const uint32_t*>(stage_base) + lane, xv);
            if (is_tangent(td.meta & 0xffu)) side = sides[gw * 32 + lane];
        }
    }
    uint32_t st = 0, sn = kAsmStages > 1 ? 1u : 0u, par_n = kAsmStages > 1 ? 0u : 1u;
    for (uint32_t k = 0, t = gw; k < count; ++k, t += nw) {
        if ((k & 31u) == 0 && k) {
            const uint32_t b = k >> 5;
            ring[((b + 1) & 1u) * 32 + lane] = pf;
            pf = load_block(b + 2);
            __syncwarp();
        }
        const uint32_t kind = td.meta & 0xffu, n_valid = td.meta >> 8;
        TileDesc tn{0, 0};
        if (k + 1 < count) {
            tn = desc(k + 1);
            mbar_wait(&bars[sn], par_n);
            if (lane < (tn.meta >> 8)) {
                gather_x(a, a.layout[tn.meta & 0xffu], tn.meta & 0xffu,
                         reinterpret_cast<const uint32_t*>(stage_base + (size_t)sn * kStageBytes) + lane, xn);
                if (is_tangent(tn.meta & 0xffu)) side_n = sides[(t + nw) * 32 + lane];
            }
        }
        if (lane < n_valid)
            assemble_slot<RES, JAC>(a, a.layout[kind], kind, reinterpret_cast<const uint32_t*>(stage_base + (size_t)st * kStageBytes) + lane,
                                    t * 32 + lane, rdst, write_jr, xv, side);
        __syncwarp();
        if (lane == 0 && k + kAsmStages < count) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(st, desc(k + kAsmStages));
        }
        td = tn;
        side = side_n;
#pragma unroll
        for (int q = 0; q < 8; ++q) xv[q] = xn[q];
        st = sn;
        if (++sn == kAsmStages) {
            sn = 0;
            par_n ^= 1u;
        }
    }
    __syncwarp();
    if (lane == 0)  // the shared memory goes back to the factorisation teams: the barrier objects end here
        for (uint32_t s2 = 0; s2 < kAsmStages; ++s2) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[s2])) : "memory");
    __syncwarp();
}

__device__ void max_abs_partial(const double* v, uint32_t count, uint32_t tid, uint32_t nth, double* partial_out,
                                double* sm) {
    double mx = __longlong_as_double(0x7ff8000000000000LL);  // NaN: identity of the NaN-ignoring max
    for (uint32_t i = tid; i < count; i += nth) mx = ezm::ez_fmax(mx, ezm::ez_abs(v[i]));
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = ezm::ez_fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *partial_out = sm[0];
    __syncthreads();
}
// Fold of the per-CTA partials, once per CTA and broadcast through shared memory: the CTA loads the partials with one
// coalesced pass (a single thread walking them in global memory pays an L2 round trip per handful of values), then a
// fixed tree reduces them (the NaN-ignoring max does not depend on the order).  Must be called by every thread of the CTA.
__device__ double fold_max(const double* partials, uint32_t count, double* sm) {
    __syncthreads();
    double mx = __longlong_as_double(0x7ff8000000000000LL);
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) mx = ezm::ez_fmax(mx, partials[i]);
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = ezm::ez_fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    const double v = sm[0];
    __syncthreads();
    return v;
}
// Deterministic block sum (fixed tree) of one value per thread -> partial_out.
__device__ void block_sum(double v, double* partial_out, double* sm) {
    sm[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x] + sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *partial_out = sm[0];
    __syncthreads();
}
// Sequential fold (index order, from +0.0) of `count` values in global memory, once per CTA: staged into shared memory
// kSmDoubles at a time by the whole CTA, added in order by thread 0.
__device__ double fold_sum(const double* partials, uint32_t count, double* sm) {
    constexpr uint32_t kStage = 4096 - 1;  // sm[kStage] carries the running sum / the result
    __syncthreads();
    double s = 0.0;
    for (uint32_t base = 0; base < count; base += kStage) {
        const uint32_t len = min(kStage, count - base);
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) sm[i] = partials[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t i = 0;
            for (; i + 8 <= len; i += 8) {
                const double s0 = sm[i], s1 = sm[i + 1], s2 = sm[i + 2], s3 = sm[i + 3];
                const double s4 = sm[i + 4], s5 = sm[i + 5], s6 = sm[i + 6], s7 = sm[i + 7];
                s = s + s0; s = s + s1; s = s + s2; s = s + s3;
                s = s + s4; s = s + s5; s = s + s6; s = s + s7;
            }
            for (; i < len; ++i) s = s + sm[i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) sm[kStage] = s;
    __syncthreads();
    const double v = sm[kStage];
    __syncthreads();
    return v;
}

// Sequential left fold of v[i]^2 (block 0 only): the CTA stages squares in shared memory chunk by chunk, one
// thread adds them in index order.
__device__ void sequential_sum_squares(const double* v, uint32_t count, double* out, double* sm, uint32_t sm_len) {
    double acc = 0.0;
    for (uint32_t base = 0; base < count; base += sm_len) {
        const uint32_t len = min(sm_len, count - base);
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
            const double r = v[base + i];
            sm[i] = r * r;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t i = 0;
            for (; i + 8 <= len; i += 8) {
                const double s0 = sm[i], s1 = sm[i + 1], s2 = sm[i + 2], s3 = sm[i + 3];
                const double s4 = sm[i + 4], s5 = sm[i + 5], s6 = sm[i + 6], s7 = sm[i + 7];
                acc = acc + s0; acc = acc + s1; acc = acc + s2; acc = acc + s3;
                acc = acc + s4; acc = acc + s5; acc = acc + s6; acc = acc + s7;
            }
            for (; i < len; ++i) acc = acc + sm[i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = acc;
}

// Chunked sum of squares for multi-CTA grids: chunk c = rows [c * chunk, ...) folded sequentially from +0.0
// into out[c]; the caller synchronises and folds the chunk sums sequentially (fold_sum).
__device__ void chunk_sum_squares(const double* v, uint32_t count, uint32_t kSumChunk, double* out, uint32_t tid, uint32_t nth, double* warp_stage) {
    // One warp per chunk: the lanes stage the chunk in shared memory with coalesced loads (kWarpStageDoubles = 512
    // values per round), lane 0 runs the dependent chain of adds from there.
    const uint32_t chunks = (count + kSumChunk - 1) / kSumChunk;
    const uint32_t lane = threadIdx.x & 31u, n_warps = nth >> 5;
    for (uint32_t c = tid >> 5; c < chunks; c += n_warps) {
        const uint32_t b = c * kSumChunk, e = min(count, b + kSumChunk);
        double acc = 0.0;
        for (uint32_t h = b; h < e; h += 512) {
            const uint32_t len = min(512u, e - h);
            for (uint32_t t = lane; t < len; t += 32) {
                const double r = v[h + t];
                warp_stage[t] = r * r;
            }
            __syncwarp();
            if (lane == 0) {
                uint32_t i = 0;
                for (; i + 8 <= len; i += 8) {
                    double q[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) q[u] = warp_stage[i + u];
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = acc + q[u];
                }
                for (; i < len; ++i) acc = acc + warp_stage[i];
            }
            __syncwarp();
        }
        if (lane == 0) out[c] = acc;
    }
}

// ---- supernodal sparse direct step solve (schedule: sparse_direct.cpp) --------------------------------------
// A = JtJ + lambda*I in elimination numbering into the panels (entries of the panels A does not have start from
// +0.0), b = -Jt r into y[].  Two phases separated by a barrier: zero fill, then the products.
__device__ void direct_zero(const LargeArgs& a, uint32_t tid, uint32_t nth) {
    double* const lv = a.vg + a.L0;  // (fields read once; 16-byte stores)
    const uint32_t n = a.nnz_l;
    const uint32_t head = (reinterpret_cast<uintptr_t>(lv) & 15u) ? 1u : 0u;
    if (n == 0) return;
    if (tid == 0 && head) lv[0] = 0.0;
    double2* const v2 = reinterpret_cast<double2*>(lv + head);
    const uint32_t n2 = (n - head) / 2;
    for (uint32_t e = tid; e < n2; e += nth) v2[e] = make_double2(0.0, 0.0);
    if (tid == 0 && ((n - head) & 1u)) lv[n - 1] = 0.0;
}
__device__ void direct_assemble(const LargeArgs& a, double lambda, uint32_t tid, uint32_t nth) {
    // (argument fields are read once: the block lives in local memory and stores force reloads, see sn_factor)
    const double* const jv = a.vg + a.J0;
    const double* const r = a.vg + a.R0;
    double* const lv = a.vg + a.L0;
    double* const yv = a.vg + a.Y0;
    const uint32_t *const aprod_ptr = a.aprod_ptr, *const aprod_a = a.aprod_a, *const aprod_b = a.aprod_b, *const aent_slot = a.aent_slot;
    const uint32_t *const perm = a.perm, *const col_ptr = a.csc_col_ptr, *const row_idx = a.csc_row_idx, *const jmap = a.jmap,
                   *const diag_slot = a.diag_slot;
    const uint32_t n_aent = a.n_aent, n = a.n;
    for (uint32_t k = tid; k < n_aent; k += nth) {
        const uint32_t qb = __ldg(aprod_ptr + k), qe = __ldg(aprod_ptr + k + 1);
        double acc = 0.0;
        for (uint32_t q = qb; q < qe; ++q) acc = __fma_rn(jv[__ldg(aprod_a + q)], jv[__ldg(aprod_b + q)], acc);
        lv[__ldg(aent_slot + k)] = acc;
    }
    for (uint32_t j = tid; j < n; j += nth) {
        const uint32_t c = __ldg(perm + j);
        double dg = 0.0, b = 0.0;
        for (uint32_t e = __ldg(col_ptr + c), ee = __ldg(col_ptr + c + 1); e < ee; ++e) {
            const double v = jv[jmap ? __ldg(jmap + e) : e];
            dg = __fma_rn(v, v, dg);
            b = __fma_rn(v, -r[__ldg(row_idx + e)], b);
        }
        lv[__ldg(diag_slot + j)] = __dadd_rn(dg, lambda);
        yv[j] = b;
    }
}

// Teams.  A supernode is factorised by a TEAM of 1 thread (panels of a few doubles, in global memory), 32 lanes
// (a warp: stages with many panels) or 512 threads (a whole CTA: the upper stages of the tree, where a handful of
// panels each receive dozens of updates).  Warps and CTAs stage the panel, the update blocks of the descendant
// panels (each followed by y of that panel's columns), the update records, the relative row positions and their
// inverse maps in shared memory (sizes in doubles / 32-bit words / 16-bit words):
template <int TEAM>
struct TeamCaps;
template <>
struct TeamCaps<1> {
    static constexpr uint32_t panel = 0, block = 0, rel = 0, inv = 0, recs = 0;
};
template <>
struct TeamCaps<32> {  // (no inverse maps: warp teams apply updates pair by pair; their space went to the block area)
    static constexpr uint32_t panel = 576, block = 576, rel = 256, inv = 0, recs = 32;
};
template <>
struct TeamCaps<128> {  // four warps: the stages of four warp teams
    static constexpr uint32_t panel = 1024, block = 3328, rel = 1024, inv = 2048, recs = 32;
};
template <>
struct TeamCaps<512> {
    static constexpr uint32_t panel = 15360, block = 4096, rel = 2048, inv = 8192, recs = 32;
};
constexpr uint32_t kYCap = 16;  // y of the panel's columns (a panel has at most 16)
template <int TEAM>
constexpr uint32_t team_stage_doubles() {  // recs = update records fetched per batch, 8 words each
    return TeamCaps<TEAM>::panel + TeamCaps<TEAM>::block + kYCap + TeamCaps<TEAM>::recs * 8 / 2 + TeamCaps<TEAM>::rel / 2 +
           TeamCaps<TEAM>::inv / 4;
}
// A warp team's stage ends with a 16-double CARRY: the records of the panel the warp factorises next (tag = its position in
// stage order), whose values and y are by then on their way into the block area (sn_factor<32>).
constexpr uint32_t kCarryDoubles = 16;
constexpr uint32_t kCarryOffset = team_stage_doubles<32>();  // in doubles from the start of the warp's stage
constexpr uint32_t kWarpStageDoubles = kCarryOffset + kCarryDoubles;  // 1,440 doubles = 11,520 bytes per warp
template <int TEAM>
__device__ __forceinline__ void team_sync() {
    if (TEAM == 32) __syncwarp();
    else if (TEAM == 128) asm volatile("bar.sync %0, 128;" ::"r"(1u + (threadIdx.x >> 7)) : "memory");  // named barrier of this quad
    else if (TEAM > 32) __syncthreads();
}


__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t* smem_dst, const uint32_t* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One update of panel P (w columns) by a block B of a descendant panel (T rows x wK columns; its first nc rows are
// columns of P's supernode, rel[] = where each row lands in P): P[i][j] = fma(-B[i][k], B[j][k], P[i][j]) for the row
// pairs i >= j with j among the first nc, k ascending; and the forward substitution of those nc rows against yK = y of
// the descendant's columns.
constexpr int kPairTeamMax = EZPZ_PAIR_TEAM_MAX;  // teams up to this size apply updates pair by pair (see sn_factor)
template <int TEAM>
__device__ __forceinline__ void sn_apply_update(double* P, uint32_t w, double* ys, const double* yK, const double* B, const uint32_t* rel,
                                                uint32_t T, uint32_t wK, uint32_t nc, uint32_t lane) {
    for (uint32_t p = lane; p < nc * T; p += TEAM) {
        const uint32_t tj = p / T, ti = p - tj * T;
        if (ti < tj) continue;
        double* dst = P + rel[ti] * w + rel[tj];
        double acc = *dst;
        if (wK == 16 && !(reinterpret_cast<uintptr_t>(B) & 15u)) {  // 16-byte loads, eight terms per round (see sn_factor)
            const double2* bi2 = reinterpret_cast<const double2*>(B + ti * 16);
            const double2* bj2 = reinterpret_cast<const double2*>(B + tj * 16);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                double2 u[4], v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    u[q] = bi2[4 * half + q];
                    v[q] = bj2[4 * half + q];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc = __fma_rn(-u[q].x, v[q].x, acc);
                    acc = __fma_rn(-u[q].y, v[q].y, acc);
                }
            }
        } else {
            for (uint32_t k = 0; k < wK; ++k) acc = __fma_rn(-B[ti * wK + k], B[tj * wK + k], acc);
        }
        *dst = acc;
    }
    for (uint32_t tj = lane; tj < nc; tj += TEAM) {
        const uint32_t c = rel[tj];
        double acc = ys[c];
        for (uint32_t k = 0; k < wK; ++k) acc = __fma_rn(-B[tj * wK + k], yK[k], acc);
        ys[c] = acc;
    }
}

// Factorisation of one supernode J fused with the forward substitution of its columns, by a team of TEAM threads
// (`lane` = index in the team, `stage` = the team's shared memory).  Left-looking, k ascending everywhere — the
// arithmetic of the column algorithm:
//   1. external updates, descendants K ascending: P[i][j] = fma(-L[i][k], L[j][k], P[i][j]) over K's columns k,
//      for every pair of K's rows (i >= j) with j a column of J;  y[j] = fma(-L[j][k], y[k], y[j]);
//   2. the panel's own columns c ascending: the same fma chains over the columns k < c of the panel, pivot
//      (fails unless > 0 and finite), 1/pivot, scaling, and y[c].
// next_pos / next2_pos: the panels this team factorises after this one (UINT32_MAX = none): their records, panel and first
// update records are pulled into L2 while this panel's own loads are in flight.
template <int TEAM>
__device__ __noinline__ void sn_factor(const LargeArgs& a, uint32_t pos, uint32_t lane, double* stage, uint32_t next_pos = UINT32_MAX,
                                       uint32_t next2_pos = UINT32_MAX) {
    using Caps = TeamCaps<TEAM>;
    // The argument block lives in local memory inside the LM kernel and every store through a generic pointer makes the
    // compiler reload its fields: everything the loops below need is read ONCE here.
    double* const lv = a.vg + a.L0;
    double* const y = a.vg + a.Y0;
    const uint32_t* const upd_rec = a.upd_rec;
    const uint32_t* const upd_rel = a.upd_rel;
    const uint32_t* const stage_rec = a.stage_rec;
    uint32_t* const fail_flag = &a.ctrl->fail;
    // Warp teams: has the previous call of this warp already started the copy of this panel (see the end of part 1)?
    uint32_t* const carry = TEAM == 32 ? reinterpret_cast<uint32_t*>(stage + kCarryOffset) : nullptr;
    const bool carried = TEAM == 32 && carry[0] == pos;
    const uint4 hdr = carried ? *reinterpret_cast<const uint4*>(carry + 4)
                              : __ldg(reinterpret_cast<const uint4*>(stage_rec) + 2 * (size_t)pos);  // j0, w, h, rows
    const uint4 hdr2 = carried ? *reinterpret_cast<const uint4*>(carry + 8)
                               : __ldg(reinterpret_cast<const uint4*>(stage_rec) + 2 * (size_t)pos + 1);  // panel, first update, updates
    const uint32_t j0 = hdr.x, w = hdr.y, h = hdr.z;
    double* const rinv_out = a.vg + a.RV0 + j0;
    double* G = lv + hdr2.x;
    const bool staged = TEAM > 1 && h * w <= Caps::panel;
    double* P = staged ? stage : G;
    // The update blocks get whatever the panel leaves of the panel + block areas (a panel of 100 doubles in a warp's stage
    // leaves 920 of the 1,024, twice the fixed split: more updates per memory round trip).
    // (CTA teams keep the fixed split for their large panels — 2D-lattice sketches measured slower with larger chunks — but
    // the small panels of the upper tree, dozens of updates each, leave them nearly the whole stage for update blocks.)
    const uint32_t pan = TEAM >= 512 ? (h * w <= 1024u ? 1024u : Caps::panel) : (staged ? ((h * w + 1u) & ~1u) : 0u);
    const uint32_t block_cap = TEAM > 1 ? Caps::panel + Caps::block - pan : 0u;
    double* kb = TEAM > 1 ? stage + pan : nullptr;
    double* ys = TEAM > 1 ? stage + Caps::panel + Caps::block : y + j0;  // w <= 16 values
    uint32_t* srec_base = TEAM > 1 ? reinterpret_cast<uint32_t*>(stage + Caps::panel + Caps::block + kYCap) : nullptr;
    uint32_t* srel = TEAM > 1 ? srec_base + Caps::recs * 8 : nullptr;
    uint16_t* inv = TEAM > 1 ? reinterpret_cast<uint16_t*>(srel + Caps::rel) : nullptr;
    if (TEAM > 1) {
        if (carried) {  // values and y arrived in the block area: panel first, then y of its columns
            cp_async_wait_all();
            team_sync<TEAM>();
            const double* cd = stage + (Caps::panel + Caps::block) - (h * w + w);  // where the previous call put them
            for (uint32_t t = lane; t < h * w; t += TEAM) P[t] = cd[t];
            for (uint32_t c = lane; c < w; c += TEAM) ys[c] = cd[h * w + c];
        } else {
            if (staged)
                for (uint32_t t = lane; t < h * w; t += TEAM) P[t] = G[t];
            for (uint32_t c = lane; c < w; c += TEAM) ys[c] = y[j0 + c];
        }
        team_sync<TEAM>();
    }
    // ---- 1. external updates
    const uint32_t ub = hdr2.y, ue = hdr2.y + hdr2.z;
    if (TEAM > 1) {
        uint32_t win = ub, win_n = 0;  // window of update records held in shared memory: [win, win + win_n)
        for (uint32_t u0 = ub; u0 < ue;) {
            // The records of the next (up to 32) updates go to shared memory; then a chunk of consecutive updates
            // whose blocks and relative positions fit the stage is fetched with asynchronous copies that are all in
            // flight together (one memory round trip per chunk instead of several per update).
            if (u0 >= win + win_n) {
                team_sync<TEAM>();  // everybody is done with the previous window
                win = u0;
                win_n = min(Caps::recs, ue - u0);
                for (uint32_t q = lane; q < win_n * 8; q += TEAM) srec_base[q] = __ldg(upd_rec + 8 * (size_t)win + q);
                team_sync<TEAM>();
            }
            uint32_t* srec = srec_base + 8 * (u0 - win);
            const uint32_t avail = win + win_n - u0;
            // Staged layout of update i of a chunk: srec[8 i + 5] = staged columns ws (all of K's columns, or a slice of
            // them), block = T rows x ws columns, row-major, followed by ws values of y.
            // Every panel entry is owned by one thread, which applies the chunk's updates to it in order: no barrier
            // between updates, no index comparisons (r >= c implies the block positions satisfy ti >= tj).
            // Warp teams instead go PAIR by pair through every update's block (positions through the staged row map), one
            // __syncwarp between updates: a lane then only touches entries an update really hits — with entry ownership half
            // of a warp's apply time went into scanning entries the update does not touch (warp 0 of CTA 0, 1M-variable
            // sketch: 27 of 53 us per 10-update panel).  Same chains either way: per entry, updates in order, k ascending.
            auto apply = [&](uint32_t cnt) {
                if constexpr (TEAM <= kPairTeamMax) {
                    for (uint32_t i = 0, boff = 0, roff = 0; i < cnt; ++i) {
                        const uint32_t T = srec[8 * i + 1], nc = srec[8 * i + 2] >> 8, ws = srec[8 * i + 5];
                        sn_apply_update<TEAM>(P, w, ys, kb + boff + T * ws, kb + boff, srel + roff, T, ws, nc, lane);
                        team_sync<TEAM>();
                        boff += T * ws + ws;
                        roff += T;
                    }
                } else {
                for (uint32_t e = lane; e < h * w; e += TEAM) {
                    const uint32_t r = e / w, c = e - r * w;
                    if (c > r) continue;
                    double acc = P[e];
                    for (uint32_t i = 0, boff = 0; i < cnt; ++i) {
                        const uint32_t T = srec[8 * i + 1], nc = srec[8 * i + 2] >> 8, ws = srec[8 * i + 5];
                        const uint32_t ti = inv[i * h + r], tj = inv[i * h + c];
                        if (ti != 0xffffu && tj < nc) {
                            const double* bi = kb + boff + ti * ws;
                            const double* bj = kb + boff + tj * ws;
                            if (ws == 16 && !(boff & 1u)) {
                                // full 16-column descendants (the upper tree): eight terms' operands per round of 16-byte
                                // loads, so the shared-memory latency is paid twice per update instead of once per term
                                const double2* bi2 = reinterpret_cast<const double2*>(bi);
                                const double2* bj2 = reinterpret_cast<const double2*>(bj);
#pragma unroll
                                for (int half = 0; half < 2; ++half) {
                                    double2 u[4], v[4];
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        u[q] = bi2[4 * half + q];
                                        v[q] = bj2[4 * half + q];
                                    }
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        acc = __fma_rn(-u[q].x, v[q].x, acc);
                                        acc = __fma_rn(-u[q].y, v[q].y, acc);
                                    }
                                }
                            } else {
                                for (uint32_t k = 0; k < ws; ++k) acc = __fma_rn(-bi[k], bj[k], acc);
                            }
                        }
                        boff += T * ws + ws;
                    }
                    P[e] = acc;
                }
                for (uint32_t c = lane; c < w; c += TEAM) {  // forward substitution against y of the descendants' columns
                    double acc = ys[c];
                    for (uint32_t i = 0, boff = 0; i < cnt; ++i) {
                        const uint32_t T = srec[8 * i + 1], nc = srec[8 * i + 2] >> 8, ws = srec[8 * i + 5];
                        const uint32_t tj = inv[i * h + c];
                        if (tj < nc) {
                            const double* bj = kb + boff + tj * ws;
                            const double* yK = kb + boff + T * ws;
                            for (uint32_t k = 0; k < ws; ++k) acc = __fma_rn(-bj[k], yK[k], acc);
                        }
                        boff += T * ws + ws;
                    }
                    ys[c] = acc;
                }
                team_sync<TEAM>();
                }
            };
            // inverse maps: inv[i * h + (panel row)] = position of that row in update i's block, 0xffff when absent
            auto build_inverse = [&](uint32_t cnt) {
                if constexpr (TEAM > kPairTeamMax) {  // (pair-by-pair apply needs no inverse maps)
                    for (uint32_t i = 0, roff = 0; i < cnt; ++i) {
                        const uint32_t T = srec[8 * i + 1];
                        for (uint32_t t = lane; t < T; t += TEAM) inv[i * h + srel[roff + t]] = (uint16_t)t;
                        roff += T;
                    }
                    team_sync<TEAM>();
                }
            };
            uint32_t cnt = 0, tot_b = 0, tot_r = 0;
            while (cnt < avail) {
                uint32_t* r = srec + 8 * cnt;
                const uint32_t T = r[1], wK = r[2] & 0xffu, len = T * wK;
                if (tot_b + len + wK > block_cap || tot_r + T > Caps::rel || (TEAM > kPairTeamMax && (cnt + 1) * h > Caps::inv) || T >= 0xffffu) break;
                for (uint32_t q = lane; q < len; q += TEAM) cp_async8(kb + tot_b + q, lv + r[0] + q);
                for (uint32_t q = lane; q < wK; q += TEAM) cp_async8(kb + tot_b + len + q, y + r[4] + q);  // y of K's columns
                for (uint32_t q = lane; q < T; q += TEAM) cp_async4(srel + tot_r + q, upd_rel + r[3] + q);
                if (lane == 0) r[5] = wK;
                tot_b += len + wK;
                tot_r += T;
                ++cnt;
            }
            if (cnt > 0) {
                if constexpr (TEAM > kPairTeamMax)
                    for (uint32_t q = lane; q < cnt * h; q += TEAM) inv[q] = 0xffffu;
                cp_async_wait_all();
                team_sync<TEAM>();
                build_inverse(cnt);
                apply(cnt);
                u0 += cnt;
                continue;
            }
            // One update whose block is larger than the stage.
            const uint32_t T = srec[1], wK = srec[2] & 0xffu, nc = srec[2] >> 8;
            if (T + 1 <= block_cap && T <= Caps::rel && (TEAM <= kPairTeamMax || h <= Caps::inv) && T < 0xffffu) {
                // Column slices of the descendant's panel, as many columns at a time as fit: every entry's fma chain
                // simply continues from slice to slice (k stays ascending).
                const uint32_t ws_max = block_cap / (T + 1);
                if constexpr (TEAM > kPairTeamMax)
                    for (uint32_t q = lane; q < h; q += TEAM) inv[q] = 0xffffu;
                for (uint32_t q = lane; q < T; q += TEAM) cp_async4(srel + q, upd_rel + srec[3] + q);
                for (uint32_t ks = 0; ks < wK; ks += ws_max) {
                    const uint32_t ws = min(ws_max, wK - ks);
                    for (uint32_t q = lane; q < T * ws; q += TEAM) {
                        const uint32_t t = q / ws, k = q - t * ws;
                        cp_async8(kb + q, lv + srec[0] + t * wK + ks + k);
                    }
                    for (uint32_t q = lane; q < ws; q += TEAM) cp_async8(kb + T * ws + q, y + srec[4] + ks + q);
                    if (lane == 0) srec[5] = ws;
                    cp_async_wait_all();
                    team_sync<TEAM>();
                    if (ks == 0) build_inverse(1);
                    apply(1);
                }
            } else {
                // pair by pair, straight from global memory
                sn_apply_update<TEAM>(P, w, ys, y + srec[4], lv + srec[0], upd_rel + srec[3], T, wK, nc, lane);
                team_sync<TEAM>();
            }
            u0 += 1;
        }
    } else {
        for (uint32_t u = ub; u < ue; ++u) {
            const uint32_t* r = upd_rec + 8 * (size_t)u;
            const uint32_t T = __ldg(r + 1), z = __ldg(r + 2);
            sn_apply_update<1>(P, w, ys, y + __ldg(r + 4), lv + __ldg(r), upd_rel + __ldg(r + 3), T, z & 0xffu, z >> 8, 0);
        }
    }
    // ---- the next panel of this warp: records now, values and y into the (now idle) block area while part 2 runs
    if (TEAM == 32) {
        team_sync<TEAM>();  // every lane is done with the block area and with the carry of this panel
        bool carry_next = false;
        if (next_pos != UINT32_MAX) {
            const uint4 nh = __ldg(reinterpret_cast<const uint4*>(stage_rec) + 2 * (size_t)next_pos);
            const uint4 nh2 = __ldg(reinterpret_cast<const uint4*>(stage_rec) + 2 * (size_t)next_pos + 1);
            const uint32_t nhw = nh.y * nh.z;
            // destination: the END of the panel + block areas — clear of this panel (pan <= 576) and of the place the next
            // call stages its panel (nhw doubles from the start)
            if (2 * nhw + nh.y <= Caps::panel + Caps::block && pan + nhw + nh.y <= Caps::panel + Caps::block) {
                double* cd = stage + (Caps::panel + Caps::block) - (nhw + nh.y);
                for (uint32_t q = lane; q < nhw; q += TEAM) cp_async8(cd + q, lv + nh2.x + q);
                for (uint32_t q = lane; q < nh.y; q += TEAM) cp_async8(cd + nhw + q, y + nh.x + q);
                if (lane == 0) {
                    carry[0] = next_pos;
                    *reinterpret_cast<uint4*>(carry + 4) = nh;
                    *reinterpret_cast<uint4*>(carry + 8) = nh2;
                }
                carry_next = true;
            } else {
                const uint32_t lines = (nhw * 8u + 127u) / 128u;
                for (uint32_t q = lane; q < lines; q += 32) prefetch_l2(lv + nh2.x + 16u * q);
            }
            if (lane == 30 && nh2.z) prefetch_l2(upd_rec + 8 * (size_t)nh2.y);
            if (lane == 29 && next2_pos != UINT32_MAX) prefetch_l2(stage_rec + 8 * (size_t)next2_pos);
        }
        if (!carry_next && lane == 0) carry[0] = UINT32_MAX;
    }
    // ---- 2. the panel's own columns
    // (A variant that keeps row r of the panel in lane r's registers and moves L[c][k] by shuffle was measured slower:
    // 9.0 ms against 8.7 ms of factor stages on the 1M-variable sketch — a warp waits on the panel's memory round trips,
    // not on the column arithmetic; profiles/r01k_lm_kernel.md.)
    // The forward substitution of column c, y[c] = (y[c] - sum_{k<c} L[c][k] y[k]) / L[c][c], rides along as one more "row"
    // (r == h) of the column: the same fma chain with other operands, on a lane that would otherwise idle, instead of a second
    // chain that the diagonal's lane ran after its pivot while the team waited at the barrier.
    for (uint32_t c = 0; c < w; ++c) {
        for (uint32_t r = c + lane; r <= h; r += TEAM) {
            const bool yrow = r == h;
            const double* pa = yrow ? P + c * w : P + r * w;
            const double* pb = yrow ? ys : P + c * w;
            double acc = yrow ? ys[c] : P[r * w + c], piv = P[c * w + c];
            for (uint32_t k = 0; k < c; ++k) {
                const double lck = P[c * w + k];
                piv = __fma_rn(-lck, lck, piv);
                acc = __fma_rn(-pa[k], pb[k], acc);
            }
            const double rinv = __ddiv_rn(1.0, __dsqrt_rn(piv));
            if (r == c) {
                if (!(piv > 0.0) || !ezm::ez_isfinite(piv)) *fail_flag = 1;
                rinv_out[c] = rinv;
            } else if (yrow) {
                ys[c] = __dmul_rn(acc, rinv);
            } else {
                P[r * w + c] = __dmul_rn(acc, rinv);
            }
        }
        team_sync<TEAM>();
    }
    if (TEAM > 1) {
        if (staged)
            for (uint32_t t = lane; t < h * w; t += TEAM) G[t] = P[t];
        for (uint32_t c = lane; c < w; c += TEAM) y[j0 + c] = ys[c];
        team_sync<TEAM>();
    }
}

// ---- Tall panels of the upper tree, split by ROWS over several CTAs --------------------------------------------------
// The upper stages of a 2D-lattice-like sketch hold a handful of panels of 16 columns x several hundred rows, each receiving
// 50-200 updates: one CTA per panel left the other ~140 SMs idle (grid_truss(100): 97 % of the solve).  Every entry of a panel
// is an independent fma chain (updates in order, k ascending), and a row below the diagonal block needs nothing but its own
// entries and the finished diagonal block.  So a panel is cut into SLICES: slice 0 owns the w x w diagonal block (its updates,
// the pivots, 1/pivot, the forward substitution of y) and publishes it with a release flag; slices 1.. own a range of the
// rows below, apply the updates to them at once, wait for the flag, and divide their rows through.  Same chains per entry as
// sn_factor, hence the same bits.  All CTAs of the launch are co-resident (cooperative / cluster launch), so waiting is safe.
constexpr uint32_t kSlicePanelCap = 10240;  // doubles of a slice's rows held in shared memory
constexpr uint32_t kSliceChunk = 32;        // update records per round
constexpr uint32_t kSliceMinRows = 48;      // panels with fewer rows below the diagonal block stay on one CTA
// doubles between staged rows of an update block of width wK (even: rows of 16 stay 16-byte aligned for the 128-bit loads)
__device__ __forceinline__ uint32_t slice_row_stride(uint32_t wK) { return (wK + 3u) & ~1u; }

__device__ __noinline__ void sn_factor_slice(const LargeArgs& a, uint32_t pos, uint32_t slice, uint32_t n_slices, uint32_t epoch,
                                             double* stage, uint32_t stage_doubles) {
    const uint32_t lane = threadIdx.x, TEAM = blockDim.x, warp = lane >> 5, wl = lane & 31u, n_warps = TEAM >> 5;
    double* const lv = a.vg + a.L0;
    double* const y = a.vg + a.Y0;
    const uint32_t* const upd_rec = a.upd_rec;
    const uint32_t* const upd_rel = a.upd_rel;
    const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(a.stage_rec) + 2 * (size_t)pos);
    const uint4 hdr2 = __ldg(reinterpret_cast<const uint4*>(a.stage_rec) + 2 * (size_t)pos + 1);
    const uint32_t j0 = hdr.x, w = hdr.y, h = hdr.z, hb = h - w;
    double* const G = lv + hdr2.x;
    const bool diag = slice == 0;
    // rows of this slice inside the panel: the diagonal block, or an even share of the rows below it
    const uint32_t r0 = diag ? 0u : w + (uint32_t)((uint64_t)hb * (slice - 1) / (n_slices - 1));
    const uint32_t r1 = diag ? w : w + (uint32_t)((uint64_t)hb * slice / (n_slices - 1));
    const uint32_t rows = r1 - r0;
    // shared memory: P[rows][w] | ys[16] | L11[w*w] + rinv[w] (slices 1..) | records | per update: offsets | blocks
    double* P = stage;
    double* ys = P + ((rows * w + 1u) & ~1u);
    double* l11 = ys + kYCap;
    double* rinv_s = l11 + 256;
    uint32_t* srec = reinterpret_cast<uint32_t*>(rinv_s + 16);   // kSliceChunk * 8 words
    uint32_t* meta = srec + kSliceChunk * 8;                      // per update: t_lo, t_hi, block offset (doubles), rel offset (words)
    uint32_t* n_fit = meta + kSliceChunk * 4;                     // [0] updates of this round that fit
    uint32_t* srel = n_fit + 4;                                   // relative positions of the staged rows (keeps kb 16-byte aligned)
    const uint32_t rel_cap = 4096;
    double* kb = reinterpret_cast<double*>(srel + rel_cap);
    const uint32_t kb_cap = stage_doubles - (uint32_t)(kb - stage);
    for (uint32_t t = lane; t < rows * w; t += TEAM) P[t] = G[(size_t)r0 * w + t];
    if (diag)
        for (uint32_t c = lane; c < w; c += TEAM) ys[c] = y[j0 + c];
    __syncthreads();
    // ---- 1. external updates
    const uint32_t ub = hdr2.y, ue = hdr2.y + hdr2.z;
    for (uint32_t u0 = ub; u0 < ue;) {
        const uint32_t win_n = min(kSliceChunk, ue - u0);
        for (uint32_t q = lane; q < win_n * 8; q += TEAM) srec[q] = __ldg(upd_rec + 8 * (size_t)u0 + q);
        __syncthreads();
        // the rows of every update's block that land in this slice: rel is ascending, so they are one range [t_lo, t_hi).
        // One WARP per update, 32-ary search (two or three round trips to memory instead of twenty dependent ones).
        for (uint32_t i = warp; i < win_n; i += n_warps) {
            const uint32_t* r = srec + 8 * i;
            const uint32_t T = r[1], nc = r[2] >> 8;
            uint32_t lo = nc, hi = nc;
            if (!diag) {
                const uint32_t* rel = upd_rel + r[3];
                auto lower_bound = [&](uint32_t from, uint32_t v) {  // first t in [from, T) with rel[t] >= v
                    uint32_t x0 = from, x1 = T;
                    while (x1 > x0) {
                        const uint32_t span = x1 - x0;
                        auto probe = [&](uint32_t l) { return x0 + (uint32_t)(((uint64_t)span * (l + 1)) / 33u); };
                        const uint32_t pl = min(probe(wl), x1 - 1);
                        const uint32_t k = __popc(__ballot_sync(0xffffffffu, __ldg(rel + pl) < v));
                        const uint32_t n0 = k ? min(probe(k - 1), x1 - 1) + 1 : x0;
                        const uint32_t n1 = k < 32 ? min(probe(k), x1 - 1) : x1;
                        x0 = n0;
                        x1 = n1;
                    }
                    return x0;
                };
                lo = lower_bound(nc, r0);
                hi = lower_bound(lo, r1);
            }
            if (wl == 0) {
                meta[4 * i] = lo;
                meta[4 * i + 1] = hi;
            }
        }
        __syncthreads();
        // inverse maps of the round, one per update: (rows + w) 16-bit positions — where each row of this slice sits in the
        // update's staged rows, and where each column of the supernode sits among its first nc rows; 0xffff = not touched
        const uint32_t inv_len = rows + w;  // 16-bit words per update
        if (lane == 0) {  // how many updates of the window fit: inverse map + [first nc rows | own rows] x wK + y of K's columns
            uint32_t tot_b = 0, tot_r = 0, cnt = 0;
            for (; cnt < win_n; ++cnt) {
                const uint32_t* r = srec + 8 * cnt;
                const uint32_t wK = r[2] & 0xffu, nc = r[2] >> 8, own = meta[4 * cnt + 1] - meta[4 * cnt];
                const uint32_t need_b = (nc + own) * slice_row_stride(wK) + wK, need_r = nc + own;
                const uint32_t inv_doubles = ((cnt + 1) * inv_len + 3u) / 4u;
                if (tot_b + ((need_b + 1u) & ~1u) + ((inv_doubles + 1u) & ~1u) > kb_cap || tot_r + need_r > rel_cap || own >= 0xffffu) break;
                meta[4 * cnt + 2] = tot_b;
                meta[4 * cnt + 3] = tot_r;
                tot_b += (need_b + 1u) & ~1u;
                tot_r += need_r;
            }
            n_fit[0] = cnt;
        }
        __syncthreads();
        const uint32_t cnt = n_fit[0];
        if (cnt == 0) {
            // one update that does not fit the stage: pair by pair straight from global memory (rare: a slice holds few rows)
            const uint32_t* r = srec;
            const uint32_t wK = r[2] & 0xffu, nc = r[2] >> 8, lo = meta[0], hi = meta[1];
            const double* B = lv + r[0];
            const uint32_t* rel = upd_rel + r[3];
            if (diag) {
                for (uint32_t p = lane; p < nc * nc; p += TEAM) {
                    const uint32_t tj = p / nc, ti = p - tj * nc;
                    if (ti < tj) continue;
                    double* dst = P + __ldg(rel + ti) * w + __ldg(rel + tj);
                    double acc = *dst;
                    for (uint32_t k = 0; k < wK; ++k) acc = __fma_rn(-B[ti * wK + k], B[tj * wK + k], acc);
                    *dst = acc;
                }
                for (uint32_t tj = lane; tj < nc; tj += TEAM) {
                    const uint32_t c = __ldg(rel + tj);
                    double acc = ys[c];
                    for (uint32_t k = 0; k < wK; ++k) acc = __fma_rn(-B[tj * wK + k], y[r[4] + k], acc);
                    ys[c] = acc;
                }
            } else {
                for (uint32_t p = lane; p < (hi - lo) * nc; p += TEAM) {
                    const uint32_t ti = lo + p / nc, tj = p % nc;
                    double* dst = P + (__ldg(rel + ti) - r0) * w + __ldg(rel + tj);
                    double acc = *dst;
                    for (uint32_t k = 0; k < wK; ++k) acc = __fma_rn(-B[(size_t)ti * wK + k], B[tj * wK + k], acc);
                    *dst = acc;
                }
            }
            __syncthreads();
            u0 += 1;
            continue;
        }
        // the inverse maps sit at the END of the block area (the blocks grow from its start)
        uint16_t* inv = reinterpret_cast<uint16_t*>(kb + kb_cap) - (size_t)cnt * inv_len;
        for (uint32_t q = lane; q < cnt * inv_len; q += TEAM) inv[q] = 0xffffu;
        // all staged rows of the round in flight together (a warp per update)
        for (uint32_t i = warp; i < cnt; i += n_warps) {
            const uint32_t* r = srec + 8 * i;
            const uint32_t wK = r[2] & 0xffu, nc = r[2] >> 8, lo = meta[4 * i], hi = meta[4 * i + 1], own = hi - lo;
            double* dst = kb + meta[4 * i + 2];
            uint32_t* rdst = srel + meta[4 * i + 3];
            const double* B = lv + r[0];
            // staged rows sit at a stride of wK + 2 doubles: at a stride of 16 the sixteen rows a warp reads for the sixteen
            // columns of the panel all start in the same bank (16-way conflicts on every operand of the apply loop)
            const uint32_t ws = slice_row_stride(wK);
            for (uint32_t q = wl, t = wl / wK, k = wl - t * wK; q < nc * wK; q += 32) {
                cp_async8(dst + t * ws + k, B + q);
                k += 32;
                while (k >= wK) {
                    k -= wK;
                    ++t;
                }
            }
            for (uint32_t q = wl, t = wl / wK, k = wl - t * wK; q < own * wK; q += 32) {
                cp_async8(dst + (nc + t) * ws + k, B + (size_t)lo * wK + q);
                k += 32;
                while (k >= wK) {
                    k -= wK;
                    ++t;
                }
            }
            for (uint32_t q = wl; q < wK; q += 32) cp_async8(dst + (nc + own) * ws + q, y + r[4] + q);
            for (uint32_t q = wl; q < nc; q += 32) cp_async4(rdst + q, upd_rel + r[3] + q);
            for (uint32_t q = wl; q < own; q += 32) cp_async4(rdst + nc + q, upd_rel + r[3] + lo + q);
        }
        cp_async_wait_all();
        __syncthreads();
        for (uint32_t i = warp; i < cnt; i += n_warps) {
            const uint32_t nc = srec[8 * i + 2] >> 8, own = meta[4 * i + 1] - meta[4 * i];
            const uint32_t* relc = srel + meta[4 * i + 3];
            uint16_t* iv = inv + (size_t)i * inv_len;  // [0, rows): this slice's rows; [rows, rows + w): the supernode's columns
            for (uint32_t t = wl; t < nc; t += 32) iv[rows + relc[t]] = (uint16_t)t;
            if (!diag)
                for (uint32_t t = wl; t < own; t += 32) iv[relc[nc + t] - r0] = (uint16_t)t;
        }
        __syncthreads();
        // every entry of the slice is owned by one thread, which applies the round's updates to it in order: no barrier
        // between updates
        for (uint32_t e = lane; e < rows * w; e += TEAM) {
            const uint32_t lr = e / w, c = e - lr * w;
            if (diag && c > lr) continue;
            double acc = P[e];
            for (uint32_t i = 0; i < cnt; ++i) {
                const uint16_t* iv = inv + (size_t)i * inv_len;
                const uint32_t tj = iv[rows + c], ti = diag ? iv[rows + lr] : iv[lr];
                if (ti == 0xffffu || tj == 0xffffu) continue;
                const uint32_t wK = srec[8 * i + 2] & 0xffu, nc = srec[8 * i + 2] >> 8, ws = slice_row_stride(wK);
                const double* Bc = kb + meta[4 * i + 2];
                const double* bi = diag ? Bc + ti * ws : Bc + (nc + ti) * ws;
                const double* bj = Bc + tj * ws;
                if (wK == 16) {
                    const double2* bi2 = reinterpret_cast<const double2*>(bi);
                    const double2* bj2 = reinterpret_cast<const double2*>(bj);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        double2 u[4], v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            u[q] = bi2[4 * half + q];
                            v[q] = bj2[4 * half + q];
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            acc = __fma_rn(-u[q].x, v[q].x, acc);
                            acc = __fma_rn(-u[q].y, v[q].y, acc);
                        }
                    }
                } else {
                    for (uint32_t k = 0; k < wK; ++k) acc = __fma_rn(-bi[k], bj[k], acc);
                }
            }
            P[e] = acc;
        }
        if (diag)
            for (uint32_t c = lane; c < w; c += TEAM) {  // forward substitution against y of the descendants' columns
                double acc = ys[c];
                for (uint32_t i = 0; i < cnt; ++i) {
                    const uint32_t tj = inv[(size_t)i * inv_len + rows + c];
                    if (tj == 0xffffu) continue;
                    const uint32_t wK = srec[8 * i + 2] & 0xffu, nc = srec[8 * i + 2] >> 8, ws = slice_row_stride(wK);
                    const double* Bc = kb + meta[4 * i + 2];
                    const double* yK = Bc + nc * ws;  // (a diagonal slice stages no rows of its own)
                    for (uint32_t k = 0; k < wK; ++k) acc = __fma_rn(-Bc[tj * ws + k], yK[k], acc);
                    ys[c] = acc;
                }
            }
        __syncthreads();
        u0 += cnt;
    }
    uint32_t* const flag = a.sn_flag + pos;
    if (diag) {
        // ---- 2. the diagonal block's own columns, the pivots and y (sn_factor, part 2, on w rows)
        double* const rinv_out = a.vg + a.RV0 + j0;
        for (uint32_t c = 0; c < w; ++c) {
            for (uint32_t r = c + lane; r <= w; r += TEAM) {
                const bool yrow = r == w;
                const double* pa = yrow ? P + c * w : P + r * w;
                const double* pb = yrow ? ys : P + c * w;
                double acc = yrow ? ys[c] : P[r * w + c], piv = P[c * w + c];
                for (uint32_t k = 0; k < c; ++k) {
                    const double lck = P[c * w + k];
                    piv = __fma_rn(-lck, lck, piv);
                    acc = __fma_rn(-pa[k], pb[k], acc);
                }
                const double rinv = __ddiv_rn(1.0, __dsqrt_rn(piv));
                if (r == c) {
                    if (!(piv > 0.0) || !ezm::ez_isfinite(piv)) a.ctrl->fail = 1;
                    rinv_out[c] = rinv;
                } else if (yrow) {
                    ys[c] = __dmul_rn(acc, rinv);
                } else {
                    P[r * w + c] = __dmul_rn(acc, rinv);
                }
            }
            __syncthreads();
        }
        for (uint32_t t = lane; t < w * w; t += TEAM) G[t] = P[t];
        for (uint32_t c = lane; c < w; c += TEAM) y[j0 + c] = ys[c];
        __threadfence();
        __syncthreads();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    } else {
        // ---- 2. wait for the diagonal block, then every row divides itself through (independent chains, c ascending)
        if (lane == 0) {
            uint32_t seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
                if (seen != epoch) __nanosleep(64);
            } while (seen != epoch);
        }
        __syncthreads();
        for (uint32_t t = lane; t < w * w; t += TEAM) l11[t] = __ldcg(G + t);
        for (uint32_t c = lane; c < w; c += TEAM) rinv_s[c] = __ldcg(a.vg + a.RV0 + j0 + c);
        __syncthreads();
        for (uint32_t r = lane; r < rows; r += TEAM) {
            double* pr = P + r * w;
            for (uint32_t c = 0; c < w; ++c) {
                double acc = pr[c];
                const double* pb = l11 + c * w;
                for (uint32_t k = 0; k < c; ++k) acc = __fma_rn(-pr[k], pb[k], acc);
                pr[c] = __dmul_rn(acc, rinv_s[c]);
            }
        }
        __syncthreads();
        for (uint32_t t = lane; t < rows * w; t += TEAM) G[(size_t)r0 * w + t] = P[t];
    }
    __syncthreads();
}

// Backward substitution of one supernode (stages descending): d[j] = (y[j] - sum_{i > j} L[i][j] d[i]) / L[j][j] with the
// rows i DESCENDING — first the rows below the supernode's diagonal block (their d is final), then the block's own rows.
// The columns of the supernode advance together: lane c of the team's first warp owns column c; the rows below the block
// are w independent chains; inside the block, step t (t = w-1 ... 0) finalises d[t] and every column c < t applies it —
// one shuffle per step instead of one sequential chain per supernode.  The final values also go to d in variable numbering.
// Warp teams pipeline their panels through the carry at the end of the warp's stage (see sn_factor): while panel k is
// substituted, the values and the row list of panel k + 1 travel into shared memory (cp.async, doubles 640.. of the stage)
// and the record of panel k + 2 into registers; half of a warp's backward time used to be the dependent round trips record ->
// values / row list -> d of the rows, of which only the last gather is left.
//   carry words: [0] panel whose values + row list are staged, [4..8] its record; [1] panel whose record is in [12..16]
constexpr uint32_t kBackCarryAt = 640, kBackCarryMaxH = 64, kBackCarryMaxHW = 384;
template <int TEAM>
__device__ __noinline__ void sn_backward(const LargeArgs& a, uint32_t pos, uint32_t lane, double* stage, uint32_t next_pos = UINT32_MAX,
                                         uint32_t next2_pos = UINT32_MAX) {
    using Caps = TeamCaps<TEAM>;
    const double* const lv = a.vg + a.L0;  // (argument fields are read once: see sn_factor)
    double* const y = a.vg + a.Y0;
    const uint32_t* const stage_rec = a.stage_rec;
    const uint32_t* const sn_rows = a.sn_rows;
    uint32_t* const carry = TEAM == 32 ? reinterpret_cast<uint32_t*>(stage + kCarryOffset) : nullptr;
    const bool carried = TEAM == 32 && carry[0] == pos;
    const uint4 hdr = carried ? *reinterpret_cast<const uint4*>(carry + 4) : __ldg(reinterpret_cast<const uint4*>(stage_rec) + 2 * (size_t)pos);
    const uint32_t j0 = hdr.x, w = hdr.y, h = hdr.z, rb = hdr.w;
    const double* G = lv + (carried ? carry[8] : __ldg(stage_rec + 8 * (size_t)pos + 4));
    const uint32_t* const rows = sn_rows + rb;
    const double* const rinv_in = a.vg + a.RV0 + j0;
    double* const d_out = a.vg + a.D0;
    const uint32_t* const perm = a.perm + j0;
    if (TEAM == 1) {  // a panel of a few doubles: one thread, straight from global memory
        for (uint32_t c = w; c-- > 0;) {
            double acc = y[j0 + c];
            for (uint32_t r = h; r-- > c + 1;) acc = __fma_rn(-G[r * w + c], y[__ldg(rows + r)], acc);
            const double v = __dmul_rn(acc, rinv_in[c]);
            y[j0 + c] = v;
            d_out[__ldg(perm + c)] = v;
        }
        return;
    }
    const bool staged = h * w <= Caps::panel && h <= Caps::block;
    const double* P = G;
    double* dv = nullptr;  // staged variant: y of the block's own rows, d of the rows below
    if (staged) {
        double* Ps = stage;
        dv = stage + Caps::panel;
        if (carried) {  // values and row list are in the stage already: one gather is all that is left
            cp_async_wait_all();
            team_sync<TEAM>();
            const double* cd = stage + kBackCarryAt;
            const uint32_t* crows = reinterpret_cast<const uint32_t*>(cd + ((h * w + 1u) & ~1u));
            for (uint32_t t = lane; t < h; t += TEAM) dv[t] = y[crows[t]];
            for (uint32_t t = lane; t < h * w; t += TEAM) Ps[t] = cd[t];
        } else {
            for (uint32_t t = lane; t < h * w; t += TEAM) Ps[t] = G[t];
            for (uint32_t t = lane; t < h; t += TEAM) dv[t] = y[__ldg(rows + t)];
        }
        P = Ps;
        team_sync<TEAM>();
    }
    uint4 n2h = make_uint4(0, 0, 0, 0);
    uint32_t n2off = 0;
    if (TEAM == 32) {
        // record of panel k + 1: carried by the previous call, else read now; record of panel k + 2: requested now, kept at the end
        bool carry_next = false;
        if (next_pos != UINT32_MAX && staged && h <= kBackCarryMaxH) {
            const bool have = carry[1] == next_pos;
            const uint4 nh = have ? *reinterpret_cast<const uint4*>(carry + 12) : __ldg(reinterpret_cast<const uint4*>(stage_rec) + 2 * (size_t)next_pos);
            const uint32_t noff = have ? carry[16] : __ldg(stage_rec + 8 * (size_t)next_pos + 4);
            const uint32_t nhw = nh.y * nh.z;
            if (nh.z <= kBackCarryMaxH && nhw <= kBackCarryMaxHW) {
                double* cd = stage + kBackCarryAt;
                uint32_t* crows = reinterpret_cast<uint32_t*>(cd + ((nhw + 1u) & ~1u));
                for (uint32_t q = lane; q < nhw; q += TEAM) cp_async8(cd + q, lv + noff + q);
                for (uint32_t q = lane; q < nh.z; q += TEAM) cp_async4(crows + q, sn_rows + nh.w + q);
                team_sync<TEAM>();  // every lane has read the carry of this panel
                if (lane == 0) {
                    carry[0] = next_pos;
                    *reinterpret_cast<uint4*>(carry + 4) = nh;
                    carry[8] = noff;
                }
                carry_next = true;
            }
        }
        if (!carry_next) {
            team_sync<TEAM>();
            if (lane == 0) carry[0] = UINT32_MAX;
        }
        if (next2_pos != UINT32_MAX) {
            n2h = __ldg(reinterpret_cast<const uint4*>(stage_rec) + 2 * (size_t)next2_pos);
            n2off = __ldg(stage_rec + 8 * (size_t)next2_pos + 4);
        }
    }
    if (lane < 32) {
        const uint32_t c = lane;
        double acc = 0.0, rinv = 0.0;
        if (c < w) {
            acc = staged ? dv[c] : y[j0 + c];
            rinv = rinv_in[c];
            for (uint32_t r = h; r-- > w;) acc = __fma_rn(-P[r * w + c], staged ? dv[r] : y[__ldg(rows + r)], acc);
        }
        double v = 0.0;
        for (uint32_t t = w; t-- > 0;) {
            const double vt = __shfl_sync(0xffffffffu, __dmul_rn(acc, rinv), t);
            if (c == t) v = vt;
            else if (c < t) acc = __fma_rn(-P[t * w + c], vt, acc);
        }
        if (c < w) {
            y[j0 + c] = v;
            d_out[__ldg(perm + c)] = v;
        }
    }
    if (TEAM == 32 && lane == 0) {
        carry[1] = next2_pos;  // UINT32_MAX when there is none
        *reinterpret_cast<uint4*>(carry + 12) = n2h;
        carry[16] = n2off;
    }
    team_sync<TEAM>();
}

// One stage of the supernode tree, run by the whole grid (or cluster, or single CTA).  Tiny panels: one per thread.
// Panels that fit a warp's stage: one per warp while the stage has more of them than the grid has CTAs, else one per
// CTA.  Larger panels: always one per CTA (cta_stage = the CTA's whole dynamic shared memory; warp_stage = this warp's
// slice of it).  (Four 8-lane teams per warp were measured slower on
// the bottom stages: diverged teams of one warp execute one after the other.)
__device__ void direct_factor_stage(const LargeArgs& a, uint32_t st, uint32_t tid, uint32_t nth, double* warp_stage, double* cta_stage,
                                    uint32_t epoch) {
    const uint32_t b0 = __ldg(a.stage_ptr + 3 * st), b1 = __ldg(a.stage_ptr + 3 * st + 1), b2 = __ldg(a.stage_ptr + 3 * st + 2),
                   b3 = __ldg(a.stage_ptr + 3 * st + 3);
    for (uint32_t k = b0 + tid; k < b1; k += nth) sn_factor<1>(a, k, 0, nullptr);
    if (b2 - b1 > a.vgrid && b2 - b1 <= 8 * a.vgrid) {
        // Between one and eight panels per CTA (the middle of the tree: every panel receives 10-20 updates): QUADS of four
        // warps — a quad stages ~15 updates per round trip where a warp's stage holds two or three, and there are still
        // enough quads for every panel of the stage.
        for (uint32_t k = b1 + (threadIdx.x >> 7) * a.vgrid + vblock_of(a); k < b2; k += 4 * a.vgrid)
            sn_factor<128>(a, k, threadIdx.x & 127u, cta_stage + (threadIdx.x >> 7) * 4 * kWarpStageDoubles);
        __syncthreads();
        for (uint32_t k = b2 + vblock_of(a); k < b3; k += a.vgrid) sn_factor<512>(a, k, threadIdx.x, cta_stage);
    } else if (b2 - b1 > a.vgrid) {
        if ((threadIdx.x & 31u) == 0) reinterpret_cast<uint32_t*>(warp_stage + kCarryOffset)[0] = UINT32_MAX;  // no carry
        __syncwarp();
        // Panels are dealt warp-major ACROSS the CTAs (panel k of the stage's cost-sorted list to CTA k mod G): the most
        // expensive panels land on different SMs, and a stage with fewer panels than warps still uses every SM.
        for (uint32_t k = b1 + (threadIdx.x >> 5) * a.vgrid + vblock_of(a), nw = nth >> 5; k < b2; k += nw)
            sn_factor<32>(a, k, threadIdx.x & 31u, warp_stage, k + nw < b2 ? k + nw : UINT32_MAX, k + 2 * nw < b2 ? k + 2 * nw : UINT32_MAX);
        __syncthreads();  // the CTA panels below reuse the warps' shared memory
        for (uint32_t k = b2 + vblock_of(a); k < b3; k += a.vgrid) sn_factor<512>(a, k, threadIdx.x, cta_stage);
    } else {
        // At most one panel per CTA.  With CTAs to spare, tall panels are cut into row slices (sn_factor_slice): CTA v takes
        // slice v / nP of panel v % nP.
        const uint32_t nP = b3 - b1, vb = vblock_of(a);
        const uint32_t s_max = nP ? a.vgrid / nP : 0u;
        if (a.sn_flag && s_max >= 2 && !a.batch) {
            const uint32_t k = b1 + vb % nP, slice = vb / nP;
            if (slice < s_max) {
                const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(a.stage_rec) + 2 * (size_t)k);
                const uint32_t w = hdr.y, hb = hdr.z - hdr.y;
                // slices of at least 8 rows; slice 0 is the diagonal block
                uint32_t S = 1u + min(s_max - 1u, (hb + 7u) / 8u);
                const uint4 hdr2 = __ldg(reinterpret_cast<const uint4*>(a.stage_rec) + 2 * (size_t)k + 1);
                // short panels stay on one CTA unless they receive many updates (the sliced code stages 32 updates per round)
                const bool worth = hb >= kSliceMinRows || (hb >= 2u && hdr2.z >= 32u);
                if (!worth || ((hb + S - 2u) / (S - 1u)) * w > kSlicePanelCap) S = 1u;
                if (S == 1u) {
                    if (slice == 0) sn_factor<512>(a, k, threadIdx.x, cta_stage);
                } else if (slice < S) {
                    sn_factor_slice(a, k, slice, S, epoch, cta_stage, (512 / 32) * kWarpStageDoubles);
                }
            }
        } else {
            for (uint32_t k = b1 + vb; k < b3; k += a.vgrid) sn_factor<512>(a, k, threadIdx.x, cta_stage);
        }
    }
}
__device__ void direct_backward_stage(const LargeArgs& a, uint32_t st, uint32_t tid, uint32_t nth, double* warp_stage, double* cta_stage) {
    const uint32_t b0 = __ldg(a.stage_ptr + 3 * st), b1 = __ldg(a.stage_ptr + 3 * st + 1), b2 = __ldg(a.stage_ptr + 3 * st + 2),
                   b3 = __ldg(a.stage_ptr + 3 * st + 3);
    for (uint32_t k = b0 + tid; k < b1; k += nth) sn_backward<1>(a, k, 0, nullptr);
    if ((threadIdx.x & 31u) == 0) {  // no carry from whatever used the stage before
        uint32_t* carry = reinterpret_cast<uint32_t*>(warp_stage + kCarryOffset);
        carry[0] = UINT32_MAX;
        carry[1] = UINT32_MAX;
    }
    __syncwarp();
    for (uint32_t k = b1 + (threadIdx.x >> 5) * a.vgrid + vblock_of(a), nw = nth >> 5; k < b2; k += nw)
        sn_backward<32>(a, k, threadIdx.x & 31u, warp_stage, k + nw < b2 ? k + nw : UINT32_MAX, k + 2 * nw < b2 ? k + 2 * nw : UINT32_MAX);
    if (b3 > b2) {
        __syncthreads();
        for (uint32_t k = b2 + vblock_of(a); k < b3; k += a.vgrid) sn_backward<512>(a, k, threadIdx.x, cta_stage);
    }
}

static_assert(team_stage_doubles<512>() <= (512 / 32) * kWarpStageDoubles, "the CTA team's stage must fit the CTA's dynamic shared memory");
static_assert(team_stage_doubles<128>() <= 4 * kWarpStageDoubles, "a quad's stage must fit the stages of its four warps");
constexpr uint32_t kBlock = 512;
constexpr uint32_t kClusterCtas = 8;        // portable maximum cluster size
constexpr size_t kSingleCtaWork = 4096;     // n + m + nnz up to which one CTA runs the whole solve
constexpr uint32_t kSmDoubles = 4096;  // 32 KB staging

// Constraint::set_from_initial_values (constraints.rs:146-193, called at lib.rs:183-186): only the two tangent
// kinds carry a side; Undefined ones are resolved from the current x (the initial guesses).
__device__ void resolve_sides(const LargeArgs& a, uint32_t tid, uint32_t nth) {
    const GX X{a.vg + a.X0};
    for (uint32_t t = tid >> 5; t < a.n_tiles; t += nth >> 5) {
        const TileDesc td = a.tiles[t];
        const uint32_t kind = td.meta & 0xffu, lane = threadIdx.x & 31u;
        if ((kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE || kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE) && lane < (td.meta >> 8)) {
            const uint32_t* rec = a.recs + (size_t)td.off16 * 4 + lane;
            const KindLayout ly = a.layout[kind];
            uint32_t side = a.side_flags[t * 32 + lane];  // as given by the caller
            if (side == EZPZ_SIDE_UNDEFINED) {
                uint32_t ids[8];
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) ids[q] = q < cl_nids[kind] ? rec[(ly.ids + q) * 32] : 0u;
                side = ezd::resolve_side(kind, side, ids, X);
            }
            a.side[t * 32 + lane] = (uint8_t)side;
        }
    }
}
__global__ void __launch_bounds__(256) resolve_sides_kernel(const LargeArgs a) {
    resolve_sides(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

constexpr size_t kLmDynamicSmem = (kBlock / 32) * kWarpStageDoubles * sizeof(double);  // 164 KB

// The body of the persistent LM kernel; `a` is the kernel's argument block itself (single system: a __grid_constant__
// parameter, read through the constant bank, never copied) or this CTA's adjusted copy of it (batch mode).
__device__ __forceinline__ void lm_large_body(const LargeArgs& a) {
    __shared__ double sm[kSmDoubles];
    extern __shared__ double warp_stage_all[];  // (kBlock / 32) * kWarpStageDoubles, see kLmDynamicSmem
    double* warp_stage = warp_stage_all + (threadIdx.x >> 5) * kWarpStageDoubles;
    cg::grid_group grid = cg::this_grid();
    // Three launch shapes: one CTA (barrier = __syncthreads), one thread-block cluster of kClusterCtas CTAs for mid-size
    // systems (hardware cluster barrier, ~0.2 us, acquire/release at cluster scope), the whole GPU (cooperative grid barrier).
    const bool single = a.vgrid == 1, clustered = a.cluster != 0;
    auto sync = [&]() {
        if (single) __syncthreads();
        else if (clustered) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        } else grid.sync();
    };
    const uint32_t tid = vblock_of(a) * blockDim.x + threadIdx.x, nth = a.vgrid * blockDim.x;
    const uint32_t G = a.vgrid;
    LargeCtrl* ctrl = a.ctrl;
    double* vg = a.vg;
    double* pm = a.partials;           // max / pAp
    double* ps1 = a.partials + G;      // rz
    double* ps2 = a.partials + 2 * G;  // rr
    const bool use_cg = !a.direct;
    unsigned long long t_mark = now_ns(), t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long t_begin = t_mark;
    auto lap = [&](int k) {  // charge the time since the previous mark to phase k (thread 0's view)
        const unsigned long long t = now_ns();
        t_acc[k] += t - t_mark;
        t_mark = t;
    };

    // pipelined assembly when a warp's slice of the dynamic shared memory holds at least two record tiles and every warp has a
    // few tiles to stream (below that the set-up of the pipeline costs more than it hides: massive 2,000 variables, 8 CTAs, 35 ->
    // 45 us; 1M-variable chain 842 -> 564 us over the 9 phases of a solve).  EZPZ_B200_PIPE_ASM=0 switches it off.
    // EZPZ_B200_PIPE_ASM=2 forces it for every size (tests, compute-sanitizer on small systems).
    const bool pipe_asm = a.pipe_asm && (a.pipe_asm == 2u || a.n_tiles >= 4u * (nth >> 5)) &&
                          2u * a.tile_bytes_max + kAsmStagesMax * 8u + 512u <= kWarpStageDoubles * 8u;
    uint32_t fact_epoch = 0;  // factorisations so far in this launch: what a diagonal slice publishes (sn_factor_slice)
    // sides from the initial guesses (lib.rs:183-186), counters
    {
        if (a.sn_flag)
            for (uint32_t k = tid; k < a.n_sn; k += nth) a.sn_flag[k] = 0;
        resolve_sides(a, tid, nth);
        for (uint32_t c = tid; c < a.n_cons; c += nth) a.degen[c] = 0;
        for (uint32_t w = tid; w < (a.n_cons + 31) / 32; w += nth) a.unsat[w] = 0;
        if (tid == 0) {
            ctrl->lambda = a.initial_lambda;
            ctrl->iterations = a.max_iterations;
            ctrl->converged = 0;
            ctrl->fail = 0;
            ctrl->any_unsat = 0;
            ctrl->any_degen = 0;
            ctrl->lin_iters = 0;
        }
    }
    sync();
    if (pipe_asm) assemble_phase_pipe<true, true>(a, a.R0, tid, nth, use_cg, warp_stage, kWarpStageDoubles * 8u);
    else assemble_phase<true, true>(a, a.R0, tid, nth, use_cg);
    sync();
    lap(0);
    // S = sum r^2 (see the header comment for the two summation orders)
    const uint32_t n_chunks = a.sum_chunk ? (a.m + a.sum_chunk - 1) / a.sum_chunk : 0u;
    auto sum_squares = [&](const double* v, double* ctrl_slot) -> double {
        if (!a.sum_chunk) {
            sequential_sum_squares(v, a.m, ctrl_slot, sm, kSmDoubles);
            __syncthreads();
            return *ctrl_slot;
        }
        chunk_sum_squares(v, a.m, a.sum_chunk, a.sumsq, tid, nth, warp_stage);
        sync();
        return fold_sum(a.sumsq, n_chunks, sm);
    };

    double lambda = a.initial_lambda, S = sum_squares(vg + a.R0, &ctrl->S);
    uint32_t iterations = a.max_iterations;
    bool converged = false;
    uint32_t lin_iters = 0;
    for (uint32_t it = 0; it < a.max_iterations; ++it) {
        max_abs_partial(vg + a.R0, a.m, tid, nth, &pm[vblock_of(a)], sm);
        sync();
        const double largest = fold_max(pm, G, sm);
        lap(1);
        if (largest <= a.residual_tolerance) {
            iterations = it;
            converged = true;
            break;
        }
        bool fail = false;
        if (!use_cg) {
            ++fact_epoch;  // (the same count in every CTA: the loop is uniform across the grid)
            direct_zero(a, tid, nth);
            sync();
            direct_assemble(a, lambda, tid, nth);
            sync();
            lap(2);
            for (uint32_t st = 0; st < a.n_levels; ++st) {
                const unsigned long long t0 = a.lvl_ns ? now_ns() : 0ull;
                direct_factor_stage(a, st, tid, nth, warp_stage, warp_stage_all, fact_epoch);
                sync();
                if (a.lvl_ns && tid == 0) a.lvl_ns[2 * st] += now_ns() - t0;
            }
            lap(3);
            for (uint32_t st = a.n_levels; st-- > 0;) {
                const unsigned long long t0 = a.lvl_ns ? now_ns() : 0ull;
                direct_backward_stage(a, st, tid, nth, warp_stage, warp_stage_all);
                sync();
                if (a.lvl_ns && tid == 0) a.lvl_ns[2 * st + 1] += now_ns() - t0;
            }
            lap(5);
            fail = ctrl->fail != 0;
        } else {
            // Jacobi-preconditioned CG on (JtJ + lambda I) d = -Jt r
            double* p = a.cgv;
            double* res = p + a.n;
            double* ap = res + a.n;
            double* dinv = ap + a.n;
            double* q = dinv + a.n;
            double* d = vg + a.D0;
            const double* jv = vg + a.J0;
            const double* r = vg + a.R0;
            double lrz = 0.0, lbb = 0.0;
            for (uint32_t j = tid; j < a.n; j += nth) {
                double b = 0.0, dg = 0.0;
                for (uint32_t e = a.csc_col_ptr[j]; e < a.csc_col_ptr[j + 1]; ++e) {
                    const double v = jv[e];
                    b = __fma_rn(v, -r[a.csc_row_idx[e]], b);
                    dg = __fma_rn(v, v, dg);
                }
                const double di = 1.0 / (dg + lambda);
                dinv[j] = di;
                d[j] = 0.0;
                res[j] = b;
                const double z = di * b;
                p[j] = z;
                lrz += b * z;
                lbb += b * b;
            }
            block_sum(lrz, &ps1[vblock_of(a)], sm);
            block_sum(lbb, &ps2[vblock_of(a)], sm);
            sync();
            double rz = fold_sum(ps1, G, sm);
            const double bb = fold_sum(ps2, G, sm);
            const double stop = a.cg_rtol * a.cg_rtol * bb;
            if (bb > 0.0) {
                for (uint32_t k = 0; k < a.cg_max_iters; ++k) {
                    for (uint32_t i = tid; i < a.m; i += nth) {  // q = J p
                        double s = 0.0;
                        for (uint32_t e = a.csr_row_ptr[i]; e < a.csr_row_ptr[i + 1]; ++e) s = __fma_rn(a.jr[e], p[a.csr_col_idx[e]], s);
                        q[i] = s;
                    }
                    sync();
                    double lpap = 0.0;
                    for (uint32_t j = tid; j < a.n; j += nth) {  // ap = Jt q + lambda p
                        double s = lambda * p[j];
                        for (uint32_t e = a.csc_col_ptr[j]; e < a.csc_col_ptr[j + 1]; ++e) s = __fma_rn(jv[e], q[a.csc_row_idx[e]], s);
                        ap[j] = s;
                        lpap += p[j] * s;
                    }
                    block_sum(lpap, &pm[vblock_of(a)], sm);
                    sync();
                    const double pap = fold_sum(pm, G, sm);
                    ++lin_iters;
                    if (!(pap > 0.0) || !ezm::ez_isfinite(pap)) {
                        fail = true;
                        break;
                    }
                    const double alpha = rz / pap;
                    double lrz2 = 0.0, lrr = 0.0;
                    for (uint32_t j = tid; j < a.n; j += nth) {
                        d[j] = d[j] + alpha * p[j];
                        const double rj = res[j] - alpha * ap[j];
                        res[j] = rj;
                        lrz2 += rj * (dinv[j] * rj);
                        lrr += rj * rj;
                    }
                    block_sum(lrz2, &ps1[vblock_of(a)], sm);
                    block_sum(lrr, &ps2[vblock_of(a)], sm);
                    sync();
                    const double rz2 = fold_sum(ps1, G, sm);
                    const double rr = fold_sum(ps2, G, sm);
                    if (rr <= stop) break;
                    const double beta = rz2 / rz;
                    rz = rz2;
                    for (uint32_t j = tid; j < a.n; j += nth) p[j] = dinv[j] * res[j] + beta * p[j];
                    sync();
                }
            }
            sync();
            lap(6);
        }
        if (fail) {
            sync();
            if (tid == 0) ctrl->fail = 0;
            lambda *= 10.0;
            sync();
            continue;
        }
        max_abs_partial(vg + a.D0, a.n, tid, nth, &pm[vblock_of(a)], sm);
        sync();
        const double step = fold_max(pm, G, sm);
        for (uint32_t j = tid; j < a.n; j += nth) vg[a.X0 + j] += vg[a.D0 + j];
        sync();
        lap(1);
        if (pipe_asm) assemble_phase_pipe<true, false>(a, a.RN0, tid, nth, false, warp_stage, kWarpStageDoubles * 8u);
        else assemble_phase<true, false>(a, a.RN0, tid, nth, false);
        sync();
        lap(0);
        const double S2 = sum_squares(vg + a.RN0, &ctrl->S2);
        lap(1);
        if (S2 < S) {
            for (uint32_t i = tid; i < a.m; i += nth) vg[a.R0 + i] = vg[a.RN0 + i];
            if (pipe_asm) assemble_phase_pipe<false, true>(a, a.R0, tid, nth, use_cg, warp_stage, kWarpStageDoubles * 8u);
            else assemble_phase<false, true>(a, a.R0, tid, nth, use_cg);
            S = S2;
            lambda *= 0.1;
        } else {
            for (uint32_t j = tid; j < a.n; j += nth) vg[a.X0 + j] -= vg[a.D0 + j];
            lambda *= 10.0;
        }
        sync();
        lap(0);
        if (step <= a.step_tolerance) {
            iterations = it;
            converged = true;
            break;
        }
    }
    // post-solve verdict (lib.rs:305-327): unweighted residuals, |r| < 1e-4 per component
    {
        for (uint32_t t = tid >> 5; t < a.n_tiles; t += nth >> 5) {
            const TileDesc td = a.tiles[t];
            const uint32_t kind = td.meta & 0xffu, lane = threadIdx.x & 31u;
            if (lane >= (td.meta >> 8)) continue;
            const uint32_t k = t * 32 + lane;
            const uint32_t* rec = a.recs + (size_t)td.off16 * 4 + lane;
            const KindLayout ly = a.layout[kind];
            double xv[8];
            gather_x(a, ly, kind, rec, xv);
            const uint32_t ident[8] = {0, 1, 2, 3, 4, 5, 6, 7};
            const RegX XR{xv};
            const uint32_t side = (kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE || kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE) ? a.side[k] : 0u;
            ezd::EvalOut o;
            ezd::eval_constraint<false>(kind, side, ident, ly.p0 != 0xff ? rec_double(rec, ly.p0) : 0.0,
                                        ly.p1 != 0xff ? rec_double(rec, ly.p1) : 0.0, XR, o);
            bool sat = ezm::ez_abs(o.res[0]) < ezd::kEps;
            if (cl_rows[kind] == 2) sat = sat && (ezm::ez_abs(o.res[1]) < ezd::kEps);
            if (!sat) {
                const uint32_t c = a.slot_orig[k];
                atomicOr(&a.unsat[c >> 5], 1u << (c & 31u));
                ctrl->any_unsat = 1;
            }
        }
    }
    if (tid == 0) {
        lap(0);
        t_acc[7] = now_ns() - t_begin;
        for (int k = 0; k < 8; ++k) ctrl->t[k] = t_acc[k];
        ctrl->iterations = iterations;
        ctrl->converged = converged ? 1u : 0u;
        ctrl->lambda = lambda;
        ctrl->S = S;
        ctrl->lin_iters = lin_iters;
    }
}

// One system per launch (one CTA, one cluster, or the whole GPU): the argument block stays where the launch put it.
__global__ void __launch_bounds__(kBlock, 1) lm_large_kernel(const __grid_constant__ LargeArgs a) { lm_large_body(a); }

// One problem per CTA (batches of mid-size systems): this CTA's slices of the per-problem arrays.
__global__ void __launch_bounds__(kBlock, 1) lm_large_batch_kernel(const LargeArgs a_in) {
    LargeArgs a = a_in;
    const size_t b = blockIdx.x;
    a.vg += b * a.vg_stride;
    a.jr += b * a.jr_stride;
    a.cgv += b * a.cgv_stride;
    a.sumsq += b * a.sumsq_stride;
    a.side += b * a.side_stride;
    a.degen += b * a.degen_stride;
    a.unsat += b * a.unsat_stride;
    a.partials += b * 3;
    a.ctrl += b;
    lm_large_body(a);

}

// ---- stand-alone kernels for throughput measurement (same device code as the phases above) ------------
// Persistent warps; every warp owns tiles gw, gw + nw, ... and keeps kAsmStages of them in flight: lane 0 arms the
// stage's mbarrier with the tile's byte count and issues one cp.async.bulk for the whole tile; the warp waits on
// the barrier's phase, evaluates its 32 constraints out of shared memory, and refills the stage.
template <int kAsmWarps, int kCtasPerSm, int kStageTiles, bool kWriteJr>
__global__ void __launch_bounds__(kAsmWarps * 32, kCtasPerSm) assemble_large_kernel(const LargeArgs a) {
    constexpr size_t kAsmWarpBytes = (size_t)kStageTiles * kTileBytesMax;
    extern __shared__ __align__(128) unsigned char asm_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    unsigned char* stage_base = asm_smem + (size_t)warp * kAsmWarpBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(asm_smem + (size_t)kAsmWarps * kAsmWarpBytes) + warp * kAsmStagesMax;
    uint2* ring = reinterpret_cast<uint2*>(asm_smem + (size_t)kAsmWarps * (kAsmWarpBytes + kAsmStagesMax * sizeof(uint64_t))) + warp * 64;
    const uint32_t gw = blockIdx.x * kAsmWarps + warp, nw = gridDim.x * kAsmWarps;
    const uint32_t kStageBytes = a.tile_bytes_max, kAsmStages = min(kAsmStagesMax, (uint32_t)(kAsmWarpBytes / kStageBytes));
    if (lane == 0) {
        for (uint32_t st = 0; st < kAsmStages; ++st) mbar_init(&bars[st], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const uint32_t count = gw < a.n_tiles ? (a.n_tiles - gw + nw - 1) / nw : 0u;
    // Tile descriptors: this warp's descriptors go through a 64-entry ring in shared memory, filled 32 at a time (lane l
    // fetches the descriptor of the warp's (32 b + l)-th tile) one block AHEAD of their use, so that neither the bulk copy
    // of tile k + kAsmStages nor the gathers of tile k + 1 wait for a dependent global load.
    auto load_block = [&](uint32_t b) {
        const uint32_t k = b * 32 + lane;
        return k < count ? *reinterpret_cast<const uint2*>(&a.tiles[gw + k * nw]) : make_uint2(0u, 0u);
    };
    auto desc = [&](uint32_t k) {
        const uint2 v = ring[k & 63u];
        return TileDesc{v.x, v.y};
    };
    ring[lane] = load_block(0);
    ring[32 + lane] = load_block(1);
    uint2 pf = load_block(2);
    __syncwarp();
    auto issue = [&](uint32_t st, const TileDesc td) {  // lane 0: start the copy of a tile into stage st
        const uint32_t bytes = (uint32_t)a.layout[td.meta & 0xffu].n_words * 128u;
        mbar_expect_tx(&bars[st], bytes);
        bulk_copy_g2s(stage_base + (size_t)st * kStageBytes, a.recs + (size_t)td.off16 * 4, bytes, &bars[st]);
    };
    if (lane == 0)
        for (uint32_t k = 0; k < kAsmStages && k < count; ++k) issue(k, desc(k));
    // The gathers of tile k + 1 (and its tangent-side byte) are issued before tile k is evaluated (its record is already in
    // shared memory, more are in flight), so their latency hides behind the arithmetic.
    auto is_tangent = [](uint32_t kind) { return kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE || kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE; };
    double xv[8], xn[8];
    uint32_t side = 0, side_n = 0;
    TileDesc td = desc(0);
    if (count) {
        mbar_wait(&bars[0], 0);
        if (lane < (td.meta >> 8)) {
            gather_x(a, a.layout[td.meta & 0xffu], td.meta & 0xffu, reinterpret_cast<const uint32_t*>(stage_base) + lane, xv);
            if (is_tangent(td.meta & 0xffu)) side = a.side[gw * 32 + lane];
        }
    }
    // st = k % kAsmStages, sn = (k + 1) % kAsmStages, par_n = ((k + 1) / kAsmStages) & 1 — kept incrementally
    uint32_t st = 0, sn = kAsmStages > 1 ? 1u : 0u, par_n = kAsmStages > 1 ? 0u : 1u;
    for (uint32_t k = 0, t = gw; k < count; ++k, t += nw) {
        if ((k & 31u) == 0 && k) {  // entering block b = k / 32: bring block b + 1 into the ring, start fetching block b + 2
            const uint32_t b = k >> 5;
            ring[((b + 1) & 1u) * 32 + lane] = pf;
            pf = load_block(b + 2);
            __syncwarp();
        }
        const uint32_t kind = td.meta & 0xffu, n_valid = td.meta >> 8;
        TileDesc tn{0, 0};
        if (k + 1 < count) {
            tn = desc(k + 1);
            mbar_wait(&bars[sn], par_n);
            if (lane < (tn.meta >> 8)) {
                gather_x(a, a.layout[tn.meta & 0xffu], tn.meta & 0xffu,
                         reinterpret_cast<const uint32_t*>(stage_base + (size_t)sn * kStageBytes) + lane, xn);
                if (is_tangent(tn.meta & 0xffu)) side_n = a.side[(t + nw) * 32 + lane];
            }
        }
        if (lane < n_valid)
            assemble_slot<true, true, !kWriteJr>(a, a.layout[kind], kind,
                                                 reinterpret_cast<const uint32_t*>(stage_base + (size_t)st * kStageBytes) + lane,
                                                 t * 32 + lane, a.R0, kWriteJr, xv, side);
        __syncwarp();
        if (lane == 0 && k + kAsmStages < count) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(st, desc(k + kAsmStages));
        }
        td = tn;
        side = side_n;
#pragma unroll
        for (int q = 0; q < 8; ++q) xv[q] = xn[q];
        st = sn;
        if (++sn == kAsmStages) {
            sn = 0;
            par_n ^= 1u;
        }
    }
}
// Launch shapes of the assembly kernel: warps per CTA x CTAs per SM (the register budget follows) x widest tiles staged
// per warp.  EZPZ_B200_ASM_VARIANT selects one for measurements; the default is the fastest measured on B200.
template <int W, int C, int T>
cudaError_t launch_assemble_as(const LargeArgs& a, bool write_jr, uint32_t n_tiles, int sm_count, cudaStream_t st) {
    constexpr size_t smem = (size_t)W * T * kTileBytesMax + (size_t)W * kAsmStagesMax * sizeof(uint64_t) + (size_t)W * 64 * sizeof(uint2);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(assemble_large_kernel<W, C, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(assemble_large_kernel<W, C, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n_tiles + W - 1) / W, (size_t)sm_count * C));
    if (write_jr) assemble_large_kernel<W, C, T, true><<<grid, W * 32, smem, st>>>(a);
    else assemble_large_kernel<W, C, T, false><<<grid, W * 32, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_assemble(const LargeArgs& a, bool write_jr, uint32_t n_tiles, int sm_count, cudaStream_t st) {
    static const int variant = [] {
        const char* e = std::getenv("EZPZ_B200_ASM_VARIANT");
        return e ? std::atoi(e) : 0;
    }();
    switch (variant) {
        case 1: return launch_assemble_as<5, 4, 2>(a, write_jr, n_tiles, sm_count, st);   // 20 warps/SM, <= 96 registers
        case 2: return launch_assemble_as<6, 4, 2>(a, write_jr, n_tiles, sm_count, st);   // 24 warps/SM, <= 80 registers
        case 3: return launch_assemble_as<4, 5, 2>(a, write_jr, n_tiles, sm_count, st);   // 20 warps/SM, <= 96 registers
        case 4: return launch_assemble_as<4, 4, 3>(a, write_jr, n_tiles, sm_count, st);   // 16 warps/SM in 4 CTAs
        default: return launch_assemble_as<8, 2, 3>(a, write_jr, n_tiles, sm_count, st);  // 16 warps/SM, 128 registers
    }
}

// Thread per row, entries ascending, one fma chain (measured faster than staging the warp's entry range through
// shared memory: both are bound by the L1's sector lookups for the x gathers, and shared memory shares that pipe).
__global__ void __launch_bounds__(256) spmv_csr_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col_idx,
                                                       const double* __restrict__ vals, const double* __restrict__ x,
                                                       double* __restrict__ y, uint32_t rows) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (uint32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) s = __fma_rn(vals[e], x[col_idx[e]], s);
        y[i] = s;
    }
}

struct LargeDevice {
    uint32_t *recs = nullptr, *slot_orig = nullptr;
    TileDesc* tiles = nullptr;
    uint8_t* side_flags = nullptr;
    KindLayout layout[EZPZ_K_COUNT];
    uint32_t n_slots = 0, n_tiles = 0, tile_bytes_max = 128;
    uint32_t *csr_row_ptr = nullptr, *csr_col_idx = nullptr, *csc_col_ptr = nullptr, *csc_row_idx = nullptr;
    uint32_t* jmap = nullptr;          // device copy of LargeProgram::jt_of_csc (direct path)
    uint32_t* direct_tables[11] = {};  // device copies of the LargeProgram arrays of the sparse direct solve
    double *vg = nullptr, *jr = nullptr, *cgv = nullptr, *partials = nullptr, *sumsq = nullptr;
    unsigned long long* lvl_ns = nullptr;
    uint8_t* side = nullptr;
    uint32_t *degen = nullptr, *unsat = nullptr;
    LargeCtrl* ctrl = nullptr;
    unsigned char* resblk = nullptr;  // [ctrl | unsat | degen], see get_large
    void* arena = nullptr;            // the one allocation all of the above point into
    uint32_t* sn_flag = nullptr;      // [n_sn], see LargeArgs
    size_t resblk_bytes = 0;
    int grid = 0;
    bool cluster = false;
    bool tables = false;
    // per-problem state of the one-CTA-per-problem batch mode (solve_large_batch), grown on demand
    struct Batch {
        double *vg = nullptr, *jr = nullptr, *cgv = nullptr, *sumsq = nullptr, *partials = nullptr;
        uint8_t* side = nullptr;
        uint32_t *degen = nullptr, *unsat = nullptr;
        LargeCtrl* ctrl = nullptr;
        uint64_t cap = 0;  // problems the buffers hold
    } batch;
};

// The device tables of one structure live in ONE allocation: the uploads are packed into a host staging block as they are
// produced and the work buffers are reserved behind them; finalize() allocates once, copies once, zeroes the work buffers once
// and hands out the pointers.  (One cudaMalloc + synchronous cudaMemcpy per table — some thirty of each — made the first solve
// of a 2,000-variable topology cost more than the CPU's whole solve.)
struct DevicePlan {
    struct Item {
        void** dst;
        size_t off;
        const void* src;
        size_t bytes;
    };
    std::vector<Item> up, raw;
    size_t up_bytes = 0, reserve_bytes = 0;
    // (the source must stay alive and unchanged until finalize() returns)
    template <class T, class A>
    int32_t upload(T** dst, const std::vector<T, A>& src) {
        up.push_back({reinterpret_cast<void**>(dst), up_bytes, src.data(), sizeof(T) * src.size()});
        up_bytes += (std::max<size_t>(sizeof(T), sizeof(T) * src.size()) + 255) / 256 * 256;
        return EZPZ_OK;
    }
    template <class T>
    void reserve(T** dst, size_t bytes) {  // zero-initialised
        raw.push_back({reinterpret_cast<void**>(dst), reserve_bytes, nullptr, 0});
        reserve_bytes += (std::max<size_t>(bytes, 8) + 255) / 256 * 256;
    }
    int32_t finalize(void** arena, cudaStream_t st, ezpz_error_detail_t* detail) {
        // stream-ordered allocation out of the device's default pool (its release threshold is lifted in context_create): a
        // topology analysed again after its structure was destroyed gets the same memory back without a trip to the driver's
        // physical allocator, which costs milliseconds for a block of a few MB
        EZ_CUDA(cudaMallocAsync(arena, std::max<size_t>(256, up_bytes + reserve_bytes), st), "cudaMallocAsync(structure tables)");
        unsigned char* base = static_cast<unsigned char*>(*arena);
        // Small structures: the tables are packed into one host block and leave with ONE copy (some thirty copies of a few
        // hundred bytes cost more than the solve of a 2,000-variable system).  Large ones: every table is copied from where
        // it lies (packing 400 MB of tables of a 1M-variable sketch into a staging block first cost more than the copies).
        std::vector<unsigned char> stage;
        if (up_bytes <= (4u << 20)) {
            stage.resize(up_bytes);
            for (const Item& it : up)
                if (it.bytes) std::memcpy(stage.data() + it.off, it.src, it.bytes);
            if (up_bytes) EZ_CUDA(cudaMemcpyAsync(base, stage.data(), up_bytes, cudaMemcpyHostToDevice, st), "cudaMemcpy(structure tables)");
        } else {
            EZ_CUDA(cudaMemsetAsync(base, 0, up_bytes, st), "cudaMemset(structure tables)");  // (the alignment gaps read as zero)
            for (const Item& it : up)
                if (it.bytes) EZ_CUDA(cudaMemcpyAsync(base + it.off, it.src, it.bytes, cudaMemcpyHostToDevice, st), "cudaMemcpy(structure table)");
        }
        if (reserve_bytes) EZ_CUDA(cudaMemsetAsync(base + up_bytes, 0, reserve_bytes, st), "cudaMemset(work buffers)");
        for (const Item& it : up) *it.dst = base + it.off;
        for (const Item& it : raw) *it.dst = base + up_bytes + it.off;
        EZ_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");  // (the sources may go after this)
        return EZPZ_OK;
    }
};
#define EZ_TRY(x)                    \
    do {                             \
        int32_t rc__ = (x);          \
        if (rc__ != EZPZ_OK) return rc__; \
    } while (0)

int32_t get_large(ezpz_context* ctx, const ezpz_structure* s, DeviceCopy* dc, LargeDevice** out, ezpz_error_detail_t* detail) {
    // one LargeDevice per context (it holds the work buffers of a solve, not only the tables); creation under the structure's
    // lock, the creating context's stream does the upload
    std::lock_guard<std::mutex> lock(const_cast<ezpz_structure*>(s)->dev_mutex);
    for (auto& e : dc->large)
        if (e.first == ctx) {
            *out = (LargeDevice*)e.second;
            return EZPZ_OK;
        }
    const LargeProgram& P = s->large;
    LargeDevice* L = new (std::nothrow) LargeDevice();
    if (!L) return EZPZ_ERR_INVALID_ARGUMENT;
    dc->large.emplace_back(ctx, L);  // owned by the device copy from here on (released with it)
    DevicePlan plan;
    ezs::uvec<uint32_t> recs, orig;  // (sources of uploads: alive until plan.finalize())
    ezs::uvec<uint8_t> flags;
    ezs::uvec<TileDesc> tiles;
    {
        // record tiles (see the comment above assemble_slot)
        static const uint8_t kHasP0[EZPZ_K_COUNT] = {0, 0, 1, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1};
        const bool with_jr = !P.direct;  // only the PCG path keeps the CSR-ordered copy of J
        for (int k = 0; k < EZPZ_K_COUNT; ++k) {
            const ezk::KindInfo& ki = ezk::kKinds[k];
            KindLayout& ly = L->layout[k];
            uint8_t w = 1;  // word 0: first residual row
            ly.ids = w; w += ki.n_ids;
            ly.p0 = kHasP0[k] ? w : 0xff; w += kHasP0[k] ? 2 : 0;
            const bool has_p1 = k == EZPZ_K_LINES_AT_ANGLE || k == EZPZ_K_ARC_ANGLE || k == EZPZ_K_POINTS_AT_ANGLE;
            ly.p1 = has_p1 ? w : 0xff; w += has_p1 ? 2 : 0;
            ly.weight = s->all_weights_one ? 0xff : w; w += s->all_weights_one ? 0 : 2;
            ly.slot0 = w; w += ki.emit_len[0];
            ly.slot1 = w; w += ki.emit_len[1];
            ly.jr = with_jr ? w : 0xff; w += with_jr ? 1 + ki.rows : 0;
            ly.n_words = w;
        }
        const uint32_t n_slots = (uint32_t)P.cons_order.size(), n_tiles = n_slots / 32;  // cons_order is padded to whole warps
        tiles.resize(n_tiles);
        orig.resize(n_slots);
        flags.resize(std::max<uint32_t>(1, n_slots));
        flags[0] = 0;
        // offsets first (a tile's size follows from its kind), then ranges of tiles on host threads, each writing its own span
        ezs::uvec<uint64_t> tile_base((size_t)n_tiles + 1);
        tile_base[0] = 0;
        for (uint32_t t = 0; t < n_tiles; ++t)  // first slot of a tile is never padding
            tile_base[t + 1] = tile_base[t] + (uint64_t)L->layout[s->dev_cons[P.cons_order[(size_t)t * 32]].kind].n_words * 32;
        if (tile_base[n_tiles] / 4 > 0xfffffff0ull) return EZPZ_ERR_TOO_LARGE;
        recs.resize(tile_base[n_tiles]);
        ezs::parallel_ranges(n_tiles, 256, [&](uint32_t tb, uint32_t te, uint32_t) {
        for (uint32_t t = tb; t < te; ++t) {
            const uint32_t c0 = P.cons_order[(size_t)t * 32];
            const uint32_t kind = s->dev_cons[c0].kind;
            const KindLayout& ly = L->layout[kind];
            const ezk::KindInfo& ki = ezk::kKinds[kind];
            const size_t base = tile_base[t];  // multiple of 32 words = 128 bytes
            std::fill(recs.begin() + base, recs.begin() + tile_base[t + 1], 0u);
            uint32_t n_valid = 0;
            for (uint32_t l = 0; l < 32; ++l) {
                const uint32_t c = P.cons_order[(size_t)t * 32 + l];
                if (c == UINT32_MAX) {  // padding sits at the end of the kind group
                    orig[(size_t)t * 32 + l] = 0;
                    flags[(size_t)t * 32 + l] = 0;
                    continue;
                }
                n_valid = l + 1;
                const DevCons& dc = s->dev_cons[c];
                auto put = [&](uint32_t word, uint32_t v) { recs[base + (size_t)word * 32 + l] = v; };
                auto put_d = [&](uint32_t word, double v) {
                    uint32_t dw[2];
                    std::memcpy(dw, &v, 8);
                    put(word, dw[0]);
                    put(word + 1, dw[1]);
                };
                orig[(size_t)t * 32 + l] = c;
                flags[(size_t)t * 32 + l] = (uint8_t)(dc.flags & 0xffu);
                put(0, dc.row0);
                for (int q = 0; q < ki.n_ids; ++q) put(ly.ids + q, dc.ids[q]);
                if (ly.p0 != 0xff) put_d(ly.p0, dc.p0);
                if (ly.p1 != 0xff) put_d(ly.p1, dc.p1);
                if (ly.weight != 0xff) put_d(ly.weight, dc.weight);
                const uint32_t jr0 = s->csr_row_ptr[dc.row0];
                if (ly.jr != 0xff) put(ly.jr, jr0);
                for (int row = 0; row < ki.rows; ++row) {
                    uint32_t code = 0;
                    for (int q = 0; q < ki.emit_len[row]; ++q) {
                        const uint32_t off = s->csc_to_csr[dc.slot[row][q] & ~kAccumulate] - jr0;  // < 16: two rows of <= 8
                        code |= (off & 15u) << (4 * q);
                        const uint32_t sl = dc.slot[row][q];  // direct path: J lives in tile order (structure.h)
                        put((row == 0 ? ly.slot0 : ly.slot1) + q,
                            P.jt_of_csc.empty() ? sl : (P.jt_of_csc[sl & ~kAccumulate] | (sl & kAccumulate)));
                    }
                    if (ly.jr != 0xff) put(ly.jr + 1 + row, code);
                }
            }
            tiles[t].off16 = (uint32_t)(base / 4);
            tiles[t].meta = kind | (n_valid << 8);
        }
        });
        EZ_TRY(plan.upload(&L->recs, recs));
        EZ_TRY(plan.upload(&L->tiles, tiles));
        EZ_TRY(plan.upload(&L->slot_orig, orig));
        EZ_TRY(plan.upload(&L->side_flags, flags));
        L->n_slots = n_slots;
        L->n_tiles = n_tiles;
        for (uint32_t t = 0; t < n_tiles; ++t)
            L->tile_bytes_max = std::max<uint32_t>(L->tile_bytes_max, (uint32_t)L->layout[tiles[t].meta & 0xffu].n_words * 128u);
    }
    EZ_TRY(plan.upload(&L->csr_row_ptr, s->csr_row_ptr));
    EZ_TRY(plan.upload(&L->csr_col_idx, s->csr_col_idx));
    EZ_TRY(plan.upload(&L->csc_col_ptr, s->csc_col_ptr));
    EZ_TRY(plan.upload(&L->csc_row_idx, s->csc_row_idx));
    if (P.direct) {
        const ezs::uvec<uint32_t>* src[11] = {&P.perm, &P.sn_rows, &P.upd_rel, &P.upd_rec, &P.stage_ptr, &P.stage_rec, &P.aent_slot,
                                                &P.aprod_ptr, &P.aprod_a, &P.aprod_b, &P.diag_slot};
        for (int k = 0; k < 11; ++k) EZ_TRY(plan.upload(&L->direct_tables[k], *src[k]));
        if (!P.jt_of_csc.empty()) EZ_TRY(plan.upload(&L->jmap, P.jt_of_csc));
    }
    const size_t nnz = s->csc_row_idx.size();
    plan.reserve(&L->vg, sizeof(double) * std::max<size_t>(1, P.VG));
    plan.reserve(&L->jr, sizeof(double) * std::max<size_t>(1, nnz));
    plan.reserve(&L->cgv, sizeof(double) * (4 * (size_t)s->n + s->m + 1));
    if (const char* dbg = std::getenv("EZPZ_B200_DEBUG"); dbg && dbg[0] == '1' && dbg[1] == '2' && P.direct)
        plan.reserve(&L->lvl_ns, sizeof(unsigned long long) * 2 * std::max<size_t>(1, P.n_levels));
    plan.reserve(&L->sumsq, sizeof(double) * ((size_t)s->m / 64 + 2));
    if (P.direct) plan.reserve(&L->sn_flag, sizeof(uint32_t) * std::max<size_t>(1, P.stage_rec.size() / 8));
    plan.reserve(&L->side, std::max<size_t>(1, L->n_slots));
    // control block, unsatisfied mask and degenerate counters side by side: a solve reads them back with one copy
    const size_t b_ctrl = align_up(sizeof(LargeCtrl), 256), b_unsat = align_up(sizeof(uint32_t) * ((s->n_cons + 31) / 32 + 1), 256),
                 b_degen = sizeof(uint32_t) * std::max<size_t>(1, s->n_cons);
    L->resblk_bytes = b_ctrl + b_unsat + b_degen;
    plan.reserve(&L->resblk, L->resblk_bytes);
    // grid: one CTA for small systems (barriers are __syncthreads), else every SM, co-resident
    const size_t work = (size_t)s->n + s->m + nnz;
    int per_sm = 0;
    EZ_CUDA(cudaFuncSetAttribute(lm_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLmDynamicSmem),
            "cudaFuncSetAttribute(lm_large_kernel)");
    EZ_CUDA(cudaFuncSetAttribute(lm_large_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLmDynamicSmem),
            "cudaFuncSetAttribute(lm_large_batch_kernel)");
    EZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lm_large_kernel, kBlock, kLmDynamicSmem), "occupancy");
    if (per_sm < 1) per_sm = 1;
    // tiny systems: one CTA; up to 65,536 values: one cluster of 8 CTAs; beyond: every SM, co-resident
    L->grid = work <= kSingleCtaWork ? 1 : (work <= 65536 ? (int)kClusterCtas : ctx->sm_count * std::min(per_sm, 2));
    L->cluster = work > kSingleCtaWork && work <= 65536;
    plan.reserve(&L->partials, sizeof(double) * 3 * (size_t)L->grid);
    EZ_TRY(plan.finalize(&L->arena, ctx->stream, detail));
    L->ctrl = reinterpret_cast<LargeCtrl*>(L->resblk);
    L->unsat = reinterpret_cast<uint32_t*>(L->resblk + b_ctrl);
    L->degen = reinterpret_cast<uint32_t*>(L->resblk + b_ctrl + b_unsat);
    uint8_t rows[EZPZ_K_COUNT], emit_len[EZPZ_K_COUNT][2], nids[EZPZ_K_COUNT];
    for (int k = 0; k < EZPZ_K_COUNT; ++k) {
        nids[k] = ezk::kKinds[k].n_ids;
        rows[k] = ezk::kKinds[k].rows;
        emit_len[k][0] = ezk::kKinds[k].emit_len[0];
        emit_len[k][1] = ezk::kKinds[k].emit_len[1];
    }
    EZ_CUDA(cudaMemcpyToSymbol(cl_rows, rows, sizeof rows), "cudaMemcpyToSymbol");
    EZ_CUDA(cudaMemcpyToSymbol(cl_emit_len, emit_len, sizeof emit_len), "cudaMemcpyToSymbol");
    EZ_CUDA(cudaMemcpyToSymbol(cl_nids, nids, sizeof nids), "cudaMemcpyToSymbol");
    *out = L;
    return EZPZ_OK;
}

void fill_args(LargeArgs& a, const ezpz_structure* s, const DeviceCopy* dc, const LargeDevice* L, const ezpz_config_t* config) {
    const LargeProgram& P = s->large;
    std::memset(&a, 0, sizeof a);
    a.recs = L->recs;
    a.tiles = L->tiles;
    a.slot_orig = L->slot_orig;
    a.side_flags = L->side_flags;
    std::memcpy(a.layout, L->layout, sizeof a.layout);
    a.n_slots = L->n_slots;
    a.n_tiles = L->n_tiles;
    a.tile_bytes_max = L->tile_bytes_max;
    a.unit_weights = s->all_weights_one ? 1u : 0u;
    a.csr_row_ptr = L->csr_row_ptr;
    a.csr_col_idx = L->csr_col_idx;
    a.csc_col_ptr = L->csc_col_ptr;
    a.csc_row_idx = L->csc_row_idx;
    a.csc_to_csr = dc->csc_to_csr;
    a.jmap = L->jmap;
    {
        const uint32_t** dst[11] = {&a.perm, &a.sn_rows, &a.upd_rel, &a.upd_rec, &a.stage_ptr, &a.stage_rec, &a.aent_slot, &a.aprod_ptr,
                                    &a.aprod_a, &a.aprod_b, &a.diag_slot};
        for (int k = 0; k < 11; ++k) *dst[k] = L->direct_tables[k];
    }
    a.sumsq = L->sumsq;
    a.lvl_ns = L->lvl_ns;
    a.vg = L->vg;
    a.jr = L->jr;
    a.cgv = L->cgv;
    a.partials = L->partials;
    a.side = L->side;
    a.degen = L->degen;
    a.unsat = L->unsat;
    a.ctrl = L->ctrl;
    a.residual_tolerance = config->residual_tolerance;
    a.step_tolerance = config->step_tolerance;
    a.initial_lambda = config->initial_lambda;
    a.cg_rtol = 1e-13;
    a.max_iterations = (uint32_t)std::min<uint64_t>(config->max_iterations, 0x7fffffffu);
    a.cg_max_iters = 20000;
    a.n_cons = s->n_cons;
    a.n = s->n;
    a.m = s->m;
    a.nnz = (uint32_t)s->csc_row_idx.size();
    a.n_levels = P.n_levels;
    a.nnz_l = P.nnz_l;
    a.n_aent = (uint32_t)P.aent_slot.size();
    a.X0 = P.X0;
    a.R0 = P.R0;
    a.RN0 = P.RN0;
    a.J0 = P.J0;
    a.L0 = P.L0;
    a.RV0 = P.RV0;
    a.Y0 = P.Y0;
    a.D0 = P.D0;
    a.direct = P.direct ? 1u : 0u;
    a.sn_flag = L->sn_flag;
    {
        const char* e = std::getenv("EZPZ_B200_PIPE_ASM");
        a.pipe_asm = e ? (e[0] == '2' ? 2u : (e[0] != '0' ? 1u : 0u)) : 1u;
    }
    a.n_sn = (uint32_t)(P.stage_rec.size() / 8);
    // the summation order belongs to the structure, not to the launch shape: a system solved alone (cluster / grid) and
    // the same system solved as one CTA's problem in a batch fold S the same way
    a.sum_chunk = ezs::sum_chunk_for(s->n, s->m, s->csc_row_idx.size());
}

// Batch mode epilogue: per problem, iterations / status out of its control block and x, masks, Jacobian out of its slices.
struct BatchOut {
    double* finals;
    uint32_t* iterations;
    uint8_t* status;
    uint32_t* unsat;
    uint32_t* degen;
    double* jac;
};
// J in CSC order for the callers that ask for it (ezpz_b200_eval, the Jacobian export of the solve calls).
__global__ void __launch_bounds__(256) export_j_csc_kernel(const LargeArgs a, double* __restrict__ out) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < a.nnz; e += gridDim.x * blockDim.x) out[e] = a.vg[a.J0 + a.jmap[e]];
}
__global__ void __launch_bounds__(256) large_batch_scatter_kernel(const LargeArgs a, const double* __restrict__ guesses) {
    const size_t b = blockIdx.x;
    double* x = a.vg + b * a.vg_stride + a.X0;
    for (uint32_t j = threadIdx.x; j < a.n; j += blockDim.x) x[j] = guesses[b * a.n + j];
}
__global__ void __launch_bounds__(256) large_batch_gather_kernel(const LargeArgs a, const BatchOut o) {
    const size_t b = blockIdx.x;
    const double* vg = a.vg + b * a.vg_stride;
    for (uint32_t j = threadIdx.x; j < a.n; j += blockDim.x) o.finals[b * a.n + j] = vg[a.X0 + j];
    const uint32_t uw = (a.n_cons + 31) / 32;
    if (o.unsat)
        for (uint32_t w = threadIdx.x; w < uw; w += blockDim.x) o.unsat[b * uw + w] = a.unsat[b * a.unsat_stride + w];
    if (o.degen)
        for (uint32_t c = threadIdx.x; c < a.n_cons; c += blockDim.x) o.degen[b * a.n_cons + c] = a.degen[b * a.degen_stride + c];
    if (o.jac)
        for (uint32_t e = threadIdx.x; e < a.nnz; e += blockDim.x) o.jac[b * a.nnz + e] = vg[a.J0 + (a.jmap ? a.jmap[e] : e)];
    if (threadIdx.x == 0) {
        const LargeCtrl* c = a.ctrl + b;
        o.iterations[b] = c->iterations;
        o.status[b] = (uint8_t)((c->converged ? EZPZ_ST_CONVERGED : 0u) | (c->any_unsat ? EZPZ_ST_UNSATISFIED : 0u) |
                                (c->any_degen ? EZPZ_ST_DEGENERATE : 0u));
    }
}

}  // namespace

namespace ezs {

void release_large(DeviceCopy* d) {
    for (auto& entry : d->large) {
    LargeDevice* L = (LargeDevice*)entry.second;
    if (!L) continue;
    if (L->arena) cudaFreeAsync(L->arena, nullptr);  // every table and work buffer of the structure (DevicePlan); back to the pool
    void* bptrs[] = {L->batch.vg, L->batch.jr, L->batch.cgv, L->batch.sumsq, L->batch.partials, L->batch.side,
                     L->batch.degen, L->batch.unsat, L->batch.ctrl};
    for (void* p : bptrs)
        if (p) cudaFree(p);
    delete L;
    }
    d->large.clear();
}

// One residual + Jacobian evaluation at x through the large path's own assembly kernel (ezpz_b200_eval).
int32_t eval_large(ezpz_context* ctx, const ezpz_structure* s, const double* x, double* r, double* jac_csc, double* jac_csr,
                   bool* have_csr, ezpz_error_detail_t* detail) {
    if (!s->large.built) return EZPZ_ERR_UNSUPPORTED;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    DeviceCopy* dc = nullptr;
    EZ_TRY(get_device_copy(ctx, s, &dc, detail));
    LargeDevice* L = nullptr;
    EZ_TRY(get_large(ctx, s, dc, &L, detail));
    ezpz_config_t cfg;
    ezpz_b200_config_default(&cfg);
    LargeArgs a;
    fill_args(a, s, dc, L, &cfg);
    cudaStream_t st = ctx->stream;
    const size_t nnz = s->csc_row_idx.size();
    EZ_CUDA(cudaMemcpyAsync(L->vg + a.X0, x, sizeof(double) * s->n, cudaMemcpyHostToDevice, st), "H2D x");
    EZ_CUDA(cudaMemsetAsync(L->degen, 0, sizeof(uint32_t) * s->n_cons, st), "memset degen");
    resolve_sides_kernel<<<(unsigned)std::max<size_t>(1, std::min<size_t>((L->n_tiles + 7) / 8, 4096)), 256, 0, st>>>(a);
    const bool with_jr = !s->large.direct;
    EZ_CUDA(launch_assemble(a, with_jr, L->n_tiles, ctx->sm_count, st), "assemble_large_kernel launch");
    ctx->launches += 2;
    if (r) EZ_CUDA(cudaMemcpyAsync(r, L->vg + a.R0, sizeof(double) * s->m, cudaMemcpyDeviceToHost, st), "D2H r");
    if (jac_csc) {
        const double* src = L->vg + a.J0;
        if (a.jmap && nnz) {  // tile order -> CSC order through the (otherwise unused on the direct path) CSR value buffer
            export_j_csc_kernel<<<(unsigned)std::min<size_t>((nnz + 255) / 256, 4096), 256, 0, st>>>(a, L->jr);
            ctx->launches += 1;
            src = L->jr;
        }
        EZ_CUDA(cudaMemcpyAsync(jac_csc, src, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st), "D2H jac");
    }
    if (jac_csr && with_jr) EZ_CUDA(cudaMemcpyAsync(jac_csr, L->jr, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st), "D2H jac csr");
    if (have_csr) *have_csr = with_jr;
    EZ_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return EZPZ_OK;
}

int32_t solve_large(ezpz_context* ctx, const ezpz_structure* s, const ezpz_config_t* config, const ezpz_one_io_t* io,
                    ezpz_error_detail_t* detail) {
    if (!s->large.built) return EZPZ_ERR_TOO_LARGE;
    if (s->n_cons == 0 || s->m == 0) return EZPZ_ERR_EMPTY_SYSTEM;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    DeviceCopy* dc = nullptr;
    EZ_TRY(get_device_copy(ctx, s, &dc, detail));
    LargeDevice* L = nullptr;
    EZ_TRY(get_large(ctx, s, dc, &L, detail));
    LargeArgs a;
    fill_args(a, s, dc, L, config);
    cudaStream_t st = ctx->stream;
    // Small systems go through the context's pinned staging buffer: one DMA in, two or three out and a single synchronisation,
    // instead of five pageable copies that each stall the host (profiles/r01j_*).
    const size_t b_x = sizeof(double) * s->n, b_j = io->jacobian ? sizeof(double) * a.nnz : 0;
    const bool staged_io = 2 * b_x + L->resblk_bytes + b_j <= ((size_t)1 << 20);
    unsigned char* pin = nullptr;
    if (staged_io) {
        EZ_TRY(ensure_pin(ctx, 2 * b_x + L->resblk_bytes + b_j, detail));
        pin = static_cast<unsigned char*>(ctx->pin);
        std::memcpy(pin, io->guesses, b_x);
    }
    EZ_CUDA(cudaMemcpyAsync(L->vg + a.X0, staged_io ? (const void*)pin : (const void*)io->guesses, b_x, cudaMemcpyHostToDevice, st), "H2D guesses");
    void* params[] = {(void*)&a};
    a.cluster = L->cluster ? 1u : 0u;
    if (L->grid == 1) {
        a.vgrid = 1;
        lm_large_kernel<<<1, kBlock, kLmDynamicSmem, st>>>(a);
    } else if (L->cluster) {
        a.vgrid = kClusterCtas;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kClusterCtas);
        cfg.blockDim = dim3(kBlock);
        cfg.dynamicSmemBytes = kLmDynamicSmem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kClusterCtas;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, lm_large_kernel, a) != cudaSuccess) {
            // no room for a cluster of this shape on this device/partition: the same 8 CTAs with the grid barrier
            (void)cudaGetLastError();
            L->cluster = false;
            a.cluster = 0;
            a.vgrid = (uint32_t)L->grid;
            EZ_CUDA(cudaLaunchCooperativeKernel((void*)lm_large_kernel, dim3(L->grid), dim3(kBlock), params, kLmDynamicSmem, st),
                    "cudaLaunchCooperativeKernel(lm_large_kernel)");
        }
    } else {
        a.vgrid = (uint32_t)L->grid;
        EZ_CUDA(cudaLaunchCooperativeKernel((void*)lm_large_kernel, dim3(L->grid), dim3(kBlock), params, kLmDynamicSmem, st),
                "cudaLaunchCooperativeKernel(lm_large_kernel)");
    }
    ctx->launches += 1;
    EZ_CUDA(cudaGetLastError(), "lm_large_kernel launch");
    LargeCtrl h;
    const size_t uw = (s->n_cons + 31) / 32;
    const double* jsrc = L->vg + a.J0;
    if (io->jacobian && a.jmap && a.nnz) {
        export_j_csc_kernel<<<(unsigned)std::min<size_t>(((size_t)a.nnz + 255) / 256, 4096), 256, 0, st>>>(a, L->jr);
        ctx->launches += 1;
        jsrc = L->jr;
    }
    if (staged_io) {
        unsigned char* p_f = pin + b_x;
        unsigned char* p_r = p_f + b_x;
        unsigned char* p_j = p_r + L->resblk_bytes;
        EZ_CUDA(cudaMemcpyAsync(p_r, L->resblk, L->resblk_bytes, cudaMemcpyDeviceToHost, st), "D2H results");
        EZ_CUDA(cudaMemcpyAsync(p_f, L->vg + a.X0, b_x, cudaMemcpyDeviceToHost, st), "D2H finals");
        if (io->jacobian) EZ_CUDA(cudaMemcpyAsync(p_j, jsrc, b_j, cudaMemcpyDeviceToHost, st), "D2H jacobian");
        EZ_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
        std::memcpy(&h, p_r, sizeof h);
        std::memcpy(io->final_values, p_f, b_x);
        if (io->unsat_mask) std::memcpy(io->unsat_mask, p_r + (reinterpret_cast<unsigned char*>(L->unsat) - L->resblk), sizeof(uint32_t) * uw);
        if (io->degen_count)
            std::memcpy(io->degen_count, p_r + (reinterpret_cast<unsigned char*>(L->degen) - L->resblk), sizeof(uint32_t) * s->n_cons);
        if (io->jacobian) std::memcpy(io->jacobian, p_j, b_j);
    } else {
        EZ_CUDA(cudaMemcpyAsync(&h, L->ctrl, sizeof h, cudaMemcpyDeviceToHost, st), "D2H ctrl");
        EZ_CUDA(cudaMemcpyAsync(io->final_values, L->vg + a.X0, b_x, cudaMemcpyDeviceToHost, st), "D2H finals");
        if (io->unsat_mask) EZ_CUDA(cudaMemcpyAsync(io->unsat_mask, L->unsat, sizeof(uint32_t) * uw, cudaMemcpyDeviceToHost, st), "D2H unsat");
        if (io->degen_count)
            EZ_CUDA(cudaMemcpyAsync(io->degen_count, L->degen, sizeof(uint32_t) * s->n_cons, cudaMemcpyDeviceToHost, st), "D2H degen");
        if (io->jacobian) EZ_CUDA(cudaMemcpyAsync(io->jacobian, jsrc, b_j, cudaMemcpyDeviceToHost, st), "D2H jacobian");
        EZ_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    }
    const char* dbg = std::getenv("EZPZ_B200_DEBUG");
    if (dbg && dbg[0] == '1' && dbg[1] == '2' && a.lvl_ns) {
        std::vector<unsigned long long> t(2 * (size_t)a.n_levels);
        cudaMemcpy(t.data(), a.lvl_ns, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost);
        const LargeProgram& P = s->large;
        for (uint32_t l = 0; l < a.n_levels; ++l)
            std::fprintf(stderr, "  stage %3u panels %7u  factor %8.1f us  backward %8.1f us (all iterations)\n", l,
                         P.stage_ptr[3 * l + 3] - P.stage_ptr[3 * l], t[2 * l] * 1e-3, t[2 * l + 1] * 1e-3);
        cudaMemset(a.lvl_ns, 0, sizeof(unsigned long long) * t.size());
    }
    if (dbg && dbg[0] == '1')
        std::fprintf(stderr,
                     "[lm_large_kernel] grid %d x %u  it %u  us: eval %.1f  sums/max %.1f  A+rhs %.1f  factor stages %.1f  "
                     "(unused) %.1f  backward stages %.1f  pcg %.1f  total %.1f\n",
                     L->grid, kBlock, h.iterations, h.t[0] * 1e-3, h.t[1] * 1e-3, h.t[2] * 1e-3, h.t[3] * 1e-3, h.t[4] * 1e-3,
                     h.t[5] * 1e-3, h.t[6] * 1e-3, h.t[7] * 1e-3);
    *io->iterations = h.iterations;
    *io->status = (uint8_t)((h.converged ? EZPZ_ST_CONVERGED : 0u) | (h.any_unsat ? EZPZ_ST_UNSATISFIED : 0u) |
                            (h.any_degen ? EZPZ_ST_DEGENERATE : 0u));
    if (io->path_used) *io->path_used = s->large.direct ? 1 : 2;
    if (io->lin_iters) *io->lin_iters = h.lin_iters;
    return EZPZ_OK;
}

// A batch of problems of one structure that does not fit the thread-per-problem kernel: the persistent LM kernel of the
// single-system path with ONE CTA PER PROBLEM (barriers are __syncthreads, the structure's tables are shared, every
// problem owns a slice of the state arrays).  Same arithmetic as solve_large on each problem, hence bit-identical to
// solving them one by one.  All io pointers are device pointers; work is enqueued on `st`.
int32_t solve_large_batch(ezpz_context* ctx, const ezpz_structure* s, const ezpz_config_t* config, uint64_t batch,
                          const ezpz_batch_io_t* io, cudaStream_t st, ezpz_error_detail_t* detail) {
    if (!s->large.built) return EZPZ_ERR_TOO_LARGE;
    if (s->n_cons == 0 || s->m == 0) return EZPZ_ERR_EMPTY_SYSTEM;
    if (io->params) {
        if (detail)
            std::snprintf(detail->message, sizeof detail->message,
                          "per-problem parameter overrides are only available on the thread-per-problem kernel");
        return EZPZ_ERR_UNSUPPORTED;
    }
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    DeviceCopy* dc = nullptr;
    EZ_TRY(get_device_copy(ctx, s, &dc, detail));
    LargeDevice* L = nullptr;
    EZ_TRY(get_large(ctx, s, dc, &L, detail));
    const LargeProgram& P = s->large;
    const size_t nnz = s->csc_row_idx.size();
    const bool use_cg = !P.direct;
    const size_t vg_stride = align_up(std::max<size_t>(1, P.VG), 32), jr_stride = use_cg ? align_up(std::max<size_t>(1, nnz), 32) : 0,
                 cgv_stride = use_cg ? align_up(4 * (size_t)s->n + s->m + 1, 32) : 0,
                 sumsq_stride = align_up((size_t)s->m / 64 + 2, 16), side_stride = align_up(std::max<size_t>(1, L->n_slots), 16),
                 degen_stride = s->n_cons, unsat_stride = (s->n_cons + 31) / 32 + 1;
    const size_t per_problem = 8 * (vg_stride + jr_stride + cgv_stride + sumsq_stride + 3) + side_stride +
                               4 * (degen_stride + unsat_stride) + sizeof(LargeCtrl);
    LargeDevice::Batch& B = L->batch;
    if (B.cap < batch) {
        // as many problems per launch as a quarter of the free memory holds (at least one wave of CTAs when it fits)
        size_t free_b = 0, total_b = 0;
        EZ_CUDA(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        void* old[] = {B.vg, B.jr, B.cgv, B.sumsq, B.partials, B.side, B.degen, B.unsat, B.ctrl};
        size_t held = (size_t)B.cap * per_problem;
        for (void* p : old)
            if (p) cudaFree(p);
        B = LargeDevice::Batch();
        uint64_t cap = std::min<uint64_t>(batch, std::max<uint64_t>(1, (free_b + held) / 4 / per_problem));
        if (const char* env = std::getenv("EZPZ_B200_LARGE_BATCH_CAP")) cap = std::max<uint64_t>(1, std::min<uint64_t>(cap, std::strtoull(env, nullptr, 10)));
        EZ_CUDA(cudaMalloc(&B.vg, 8 * vg_stride * cap), "cudaMalloc(batch vg)");
        EZ_CUDA(cudaMemsetAsync(B.vg, 0, 8 * vg_stride * cap, st), "cudaMemset(batch vg)");
        EZ_CUDA(cudaMalloc(&B.jr, std::max<size_t>(8, 8 * jr_stride * cap)), "cudaMalloc(batch jr)");
        EZ_CUDA(cudaMalloc(&B.cgv, std::max<size_t>(8, 8 * cgv_stride * cap)), "cudaMalloc(batch cgv)");
        EZ_CUDA(cudaMalloc(&B.sumsq, 8 * sumsq_stride * cap), "cudaMalloc(batch sumsq)");
        EZ_CUDA(cudaMalloc(&B.partials, 8 * 3 * cap), "cudaMalloc(batch partials)");
        EZ_CUDA(cudaMalloc(&B.side, side_stride * cap), "cudaMalloc(batch side)");
        EZ_CUDA(cudaMalloc(&B.degen, 4 * std::max<size_t>(1, degen_stride) * cap), "cudaMalloc(batch degen)");
        EZ_CUDA(cudaMalloc(&B.unsat, 4 * unsat_stride * cap), "cudaMalloc(batch unsat)");
        EZ_CUDA(cudaMalloc(&B.ctrl, sizeof(LargeCtrl) * cap), "cudaMalloc(batch ctrl)");
        EZ_CUDA(cudaMemsetAsync(B.ctrl, 0, sizeof(LargeCtrl) * cap, st), "cudaMemset(batch ctrl)");
        B.cap = cap;
    }
    LargeArgs a;
    fill_args(a, s, dc, L, config);
    a.vg = B.vg;
    a.jr = B.jr;
    a.cgv = B.cgv;
    a.sumsq = B.sumsq;
    a.partials = B.partials;
    a.side = B.side;
    a.degen = B.degen;
    a.unsat = B.unsat;
    a.ctrl = B.ctrl;
    a.lvl_ns = nullptr;
    a.cluster = 0;
    a.batch = 1;
    a.vg_stride = vg_stride;
    a.jr_stride = jr_stride;
    a.cgv_stride = cgv_stride;
    a.sumsq_stride = sumsq_stride;
    a.side_stride = side_stride;
    a.degen_stride = degen_stride;
    a.unsat_stride = unsat_stride;
    const uint32_t uw = (s->n_cons + 31) / 32;
    for (uint64_t b0 = 0; b0 < batch; b0 += B.cap) {
        const unsigned cnt = (unsigned)std::min<uint64_t>(B.cap, batch - b0);
        large_batch_scatter_kernel<<<cnt, 256, 0, st>>>(a, io->guesses + b0 * s->n);
        a.vgrid = 1;
        lm_large_batch_kernel<<<cnt, kBlock, kLmDynamicSmem, st>>>(a);
        BatchOut o;
        o.finals = io->final_values + b0 * s->n;
        o.iterations = io->iterations + b0;
        o.status = io->status + b0;
        o.unsat = io->unsat_mask ? io->unsat_mask + b0 * uw : nullptr;
        o.degen = io->degen_count ? io->degen_count + b0 * s->n_cons : nullptr;
        o.jac = io->jacobian ? io->jacobian + b0 * nnz : nullptr;
        large_batch_gather_kernel<<<cnt, 256, 0, st>>>(a, o);
        ctx->launches += 3;
        EZ_CUDA(cudaGetLastError(), "lm_large_kernel batch launch");
    }
    return EZPZ_OK;
}

}  // namespace ezs

// Stand-alone launches of the assembly and SpMV kernels on a structure's large-system buffers, timed with
// CUDA events: reps launches each, returns mean microseconds per launch and the algorithmic bytes moved
// (SURVEY.md §8d formulas).  x must hold n doubles.  which: 0 assembly (J in CSC order, the direct path's), 1 SpMV
// y = J p (CSR), 2 SpMV z = Jt q (CSC), 3 assembly writing both the CSC and the CSR-ordered copy (PCG path's).
extern "C" int32_t ezpz_b200_large_bench(ezpz_context_t* ctx, const ezpz_structure_t* s, const double* x, int32_t which,
                                         int32_t reps, double* mean_us, double* algorithmic_bytes,
                                         ezpz_error_detail_t* detail) {
    if (!ctx || !s || !x || !mean_us || reps < 1) return EZPZ_ERR_INVALID_ARGUMENT;
    if (detail) std::memset(detail, 0, sizeof *detail);
    if (!s->large.built) return EZPZ_ERR_UNSUPPORTED;
    EZ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    DeviceCopy* dc = nullptr;
    EZ_TRY(get_device_copy(ctx, s, &dc, detail));
    LargeDevice* L = nullptr;
    EZ_TRY(get_large(ctx, s, dc, &L, detail));
    ezpz_config_t cfg;
    ezpz_b200_config_default(&cfg);
    LargeArgs a;
    fill_args(a, s, dc, L, &cfg);
    cudaStream_t st = ctx->stream;
    EZ_CUDA(cudaMemcpyAsync(L->vg + a.X0, x, sizeof(double) * s->n, cudaMemcpyHostToDevice, st), "H2D x");
    EZ_CUDA(cudaMemsetAsync(L->side, 1, L->n_slots, st), "memset side");
    EZ_CUDA(cudaMemsetAsync(L->degen, 0, sizeof(uint32_t) * s->n_cons, st), "memset degen");
    const unsigned grid = (unsigned)ctx->sm_count * 8;
    const double n = s->n, m = s->m, nnz = (double)s->csc_row_idx.size(), C = s->n_cons;
    cudaEvent_t e0, e1;
    EZ_CUDA(cudaEventCreate(&e0), "cudaEventCreate");
    EZ_CUDA(cudaEventCreate(&e1), "cudaEventCreate");
    // one untimed launch first (also fills jr for the SpMVs)
    EZ_CUDA(launch_assemble(a, true, L->n_tiles, ctx->sm_count, st), "assemble_large_kernel launch");
    ctx->launches += 1;
    // Launches run back to back inside ONE event pair (a per-launch pair adds ~5 us of launch latency to kernels that
    // take 15-50 us).  For an HBM figure the caller passes a system whose working set exceeds the 126 MB L2
    // (profiles/large_bench.py: 2.08M variables, ~270 MB), so every launch streams from HBM; EZPZ_B200_BENCH_FLUSH_L2=1
    // instead times each launch on its own after a 256 MiB fill.
    const char* fl = std::getenv("EZPZ_B200_BENCH_FLUSH_L2");
    const bool flush = fl && fl[0] == '1';
    void* flush_buf = nullptr;
    const size_t flush_bytes = 256u << 20;
    if (flush) EZ_CUDA(cudaMalloc(&flush_buf, flush_bytes), "cudaMalloc(flush)");
    float ms = 0.f;
    auto launch = [&]() {
        if (which == 0 || which == 3) (void)launch_assemble(a, which == 3, L->n_tiles, ctx->sm_count, st);
        else if (which == 1) spmv_csr_kernel<<<grid, 256, 0, st>>>(L->csr_row_ptr, L->csr_col_idx, L->jr, L->vg + a.X0, L->cgv + 4 * (size_t)s->n, s->m);
        else spmv_csr_kernel<<<grid, 256, 0, st>>>(L->csc_col_ptr, L->csc_row_idx, L->vg + a.J0, L->vg + a.R0, L->cgv, s->n);
        ctx->launches += 1;
    };
    if (flush) {
        for (int k = 0; k < reps; ++k) {
            EZ_CUDA(cudaMemsetAsync(flush_buf, k & 0xff, flush_bytes, st), "memset flush");
            EZ_CUDA(cudaEventRecord(e0, st), "cudaEventRecord");
            launch();
            EZ_CUDA(cudaEventRecord(e1, st), "cudaEventRecord");
            EZ_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
            float one = 0.f;
            EZ_CUDA(cudaEventElapsedTime(&one, e0, e1), "cudaEventElapsedTime");
            ms += one;
        }
    } else {
        EZ_CUDA(cudaEventRecord(e0, st), "cudaEventRecord");
        for (int k = 0; k < reps; ++k) launch();
        EZ_CUDA(cudaEventRecord(e1, st), "cudaEventRecord");
        EZ_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
        EZ_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
    }
    EZ_CUDA(cudaGetLastError(), "bench kernels");
    if (flush_buf) cudaFree(flush_buf);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *mean_us = (double)ms * 1e3 / reps;
    if (algorithmic_bytes) {
        // 64-byte records + x + r + J values + scatter slots (+ the CSR-ordered copy, which == 3)
        if (which == 0) *algorithmic_bytes = 64.0 * C + 8 * n + 8 * m + 8 * nnz + 4 * nnz;
        else if (which == 3) *algorithmic_bytes = 64.0 * C + 8 * n + 8 * m + 2 * 8 * nnz + 4 * nnz;
        else if (which == 1) *algorithmic_bytes = 12 * nnz + 4 * (m + 1) + 8 * n + 8 * m;
        else *algorithmic_bytes = 12 * nnz + 4 * (n + 1) + 8 * m + 8 * n;
    }
    return EZPZ_OK;
}
