// structure.h — host-side analysis of one sketch topology (internal C++ view of ezpz_structure_t).
//
// Replaces, once per topology instead of once per solve, the work of Model::new
// (ezpz/src/solver.rs:192-300): id validation (:142-189), the sorted + deduplicated sparsity pattern of
// J (:217-265), and the symbolic factorisation of A = JtJ + lambda*I (:289-300).  On top of that it
// precomputes what the device needs so that no kernel ever searches: the scatter slot of every partial
// derivative (the reference searches the column linearly per nonzero, solver.rs:412-418) and the
// operation tapes of the batched small-system kernel.
#pragma once
#include <cstdint>
#include <mutex>
#include <vector>

#include "../../include/ezpz_b200.h"
#include "host_parallel.h"

struct ezpz_structure;

namespace ezs {

// Analysed constraint as the device reads it (112 bytes, 16-byte aligned).
struct DevCons {
    double p0, p1, weight;
    uint32_t kind, flags, row0, side_slot;  // side_slot: index of the per-problem side value, or 0xffffffff
    uint32_t ids[8];
    // Scatter slots (index into the Jacobian value array, CSC order) of each emitted partial, emission
    // order; bit 31 set = this slot was already written by an earlier partial of the same row, accumulate.
    uint32_t slot[2][8];
};
static_assert(sizeof(DevCons) == 136, "DevCons layout");

constexpr uint32_t kAccumulate = 0x80000000u;

// Ops of the sequential small-system tape (SmallProgram::tape), host format, 3 + pairs words each:
//   {dst | pairs << 16, fin | code << 16, added}, then one word a | b << 16 per multiply-add.
// acc starts from V[dst] (OP_INIT_DST) or +0.0; the first `added` pairs are accumulated as fma(V[a], V[b], acc); lambda is
// added behind them when OP_MID_LAMBDA is set; the remaining pairs as fma(-V[a], V[b], acc); then the finalisation.
// Three forms occur (build_small_program), which the device tape names as shapes:
//   TAPE_PIVOT     +0.0, added pairs (A[j][j]), + lambda, subtracted pairs (L[j][k]^2), pivot: fail unless acc > 0, 1/sqrt
//   TAPE_ENTRY     +0.0, added pairs, subtracted pairs, * V[fin]       (L[i][j] with its A[i][j]; y[i] with its b[i])
//   TAPE_BACKWARD  V[dst], subtracted pairs, * V[fin]                  (backward substitution, in place)
enum : uint32_t {
    OP_INIT_DST = 1u,
    OP_FIN_SHIFT = 2u,  // bits 2..3: finalisation
    OP_FIN_NONE = 0u, OP_FIN_MUL = 2u, OP_FIN_PIVOT = 3u,
    OP_MID_LAMBDA = 32u
};
enum : uint32_t { TAPE_BARRIER = 0u, TAPE_ENTRY = 1u, TAPE_PIVOT = 2u, TAPE_BACKWARD = 3u };

// The batched small-system program: slot map of the per-problem value array V and the op tape.
struct SmallProgram {
    bool valid = false;       // false when the system does not fit the thread-per-problem kernel
    uint32_t W = 0;           // doubles per problem
    uint32_t X0 = 0, R0 = 0, RN0 = 0, J0 = 0, L0 = 0, D0 = 0, S0 = 0;
    uint32_t F0 = 0;          // one slot of flags shared by the warps that cooperate on a problem (device.cu)
    uint32_t n_side = 0;
    uint32_t n_ops = 0;       // ops executed per LM iteration (assemble, rhs, factor, forward, backward)
    uint64_t n_pairs = 0;     // multiply-add pairs per LM iteration
    std::vector<uint32_t> tape;
};

// The small program compiled for R cooperating warps ("roles") per 32 problems and one shared-memory stride: what
// lm_roles_kernel (device.cu) stages in shared memory.  32-bit words:
//   [R x kRoleHdrWords]  per role: {cons list offset, count, tape offset, ops, x_lo, x_hi, r_lo, r_hi, j_lo, j_hi, 0, 0}
//   [n_cons x DevCons]   the analysed constraints, input order
//   per role             the constraints it evaluates (indices, ascending)
//   per role             its tape: per op {dst byte offset, fin byte offset, added | subtracted << 16, shape | words to the
//                        next header << 8} + {a, b} byte offsets per pair (+ two pad words after an odd number of pairs:
//                        headers and pair quads are 16-byte aligned), and one pad header at the end;
//                        shape TAPE_BARRIER = all roles of the group meet here (a cross-role dependency follows)
// Every op of the sequential tape (SmallProgram::tape) appears in exactly one role's tape with its pairs in the same order, so
// every value is produced by the same chain of roundings whatever R is.
constexpr uint32_t kRoleHdrWords = 12;
struct RoleBlob {
    uint32_t R = 0, stride = 0;
    std::vector<uint32_t> words;
    uint32_t cons_word = 0;       // word offset of the DevCons array
    uint32_t tape_barriers = 0;   // barriers inside one execution of the tape
    double busiest_share = 1.0;   // busiest role's share of the tape's cost (1/R = perfectly balanced)
    // Cost model of one LM iteration on the busiest role, in units of one tape word pair (device.cu picks the number of
    // roles with it): its tape (3 per op + 2 per multiply-add), its constraints (kEvalCost) and what every role repeats
    // (the folds over r and d, the barriers).
    double critical_cost = 0.0;
};
// structure.cpp
void build_role_blob(const ezpz_structure& S, uint32_t R, uint32_t stride, RoleBlob& out);

// The single-large-system programme (large.cu): the processing order of the assembly phase and, unless the
// factor would be too large, the supernodal sparse direct solve built by sparse_direct.cpp.  Slots index one
// global value array  VG = [x | r | r_next | J | L panels | 1/pivot | y | d].  J is stored in CSC order on the PCG path (its
// SpMVs walk columns) and in TILE ORDER on the direct path (jt_of_csc): the partial q of lane l of record tile t lives at
// jt_base(t) + q * 32 + l, so a warp's store of one partial is 256 contiguous bytes instead of 32 scattered sectors.
struct LargeProgram {
    using u32v = uvec<uint32_t>;  // (filled by host threads: no serial zero-fill, host_parallel.h)
    bool built = false;
    bool direct = false;          // sparse direct solve available (otherwise the PCG path runs)
    bool nested = false;          // perm is a nested-dissection order (false: natural order 0..n-1)
    uint32_t X0 = 0, R0 = 0, RN0 = 0, J0 = 0, L0 = 0, RV0 = 0, Y0 = 0, D0 = 0;
    uint64_t VG = 0;
    uint32_t n_levels = 0;        // stages = height of the supernode tree
    uint32_t nnz_l = 0;           // doubles of panel storage
    u32v cons_order;                      // processing slots of the assembly phase -> constraint
                                                           // index (tile-local kind sort, UINT32_MAX = padding)
    u32v perm;                            // elimination position -> variable
    u32v jt_of_csc;                       // direct path: position in the J region of each CSC entry (empty = CSC order)
    uint32_t n_j = 0;                                      // doubles of the J region (nnz in CSC order; more in tile order: idle lanes)
    // Supernode s = columns [sn_ptr[s], sn_ptr[s+1]) (a chain of the elimination tree, <= 16 columns).  Its panel is
    // dense, row-major, h x w doubles at VG[L0 + panel_off[s]]: rows sn_rows[sn_row_ptr[s] ...) = the w own columns
    // (the diagonal block) followed by the sorted union of the columns' sub-diagonal rows.
    u32v sn_ptr, sn_row_ptr, sn_rows, panel_off;
    // Updates of supernode J by its descendants, ascending: upd_sn[u] = K, whose panel rows upd_rbegin[u].. (to the end
    // of K's row list) all lie in J's panel; the first upd_ncols[u] of them are columns of J.  upd_rel[upd_rel_ptr[u] + t]
    // = position in J's row list of K's row upd_rbegin[u] + t.
    // The block of K's panel an update reads (rows upd_rbegin[u] to the end, all of K's columns) is contiguous.
    // upd_rec = the 8-word record per update the device reads (sparse_direct.cpp).
    u32v upd_ptr, upd_sn, upd_rbegin, upd_ncols, upd_rel_ptr, upd_rel, upd_rec;
    // Stage k (height in the supernode tree): stage_sn[stage_ptr[3k] .. stage_ptr[3k+1]) = panels of a few doubles (one
    // thread each), [3k+1 .. 3k+2) = panels that fit a warp's shared-memory stage, [3k+2 .. 3k+3) = larger ones (one CTA each).
    // stage_rec = the 8-word record per supernode in stage order that the device reads (sparse_direct.cpp).
    u32v stage_ptr, stage_sn, stage_rec;
    // A = JtJ: for every entry A has (strictly lower), its panel slot and the products J[r][i] * J[r][j] over shared
    // rows r ascending as pairs of positions in the CSC value array of J; diag_slot[j] = panel slot of A[j][j].
    u32v aent_slot, aprod_ptr, aprod_a, aprod_b, diag_slot;
    // host only: the entries of A per column of the elimination numbering (aent_colptr, n + 1) and their rows (aent_row), from
    // which the product lists are rebuilt when constraints are added without a new coupling (ezpz_b200_structure_extend)
    u32v aent_colptr, aent_row;
};

struct DeviceCopy;  // defined in device.h

}  // namespace ezs

struct ezpz_structure {
    uint32_t n_cons = 0, n = 0, m = 0;
    ezs::uvec<ezpz_constraint_t> cons;  // (uvec: filled by host threads, host_parallel.h)
    std::vector<uint32_t> cons_row0;  // n_cons + 1
    // J pattern, both orientations, and the permutations between their value orders
    std::vector<uint32_t> csc_col_ptr, csr_row_ptr;
    ezs::uvec<uint32_t> csc_row_idx, csr_col_idx, csr_to_csc, csc_to_csr;
    // lower(A) and L patterns, CSC with the diagonal first in every column
    std::vector<uint32_t> a_col_ptr, l_col_ptr, l_row_idx;
    ezs::uvec<uint32_t> a_row_idx;
    // connected components of the graph of A: comp_of[var]
    std::vector<uint32_t> comp_of;
    uint32_t n_components = 0;
    uint32_t max_component = 0;  // vars in the largest component
    ezs::uvec<ezs::DevCons> dev_cons;
    uint32_t n_side = 0;
    bool all_weights_one = true;
    ezs::SmallProgram small;
    ezs::LargeProgram large;
    bool have_l_pattern = false;  // false when the symbolic factorisation was skipped (very large systems)
    // ids of the initial guesses (ezpz_b200_structure_create's var_ids) as a membership table, empty when the ids were 0..n-1:
    // what validate_variables needs again when constraints are added (ezpz_b200_structure_extend)
    std::vector<uint8_t> var_present;
    bool l_pattern_built = false; // l_col_ptr / l_row_idx are filled (at creation when the batched kernel may run, else on request)
    // device copies, one per CUDA device ordinal, created lazily
    std::mutex dev_mutex;
    std::vector<ezs::DeviceCopy*> dev;
    std::vector<ezs::RoleBlob*> role_probes;  // stride-1 role programmes, one per role count tried: table size and cost model
};

namespace ezs {
// Summation order of S = sum r^2 on the large path: 0 = one sequential fold (systems of at most 4,096
// values, which one CTA solves), otherwise rows are folded sequentially inside chunks of this many rows and the chunk sums
// are folded sequentially: the power of two next to sqrt(m), within [64, 1024], so that neither chain is long.
inline uint32_t sum_chunk_for(uint32_t n, uint32_t m, size_t nnz) {
    if ((size_t)n + m + nnz <= 4096) return 0;
    uint32_t c = 64;
    while (c < 1024 && (uint64_t)c * c < m) c *= 2;
    return c;
}
// Implemented in device.cu; called by ezpz_b200_structure_destroy.
void release_device_copies(ezpz_structure* s);
// sparse_direct.cpp: ordering, symbolic factorisation and level schedule of the large-system direct solve.
// `order_hint` (n entries, elimination position -> variable, or nullptr): an elimination order to keep instead of choosing one
// (ezpz_b200_structure_extend hands over the base structure's).
// `same_a`: an analysed structure over the same variables whose A has exactly this structure's pattern: its sparse-direct
// schedule is taken over and only the product lists are rebuilt.
void build_sparse_direct(ezpz_structure& S, const uint32_t* order_hint = nullptr, bool hint_nested = false,
                         const ezpz_structure* same_a = nullptr);
}  // namespace ezs
