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

// Tape op codes (see batch_small.cu: run_tape).
enum : uint32_t {
    OP_INIT_DST = 1u,   // acc starts from V[dst] instead of +0.0
    OP_NEGATE = 2u,     // acc = fma(-V[a], V[b], acc) instead of fma(V[a], V[b], acc)
    OP_FIN_SHIFT = 2u,  // bits 2..3: 0 none, 1 acc += lambda, 2 acc *= V[fin], 3 pivot: fail unless acc > 0, acc = 1/sqrt(acc)
    OP_FIN_NONE = 0u, OP_FIN_LAMBDA = 1u, OP_FIN_MUL = 2u, OP_FIN_PIVOT = 3u
};

// The batched small-system program: slot map of the per-problem value array V and the op tape.
struct SmallProgram {
    bool valid = false;       // false when the system does not fit the thread-per-problem kernel
    uint32_t W = 0;           // doubles per problem
    uint32_t X0 = 0, R0 = 0, RN0 = 0, J0 = 0, L0 = 0, D0 = 0, S0 = 0;
    uint32_t n_side = 0;
    uint32_t n_ops = 0;       // ops executed per LM iteration (assemble, rhs, factor, forward, backward)
    uint64_t n_pairs = 0;     // multiply-add pairs per LM iteration
    std::vector<uint32_t> tape;
};

// The single-large-system programme (large.cu): the processing order of the assembly phase and, unless the
// factor would be too large, the level-scheduled sparse direct solve built by sparse_direct.cpp.  Slots index
// one global value array  VG = [x | r | r_next | J (CSC order) | L (by rows) | diag(A) | 1/pivot | y | d].
constexpr uint32_t kEntryInA = 0x80000000u;  // ent_slot flag: A = JtJ itself has this entry (else pure fill)

struct LargeProgram {
    bool built = false;
    bool direct = false;          // sparse direct solve available (otherwise the PCG path runs)
    bool nested = false;          // perm is a nested-dissection order (false: natural order 0..n-1)
    uint32_t X0 = 0, R0 = 0, RN0 = 0, J0 = 0, L0 = 0, DG0 = 0, RV0 = 0, Y0 = 0, D0 = 0;
    uint64_t VG = 0;
    uint32_t n_levels = 0, solo_level = 0, nnz_l = 0;
    std::vector<uint32_t> cons_order;                      // processing slots of the assembly phase -> constraint
                                                           // index (tile-local kind sort, UINT32_MAX = padding)
    std::vector<uint32_t> perm;                            // elimination position -> variable
    std::vector<uint32_t> lr_ptr, lr_col;                  // strictly-lower L by rows, columns ascending; the value
                                                           // of entry e lives at VG[L0 + e]
    std::vector<uint32_t> lvl_ptr;                         // n_levels + 1: ranges of lvl_cols (etree height levels)
    std::vector<uint32_t> lvl_cols;                        // columns ordered by (level, index)
    std::vector<uint32_t> lvl_maxrow;                      // longest row of L among each level's columns
    std::vector<uint32_t> ent_ptr;                         // n + 1, by POSITION in lvl_cols: ranges of ent_*
    std::vector<uint32_t> ent_row, ent_col, ent_slot;      // sub-diagonal entries of each column, rows ascending;
                                                           // ent_slot = position in row order | kEntryInA
    // Static row intersections: for entry e = (i, j), bit t of the first ceil(len_j / 32) words at
    // ent_mask[ent_mask_ptr[e]] says that column lr_col[lr_ptr[j] + t] of row j is also in row i; the following
    // ceil(pre_i / 32) words mark the matching positions of row i's prefix (pre_i = entries of row i left of j).
    // The q-th set bits of the two masks pair up, so the factorisation never compares column indices.
    std::vector<uint32_t> ent_mask_ptr;                    // nnz_l + 1
    std::vector<uint32_t> ent_mask;
    // A = JtJ: for every L entry flagged kEntryInA (in ent_* order), the products J[r][i] * J[r][j] over shared rows
    // r ascending, as pairs of positions in the CSC value array of J.
    std::vector<uint32_t> aent;                            // indices into ent_* of the entries A has
    std::vector<uint32_t> aprod_ptr;                       // aent.size() + 1
    std::vector<uint32_t> aprod_a, aprod_b;
};

struct DeviceCopy;  // defined in device.h

}  // namespace ezs

struct ezpz_structure {
    uint32_t n_cons = 0, n = 0, m = 0;
    std::vector<ezpz_constraint_t> cons;
    std::vector<uint32_t> cons_row0;  // n_cons + 1
    // J pattern, both orientations, and the permutations between their value orders
    std::vector<uint32_t> csc_col_ptr, csc_row_idx, csr_row_ptr, csr_col_idx, csr_to_csc, csc_to_csr;
    // lower(A) and L patterns, CSC with the diagonal first in every column
    std::vector<uint32_t> a_col_ptr, a_row_idx, l_col_ptr, l_row_idx;
    // connected components of the graph of A: comp_of[var]
    std::vector<uint32_t> comp_of;
    uint32_t n_components = 0;
    uint32_t max_component = 0;  // vars in the largest component
    std::vector<ezs::DevCons> dev_cons;
    uint32_t n_side = 0;
    bool all_weights_one = true;
    ezs::SmallProgram small;
    ezs::LargeProgram large;
    bool have_l_pattern = false;  // false when the symbolic factorisation was skipped (very large systems)
    // device copies, one per CUDA device ordinal, created lazily
    std::mutex dev_mutex;
    std::vector<ezs::DeviceCopy*> dev;
};

namespace ezs {
// Implemented in device.cu; called by ezpz_b200_structure_destroy.
void release_device_copies(ezpz_structure* s);
// sparse_direct.cpp: ordering, symbolic factorisation and level schedule of the large-system direct solve.
void build_sparse_direct(ezpz_structure& S);
}  // namespace ezs
