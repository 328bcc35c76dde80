/* ezpz_b200.h — C ABI of the B200-native ezpz solve path (libezpz_b200.so).
 *
 * This is the drop-in boundary for the hot path of KittyCAD/ezpz 0.2.27: the two calls
 * `Model::new(...)` and `model.solve_levenberg_marquardt(...)` inside `solve_inner`
 * (ezpz/src/lib.rs:279-292), plus the post-solve unsatisfied check (lib.rs:305-327) and the
 * optional freedom analysis (lib.rs:328, solver/find_dof.rs:15-104).  The reference has no FFI of
 * its own; these entry points are what a Rust `extern "C"` block added to
 * ezpz/src/lib.rs would bind (INTEGRATION.md shows the shim).
 *
 * Rules: plain pointers and sizes only; all buffers caller-owned; every function returns an
 * int32 status (EZPZ_OK == 0), never aborts, never throws across the boundary (the fuzz target's
 * contract, fuzz/fuzz_targets/fuzz_target_1.rs:7-23).  There is no CPU fallback: functions that
 * need the GPU return EZPZ_ERR_NO_DEVICE / EZPZ_ERR_CUDA when it is missing.
 */
#ifndef EZPZ_B200_H
#define EZPZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EZPZ_B200_ABI_VERSION 4

/* ---------------------------------------------------------------------------------------------
 * Constraint record.  One 64-byte record holds any of the 25 variants of `enum Constraint`
 * (ezpz/src/constraints.rs:37-93).  `kind` follows the enum's declaration order.
 *
 * `ids[]` lists the variable ids of the variant's datums in field order, each datum in its
 * `all_variables()` order (datatypes/inputs.rs:39,104,138,160,183): point = x,y; line = p0.x,p0.y,
 * p1.x,p1.y; circle = cx,cy,r; arc = start.x,start.y,end.x,end.y,center.x,center.y.
 *
 *   kind                              ids[]                                   p0        p1   flags
 *   0  LINE_TANGENT_TO_CIRCLE         line(4) circle(3)                       -         -    line side
 *   1  CIRCLE_TANGENT_TO_CIRCLE       circle a(3) circle b(3)                 -         -    circle side
 *   2  DISTANCE                       p0(2) p1(2)                             distance
 *   3  DISTANCE_VAR                   p(2) q(2) d(1)
 *   4  VERTICAL_DISTANCE              p0(2) p1(2)                             distance
 *   5  HORIZONTAL_DISTANCE            p0(2) p1(2)                             distance
 *   6  VERTICAL                       line(4)
 *   7  HORIZONTAL                     line(4)
 *   8  LINES_AT_ANGLE                 line0(4) line1(4)                       cos       sin  angle kind
 *   9  FIXED                          id                                      value
 *   10 SCALAR_EQUAL                   a b
 *   11 POINTS_COINCIDENT              p0(2) p1(2)
 *   12 CIRCLE_RADIUS                  circle(3)                               radius
 *   13 LINES_EQUAL_LENGTH             line0(4) line1(4)
 *   14 ARC_RADIUS                     arc(6)                                  radius
 *   15 ARC                            arc(6)
 *   16 MIDPOINT                       line(4) point(2)
 *   17 POINT_LINE_DISTANCE            point(2) line(4)                        distance
 *   18 VERTICAL_POINT_LINE_DISTANCE   point(2) line(4)                        distance
 *   19 HORIZONTAL_POINT_LINE_DISTANCE point(2) line(4)                        distance
 *   20 SYMMETRIC                      line(4) a(2) b(2)
 *   21 POINT_ARC_COINCIDENT           arc(6) point(2)
 *   22 ARC_LENGTH                     arc(6)                                  length
 *   23 ARC_ANGLE                      arc(6)                                  cos       sin
 *   24 POINTS_AT_ANGLE                p0(2) p1(2) p2(2)                       cos       sin  angle kind
 *
 * Angles: the binding computes (cos, sin) once per constraint exactly as
 * `rotation_for_angle_kind` does (constraints.rs:2641-2647): Parallel = (1, 0), Perpendicular =
 * (0, 1), Other(a) = libm::sincos(a.to_radians()).  `ezpz_b200_angle_sincos` is provided so a
 * binding without libm gets the same bits.
 * `weight` multiplies the residual and Jacobian rows (solver.rs:353,403); 1.0 by default.
 * Unused ids must be 0.
 */
typedef enum ezpz_kind {
    EZPZ_K_LINE_TANGENT_TO_CIRCLE = 0,
    EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE = 1,
    EZPZ_K_DISTANCE = 2,
    EZPZ_K_DISTANCE_VAR = 3,
    EZPZ_K_VERTICAL_DISTANCE = 4,
    EZPZ_K_HORIZONTAL_DISTANCE = 5,
    EZPZ_K_VERTICAL = 6,
    EZPZ_K_HORIZONTAL = 7,
    EZPZ_K_LINES_AT_ANGLE = 8,
    EZPZ_K_FIXED = 9,
    EZPZ_K_SCALAR_EQUAL = 10,
    EZPZ_K_POINTS_COINCIDENT = 11,
    EZPZ_K_CIRCLE_RADIUS = 12,
    EZPZ_K_LINES_EQUAL_LENGTH = 13,
    EZPZ_K_ARC_RADIUS = 14,
    EZPZ_K_ARC = 15,
    EZPZ_K_MIDPOINT = 16,
    EZPZ_K_POINT_LINE_DISTANCE = 17,
    EZPZ_K_VERTICAL_POINT_LINE_DISTANCE = 18,
    EZPZ_K_HORIZONTAL_POINT_LINE_DISTANCE = 19,
    EZPZ_K_SYMMETRIC = 20,
    EZPZ_K_POINT_ARC_COINCIDENT = 21,
    EZPZ_K_ARC_LENGTH = 22,
    EZPZ_K_ARC_ANGLE = 23,
    EZPZ_K_POINTS_AT_ANGLE = 24,
    EZPZ_K_COUNT = 25
} ezpz_kind_t;

/* `flags` values.  LineSide / CircleSide (constraints.rs:109-129): UNDEFINED is resolved per
 * problem from that problem's initial guesses, as `set_from_initial_values` does
 * (constraints.rs:146-193). */
#define EZPZ_SIDE_UNDEFINED 0u
#define EZPZ_LINE_SIDE_LEFT 1u
#define EZPZ_LINE_SIDE_RIGHT 2u
#define EZPZ_CIRCLE_SIDE_EXTERIOR 1u
#define EZPZ_CIRCLE_SIDE_INTERIOR 2u
/* AngleKind (datatypes.rs:9-16); informational, the arithmetic uses p0/p1. */
#define EZPZ_ANGLE_PARALLEL 0u
#define EZPZ_ANGLE_PERPENDICULAR 1u
#define EZPZ_ANGLE_OTHER 2u

typedef struct ezpz_constraint {
    uint32_t kind;
    uint32_t flags;
    uint32_t ids[8];
    double p0;
    double p1;
    double weight;
} ezpz_constraint_t; /* sizeof == 64 */

/* `Config` (solver.rs:31-81).  Defaults: 35, 1e-8, 1e-12, 1e-9. */
typedef struct ezpz_config {
    uint64_t max_iterations;
    double residual_tolerance;
    double step_tolerance;
    double initial_lambda;
} ezpz_config_t;

/* Status codes.  1..8 mirror `NonLinearSystemError` (error.rs:35-86); 20.. mirror `TextualError`
 * (error.rs:11-32). */
typedef enum ezpz_status {
    EZPZ_OK = 0,
    EZPZ_ERR_NOT_FOUND = 1,
    EZPZ_ERR_WRONG_NUMBER_GUESSES = 2,
    EZPZ_ERR_MISSING_GUESS = 3,
    EZPZ_ERR_MATRIX = 4,        /* FaerMatrix: a column index is outside 0..n_vars */
    EZPZ_ERR_FAER = 5,          /* kept for completeness; not produced */
    EZPZ_ERR_SOLVE = 6,         /* FaerSolve: non-numeric factorisation failure */
    EZPZ_ERR_SVD = 7,           /* kept for completeness; not produced */
    EZPZ_ERR_EMPTY_SYSTEM = 8,  /* EmptySystemNotAllowed */
    EZPZ_ERR_INVALID_ARGUMENT = 10,
    EZPZ_ERR_NO_DEVICE = 11,
    EZPZ_ERR_CUDA = 12,
    EZPZ_ERR_UNSUPPORTED = 13,
    EZPZ_ERR_TOO_LARGE = 14,
    EZPZ_ERR_PARSE = 20,
    EZPZ_ERR_TEXT_MISSING_GUESS = 21,
    EZPZ_ERR_TEXT_UNUSED_GUESSES = 22,
    EZPZ_ERR_TEXT_UNDEFINED_POINT = 23
} ezpz_status_t;

/* Filled on error when non-NULL. */
typedef struct ezpz_error_detail {
    uint64_t constraint_id; /* MissingGuess.constraint_id */
    uint32_t variable;      /* MissingGuess.variable / NotFound */
    uint32_t reserved;
    uint64_t a, b;          /* WrongNumberGuesses{labels=a, guesses=b}; CUDA error code in a */
    char message[192];
} ezpz_error_detail_t;

/* Per-problem status bits written by the solve kernels. */
#define EZPZ_ST_CONVERGED 0x01u     /* SuccessfulSolve.converged (newton.rs:19-24) */
#define EZPZ_ST_UNSATISFIED 0x02u   /* at least one constraint failed lib.rs:358-370 */
#define EZPZ_ST_DEGENERATE 0x04u    /* at least one Warning::Degenerate was raised */
#define EZPZ_ST_SOLVE_ERROR 0x08u   /* non-numeric factorisation failure (FaerSolve) */

/* ---------------------------------------------------------------------------------------------
 * Structure: the result of analysing one sketch topology once (replaces the per-solve work of
 * Model::new, solver.rs:192-300): validation, the deduplicated sorted sparsity pattern of J in
 * CSC and CSR, per-partial scatter slots, the pattern of A = JtJ + lambda*I, its symbolic Cholesky
 * and the operation tapes the device executes.  Immutable after creation; may be shared by
 * threads and contexts (device tables are created per device on first use, the work buffers of the
 * single-large-system path per context, and both live until the structure is destroyed).
 */
typedef struct ezpz_structure ezpz_structure_t;

/* `var_ids`: the ids of the initial guesses in guess order (lib.rs:275), or NULL for 0..n_vars-1.
 * Validation order and errors follow validate_variables (solver.rs:142-189) then the CSC
 * construction (solver.rs:256): MISSING_GUESS, then MATRIX when an id >= n_vars.
 * EMPTY_SYSTEM is NOT raised here (the reference raises it inside the loop, newton.rs:54). */
int32_t ezpz_b200_structure_create(const ezpz_constraint_t* cons, uint32_t n_cons,
                                   const uint32_t* var_ids, uint32_t n_vars,
                                   ezpz_structure_t** out, ezpz_error_detail_t* detail);
void ezpz_b200_structure_destroy(ezpz_structure_t* s);
/* A structure for the constraints of `base` followed by `extra` over the same variables: what a caller that adds a
 * constraint to a solved sketch and solves again needs (the trim workflows of the reference's tests, tests.rs:748-897; the
 * reference re-runs Model::new on the longer list, lib.rs:279).  Patterns, scatter slots and tapes are those
 * ezpz_b200_structure_create gives for the concatenated list; a system of the sparse-direct path KEEPS base's elimination
 * order instead of dissecting the graph again (any order of the variables is valid; the arithmetic is the oracle's in that
 * order, as ezpz_b200_structure_ordering reports it).  Ids are validated against the guess ids `base` was created with.
 * `base` is not modified and may be destroyed afterwards. */
int32_t ezpz_b200_structure_extend(const ezpz_structure_t* base, const ezpz_constraint_t* extra, uint32_t n_extra,
                                   ezpz_structure_t** out, ezpz_error_detail_t* detail);

/* Sizes: rows m (= sum residual_dim, constraints.rs:954-993), vars n, nnz(J), nnz(lower A),
 * nnz(L), number of connected components of A. */
int32_t ezpz_b200_structure_dims(const ezpz_structure_t* s, uint32_t* m, uint32_t* n,
                                 uint64_t* nnz_j, uint64_t* nnz_a, uint64_t* nnz_l,
                                 uint32_t* n_components);

/* The parity artefact: the sorted, deduplicated pattern of J in both orientations.  Pointers stay
 * valid until the structure is destroyed.  CSC is what faer builds (solver.rs:256). */
int32_t ezpz_b200_structure_pattern(const ezpz_structure_t* s, const uint32_t** csc_col_ptr,
                                    const uint32_t** csc_row_idx, const uint32_t** csr_row_ptr,
                                    const uint32_t** csr_col_idx);
/* Pattern of the lower triangle of A (CSC, sorted) and of its Cholesky factor L (natural order). */
int32_t ezpz_b200_structure_pattern_a(const ezpz_structure_t* s, const uint32_t** a_col_ptr,
                                      const uint32_t** a_row_idx, const uint32_t** l_col_ptr,
                                      const uint32_t** l_row_idx);
/* Introspection of the batched kernel's programme (tests): the tables lm_small_kernel stages in shared memory for
 * `roles` cooperating warps per 32 problems and `stride` problems per CTA.  Layout: ezpz_b200/csrc/structure.h (RoleBlob).
 * Copies at most `cap` 32-bit words to `words`; *n_words is the full length, *cons_word the offset of the constraint
 * array, *dims = {W (doubles of state per problem), n_cons, barriers inside one tape execution, modelled cost of the
 * busiest role}.
 * EZPZ_ERR_UNSUPPORTED when the structure does not run on the thread-per-problem kernel. */
int32_t ezpz_b200_structure_role_program(const ezpz_structure_t* s, uint32_t roles, uint32_t stride,
                                         uint32_t* words, uint64_t cap, uint64_t* n_words,
                                         uint32_t* cons_word, uint32_t* dims);
/* The launch shape the batched kernel would take for `batch` problems of this structure on a device with `sm_count` SMs
 * and `smem_per_block` bytes of opt-in shared memory per block: cooperating warps per 32 problems and problems per CTA.
 * Host arithmetic only (tests, tuning). */
int32_t ezpz_b200_structure_batch_shape(const ezpz_structure_t* s, uint64_t batch, uint32_t sm_count,
                                        uint64_t smem_per_block, uint32_t* roles, uint32_t* problems_per_cta);
/* First row of each constraint in J (n_cons + 1 entries). */
int32_t ezpz_b200_structure_rows(const ezpz_structure_t* s, const uint32_t** cons_row0);
/* A 64-bit hash of everything the host analysis produced (patterns, scatter slots, operation tapes, elimination order,
 * supernode schedule, product lists): structures with equal fingerprints drive the device through the same arithmetic.
 * Tests hold the threaded analysis phases and ezpz_b200_structure_extend to it.  0 for NULL. */
uint64_t ezpz_b200_structure_fingerprint(const ezpz_structure_t* s);

/* ---------------------------------------------------------------------------------------------
 * Context: one CUDA device + stream + workspace pool.  One per thread that solves.
 */
typedef struct ezpz_context ezpz_context_t;
int32_t ezpz_b200_context_create(int32_t device, ezpz_context_t** out, ezpz_error_detail_t* detail);
void ezpz_b200_context_destroy(ezpz_context_t* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t ezpz_b200_context_launches(const ezpz_context_t* ctx);
int32_t ezpz_b200_context_synchronize(ezpz_context_t* ctx);
/* Drops the analysed structures ezpz_b200_solve keeps for this context (its topology cache: the last 8 distinct
 * constraint lists, at most 2^19 variables in total) together with their device tables. */
void ezpz_b200_context_clear_cache(ezpz_context_t* ctx);

/* ---------------------------------------------------------------------------------------------
 * Batched solve: `batch` independent problems sharing one structure, differing in their initial
 * guesses (and optionally in the scalar target p0 of each constraint).  Replaces `batch` calls of
 * solve_inner (lib.rs:265-356) at one priority level.  The whole Levenberg–Marquardt loop
 * (newton.rs:29-145) runs on the device; there is no host round trip per iteration.
 * Structures of up to 880 values per problem run one THREAD per problem (state in shared memory); larger
 * ones run one CTA per problem through the persistent kernel of the single-system path — every problem
 * then gets exactly what ezpz_b200_solve_one would give it.  `params` is only available on the
 * thread-per-problem kernel (EZPZ_ERR_UNSUPPORTED otherwise).
 *
 * Host-buffer form: copies guesses in and results out inside the call (pinned staging).
 *   guesses        [batch * n]       row-major, one row per problem, in id order
 *   params         [batch * n_cons]  optional (NULL): per-problem override of each constraint's p0
 *   final_values   [batch * n]
 *   iterations     [batch]           SuccessfulSolve.iterations
 *   status         [batch]           EZPZ_ST_* bits
 *   unsat_mask     [batch * ceil(n_cons/32)] u32 words, optional: bit c set = constraint c
 *                                    unsatisfied (lib.rs:305-327)
 *   degen_count    [batch * n_cons]  optional: how many Warning::Degenerate were raised per
 *                                    constraint (solver.rs:340-346,385-391)
 *   jacobian       [batch * nnz]     optional: the cached Jacobian values (CSC order) at the last
 *                                    accepted point, i.e. what freedom_analysis reads
 *                                    (find_dof.rs:16-18)
 *   under_mask     [batch * ceil(n/32)] u32 words, optional: bit j set = variable j underconstrained — the freedom analysis
 *                                    of solve_analysis (lib.rs:134-144, find_dof.rs:15-104) fused into the call: it reads the
 *                                    Jacobians the solve kernel left on the device, no host round trip in between
 */
typedef struct ezpz_batch_io {
    const double* guesses;
    const double* params;
    double* final_values;
    uint32_t* iterations;
    uint8_t* status;
    uint32_t* unsat_mask;
    uint32_t* degen_count;
    double* jacobian;
    uint32_t* under_mask;
} ezpz_batch_io_t;

int32_t ezpz_b200_solve_batch(ezpz_context_t* ctx, const ezpz_structure_t* s,
                              const ezpz_config_t* config, uint64_t batch,
                              const ezpz_batch_io_t* io, ezpz_error_detail_t* detail);

/* Device-buffer form: every pointer in `io` is a device pointer on the context's device; the
 * kernels are enqueued on `cuda_stream` (a cudaStream_t; NULL = the context's own stream) and the
 * call returns without synchronising. */
int32_t ezpz_b200_solve_batch_device(ezpz_context_t* ctx, const ezpz_structure_t* s,
                                     const ezpz_config_t* config, uint64_t batch,
                                     const ezpz_batch_io_t* io, void* cuda_stream,
                                     ezpz_error_detail_t* detail);

/* ---------------------------------------------------------------------------------------------
 * One call, several GPUs.  The reference's seam is one synchronous in-process call (lib.rs:80-87), so the sharding of a
 * batch over the GPUs of a box happens inside the library: a multi-context owns one worker thread + context per device;
 * ezpz_b200_solve_batch_multi cuts the batch into contiguous shards of whole 32-problem groups, one per device, runs
 * ezpz_b200_solve_batch on each shard of the CALLER's host buffers concurrently and returns when all are done.  Problems
 * are independent: no exchange step, no collective, and a problem's result does not depend on the number of devices.
 *   devices   [n_devices] CUDA ordinals, or NULL: the first n_devices visible devices (n_devices <= 0: all of them)
 * Buffers as in ezpz_batch_io_t (host pointers).  Page-locked buffers (ezpz_b200_host_alloc / ezpz_b200_host_register,
 * cudaHostAlloc, ...) are read and written by the kernels directly across PCIe; pageable buffers work through staged copies.
 */
typedef struct ezpz_multi ezpz_multi_t;
int32_t ezpz_b200_multi_create(const int32_t* devices, int32_t n_devices, ezpz_multi_t** out,
                               ezpz_error_detail_t* detail);
void ezpz_b200_multi_destroy(ezpz_multi_t* mg);
int32_t ezpz_b200_multi_device_count(const ezpz_multi_t* mg);
/* The context of device `index` of the multi-context (for single-device calls on the same workspaces); NULL if out of range. */
ezpz_context_t* ezpz_b200_multi_context(ezpz_multi_t* mg, int32_t index);
/* Kernels launched so far by all devices of the multi-context. */
uint64_t ezpz_b200_multi_launches(const ezpz_multi_t* mg);
int32_t ezpz_b200_solve_batch_multi(ezpz_multi_t* mg, const ezpz_structure_t* s, const ezpz_config_t* config,
                                    uint64_t batch, const ezpz_batch_io_t* io, ezpz_error_detail_t* detail);

/* Several batches in one call: the structure-homogeneous sub-batches of a MIXED workload (problems of different topologies
 * cannot share a launch).  Jobs of 65,536 problems and more are cut over all workers like a single batch call, the others go
 * whole to the least loaded worker (largest first); each worker runs its list in order.  A multi-context may hold several
 * workers per device (list the device more than once in ezpz_b200_multi_create): their streams overlap on the GPU, which keeps
 * it busy when the sub-batches are small.  `status` of every job is set; the call returns the first failure. */
typedef struct ezpz_batch_job {
    const ezpz_structure_t* structure;
    uint64_t batch;
    ezpz_batch_io_t io;  /* host pointers, as for ezpz_b200_solve_batch */
    int32_t status;
    int32_t reserved;
} ezpz_batch_job_t;
int32_t ezpz_b200_solve_jobs_multi(ezpz_multi_t* mg, const ezpz_config_t* config, ezpz_batch_job_t* jobs, uint32_t n_jobs,
                                   ezpz_error_detail_t* detail);

/* Page-locked host memory every device can address (what a Rust caller would wrap its Vec<f64> buffers with): register an
 * existing allocation, or allocate one. */
int32_t ezpz_b200_host_register(void* ptr, uint64_t bytes);
int32_t ezpz_b200_host_unregister(void* ptr);
int32_t ezpz_b200_host_alloc(uint64_t bytes, void** out);
void ezpz_b200_host_free(void* ptr);

/* Contiguous shard of a batch for rank `rank` of `world` (one process per GPU; no collective). */
void ezpz_b200_shard_range(uint64_t batch, uint32_t rank, uint32_t world, uint64_t* begin,
                           uint64_t* end);

/* ---------------------------------------------------------------------------------------------
 * Single system of any size (replaces one solve_inner call).  Small systems run through the batch
 * kernel with batch == 1; large systems through the sparse path (assembly kernel, SpMV, device-side
 * LM control).  `path_used`: 0 = batched-small kernel, 1 = sparse block-Cholesky path, 2 = sparse
 * PCG path.
 */
typedef struct ezpz_one_io {
    const double* guesses; /* [n] */
    double* final_values;  /* [n] */
    uint32_t* iterations;
    uint8_t* status;
    uint32_t* unsat_mask;  /* optional [ceil(n_cons/32)] */
    uint32_t* degen_count; /* optional [n_cons] */
    double* jacobian;      /* optional [nnz], CSC order, values at the last accepted point */
    int32_t* path_used;    /* optional */
    uint32_t* lin_iters;   /* optional: total inner (PCG) iterations, 0 for direct paths */
} ezpz_one_io_t;

int32_t ezpz_b200_solve_one(ezpz_context_t* ctx, const ezpz_structure_t* s,
                            const ezpz_config_t* config, const ezpz_one_io_t* io,
                            ezpz_error_detail_t* detail);

/* How a structure that takes the large path solves its damped step (replaces what faer's
 * SymbolicLlt::try_new decides inside precompute_symbolic_cholesky, solver.rs:289-300):
 *   *path        0 = batched-small kernel (no large programme), 1 = sparse direct, 2 = PCG
 *   *elim_order  [n] elimination position -> variable (NULL unless path 1); natural order 0..n-1 or
 *                nested dissection (*nested != 0)
 *   *n_levels    height of the supernode tree = number of parallel factorisation stages
 *   *nnz_l       doubles of panel storage of the factor (dense supernode panels, explicit zeros included)
 *   *sum_chunk   rows per chunk of the sum-of-squares fold on this structure (0 = one sequential fold)
 * The arithmetic is the oracle's applied to P A Pt; parity tests hand `elim_order` and `sum_chunk` to the oracle. */
int32_t ezpz_b200_structure_ordering(const ezpz_structure_t* s, int32_t* path, const uint32_t** elim_order,
                                     int32_t* nested, uint32_t* n_levels, uint64_t* nnz_l,
                                     uint32_t* sum_chunk);

/* Stand-alone timing of the large-system kernels on a structure's device buffers (CUDA events, `reps`
 * launches): which = 0 fused assembly (residual + Jacobian + scatter into CSC order), 1 SpMV y = J p (CSR),
 * 2 SpMV z = Jt q (CSC), 3 the assembly that also writes the CSR-ordered copy (PCG path).  Returns mean microseconds per launch and the algorithmic bytes one launch moves
 * (SURVEY.md §8d).  Only for structures that take the large path. */
int32_t ezpz_b200_large_bench(ezpz_context_t* ctx, const ezpz_structure_t* s, const double* x,
                              int32_t which, int32_t reps, double* mean_us, double* algorithmic_bytes,
                              ezpz_error_detail_t* detail);

/* ---------------------------------------------------------------------------------------------
 * Parity / debugging entry (the reference's `dbg-jac` feature, solver.rs:370-439): one evaluation
 * of the residual vector and the Jacobian values at `x` through the device assembly kernel.
 *   r        [m]      weighted residuals (Model::residual, solver.rs:318-356)
 *   jac_csc  [nnz]    Jacobian values in CSC order (Model::refresh_jacobian, solver.rs:359-440)
 *   jac_csr  [nnz]    the same values in CSR order (optional)
 *   degen    [n_cons] optional: bit0 residual raised Degenerate, bit1 Jacobian raised Degenerate
 * Undefined sides are resolved from `x` itself.
 */
int32_t ezpz_b200_eval(ezpz_context_t* ctx, const ezpz_structure_t* s, const double* x, double* r,
                       double* jac_csc, double* jac_csr, uint8_t* degen,
                       ezpz_error_detail_t* detail);

/* Freedom analysis (find_dof.rs:15-104) for `batch` problems of one structure: dense column-pivoted
 * QR of each problem's Jacobian (`jacobian`: [batch * nnz] values in CSC order, as exported by the
 * solve calls), rank by |R_ii| > 1e-8 * max|R_ii|, nullspace participation per variable; variable j is
 * underconstrained iff its participation exceeds (1e-3 * max participation)^2.  Bit j of
 * under_mask[problem * ceil(n/32) + j/32] is set for underconstrained variables.  Runs on the device for any
 * number of variables (dense, like the reference's: O(m n) memory, O(m n^2) work): a warp or CTA per problem for
 * batches of small sketches, the whole GPU on one system after the other for large ones.  Host pointers. */
int32_t ezpz_b200_freedom_analysis(ezpz_context_t* ctx, const ezpz_structure_t* s, uint64_t batch,
                                   const double* jacobian, uint32_t* under_mask,
                                   ezpz_error_detail_t* detail);
/* Device-pointer form: `jacobian` and `under_mask` live on the context's device, the kernels are enqueued on
 * `cuda_stream` (NULL = the context's own stream) and the call returns without synchronising. */
int32_t ezpz_b200_freedom_analysis_device(ezpz_context_t* ctx, const ezpz_structure_t* s, uint64_t batch,
                                          const double* jacobian, uint32_t* under_mask, void* cuda_stream,
                                          ezpz_error_detail_t* detail);

/* ---------------------------------------------------------------------------------------------
 * ezpz::solve / ezpz::solve_analysis (lib.rs:80-144) with flat arguments: the priority loop
 * (lib.rs:148-263), the lint (warnings.rs:34-59), one structure analysis + one device solve per
 * priority level, the unsatisfied list in ORIGINAL request indices, and (analysis != 0) the
 * underconstrained variable list.  This is what the Rust shim's `solve` calls.
 *   priorities  [n_cons] or NULL (all 0)
 *   angles_deg  [n_cons] or NULL: for LinesAtAngle(Other(a)) the angle in degrees, NaN otherwise
 *               (only used by the lint)
 *   var_ids     [n_vars] or NULL (ids 0..n_vars-1)
 * On error the status is returned and `outcome` carries what FailureOutcome does: warnings, num_vars,
 * num_eqs (solve_outcome.rs:136-181).
 */
typedef struct ezpz_warning {
    int64_t about_constraint; /* -1 = None */
    uint32_t kind;            /* 0 Degenerate, 1 ShouldBeParallel, 2 ShouldBePerpendicular */
    uint32_t count;           /* Degenerate: how many times it was raised */
    double angle_deg;
} ezpz_warning_t;

typedef struct ezpz_outcome {
    double* final_values;       /* [n_vars] */
    uint64_t* unsatisfied;      /* [n_cons] */
    uint32_t* underconstrained; /* [n_vars], may be NULL when analysis == 0 */
    ezpz_warning_t* warnings;   /* [warnings_cap], may be NULL */
    uint32_t warnings_cap;
    uint32_t n_warnings;        /* produced (may exceed warnings_cap; only the first cap are stored) */
    uint32_t n_unsatisfied;
    uint32_t n_underconstrained;
    uint64_t iterations;
    uint32_t converged;
    uint32_t priority_solved;
    uint32_t num_vars;
    uint32_t num_eqs;
    int32_t path_used;
    uint32_t reserved;
} ezpz_outcome_t;

int32_t ezpz_b200_solve(ezpz_context_t* ctx, const ezpz_constraint_t* cons, const uint32_t* priorities,
                        const double* angles_deg, uint32_t n_cons, const uint32_t* var_ids,
                        const double* guesses, uint32_t n_vars, const ezpz_config_t* config,
                        int32_t analysis, ezpz_outcome_t* outcome, ezpz_error_detail_t* detail);

/* ---------------------------------------------------------------------------------------------
 * The priority loop of ezpz::solve (lib.rs:199-246) for a BATCH of problems of one topology: the
 * constraints, their priorities and weights are shared, guesses (and optionally the p0 targets) differ
 * per problem.  For every distinct priority level, ascending: ONE structure analysis of the constraints
 * with priority <= level and ONE batched solve of all still-undecided problems from their ORIGINAL
 * guesses.  A problem keeps the result of the last level that left no constraint unsatisfied; the first
 * level with unsatisfied constraints stops it (its own result is kept only if it is the first level) —
 * exactly what the per-problem loop does.
 *   priorities       [n_cons] (NULL = all 0)
 *   guesses          [batch * n_vars]
 *   params           [batch * n_cons] per-problem p0 overrides in ORIGINAL request order, or NULL
 *   final_values     [batch * n_vars]
 *   iterations       [batch]   LM iterations of the level whose result is kept
 *   status           [batch]   EZPZ_ST_* of that level
 *   priority_solved  [batch]   highest priority among the constraints of that level
 *   unsat_mask       [batch * ceil(n_cons / 32)] bit c = ORIGINAL request c unsatisfied (optional)
 */
int32_t ezpz_b200_solve_batch_priorities(ezpz_context_t* ctx, const ezpz_constraint_t* cons,
                                         const uint32_t* priorities, uint32_t n_cons, uint32_t n_vars,
                                         const ezpz_config_t* config, uint64_t batch, const double* guesses,
                                         const double* params, double* final_values, uint32_t* iterations,
                                         uint8_t* status, uint32_t* priority_solved, uint32_t* unsat_mask,
                                         ezpz_error_detail_t* detail);

/* ---------------------------------------------------------------------------------------------
 * Scalar helpers with the bits of libm 0.2.16 (see ezpz_b200/csrc/dmath.cuh). */
void ezpz_b200_angle_sincos(double radians, double* sin_out, double* cos_out);
double ezpz_b200_hypot(double x, double y);
void ezpz_b200_config_default(ezpz_config_t* cfg);
uint32_t ezpz_b200_abi_version(void);
const char* ezpz_b200_status_name(int32_t status);

/* ---------------------------------------------------------------------------------------------
 * Text problem format (ezpz/src/textual.rs, textual/parser.rs:29-555, textual/executor.rs:40-445):
 * "# constraints ... # guesses ..." -> constraint records + guesses.  Host only.
 */
typedef struct ezpz_problem ezpz_problem_t;

/* Problem::from_str (textual.rs:43-49).  On EZPZ_ERR_PARSE `detail->message` says where. */
int32_t ezpz_b200_problem_parse(const char* text, uint64_t len, ezpz_problem_t** out,
                                ezpz_error_detail_t* detail);
void ezpz_b200_problem_destroy(ezpz_problem_t* p);
/* Problem::to_constraint_system (executor.rs:40-445).  Pointers are owned by the problem. */
int32_t ezpz_b200_problem_system(ezpz_problem_t* p, const ezpz_constraint_t** cons,
                                 uint32_t* n_cons, const double** guesses, uint32_t* n_vars,
                                 ezpz_error_detail_t* detail);
/* Declared geometry, in declaration order; kind 0 = points, 1 = circles, 2 = arcs.  Variable ids:
 * points 2i,2i+1; then circles cx,cy,r; then arcs a.x,a.y,b.x,b.y,c.x,c.y (executor.rs:525-566). */
uint32_t ezpz_b200_problem_count(const ezpz_problem_t* p, int32_t kind);
const char* ezpz_b200_problem_label(const ezpz_problem_t* p, int32_t kind, uint32_t index);
/* For each constraint, the angle as written in the text (degrees) when it is LinesAtAngle(Other),
 * NaN otherwise; used by the lint (warnings.rs:34-59). */
int32_t ezpz_b200_problem_angles_deg(const ezpz_problem_t* p, const double** angles_deg);

#ifdef __cplusplus
}
#endif
#endif /* EZPZ_B200_H */
