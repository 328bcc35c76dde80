//! `ezpz/src/gpu.rs` — the Rust side of the C ABI of libezpz_b200.so (include/ezpz_b200.h, ABI version 3).
//!
//! Every declaration names what it replaces in the reference.  No Rust toolchain exists in the image this was written in:
//! the file ships as source, the same entry points are exercised from C (tests/c_abi/solve_tiny.c), C++ (host_api.cpp,
//! cli.cpp) and Python (ezpz_b200/native.py), and tests/test_host.py checks that the functions declared here are exactly
//! the header's.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_void};

pub const EZPZ_B200_ABI_VERSION: u32 = 4;

/// One `Constraint` in the flat 64-byte form (constraints.rs:37-93; layout table in the header).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct EzpzConstraint {
    pub kind: u32,
    pub flags: u32,
    pub ids: [u32; 8],
    pub p0: f64,
    pub p1: f64,
    pub weight: f64,
}

/// `Config` (solver.rs:31-81).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct EzpzConfig {
    pub max_iterations: u64,
    pub residual_tolerance: f64,
    pub step_tolerance: f64,
    pub initial_lambda: f64,
}

#[repr(C)]
pub struct EzpzErrorDetail {
    pub constraint_id: u64,
    pub variable: u32,
    pub reserved: u32,
    pub a: u64,
    pub b: u64,
    pub message: [u8; 192],
}

#[repr(C)]
pub struct EzpzBatchIo {
    pub guesses: *const f64,
    pub params: *const f64,
    pub final_values: *mut f64,
    pub iterations: *mut u32,
    pub status: *mut u8,
    pub unsat_mask: *mut u32,
    pub degen_count: *mut u32,
    pub jacobian: *mut f64,
    pub under_mask: *mut u32,
}

#[repr(C)]
pub struct EzpzOneIo {
    pub guesses: *const f64,
    pub final_values: *mut f64,
    pub iterations: *mut u32,
    pub status: *mut u8,
    pub unsat_mask: *mut u32,
    pub degen_count: *mut u32,
    pub jacobian: *mut f64,
    pub path_used: *mut i32,
    pub lin_iters: *mut u32,
}

#[repr(C)]
pub struct EzpzWarning {
    pub about_constraint: i64,
    pub kind: u32,
    pub count: u32,
    pub angle_deg: f64,
}

#[repr(C)]
pub struct EzpzOutcome {
    pub final_values: *mut f64,
    pub unsatisfied: *mut u64,
    pub underconstrained: *mut u32,
    pub warnings: *mut EzpzWarning,
    pub warnings_cap: u32,
    pub n_warnings: u32,
    pub n_unsatisfied: u32,
    pub n_underconstrained: u32,
    pub iterations: u64,
    pub converged: u32,
    pub priority_solved: u32,
    pub num_vars: u32,
    pub num_eqs: u32,
    pub path_used: i32,
    pub reserved: u32,
}

#[repr(C)]
pub struct EzpzBatchJob {
    pub structure: *const EzpzStructure,
    pub batch: u64,
    pub io: EzpzBatchIo,
    pub status: i32,
    pub reserved: i32,
}

pub enum EzpzStructure {}
pub enum EzpzContext {}
pub enum EzpzMulti {}
pub enum EzpzProblem {}

pub const EZPZ_ST_CONVERGED: u8 = 0x01;
pub const EZPZ_ST_UNSATISFIED: u8 = 0x02;
pub const EZPZ_ST_DEGENERATE: u8 = 0x04;
pub const EZPZ_ST_SOLVE_ERROR: u8 = 0x08;

extern "C" {
    // ---- structure: replaces Model::new (solver.rs:192-300), once per topology
    pub fn ezpz_b200_structure_create(cons: *const EzpzConstraint, n_cons: u32, var_ids: *const u32, n_vars: u32,
                                      out: *mut *mut EzpzStructure, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_structure_destroy(s: *mut EzpzStructure);
    pub fn ezpz_b200_structure_extend(base: *const EzpzStructure, extra: *const EzpzConstraint, n_extra: u32,
                                      out: *mut *mut EzpzStructure, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_structure_dims(s: *const EzpzStructure, m: *mut u32, n: *mut u32, nnz_j: *mut u64, nnz_a: *mut u64,
                                    nnz_l: *mut u64, n_components: *mut u32) -> i32;
    pub fn ezpz_b200_structure_pattern(s: *const EzpzStructure, csc_col_ptr: *mut *const u32, csc_row_idx: *mut *const u32,
                                       csr_row_ptr: *mut *const u32, csr_col_idx: *mut *const u32) -> i32;
    pub fn ezpz_b200_structure_pattern_a(s: *const EzpzStructure, a_col_ptr: *mut *const u32, a_row_idx: *mut *const u32,
                                         l_col_ptr: *mut *const u32, l_row_idx: *mut *const u32) -> i32;
    pub fn ezpz_b200_structure_role_program(s: *const EzpzStructure, roles: u32, stride: u32, words: *mut u32, cap: u64,
                                            n_words: *mut u64, cons_word: *mut u32, dims: *mut u32) -> i32;
    pub fn ezpz_b200_structure_batch_shape(s: *const EzpzStructure, batch: u64, sm_count: u32, smem_per_block: u64,
                                           roles: *mut u32, problems_per_cta: *mut u32) -> i32;
    pub fn ezpz_b200_structure_rows(s: *const EzpzStructure, cons_row0: *mut *const u32) -> i32;
    pub fn ezpz_b200_structure_fingerprint(s: *const EzpzStructure) -> u64;
    // what faer's SymbolicLlt::try_new decides (solver.rs:289-300), reported
    pub fn ezpz_b200_structure_ordering(s: *const EzpzStructure, path: *mut i32, elim_order: *mut *const u32, nested: *mut i32,
                                        n_levels: *mut u32, nnz_l: *mut u64, sum_chunk: *mut u32) -> i32;

    // ---- contexts
    pub fn ezpz_b200_context_create(device: i32, out: *mut *mut EzpzContext, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_context_destroy(ctx: *mut EzpzContext);
    pub fn ezpz_b200_context_launches(ctx: *const EzpzContext) -> u64;
    pub fn ezpz_b200_context_synchronize(ctx: *mut EzpzContext) -> i32;
    pub fn ezpz_b200_context_clear_cache(ctx: *mut EzpzContext);

    // ---- solves: replace model.solve_levenberg_marquardt + the unsatisfied check (lib.rs:292-327, newton.rs:29-145)
    pub fn ezpz_b200_solve_one(ctx: *mut EzpzContext, s: *const EzpzStructure, cfg: *const EzpzConfig, io: *const EzpzOneIo,
                               detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_solve_batch(ctx: *mut EzpzContext, s: *const EzpzStructure, cfg: *const EzpzConfig, batch: u64,
                                 io: *const EzpzBatchIo, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_solve_batch_device(ctx: *mut EzpzContext, s: *const EzpzStructure, cfg: *const EzpzConfig, batch: u64,
                                        io: *const EzpzBatchIo, cuda_stream: *mut c_void, detail: *mut EzpzErrorDetail) -> i32;
    // one call, every GPU of the box (north-star item 5)
    pub fn ezpz_b200_multi_create(devices: *const i32, n_devices: i32, out: *mut *mut EzpzMulti, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_multi_destroy(mg: *mut EzpzMulti);
    pub fn ezpz_b200_multi_device_count(mg: *const EzpzMulti) -> i32;
    pub fn ezpz_b200_multi_context(mg: *mut EzpzMulti, index: i32) -> *mut EzpzContext;
    pub fn ezpz_b200_multi_launches(mg: *const EzpzMulti) -> u64;
    pub fn ezpz_b200_solve_batch_multi(mg: *mut EzpzMulti, s: *const EzpzStructure, cfg: *const EzpzConfig, batch: u64,
                                       io: *const EzpzBatchIo, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_solve_jobs_multi(mg: *mut EzpzMulti, cfg: *const EzpzConfig, jobs: *mut EzpzBatchJob, n_jobs: u32,
                                      detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_host_register(ptr: *mut c_void, bytes: u64) -> i32;
    pub fn ezpz_b200_host_unregister(ptr: *mut c_void) -> i32;
    pub fn ezpz_b200_host_alloc(bytes: u64, out: *mut *mut c_void) -> i32;
    pub fn ezpz_b200_host_free(ptr: *mut c_void);
    pub fn ezpz_b200_shard_range(batch: u64, rank: u32, world: u32, begin: *mut u64, end: *mut u64);
    // the priority loop of solve_with_priority_inner (lib.rs:199-246) for a batch of problems of one topology
    pub fn ezpz_b200_solve_batch_priorities(ctx: *mut EzpzContext, cons: *const EzpzConstraint, priorities: *const u32, n_cons: u32,
                                            n_vars: u32, cfg: *const EzpzConfig, batch: u64, guesses: *const f64, params: *const f64,
                                            final_values: *mut f64, iterations: *mut u32, status: *mut u8,
                                            priority_solved: *mut u32, unsat_mask: *mut u32, detail: *mut EzpzErrorDetail) -> i32;
    // ezpz::solve / solve_analysis with flat arguments (lib.rs:80-144)
    pub fn ezpz_b200_solve(ctx: *mut EzpzContext, cons: *const EzpzConstraint, priorities: *const u32, angles_deg: *const f64,
                           n_cons: u32, var_ids: *const u32, guesses: *const f64, n_vars: u32, cfg: *const EzpzConfig,
                           analysis: i32, outcome: *mut EzpzOutcome, detail: *mut EzpzErrorDetail) -> i32;

    // ---- analysis / debugging
    // replaces FreedomAnalysis::analyze -> Model::freedom_analysis (analysis.rs:34-37, find_dof.rs:15-104)
    pub fn ezpz_b200_freedom_analysis(ctx: *mut EzpzContext, s: *const EzpzStructure, batch: u64, jacobian: *const f64,
                                      under_mask: *mut u32, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_freedom_analysis_device(ctx: *mut EzpzContext, s: *const EzpzStructure, batch: u64, jacobian: *const f64,
                                             under_mask: *mut u32, cuda_stream: *mut c_void, detail: *mut EzpzErrorDetail) -> i32;
    // the `dbg-jac` feature (solver.rs:370-439)
    pub fn ezpz_b200_eval(ctx: *mut EzpzContext, s: *const EzpzStructure, x: *const f64, r: *mut f64, jac_csc: *mut f64,
                          jac_csr: *mut f64, degen: *mut u8, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_large_bench(ctx: *mut EzpzContext, s: *const EzpzStructure, x: *const f64, which: i32, reps: i32,
                                 mean_us: *mut f64, algorithmic_bytes: *mut f64, detail: *mut EzpzErrorDetail) -> i32;

    // ---- scalar helpers with libm 0.2.16's bits
    pub fn ezpz_b200_angle_sincos(radians: f64, sin_out: *mut f64, cos_out: *mut f64);
    pub fn ezpz_b200_hypot(x: f64, y: f64) -> f64;
    pub fn ezpz_b200_config_default(cfg: *mut EzpzConfig);
    pub fn ezpz_b200_abi_version() -> u32;
    pub fn ezpz_b200_status_name(status: i32) -> *const c_char;

    // ---- text format (the crate keeps its own winnow parser; these exist for the C++ CLI twin and for cross-checks)
    pub fn ezpz_b200_problem_parse(text: *const c_char, len: u64, out: *mut *mut EzpzProblem, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_problem_destroy(p: *mut EzpzProblem);
    pub fn ezpz_b200_problem_system(p: *mut EzpzProblem, cons: *mut *const EzpzConstraint, n_cons: *mut u32,
                                    guesses: *mut *const f64, n_vars: *mut u32, detail: *mut EzpzErrorDetail) -> i32;
    pub fn ezpz_b200_problem_count(p: *const EzpzProblem, kind: i32) -> u32;
    pub fn ezpz_b200_problem_label(p: *const EzpzProblem, kind: i32, index: u32) -> *const c_char;
    pub fn ezpz_b200_problem_angles_deg(p: *const EzpzProblem, angles_deg: *mut *const f64) -> i32;
}

impl From<crate::Config> for EzpzConfig {
    fn from(c: crate::Config) -> Self {
        EzpzConfig {
            max_iterations: c.max_iterations as u64,
            residual_tolerance: c.residual_tolerance,
            step_tolerance: c.step_tolerance,
            initial_lambda: c.initial_lambda,
        }
    }
}

/// An analysed sketch topology, reusable for any number of solves (`Model::new` hoisted out of `solve`).
pub struct Structure(*mut EzpzStructure);
unsafe impl Send for Structure {}
unsafe impl Sync for Structure {}
impl Drop for Structure {
    fn drop(&mut self) {
        unsafe { ezpz_b200_structure_destroy(self.0) }
    }
}

impl Structure {
    /// Analyse a constraint list over variables `0..n_vars` once (`ezpz_b200_structure_create`).
    pub fn new(cons: &[EzpzConstraint], n_vars: u32) -> Result<Self, i32> {
        let mut out: *mut EzpzStructure = std::ptr::null_mut();
        let rc = unsafe {
            ezpz_b200_structure_create(cons.as_ptr(), cons.len() as u32, std::ptr::null(), n_vars, &mut out, std::ptr::null_mut())
        };
        if rc == 0 { Ok(Structure(out)) } else { Err(rc) }
    }

    /// The structure of this one's constraints followed by `extra` (`ezpz_b200_structure_extend`): what a caller that adds a
    /// constraint to a sketch it has solved uses instead of analysing the longer list from scratch.  `self` stays valid.
    pub fn extend(&self, extra: &[EzpzConstraint]) -> Result<Self, i32> {
        let mut out: *mut EzpzStructure = std::ptr::null_mut();
        let rc = unsafe { ezpz_b200_structure_extend(self.0, extra.as_ptr(), extra.len() as u32, &mut out, std::ptr::null_mut()) };
        if rc == 0 { Ok(Structure(out)) } else { Err(rc) }
    }
}

/// The additive batch API: `batch` problems of one topology in one call, sharded over every GPU of the box.
/// `guesses` is row-major `batch x n_vars`; results come back as flat vectors (finals, iterations, EZPZ_ST_* bits).
/// (`Vec` memory is pageable: the library stages it — with its own host threads up to 8 MB per call, through the driver's
/// staged copies above.  A caller that re-solves at full rate keeps its buffers in `ezpz_b200_host_alloc` memory or registers them
/// once with `ezpz_b200_host_register`: 0.35 ms instead of 1.6 ms per 65,536 sketches.)
pub fn solve_batch(mg: *mut EzpzMulti, st: &Structure, n_vars: usize, guesses: &[f64], config: crate::Config)
    -> Result<(Vec<f64>, Vec<u32>, Vec<u8>), i32> {
    let batch = guesses.len() / n_vars;
    let (mut finals, mut iters, mut status) = (vec![0.0; guesses.len()], vec![0u32; batch], vec![0u8; batch]);
    let io = EzpzBatchIo {
        guesses: guesses.as_ptr(), params: std::ptr::null(), final_values: finals.as_mut_ptr(), iterations: iters.as_mut_ptr(),
        status: status.as_mut_ptr(), unsat_mask: std::ptr::null_mut(), degen_count: std::ptr::null_mut(),
        jacobian: std::ptr::null_mut(), under_mask: std::ptr::null_mut(),
    };
    let cfg: EzpzConfig = config.into();
    let rc = unsafe { ezpz_b200_solve_batch_multi(mg, st.0, &cfg, batch as u64, &io, std::ptr::null_mut()) };
    if rc == 0 { Ok((finals, iters, status)) } else { Err(rc) }
}
