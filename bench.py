#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native ezpz solve path.

Metric (BASELINE.json): constraint solves/sec, batched.  Workload at every N: config 2 of BASELINE.json,
65,536 perturbed-guess copies of test_cases/two_rectangles PER GPU (weak scaling; problems are
independent, so GPUs share nothing and there is no data-path collective).  A "step" is one pass of the
hot path over that batch: structure analysed once outside the loop (that is the design: one analysis per
topology), then per step the whole Levenberg-Marquardt solve of all problems, one kernel launch per GPU.

  value     solves/s with guesses already resident in HBM (device-pointer C-ABI entry,
            ezpz_b200_solve_batch_device), CUDA-event time on the launching stream, L2 flushed between
            steps, max over ranks.
  e2e       the same metric through the host-buffer C-ABI call a user makes: ONE call of
            ezpz_b200_solve_batch_multi from ONE process (rank 0) shards the N x 65,536 problems of the caller's
            page-locked host buffers over the N GPUs; guesses cross PCIe to the GPU and finals / iterations /
            status / unsatisfied masks come back inside the timed region.  `e2e.single_context_per_rank` is the
            round-1 way (one process per GPU, each calling ezpz_b200_solve_batch on its shard), `e2e.pageable`
            the same call on ordinary (not page-locked) memory.
  roofline  HBM model of the dominant kernel: 264 algorithmic bytes per solve (8n in + 8n out + 8
            iterations/status, n = 16) over the kernel's mean launch duration, against the measured
            copy bandwidth in MEASURED_PEAKS.json.  This path is bound by FP64 issue, shared-memory
            bandwidth and the Cholesky dependency chain, not by HBM, so the fraction is
            small by construction; `fp64` reports the arithmetic side.
  cpu_baseline  the CPU oracle (a port of the reference algorithm; the Rust reference cannot be built
            here) on all host cores: `value` with the per-solve structure analysis the reference repeats
            (lib.rs:279), `value_hoisted` with the analysis done once per thread — the like-for-like figure
            for a GPU arm that analyses once.  `speedups` prints the ratios against both.
  strong_scaling  the literal BASELINE target: 65,536 problems IN TOTAL over the N GPUs (device-timed shards,
            max over ranks; and one multi call on host buffers).
  mixed_sweep   BASELINE.json configs[4]: 1K..1M mixed sketches (eight structures incl. inconsistent and
            underconstrained ones) through the host-buffer call, EVERY problem's verdict (iterations, converged,
            unsatisfied set, underconstrained set) and coordinates compared with the oracle.
  large_system  (N = 1) configs 3 and 4 and the reference's own bench workloads (solver_bench.rs): one large
            system per solve, GPU time through the C ABI next to the CPU port, first-solve latency with the host
            analysis included, a 2D-lattice sketch, plus the HBM figures of the assembly and SpMV kernels.

`--impl reference` times the CPU port alone on the same config and prints the same JSON shape; it loads nothing
of the product (records come from the oracle-side reader of the text format, tests/textual_twin.py).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BATCH_PER_GPU = 65536
ALGO_BYTES_PER_SOLVE = 264  # 8*16 guesses in + 8*16 finals out + 4 iterations + 1 status (+3 pad): SURVEY.md §8d
METRIC = "constraint solves/sec (batched two_rectangles, 65,536 per GPU)"
SEED = 0xE2B200D5EED00000


def make_config(world):
    """The `config` object of the JSON line — identical in both arms."""
    return {"workload": "65,536 perturbed-guess copies of test_cases/two_rectangles per GPU "
                        "(BASELINE.json configs[1]; n=16 vars, m=16 rows, nnz(J)=36)",
            "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * world,
            "parallelism": f"independent shards x{world}, no collective",
            "solver_config": "Config::default() (35 iterations, 1e-8, 1e-12, lambda 1e-9)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm
def cpu_arm(steps, warmup, sample_batch):
    """The CPU port on all host cores.  `value`: per-solve structure analysis repeated, as the reference does
    (Model::new inside every solve, lib.rs:279); `value_hoisted`: analysis once per thread."""
    import orc
    import workloads as wl
    recs, n, g = wl.perturbed_batch("two_rectangles", sample_batch, SEED)
    cores = int(orc.lib().orc_hardware_threads())
    for _ in range(warmup):
        orc.solve_batch(recs, n, g[:4096], nthreads=cores, hoist=False)
    times, times_h = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.solve_batch(recs, n, g, nthreads=cores, hoist=False)
        times.append(time.perf_counter() - t0)
    for _ in range(max(1, min(steps, 3))):
        t0 = time.perf_counter()
        orc.solve_batch(recs, n, g, nthreads=cores, hoist=True)
        times_h.append(time.perf_counter() - t0)
    t, th = statistics.median(times), statistics.median(times_h)
    return {"value": sample_batch / t, "value_hoisted": sample_batch / th, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{sample_batch} two_rectangles problems per step on {cores} host threads, median of {steps} steps; "
                      f"`value` repeats the structure analysis per solve as the reference does, `value_hoisted` analyses "
                      f"once per thread",
            "ms_per_step": t * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    base = cpu_arm(max(1, args.steps), max(1, min(args.warmup, 2)), BATCH_PER_GPU * max(1, world))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(world),
            "cpu_baseline": {k: base[k] for k in ("value", "value_hoisted", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU port of the reference algorithm (oracle/), all host threads; the Rust reference cannot be built "
                    "in this image (no cargo/rustc, faer un-vendored)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# GPU arm helpers
def device_timed(ctx, st, g, steps, warmup, dev, flush):
    """CUDA-event times (ms) of `steps` ezpz_b200_solve_batch_device launches on a dedicated stream, L2 flushed between
    steps (not timed).  Returns (list of ms, iterations array, status array)."""
    import torch
    B, n = g.shape
    d_g = torch.from_numpy(g).to(dev)
    d_f = torch.empty((B, n), dtype=torch.float64, device=dev)
    d_it = torch.empty(B, dtype=torch.int32, device=dev)
    d_st = torch.empty(B, dtype=torch.uint8, device=dev)
    io = {"guesses": d_g.data_ptr(), "final_values": d_f.data_ptr(), "iterations": d_it.data_ptr(),
          "status": d_st.data_ptr()}
    # A dedicated (non-default) torch stream: the kernel is launched on it through the C ABI and the CUDA events that
    # time it are recorded on the same stream.
    tstream = torch.cuda.Stream(device=dev)
    stream = tstream.cuda_stream
    assert stream != 0
    torch.cuda.synchronize()
    for _ in range(warmup):
        ctx.solve_batch_device(st, io, B, stream=stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(tstream):
        for k in range(steps):
            flush.fill_(k & 0xFF)  # evict the batch from L2 between timed steps (not timed)
            ev[k][0].record(tstream)
            ctx.solve_batch_device(st, io, B, stream=stream)
            ev[k][1].record(tstream)
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev], d_it.cpu().numpy(), d_st.cpu().numpy()


def timed_calls(fn, steps, warm, before=None):
    ts = []
    for k in range(warm + steps):
        if before:
            before()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if k >= warm:
            ts.append(dt)
    return ts


def st_pairs(st):
    """multiply-add pairs per LM iteration of the structure's tape (assemble + rhs + factor + solves)."""
    pa = st.pattern_a()
    lcp, lri = pa["l_col_ptr"], pa["l_row_idx"]
    n = st.n
    nnz_l = len(lri)
    rows = [[] for _ in range(n)]
    for j in range(n):
        for p in range(lcp[j] + 1, lcp[j + 1]):
            rows[lri[p]].append(j)
    chol = 0
    for j in range(n):
        chol += len(rows[j])
        sj = set(rows[j])
        for p in range(lcp[j] + 1, lcp[j + 1]):
            chol += sum(1 for k in rows[lri[p]] if k < j and k in sj)
    solves = 2 * (nnz_l - n)
    pat = st.pattern()
    jtj = 0
    rp = pat["csr_row_ptr"]
    for r in range(st.m):
        k = rp[r + 1] - rp[r]
        jtj += k * (k + 1) // 2
    return int(jtj + st.nnz + chol + solves)


def mixed_sweep(multi, sizes, reps=3):
    """BASELINE.json configs[4]: per total batch size the eight structure-homogeneous sub-batches of the mix go through ONE
    ezpz_b200_solve_jobs_multi call on page-locked host buffers (four workers per GPU, so the sub-batches' kernels overlap;
    verdict outputs on: unsatisfied masks and the underconstrained masks of the fused freedom analysis) and EVERY problem is
    compared with the oracle.  `ms_per_pass_sequential_calls`: the same sub-batches as eight ezpz_b200_solve_batch_multi calls."""
    import numpy as np

    import ezpz_b200 as ez
    import orc
    import workloads as wl
    out = []
    structures = {}
    for total in sizes:
        subs = []
        for name, recs, n, g in wl.mixed_batches(total):
            if name not in structures:
                structures[name] = ez.Structure(recs, n)
            st = structures[name]
            hg, res, owners = ez.pinned_batch_buffers(st, len(g), want_unsat=True, want_under=True)
            hg[:] = g
            subs.append((name, recs, n, g, st, hg, res, owners))

        def sequential_pass():
            for name, recs, n, g, st, hg, res, owners in subs:
                multi["one"].solve_batch(st, hg, out=res)

        def one_pass():
            multi["jobs"].solve_jobs([(st, hg, res) for name, recs, n, g, st, hg, res, owners in subs])

        sequential_pass()
        ts_seq = timed_calls(sequential_pass, reps, 1)
        for name, recs, n, g, st, hg, res, owners in subs:  # the checked results below must come from the jobs call
            res.final_values[:] = 0
            res.under_mask[:] = 0xFFFFFFFF
        one_pass()
        ts = timed_calls(one_pass, reps, 1)
        n_total = sum(len(s[3]) for s in subs)
        bad = {"iterations": 0, "converged": 0, "unsatisfied_set": 0, "underconstrained_set": 0, "coordinates_bits": 0}
        verdicts = {}
        t0 = time.perf_counter()
        for name, recs, n, g, st, hg, res, owners in subs:
            fin, it, status, um, vm = orc.solve_batch(recs, n, g, hoist=True, verdicts=True)
            bad["iterations"] += int((res.iterations != it).sum())
            bad["converged"] += int(((res.status & 1) != (status & 1)).sum())
            bad["unsatisfied_set"] += int((res.unsat_mask != um).any(axis=1).sum())
            bad["underconstrained_set"] += int((res.under_mask != vm).any(axis=1).sum())
            bad["coordinates_bits"] += int((res.final_values.view(np.uint64) != fin.view(np.uint64)).any(axis=1).sum())
            verdicts[name] = {"of": len(g), "converged": int((status & 1).sum()), "inconsistent": int(((status >> 1) & 1).sum()),
                              "underconstrained": int(vm.any(axis=1).sum())}
        cpu_s = time.perf_counter() - t0
        t = statistics.median(ts)
        out.append({"batch_total": n_total, "ms_per_pass_e2e": t * 1e3, "solves_per_s_e2e": n_total / t,
                    "ms_per_pass_sequential_calls": statistics.median(ts_seq) * 1e3,
                    "jobs_per_call": len(subs), "verdict_mismatches_vs_oracle": bad, "problems_checked": n_total,
                    "oracle_with_analysis_s": cpu_s, "verdicts": verdicts})
        del subs
    return out


def large_system_report(ctx, peaks):
    """Configs 3 and 4 of BASELINE.json and the reference's own large bench workloads (rank 0, N = 1 only)."""
    import numpy as np
    import torch

    import ezpz_b200 as ez
    import orc
    import workloads as wl

    out = {}
    path_name = {0: "batched-small", 1: "sparse direct", 2: "PCG"}

    def one_system(recs, n, g, reps, cpu_reps, pinned=False, ordered_cpu=False):
        t0 = time.perf_counter()
        st = ez.Structure(recs, n)
        analysis_s = time.perf_counter() - t0
        hg, hf = g, None
        if pinned:
            hg = torch.from_numpy(g).pin_memory().numpy()
            hf = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        times, it, status, path = ctx.time_solve_one(st, hg, reps=reps, final_values=hf)
        first_s = analysis_s + times[0]
        cpu = []
        for _ in range(cpu_reps):
            t0 = time.perf_counter()
            o = orc.solve_inner(recs, g)
            cpu.append(time.perf_counter() - t0)
        d = {"n": n, "m": st.m, "nnz": st.nnz, "gpu_solve_us": statistics.median(times[1:]) * 1e6,
             "cpu_port_solve_us": statistics.median(cpu) * 1e6, "cpu_cores": 1, "lm_iterations": it,
             "cpu_lm_iterations": int(o.iterations), "converged": bool(status & 1), "path": path_name[path],
             "host_analysis_us_once_per_topology": analysis_s * 1e6,
             "first_solve_us_analysis_included": first_s * 1e6}
        if ordered_cpu:
            od = st.ordering()
            t0 = time.perf_counter()
            oo = orc.solve_inner_ordered(recs, g, od["elim_order"], od["sum_chunk"])
            d["cpu_port_with_the_gpu_paths_elimination_order_us"] = (time.perf_counter() - t0) * 1e6
            d["elimination_tree_levels"] = od["n_levels"]
            d["nnz_l"] = od["nnz_l"]
            assert int(oo.iterations) == it
        return d, st, o, hf

    # config 3 (+ solver_bench.rs:174-200 sizes): massive_parallel_system with 200 / 500 / 600 lines
    for lines, key in ((200, "massive_parallel_system_800x800"), (500, "massive_parallel_system_2000x2000"),
                       (600, "massive_parallel_system_2400x2400")):
        recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(lines, False))
        d, st, o, _ = one_system(recs, n, g, 30, 5)
        if lines == 500:
            d["readme_reference_us"] = 2943
        out[key] = d
    out["massive_parallel_system_2000x2000"]["note"] = (
        "ezpz_b200_solve_one, host buffers, H2D + one persistent kernel + D2H per call; CPU = oracle port incl. its "
        "per-solve analysis, as ezpz-cli times it; README figure is the reference's own (hardware not stated)")
    # solver_bench.rs:42-140: the small bench workloads through ezpz_b200_solve (priority loop, lint, cached topology)
    small = {}
    for name in ("inconsistent", "nonsquare", "two_rectangles"):
        recs, n, g, _ = wl.system_from_text(wl.fixture_text(name))
        st = ez.Structure(recs, n)
        times, it, status, path = ctx.time_solve_one(st, g, reps=50)
        cpu = []
        for _ in range(50):
            t0 = time.perf_counter()
            orc.solve_inner(recs, g)
            cpu.append(time.perf_counter() - t0)
        small[name] = {"gpu_solve_one_us": statistics.median(times[5:]) * 1e6, "cpu_port_us": statistics.median(cpu) * 1e6,
                       "lm_iterations": it}
    out["single_small_sketch_latency"] = dict(small, note="one sketch per call is launch-latency bound on a GPU; the batched "
                                              "call is the product's answer to many small sketches")
    # config 4: synthetic 1,001,000-variable sketch
    recs, n, g, exact = wl.chain_sketch(77000)
    d, st, o, hf = one_system(recs, n, g, 6, 1, pinned=True)
    scale = np.maximum(1.0, np.abs(o.final_values))
    d["max_rel_diff_vs_cpu_port"] = float((np.abs(hf - o.final_values) / scale).max())
    d["gpu_solve_ms"], d["cpu_port_solve_ms"] = d.pop("gpu_solve_us") / 1e3, d.pop("cpu_port_solve_us") / 1e3
    od = st.ordering()
    d["elimination_tree_levels"], d["nnz_l"] = od["n_levels"], od["nnz_l"]
    t0 = time.perf_counter()
    st2 = st.extend(recs[-1:])  # ezpz_b200_structure_extend: one constraint added, the elimination order kept
    d["host_reanalysis_us_after_adding_one_constraint"] = (time.perf_counter() - t0) * 1e6
    del st2
    out["synthetic_1M_variable_sketch"] = d
    del st
    # a 2D lattice (separators of ~N points: tall panels, the opposite regime of the chain)
    recs, n, g, exact = wl.grid_truss(100)
    d, st, o, _ = one_system(recs, n, g, 4, 1, ordered_cpu=True)
    d["gpu_solve_ms"], d["cpu_port_solve_ms"] = d.pop("gpu_solve_us") / 1e3, d.pop("cpu_port_solve_us") / 1e3
    d["note"] = ("cpu_port_solve_ms factorises in natural order (the oracle's default; faer would pick a fill-reducing order); "
                 "the `..._elimination_order_us` figure is the same CPU port using the GPU path's nested-dissection order")
    out["grid_truss_100x100_lattice"] = d
    del st
    # kernels on 2.08M variables (~270 MB working set > 126 MB L2), launches back to back
    recs, n, g, _ = wl.chain_sketch(160000)
    st = ez.Structure(recs, n)
    ks = {}
    for which, name in ((0, "assemble_large_kernel"), (2, "spmv_csr_kernel (z = Jt q)"), (1, "spmv_csr_kernel (y = J p)")):
        us, by = ctx.large_bench(st, g, which, 20)
        ks[name] = {"us_per_launch": us, "algorithmic_MB": by / 1e6, "algorithmic_GB_s": by / us / 1e3,
                    "frac_of_hbm_peak_on_algorithmic_bytes": by / us / 1e3 / peaks["hbm_gbs"]}
    ks["note"] = ("algorithmic bytes are SURVEY.md §8d's (64-byte records); the shipped assembly kernel reads packed ~50-byte "
                  "record tiles and ncu measured 163 MB of DRAM traffic for 268 MB algorithmic "
                  "(profiles/r01j_assembly_tile_order.md), i.e. its real DRAM utilisation is ~0.6x the fraction above")
    out["kernels_on_2M_variable_sketch"] = ks
    return out


def reference_benches(ctx):
    """The reference's own criterion benches (ezpz/benches/solver_bench.rs:42-209), one solve per call through
    ezpz_b200_solve (priority loop, lint, topology cache warm — the reference re-analyses every call) next to the CPU port:
    solve_inconsistent, solve_nonsquare, solve_two_rectangles, solve_two_rectangles_dependent (14 variables),
    solve_massive (200 and 600 lines), solve_massive_analysis (200 lines), solve_nonsquare_analysis."""
    import ctypes as C

    import numpy as np

    import ezpz_b200 as ez
    import orc
    import textual_twin
    import workloads as wl
    from ezpz_b200 import native

    def rec(kind, ids, p0=0.0):
        r = np.zeros(1, dtype=textual_twin.REC_DTYPE)
        r["kind"], r["p0"], r["weight"] = kind, p0, 1.0
        r["ids"][0, :len(ids)] = ids
        return r

    def c_solve(recs, g, analysis):
        n_cons, n_vars = len(recs), len(g)
        fv, un, uc = np.zeros(n_vars), np.zeros(n_cons, np.uint64), np.zeros(n_vars, np.uint32)
        warr = (native.WarningRec * (2 * n_cons + 8))()
        out = native.OutcomeRec()
        out.final_values, out.unsatisfied, out.underconstrained = fv.ctypes.data, un.ctypes.data, uc.ctypes.data
        out.warnings, out.warnings_cap = C.addressof(warr), 2 * n_cons + 8
        det = native.ErrorDetail()
        cfg = ez.Config()._native()
        args = (ctx.handle, native.ptr(recs), None, None, n_cons, None, native.ptr(g), n_vars, C.byref(cfg), 1 if analysis else 0,
                C.byref(out), C.byref(det))
        keep = (fv, un, uc, warr, out, det, cfg, recs, g)

        def go():
            rc = native.lib().ezpz_b200_solve(*args)
            assert rc == 0 and keep
            return out
        return go

    def case(recs, g, analysis, reps):
        recs = np.ascontiguousarray(recs)
        g = np.ascontiguousarray(g, dtype=np.float64)
        go = c_solve(recs, g, analysis)
        o = go()
        gpu = statistics.median(timed_calls(go, reps, 3)) * 1e6
        cpu_reps = max(1, min(reps, 20))
        ref = orc.solve(recs, g, analysis=analysis)
        cpu = statistics.median(timed_calls(lambda: orc.solve(recs, g, analysis=analysis), cpu_reps, 1)) * 1e6
        d = {"gpu_us": gpu, "cpu_port_us": cpu, "lm_iterations": int(o.iterations), "cpu_lm_iterations": int(ref.iterations),
             "n_vars": len(g), "n_eqs": int(o.num_eqs)}
        if analysis:
            d["underconstrained_agrees"] = int(o.n_underconstrained) == len(ref.underconstrained)
        return d

    out = {}
    for name in ("inconsistent", "nonsquare", "two_rectangles"):
        recs, n, g, _ = wl.system_from_text(wl.fixture_text(name))
        out[f"solve_{name}"] = case(recs, g, False, 50)
    # solve_two_rectangles_dependent: points p0..p3 = ids 0..7, p5..p7 = ids 8..13; the second rectangle hangs on p2
    P = lambda k: [2 * k, 2 * k + 1]
    p0, p1, p2, p3, p5, p6, p7 = P(0), P(1), P(2), P(3), P(4), P(5), P(6)
    dep = [rec(9, [0], 1.0), rec(9, [1], 1.0), rec(7, p0 + p1), rec(7, p2 + p3), rec(6, p3 + p0), rec(6, p1 + p2),
           rec(2, p0 + p1, 4.0), rec(2, p0 + p3, 3.0),
           rec(7, p2 + p5), rec(7, p6 + p7), rec(6, p7 + p2), rec(6, p5 + p6), rec(2, p2 + p5, 4.0), rec(2, p2 + p7, 4.0)]
    g_dep = [1.0, 1.0, 4.5, 1.5, 4.0, 3.5, 1.5, 3.0, 5.5, 3.5, 5.0, 4.5, 2.5, 4.0]
    out["solve_two_rectangles_dependent"] = case(np.concatenate(dep), g_dep, False, 50)
    for lines in (200, 600):
        recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(lines, False))
        out[f"solve_massive_{lines}_lines"] = case(recs, g, False, 30)
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(200, False))
    out["solve_massive_analysis_200_lines"] = case(recs, g, True, 3)
    recs, n, g, _ = wl.system_from_text(wl.fixture_text("nonsquare"))
    out["solve_nonsquare_analysis"] = case(recs, g, True, 50)
    out["note"] = ("one sketch per call is launch-latency bound on a GPU (the batched call is the product's answer to many small "
                   "sketches); the CPU port re-analyses the structure every call as the reference does, the GPU side hits its "
                   "topology cache")
    return out


def freedom_report(ctx):
    """Freedom analysis (find_dof.rs:15-104) throughput: batched small sketches fused into the solve call, and the
    reference's own analysis bench sizes (solver_bench.rs:146-171: massive with analysis)."""
    import numpy as np

    import ezpz_b200 as ez
    import orc
    import workloads as wl
    out = {}
    for name in ("two_rectangles", "underconstrained", "parc_coincident"):
        recs, n, g = wl.perturbed_batch(name, 65536, SEED + (3 << 40))
        st = ez.Structure(recs, n)
        hg, res, owners = ez.pinned_batch_buffers(st, len(g), want_unsat=True, want_under=True)
        hg[:] = g
        ts = timed_calls(lambda: ctx.solve_batch(st, hg, out=res), 5, 2)
        hg2, res2, owners2 = ez.pinned_batch_buffers(st, len(g), want_unsat=True, want_under=False)
        hg2[:] = g
        ts2 = timed_calls(lambda: ctx.solve_batch(st, hg2, out=res2), 5, 2)
        t0 = time.perf_counter()
        fin, it, status, um, vm = orc.solve_batch(recs, n, g, hoist=True, verdicts=True)
        cpu_s = time.perf_counter() - t0
        out[name] = {"batch": len(g), "solve_plus_analysis_per_s_e2e": len(g) / statistics.median(ts),
                     "solve_only_per_s_e2e": len(g) / statistics.median(ts2),
                     "cpu_port_solve_plus_analysis_per_s": len(g) / cpu_s, "cpu_cores": int(orc.lib().orc_hardware_threads()),
                     "underconstrained_set_mismatches": int((res.under_mask != vm).any(axis=1).sum())}
    for lines in (50, 200, 600):
        recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(lines, False))
        st = ez.Structure(recs, n)
        one = ctx.solve_one(st, g, want_jacobian=True)
        ts = timed_calls(lambda: ctx.freedom_analysis(st, one.jacobian), 3, 1)
        d = {"n": n, "m": st.m, "gpu_analysis_ms": statistics.median(ts) * 1e3}
        if lines <= 200:
            t0 = time.perf_counter()
            o = orc.solve_inner(recs, g, analysis=True)
            t1 = time.perf_counter()
            orc.solve_inner(recs, g, analysis=False)
            t2 = time.perf_counter()
            d["cpu_port_analysis_ms"] = ((t1 - t0) - (t2 - t1)) * 1e3
            mask = ctx.freedom_analysis(st, one.jacobian)
            got = np.flatnonzero(np.unpackbits(mask.view(np.uint8), bitorder="little")[:n]).tolist()
            d["matches_cpu_port"] = got == o.underconstrained
        out[f"massive_parallel_system_{4 * lines}_variables"] = d
    return out


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import ezpz_b200 as ez
    import workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # stdout carries the ONE JSON line: whatever native libraries print while the job runs (NCCL's version banner goes
        # to file descriptor 1 whatever NCCL_DEBUG_FILE says) is sent to stderr; the descriptor comes back for the JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")

    def barrier():
        if world > 1:
            dist.barrier()

    def host_barrier():
        """A barrier on the host only (gloo): while rank 0 drives every GPU of the box through the multi-GPU call the other
        ranks must leave their GPUs alone — a rank parked in an NCCL barrier keeps a kernel spinning on its device, and two
        processes' work on one GPU is time-sliced, not concurrent."""
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=host_group)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = BATCH_PER_GPU
    steps, warmup = args.steps, max(3, args.warmup)
    # each rank's shard of the weak-scaled job: ranks perturb with disjoint seeds
    recs, n, g = wl.perturbed_batch("two_rectangles", B, SEED + (rank << 32))
    ctx = ez.Context(local)
    st = ez.Structure(recs, n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- value: device-timed, guesses resident in HBM
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    wall0 = time.perf_counter()
    times_ms, it_host, st_host = device_timed(ctx, st, g, steps, warmup, dev, flush)
    launches = ctx.launches - launches0 - warmup
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms_max = max_over_ranks(sum(times_ms))
    assert (st_host & 1).all() and not (st_host & 2).any(), "benchmark batch did not converge/satisfy"
    mean_iters = float(it_host.mean())

    # ---- strong scaling: 65,536 problems in total, each rank its contiguous shard (device-timed)
    sb, se = ez.shard_range(B, rank, world)
    recs0, _, g_all = wl.perturbed_batch("two_rectangles", B, SEED)
    strong_ms, _, strong_status = device_timed(ctx, st, np.ascontiguousarray(g_all[sb:se]), steps, warmup, dev, flush)
    strong_ms_max = max_over_ranks(sum(strong_ms))
    assert (strong_status & 1).all()

    # ---- e2e, the round-1 way: one process per GPU, each calling ezpz_b200_solve_batch on page-locked host buffers
    e2e_warm = max(10, args.warmup)  # first DMA passes over fresh pinned pages are slow on some boxes of the pool
    hg, res, owners = ez.pinned_batch_buffers(st, B, want_unsat=True)
    hg[:] = g
    ts = timed_calls(lambda: ctx.solve_batch(st, hg, out=res), steps, e2e_warm, before=barrier)
    assert (res.status & 1).all()
    per_rank_s_max = max_over_ranks(sum(ts))
    pageable = None
    if world == 1:
        tp = timed_calls(lambda: ctx.solve_batch(st, g), steps, 3)
        pageable = {"value": B * len(tp) / sum(tp), "unit": "solves/s", "ms_per_step": sum(tp) / len(tp) * 1e3,
                    "api": "ezpz_b200_solve_batch on ordinary numpy arrays (what a Rust Vec<f64> is), output arrays allocated "
                           "per call: at this size (17 MB per call) the copies are the driver's staged ones; calls of up to 8 MB go "
                           "through the library's page-locked block on its host threads (profiles/r02v_pageable_buffers.log)"}
    del hg, res, owners
    host_barrier()

    # ---- e2e, the headline: ONE call from ONE process drives all N GPUs (rank 0; the other ranks wait at a host barrier)
    one_call = strong_one_call = sweep = None
    if rank == 0:
        multi = ez.MultiContext(devices=list(range(world)))
        total = B * world
        parts = [g] + [wl.perturbed_batch("two_rectangles", B, SEED + (r << 32))[2] for r in range(1, world)]
        hg, res, owners = ez.pinned_batch_buffers(st, total, want_unsat=True)
        for r, part in enumerate(parts):
            hg[r * B:(r + 1) * B] = part
        l0 = multi.launches
        ts = timed_calls(lambda: multi.solve_batch(st, hg, out=res), steps, e2e_warm)
        multi_launches_per_call = (multi.launches - l0) / (steps + e2e_warm)
        assert (res.status & 1).all() and not (res.status & 2).any()
        one_call = {"s": sum(ts), "n": len(ts), "total": total}
        del hg, res, owners
        hg, res, owners = ez.pinned_batch_buffers(st, B, want_unsat=True)
        hg[:] = g_all
        ts = timed_calls(lambda: multi.solve_batch(st, hg, out=res), steps, e2e_warm)
        strong_one_call = {"s": sum(ts), "n": len(ts)}
        del hg, res, owners
        if not args.no_extras:
            sizes = [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20] if world == 1 else [1 << 16, 1 << 20]
            jobs_multi = ez.MultiContext(devices=[d for d in range(world) for _ in range(4)])
            sweep = mixed_sweep({"one": multi, "jobs": jobs_multi}, sizes)
            del jobs_multi
    host_barrier()
    clocks = sampler.stop()

    if rank == 0:
        peaks, peak_src = load_peaks()
        total = B * world
        ms_per_step = dev_ms_max / steps
        value = total * steps / (dev_ms_max * 1e-3)
        kernel_ms = statistics.mean(times_ms)
        achieved = ALGO_BYTES_PER_SOLVE * B / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        prof = os.path.join(ROOT, "profiles", "lm_small_kernel_traffic.json")
        if os.path.exists(prof):
            with open(prof) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        # arithmetic side: multiply-add pairs of the tape + ~270 flops of constraint evaluation per iteration
        flops_per_iter = 2 * st_pairs(st) + 270
        fp64 = {"flops_per_lm_iteration": flops_per_iter, "mean_lm_iterations": mean_iters,
                "achieved_gflops": B * (mean_iters + 1) * flops_per_iter / (kernel_ms * 1e-3) / 1e9,
                "note": "nominal B200 FP64 peak ~37 TFLOP/s; this kernel is bound by shared-memory bandwidth "
                        "(2 LDS per FMA) and dependent-issue latency"}
        e2e_value = one_call["total"] * one_call["n"] / one_call["s"]
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(world),
            "timing": {"l2": "256 MiB fill between timed steps (not timed)",
                       "how": "CUDA events per step on the launching stream, max over ranks", "wall_s_timed_region": wall},
            "e2e": {"value": e2e_value, "unit": "solves/s",
                    "h2d_bytes_per_step": int(total * n * 8), "d2h_bytes_per_step": int(total * (n * 8 + 4 + 1 + 4)),
                    "ms_per_step": one_call["s"] / one_call["n"] * 1e3,
                    "api": f"ONE ezpz_b200_solve_batch_multi call per step from one process over {world} GPU(s): caller's "
                           "page-locked host buffers -> PCIe -> kernel -> PCIe -> caller's buffers, synchronous",
                    "kernel_launches_per_call": multi_launches_per_call,
                    "single_context_per_rank": {"value": total * steps / per_rank_s_max, "ms_per_step": per_rank_s_max / steps * 1e3,
                                                "api": "one process per GPU, each ezpz_b200_solve_batch on its shard, barrier "
                                                       "before every call, max over ranks"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "lm_small_kernel",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SOLVE * B,
                         "peak_source": peak_src,
                         "note": "HBM is not the binding resource of the batched small-system path (SURVEY.md §8d)"},
            "fp64": fp64,
            "strong_scaling": {"workload": "65,536 two_rectangles problems IN TOTAL (the literal BASELINE target config)",
                               "problems_per_gpu": int(se - sb),
                               "value": B * steps / (strong_ms_max * 1e-3), "ms_per_step": strong_ms_max / steps,
                               "e2e_value": B * strong_one_call["n"] / strong_one_call["s"],
                               "e2e_ms_per_step": strong_one_call["s"] / strong_one_call["n"] * 1e3, "unit": "solves/s"},
        }
        if pageable:
            line["e2e"]["pageable"] = pageable
        if sweep is not None:
            line["mixed_sweep"] = {"config": "BASELINE.json configs[4]: 50% two_rectangles, 15% square, 10% circle_tangent, 5% each "
                                             "arc_length, parc_coincident, inconsistent, underconstrained, perpendicular",
                                   "api": "ONE ezpz_b200_solve_jobs_multi call per pass (eight structure-homogeneous jobs, four workers per "
                                          "GPU) on page-locked host buffers, unsatisfied and underconstrained masks on", "sizes": sweep}
        if world == 1:
            cb = cpu_arm(3, 1, BATCH_PER_GPU)
            line["cpu_baseline"] = {k: v for k, v in cb.items() if k != "ms_per_step"}
            line["speedups"] = {"e2e_vs_cpu_as_the_reference_runs": e2e_value / cb["value"],
                                "e2e_vs_cpu_analysis_hoisted": e2e_value / cb["value_hoisted"],
                                "device_vs_cpu_as_the_reference_runs": value / cb["value"],
                                "device_vs_cpu_analysis_hoisted": value / cb["value_hoisted"],
                                "note": "the GPU arm analyses the structure once outside the timed loop; the hoisted CPU figure is "
                                        "the like-for-like one"}
            if not args.no_extras:
                line["freedom_analysis"] = freedom_report(ctx)
                line["reference_benches"] = reference_benches(ctx)
            if not args.no_large and not args.no_extras:
                line["large_system"] = large_system_report(ctx, peaks)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        if saved_stdout is not None and rank == 0:
            os.dup2(2, 1)  # (teardown chatter, if any, off stdout again)
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-large", action="store_true", help="skip the single-large-system report (configs 3 and 4)")
    ap.add_argument("--no-extras", action="store_true", help="headline, e2e, strong scaling and CPU baseline only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
