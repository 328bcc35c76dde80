#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native ezpz solve path.

Metric (BASELINE.json): constraint solves/sec, batched.  Workload at every N: config 2 of BASELINE.json,
65,536 perturbed-guess copies of test_cases/two_rectangles PER GPU (weak scaling; problems are
independent, so ranks share nothing and there is no data-path collective).  A "step" is one pass of the
hot path over that batch: structure analysed once outside the loop (that is the design: one analysis per
topology), then per step the whole Levenberg-Marquardt solve of all 65,536 problems in one kernel launch.

  value     solves/s with guesses already resident in HBM (device-pointer C-ABI entry,
            ezpz_b200_solve_batch_device), CUDA-event time on the launching stream, L2 flushed between
            steps, max over ranks.
  e2e       the same metric through the host-buffer C-ABI call a user makes (ezpz_b200_solve_batch):
            pinned host guesses -> H2D -> kernel -> D2H of finals/iterations/status inside the timed
            region.
  roofline  HBM model of the dominant kernel: 264 algorithmic bytes per solve (8n in + 8n out + 8
            iterations/status, n = 16) over the kernel's mean launch duration, against the measured
            copy bandwidth in MEASURED_PEAKS.json.  This path is bound by FP64 issue, shared-memory
            bandwidth and the Cholesky dependency chain, not by HBM, so the fraction is
            small by construction; `fp64` reports the arithmetic side.
  cpu_baseline  the CPU oracle (a port of the reference algorithm; the Rust reference cannot be built
            here) on all host cores, per-solve structure analysis included as the reference does.

  large_system  (N = 1) BASELINE.json configs 3 and 4: one large system per solve, GPU time through the C ABI next
            to the CPU port, plus the HBM figures of the assembly and SpMV kernels on a system larger than L2.

`--impl reference` times that CPU port alone on the same config and prints the same JSON shape.
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


def cpu_arm(steps, warmup, sample_batch):
    """The CPU port on all host cores: per-solve analysis repeated, as the reference does (lib.rs:279)."""
    import orc
    import workloads as wl
    recs, n, g = wl.two_rectangles_batch(sample_batch)
    cores = int(orc.lib().orc_hardware_threads())
    for _ in range(warmup):
        orc.solve_batch(recs, n, g[:4096], nthreads=cores, hoist=False)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.solve_batch(recs, n, g, nthreads=cores, hoist=False)
        times.append(time.perf_counter() - t0)
    t_hoist0 = time.perf_counter()
    orc.solve_batch(recs, n, g, nthreads=cores, hoist=True)
    t_hoist = time.perf_counter() - t_hoist0
    best = statistics.median(times)
    return {"value": sample_batch / best, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{sample_batch} of the 65,536 two_rectangles problems per step, median of {steps} steps, "
                      f"structure analysis repeated per solve as the reference does; with the analysis hoisted "
                      f"once per thread: {sample_batch / t_hoist:.0f} solves/s",
            "ms_per_step": best * 1e3}


def large_system_report(ctx, peaks):
    """Configs 3 and 4 of BASELINE.json (one large system per solve; rank 0, N = 1 only): solve time through the
    C ABI with host buffers next to the CPU port on the same inputs, and the HBM figures of the assembly and SpMV
    kernels on a system whose working set exceeds L2."""
    import numpy as np
    import torch

    import ezpz_b200 as ez
    import orc
    import workloads as wl

    out = {}
    # config 3: massive_parallel_system, 500 lines = 2,000 rows x 2,000 vars (README.md:36-40)
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500, False))
    st = ez.Structure(recs, n)
    times, it, status, path = ctx.time_solve_one(st, g, reps=30)
    gpu_us = statistics.median(times[5:]) * 1e6
    cpu = []
    for _ in range(7):
        t0 = time.perf_counter()
        o = orc.solve_inner(recs, g)
        cpu.append(time.perf_counter() - t0)
    out["massive_parallel_system_2000x2000"] = {
        "gpu_solve_us": gpu_us, "cpu_port_solve_us": statistics.median(cpu) * 1e6, "cpu_cores": 1,
        "readme_reference_us": 2943, "lm_iterations": it, "cpu_lm_iterations": int(o.iterations),
        "converged": bool(status & 1), "path": {0: "batched-small", 1: "sparse direct", 2: "PCG"}[path],
        "note": "ezpz_b200_solve_one, host buffers, H2D + one persistent kernel + D2H per call; CPU = oracle port incl. its "
                "per-solve analysis, as ezpz-cli times it; README figure is the reference's own (hardware not stated)"}
    # config 4: synthetic 1,001,000-variable sketch
    recs, n, g, exact = wl.chain_sketch(77000)
    t0 = time.perf_counter()
    st = ez.Structure(recs, n)
    analysis_s = time.perf_counter() - t0
    od = st.ordering()
    hg = torch.from_numpy(g).pin_memory().numpy()
    hf = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    times, it, status, path = ctx.time_solve_one(st, hg, reps=6, final_values=hf)
    gpu_ms = statistics.median(times[1:]) * 1e3
    t0 = time.perf_counter()
    o = orc.solve_inner(recs, g)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    scale = np.maximum(1.0, np.abs(o.final_values))
    out["synthetic_1M_variable_sketch"] = {
        "n": n, "m": st.m, "nnz": st.nnz, "gpu_solve_ms": gpu_ms, "cpu_port_solve_ms": cpu_ms, "cpu_cores": 1,
        "lm_iterations": it, "cpu_lm_iterations": int(o.iterations), "converged": bool(status & 1),
        "max_rel_diff_vs_cpu_port": float((np.abs(hf - o.final_values) / scale).max()),
        "path": {0: "batched-small", 1: "sparse direct", 2: "PCG"}[path], "elimination_tree_levels": od["n_levels"],
        "nnz_l": od["nnz_l"], "host_analysis_s_once_per_topology": analysis_s}
    del st
    # kernels on 2.08M variables (~270 MB working set > 126 MB L2), launches back to back
    recs, n, g, _ = wl.chain_sketch(160000)
    st = ez.Structure(recs, n)
    ks = {}
    for which, name in ((0, "assemble_large_kernel"), (2, "spmv_csr_kernel (z = Jt q)"), (1, "spmv_csr_kernel (y = J p)")):
        us, by = ctx.large_bench(st, g, which, 20)
        ks[name] = {"us_per_launch": us, "algorithmic_MB": by / 1e6, "GB_s": by / us / 1e3,
                    "frac_of_hbm_peak": by / us / 1e3 / peaks["hbm_gbs"]}
    out["kernels_on_2M_variable_sketch"] = ks
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_arm(max(1, args.steps), max(1, min(args.warmup, 2)), BATCH_PER_GPU)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "65,536 perturbed-guess copies of test_cases/two_rectangles (n=16, m=16, nnz=36)",
                       "batch": BATCH_PER_GPU, "note": "CPU port of the reference algorithm (oracle/), all host threads; "
                       "the Rust reference cannot be built in this image"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


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

    B = BATCH_PER_GPU
    # each rank's shard of the weak-scaled job: ranks perturb with disjoint seeds
    recs, n, g = wl.perturbed_batch("two_rectangles", B, 0xE2B200D5EED00000 + (rank << 32))
    ctx = ez.Context(local)
    st = ez.Structure(recs, n)

    d_g = torch.from_numpy(g).to(dev)
    d_f = torch.empty((B, n), dtype=torch.float64, device=dev)
    d_it = torch.empty(B, dtype=torch.int32, device=dev)
    d_st = torch.empty(B, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    io = {"guesses": d_g.data_ptr(), "final_values": d_f.data_ptr(), "iterations": d_it.data_ptr(),
          "status": d_st.data_ptr()}
    # A dedicated (non-default) torch stream: the kernel is launched on it through the C ABI and the CUDA
    # events that time it are recorded on the same stream.
    tstream = torch.cuda.Stream(device=dev)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device():
        ctx.solve_batch_device(st, io, B, stream=stream)

    torch.cuda.synchronize()
    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    wall0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with torch.cuda.stream(tstream):
        for k in range(args.steps):
            flush.fill_(k & 0xFF)  # evict the batch from L2 between timed steps (not timed)
            ev[k][0].record(tstream)
            step_device()
            ev[k][1].record(tstream)
    torch.cuda.synchronize()
    launches = ctx.launches - launches0
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    times_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(times_ms)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())

    # ---- e2e through the host-buffer C-ABI call, pinned host memory
    h_g = torch.from_numpy(g).pin_memory()
    pinned = ez.BatchResult()
    pinned.final_values = torch.empty((B, n), dtype=torch.float64).pin_memory().numpy()
    pinned.iterations = torch.empty(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    pinned.status = torch.empty(B, dtype=torch.uint8).pin_memory().numpy()
    pinned.unsat_mask = torch.empty((B, (st.n_cons + 31) // 32), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    pinned.degen_count = None
    pinned.jacobian = None
    h_g_np = h_g.numpy()
    e2e_times = []
    # untimed warm-up calls first: on some boxes of the pool the first DMA passes over freshly pinned host pages run at a
    # quarter of the link rate (profiles/r01j_pcie_probe.log: 13.6 GB/s, then 54 GB/s)
    e2e_warm = max(10, args.warmup)
    for k in range(e2e_warm + args.steps):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ctx.solve_batch(st, h_g_np, out=pinned)
        dt = time.perf_counter() - t0
        if k >= e2e_warm:
            e2e_times.append(dt)
    assert (pinned.status & 1).all()
    e2e_s = sum(e2e_times)
    t2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_s_max = float(t2.item())
    clocks = sampler.stop()

    # sanity: the timed work really solved the batch
    it_host = d_it.cpu().numpy()
    st_host = d_st.cpu().numpy()
    assert (st_host & 1).all() and not (st_host & 2).any(), "benchmark batch did not converge/satisfy"
    mean_iters = float(it_host.mean())
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks, peak_src = load_peaks()
        total = B * world
        ms_per_step = dev_ms_max / args.steps
        value = total * args.steps / (dev_ms_max * 1e-3)
        kernel_ms = statistics.mean(times_ms)
        achieved = ALGO_BYTES_PER_SOLVE * B / (kernel_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "lm_small_kernel_traffic.json")
        if os.path.exists(prof):
            with open(prof) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        # arithmetic side: multiply-add pairs of the tape + ~270 flops of constraint evaluation per iteration
        flops_per_iter = 2 * st_pairs(st) + 270
        fp64 = {"flops_per_lm_iteration": flops_per_iter, "mean_lm_iterations": mean_iters,
                "achieved_gflops": B * (mean_iters + 1) * flops_per_iter / (kernel_ms * 1e-3) / 1e9,
                "note": "nominal B200 FP64 peak ~37 TFLOP/s; this kernel is bound by shared-memory bandwidth "
                        "(2 LDS per FMA) and dependent-issue latency"}
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "65,536 perturbed-guess copies of test_cases/two_rectangles per GPU "
                                   "(BASELINE.json configs[1]; n=16 vars, m=16 rows, nnz(J)=36)",
                       "batch_per_gpu": B, "global_batch": total, "parallelism": f"independent shards x{world}, no collective",
                       "l2": "256 MiB fill between timed steps (not timed)", "timing": "CUDA events per step on the launching stream, max over ranks",
                       "wall_s_timed_region": wall},
            "e2e": {"value": total * len(e2e_times) / e2e_s_max, "unit": "solves/s",
                    "h2d_bytes_per_step": int(B * n * 8), "d2h_bytes_per_step": int(B * n * 8 + B * 4 + B + B * 4),
                    "ms_per_step": e2e_s_max / len(e2e_times) * 1e3,
                    "api": "ezpz_b200_solve_batch (host buffers, pinned), H2D + kernel + D2H + sync per step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "kernel": "lm_small_kernel",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SOLVE * B,
                         "peak_source": peak_src,
                         "note": "HBM is not the binding resource of the batched small-system path (SURVEY.md §8d)"},
            "fp64": fp64,
        }
        if world == 1:
            sample = BATCH_PER_GPU
            line["cpu_baseline"] = {k: v for k, v in cpu_arm(3, 1, sample).items() if k != "ms_per_step"}
            if not args.no_large:
                line["large_system"] = large_system_report(ctx, peaks)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        if saved_stdout is not None and rank == 0:
            os.dup2(2, 1)  # (teardown chatter, if any, off stdout again)
        dist.destroy_process_group()


def st_pairs(st):
    """multiply-add pairs per LM iteration of the structure's tape (assemble + rhs + factor + solves)."""
    # JtJ products + Jt r + Cholesky + two triangular solves, counted from the patterns
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
    rp, ci = pat["csr_row_ptr"], pat["csr_col_idx"]
    for r in range(st.m):
        k = rp[r + 1] - rp[r]
        jtj += k * (k + 1) // 2
    return int(jtj + st.nnz + chol + solves)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-large", action="store_true", help="skip the single-large-system report (configs 3 and 4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
