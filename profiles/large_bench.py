"""Stand-alone timing of the large-path kernels (configs 3 and 4 of BASELINE.json) with CUDA events:
python profiles/large_bench.py [cells]    (default 77000 cells = 1,001,000 variables)
Prints one JSON line per kernel with algorithmic GB/s against MEASURED_PEAKS.json, plus whole-solve times."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 77000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    ctx = ez.Context(0)
    recs, n, g, exact = wl.chain_sketch(cells)
    t = time.perf_counter()
    st = ez.Structure(recs, n)
    t_analysis = time.perf_counter() - t
    print(json.dumps({"system": f"chain_sketch({cells})", "n": n, "m": st.m, "nnz": st.nnz, "constraints": st.n_cons,
                      "host_analysis_s": t_analysis}))
    for which, name in ((0, "assemble_large_kernel (J in CSC order)"), (3, "assemble_large_kernel (CSC + CSR-ordered copy)"),
                        (1, "spmv_csr_kernel (y = J p)"), (2, "spmv_csr_kernel (z = Jt q)")):
        us, by = ctx.large_bench(st, g, which, reps)
        print(json.dumps({"kernel": name, "us_per_launch": us, "algorithmic_MB": by / 1e6, "GB_s": by / us / 1e3,
                          "frac_of_measured_hbm": by / us / 1e3 / peak}))
    if "--solve" in sys.argv:
        out = ctx.solve_one(st, g)
        t = time.perf_counter()
        out = ctx.solve_one(st, g)
        dt = time.perf_counter() - t
        print(json.dumps({"solve_one_ms": dt * 1e3, "path": out.path_used, "lm_iterations": out.iterations,
                          "converged": out.converged, "cg_iterations": out.lin_iters,
                          "max_err_vs_constructed_solution": float(np.abs(out.final_values - exact).max())}))


if __name__ == "__main__":
    main()
