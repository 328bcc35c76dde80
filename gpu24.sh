ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r1f.log 2>&1
tail -2 gpurun_out/ncu_bench_r1f.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:"assemble_large" -s 2 -c 1 -o gpurun_out/prof_asm_r1f python profiles/large_bench.py 160000 2 > gpurun_out/ncu_asm_r1f.log 2>&1
tail -2 gpurun_out/ncu_asm_r1f.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"spmv_csr" -s 0 -c 2 -o gpurun_out/prof_spmv_r1f python profiles/large_bench.py 160000 1 > gpurun_out/ncu_spmv_r1f.log 2>&1
tail -2 gpurun_out/ncu_spmv_r1f.log | cut -c1-200
cat > /tmp/lm8192.py <<'PY'
import sys
sys.path.insert(0,'tests')
import ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
recs, n, g, _ = wl.chain_sketch(8192)
st = ez.Structure(recs, n)
out = ctx.solve_one(st, g)
print(out.iterations, out.converged)
PY
ncu --set full --clock-control none --import-source on -k regex:"lm_large" -c 1 -o gpurun_out/prof_lm_large_r1f python /tmp/lm8192.py > gpurun_out/ncu_lm_large_r1f.log 2>&1
tail -2 gpurun_out/ncu_lm_large_r1f.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"lm_small" -s 3 -c 1 -o gpurun_out/prof_lm_small_r1f python bench.py --no-large --steps 2 --warmup 1 > gpurun_out/ncu_lm_small_r1f.log 2>&1
tail -2 gpurun_out/ncu_lm_small_r1f.log | cut -c1-200
