set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -3 gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1a.json 2>&1; cat gpurun_out/bench_ref_r1a.json
nproc; lscpu | grep "Model name"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lm_small -s 3 -c 2 -o gpurun_out/prof_lm_small_r1a python bench.py --steps 2 --warmup 1 > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log
ls -la gpurun_out
