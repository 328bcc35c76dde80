python bench.py > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err
tail -c 600 gpurun_out/bench_r1j.err
python bench.py --impl reference > gpurun_out/bench_ref_r1j.json 2>> gpurun_out/bench_r1j.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r1j.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -2
