python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err
tail -c 300 gpurun_out/bench_r1l.err
python bench.py --impl reference > gpurun_out/bench_ref_r1l.json 2>> gpurun_out/bench_r1l.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r1l.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -1
