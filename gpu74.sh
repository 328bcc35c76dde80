python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err
tail -c 300 gpurun_out/bench_r1k.err
python bench.py --impl reference > gpurun_out/bench_ref_r1k.json 2>> gpurun_out/bench_r1k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r1k.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lm_large_kernel -c 1 -o gpurun_out/prof_lm_large_r1k -f python profiles/lm_large_once.py 77000 > gpurun_out/ncu_lm_large_r1k.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -1
