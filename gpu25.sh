python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
python profiles/large_bench.py 160000 20 | grep assemble | cut -c1-200
python profiles/large_bench.py 77000 20 | grep assemble | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"lm_large" -c 1 -o gpurun_out/prof_lm_large_r1f python profiles/lm_large_once.py > gpurun_out/ncu_lm_large_r1f.log 2>&1
tail -2 gpurun_out/ncu_lm_large_r1f.log | cut -c1-200
