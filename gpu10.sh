set -x
python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -15
python profiles/large_bench.py 77000 20 --solve
ncu --set full --clock-control none --import-source on -k regex:"assemble_large" -s 1 -c 1 -o gpurun_out/prof_large_r1c python profiles/large_bench.py 77000 1 > gpurun_out/ncu_large_r1c.log 2>&1
tail -3 gpurun_out/ncu_large_r1c.log
