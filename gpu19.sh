python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -5
python profiles/large_bench.py 77000 20
python profiles/large_bench.py 160000 20
EZPZ_B200_FORCE_PCG=1 python profiles/large_bench.py 160000 20 --solve
