python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "eval" 2>&1 | tail -3
python profiles/large_bench.py 160000 20 | grep assemble | cut -c1-200
python profiles/large_bench.py 77000 20 | grep assemble | cut -c1-200
