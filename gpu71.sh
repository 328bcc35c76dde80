python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
for round in 1 2; do
EZPZ_B200_DEBUG=1 python profiles/lm_large_once.py 77000 2>&1 | grep lm_large_kernel | tail -1 | cut -c30-
done
python profiles/large_bench.py 160000 30 | grep "assemble_large_kernel (J in CSC" | cut -c1-220
