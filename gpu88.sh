python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
for r in 1 2; do
EZPZ_B200_DEBUG=1 python profiles/lm_large_once.py 77000 2>&1 | grep "lm_large_kernel" | tail -1 | cut -c1-200
done
