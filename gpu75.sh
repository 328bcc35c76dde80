python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
for round in 1 2; do
for v in 1 0; do
echo "== sort $v"
EZPZ_B200_STAGE_SORT=$v EZPZ_B200_DEBUG=12 python profiles/lm_large_once.py 77000 2>&1 | grep "stage   [0-9] \|lm_large_kernel" | tail -11 | cut -c1-200
done
done
