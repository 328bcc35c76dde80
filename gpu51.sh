python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "eval" 2>&1 | tail -3
for v in 0 4; do
  echo "== variant $v"
  EZPZ_B200_ASM_VARIANT=$v python profiles/large_bench.py 160000 30 | grep "assemble_large_kernel" | cut -c1-220
  EZPZ_B200_ASM_VARIANT=$v python profiles/large_bench.py 77000 30 | grep "assemble_large_kernel (J in CSC" | cut -c1-220
done
