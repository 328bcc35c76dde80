L=ezpz_b200/_lib
for round in 1 2; do
for v in H J K; do
  cp $L/lib$v.so $L/libezpz_b200.so
  echo "== $v (round $round)"
  EZPZ_B200_DEBUG=1 python profiles/lm_large_once.py 77000 2>&1 | grep lm_large_kernel | tail -1 | cut -c30-
done
done
cp $L/libK.so $L/libezpz_b200.so
python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
