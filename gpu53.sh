python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "eval" 2>&1 | tail -3
L=ezpz_b200/_lib
cp $L/libezpz_b200.so /tmp/new.so
for round in 1 2; do
  for which in new old; do
    if [ $which = old ]; then cp $L/libezpz_b200_old.so $L/libezpz_b200.so; else cp /tmp/new.so $L/libezpz_b200.so; fi
    echo "== $which (round $round)"
    python profiles/large_bench.py 160000 30 | grep "assemble_large_kernel (J in CSC" | cut -c1-220
    python profiles/large_bench.py 77000 30 | grep "assemble_large_kernel (J in CSC" | cut -c1-220
  done
done
cp /tmp/new.so $L/libezpz_b200.so
ncu --set full --import-source on --clock-control none -k regex:assemble_large -s 3 -c 1 -o gpurun_out/prof_asm_r1k -f python profiles/large_bench.py 160000 5 > gpurun_out/ncu_asm_r1k.log 2>&1
