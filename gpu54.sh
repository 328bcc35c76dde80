python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -5
L=ezpz_b200/_lib
cp $L/libezpz_b200.so /tmp/new.so
for which in new old; do
    if [ $which = old ]; then cp $L/libezpz_b200_old.so $L/libezpz_b200.so; else cp /tmp/new.so $L/libezpz_b200.so; fi
    echo "== $which"
    python profiles/large_bench.py 160000 30 | grep "assemble_large_kernel (J in CSC" | cut -c1-220
    python profiles/large_bench.py 77000 30 --solve | grep "assemble_large_kernel (J in CSC\|solve_one" | cut -c1-220
done
cp /tmp/new.so $L/libezpz_b200.so
EZPZ_B200_DEBUG=1 python profiles/lm_large_once.py 77000 2>&1 | grep lm_large_kernel | tail -1
ncu --set full --import-source on --clock-control none -k regex:assemble_large -s 3 -c 1 -o gpurun_out/prof_asm_r1l -f python profiles/large_bench.py 160000 5 > gpurun_out/ncu_asm_r1l.log 2>&1
