ncu --set full --clock-control none --import-source on -k regex:"assemble_large" -s 2 -c 1 -o gpurun_out/prof_large_r1d python profiles/large_bench.py 77000 2 > gpurun_out/ncu_large_r1d.log 2>&1
tail -3 gpurun_out/ncu_large_r1d.log
