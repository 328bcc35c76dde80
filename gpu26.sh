python profiles/sweep_mixed.py > gpurun_out/sweep_mixed_n1.jsonl 2> gpurun_out/sweep_mixed_n1.err
tail -3 gpurun_out/sweep_mixed_n1.err
cut -c1-330 gpurun_out/sweep_mixed_n1.jsonl
