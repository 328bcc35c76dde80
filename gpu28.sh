N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 profiles/sweep_mixed.py > gpurun_out/sweep_mixed_n$N.jsonl 2> gpurun_out/sweep_mixed_n$N.err
tail -2 gpurun_out/sweep_mixed_n$N.err
python - <<PY
import json
for l in open('gpurun_out/sweep_mixed_n$N.jsonl'):
    d=json.loads(l); print(d['n_gpus'], d['batch_total'], '%.3f ms'%d['ms_per_pass'], '%.1f M/s'%(d['solves_per_s']/1e6), 'mismatch', d['parity_mismatches_in_sample'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r1f_n$N.json
python -c "import json; d=json.load(open('gpurun_out/bench_r1f_n$N.json')); print('bench', d['n_gpus'], '%.1f M/s'%(d['value']/1e6), 'e2e %.1f M/s'%(d['e2e']['value']/1e6), d['ms_per_step'])"
