set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r1f_n2.json 2> gpurun_out/bench_r1f_n2.err; tail -5 gpurun_out/bench_r1f_n2.err
cat gpurun_out/bench_r1f_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | tail -1 | cut -c1-300
