set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; tail -3 gpurun_out/bench_r1h.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1h.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['roofline']['traffic'])
print(json.dumps(d['large_system'], indent=1))
"
EZPZ_B200_DEBUG=12 python profiles/lm_large_once.py 77000 2>&1 | grep -E "stage +[0-9]+ panels|lm_large|sparse_direct\]" > gpurun_out/stages_r1h.log; tail -3 gpurun_out/stages_r1h.log
