set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -3 gpurun_out/bench_r1f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print(json.dumps(d['large_system'], indent=1))
"
python __graft_entry__.py --smoke 2>&1 | tail -2
