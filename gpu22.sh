python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for c in default 1 2 3 4 6 8; do
  if [ $c = default ]; then unset EZPZ_B200_CHUNKS; else export EZPZ_B200_CHUNKS=$c; fi
  python bench.py --no-large --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $c', 'value %.1fM'%(d['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6), 'e2e ms %.3f'%d['e2e']['ms_per_step'])"
done
