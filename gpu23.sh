for c in 3 default 3 default 5; do
  if [ $c = default ]; then unset EZPZ_B200_CHUNKS; else export EZPZ_B200_CHUNKS=$c; fi
  python bench.py --no-large --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $c', 'value %.1fM'%(d['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6), 'e2e ms %.3f'%d['e2e']['ms_per_step'])"
done
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv
