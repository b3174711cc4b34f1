set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for i in 1 2 3; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL   ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF', d['roofline'].get('achieved'), 'warm', d['warm_bank']['value'])"
  PICOPOSE_B200_PDL=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PLAIN ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF', d['roofline'].get('achieved'), 'warm', d['warm_bank']['value'])"
done
