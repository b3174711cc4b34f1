set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
for i in 1 2; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('FUSED ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'gemm', d['roofline'].get('achieved'))"
  PICOPOSE_B200_FUSED_FINALIZE=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('SEPARATE ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'gemm', d['roofline'].get('achieved'))"
  (cd _ab_old && timeout 300 python bench.py --steps 200 --warmup 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('R1 ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'gemm', d['roofline'].get('achieved'))")
done
