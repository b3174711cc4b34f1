# same-box A/B of the working tree against the committed HEAD checked out in _ab_prev/ (both built beforehand)
set -x
for i in 1 2; do
  timeout 300 python tools/bench_sustained.py --rounds 1 2>/dev/null | grep ours_issued | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('NEW batch ms', d['ours_issued']['ms_per_call'], d['ours_issued']['tflops'], 'cublas', d['cublas_8192']['tflops'])"
  (cd _ab_prev && timeout 300 python tools/bench_sustained.py --rounds 1 2>/dev/null | grep ours_issued | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('OLD batch ms', d['ours_issued']['ms_per_call'], d['ours_issued']['tflops'], 'cublas', d['cublas_8192']['tflops'])")
done
for i in 1 2; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NEW ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF', d['roofline'].get('achieved'))"
  (cd _ab_prev && timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('OLD ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF', d['roofline'].get('achieved'))")
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matching" 2>&1 | tail -3
