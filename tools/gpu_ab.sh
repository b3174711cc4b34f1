#!/bin/bash
# Same-box A/B of the working tree against a committed revision (boxes of the pool differ by up to 8 % on the contraction,
# so a kernel change is only ever judged inside ONE gpurun call).  Prepare here, on the CPU box:
#   git worktree add _ab_prev <commit> && (cd _ab_prev && python -m picopose_b200.build) && python -m picopose_b200.build
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab.sh > gpurun_out/ab.log 2>&1; grep -v "^+" gpurun_out/ab.log'
#   git worktree remove --force _ab_prev          # it travels with every gpurun snapshot otherwise
set -x
line() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm TF', d['roofline'].get('achieved'))"; }
batch() { grep ours_issued | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1 batch ms', d['ours_issued']['ms_per_call'], d['ours_issued']['tflops'], 'cublas', d['cublas_8192']['tflops'])"; }
for i in 1 2; do
  timeout 300 python tools/bench_sustained.py --rounds 1 2>/dev/null | batch NEW
  (cd _ab_prev && timeout 300 python tools/bench_sustained.py --rounds 1 2>/dev/null | batch OLD)
done
for i in 1 2 3; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | tail -1 | line NEW
  (cd _ab_prev && timeout 300 python bench.py --steps 200 --warmup 20 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | tail -1 | line OLD)
done
