#!/bin/bash
# 2-GPU call: two-rank IPC tests, then bench.py --gpus 2 with the query gather issued through separate library calls
# (PICOPOSE_B200_FUSED_GATHER=0) against the one-call form (pp_xchg_push_signal), alternating; RUNS="1" for one run.
# The push-chain microbenchmark per stream count: tools/microbench/push_chain.py.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_push_ab.sh'
set -x
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_two_ranks.py "tests/test_gpu_parity.py::test_peer_gather_primitives_two_ranks_on_one_gpu" -m gpu -q 2>&1 | tail -3
rm -f gpurun_out/bench_gather_ab_2gpu.jsonl
for v in ${RUNS:-0 1 0 1}; do
PICOPOSE_B200_FUSED_GATHER=$v timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2958$v bench.py --gpus 2 --no-config-blocks --no-cpu-baseline 2>/dev/null | tail -1 | tee -a gpurun_out/bench_gather_ab_2gpu.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('FUSED=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'host', d['e2e']['host_enqueue_ms_per_step'], d['sharded_equals_single_gpu'][:30])
"
done
