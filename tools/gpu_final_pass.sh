#!/bin/bash
# One gpurun call that re-validates and re-measures everything on a single B200: GPU tests, smoke, both bench arms,
# the lookup sweep in both volume layouts with ncu-measured DRAM traffic, the sustained-load comparison with cuBLAS, the
# per-function comparison with torch eager, the ncu launch list and an `ncu --set full` capture of one whole timed step
# (outputs in gpurun_out/, tagged).
#   TAG=r2x gpurun --timeout 2400 -- 'TAG=r2x bash tools/gpu_final_pass.sh'
# Multi-GPU lines (N = 2, 4, 8; one JSON line on stdout, NCCL's banner goes to stderr):
#   gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#       --master-port 29500 bench.py --gpus N > gpurun_out/bench_${TAG}_${N}gpu.json'
#   two real ranks over CUDA IPC: gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_two_ranks.py -m gpu -q'
# Microbenchmarks behind DESIGN 3.1 / 3.4: tools/microbench/{dram_gran,partition,subset_stream}.cu (+ run.sh).
set -x
cd "$(dirname "$0")/.."
T=${TAG:-final}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 1500 gpurun_out/bench_${T}.json
timeout 900 python bench.py --no-config-blocks --no-lookup-roofline --no-cpu-baseline > gpurun_out/bench_${T}_200steps.json 2> /dev/null; tail -c 600 gpurun_out/bench_${T}_200steps.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${T}.json 2> gpurun_out/bench_ref_${T}.err; tail -c 700 gpurun_out/bench_ref_${T}.json
timeout 300 python tools/bench_lookup.py --json gpurun_out/lookup_sweep_${T}.json 2>&1 | tail -11
timeout 300 python tools/bench_lookup.py --once --radii 4 8 > /dev/null 2>&1 && timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:corr_lookup --csv --log-file gpurun_out/${T}_lookup_traffic.csv python tools/bench_lookup.py --once --radii 4 8 > gpurun_out/ncu_lookup.log 2>&1
timeout 600 python tools/bench_sustained.py --json gpurun_out/sustained_${T}.json 2>&1 | tail -3
timeout 600 python tools/bench_vs_torch_eager.py --json gpurun_out/vs_torch_eager_${T}.json 2>&1 | tail -12 | cut -c1-200
timeout 600 python tools/bench_stage3.py --batch 16 --json gpurun_out/stage3_${T}.json 2>&1 | tail -4 | cut -c1-400
timeout 300 python bench.py --steps 2 --warmup 3 --no-config-blocks --no-lookup-roofline --no-cpu-baseline > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 3 --no-config-blocks --no-lookup-roofline --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 28 -c 7 -o gpurun_out/prof_step_${T} -f python bench.py --steps 2 --warmup 3 --no-config-blocks --no-lookup-roofline --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_gemm_kernel -s 1 -c 1 -o gpurun_out/prof_gemm_batch_${T} -f python tools/bench_sustained.py --seconds 0.05 --rounds 1 > gpurun_out/ncu_gemm_batch.log 2>&1; tail -2 gpurun_out/ncu_gemm_batch.log
