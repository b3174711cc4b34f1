#!/bin/bash
# One gpurun call that re-validates and re-measures everything on a single B200: GPU tests, smoke, both bench arms,
# the lookup sweep, the ncu launch list and an `ncu --set full` capture of one whole timed step (outputs in gpurun_out/).
#   gpurun --timeout 1800 -- 'bash tools/gpu_final_pass.sh'
# Multi-GPU lines (N = 2, 4, 8; one JSON line on stdout, NCCL's banner goes to stderr):
#   gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#       --master-port 29500 bench.py --gpus N > gpurun_out/bench_${TAG}_${N}gpu.json'
#   ... tools/bench_configs.py --detections 512 --json gpurun_out/config5_8gpu_${TAG}.json        (configs[4], 8 GPUs)
# Microbenchmarks behind DESIGN 3.1 / 3.4: tools/microbench/{dram_gran,partition,subset_stream}.cu (+ run.sh).
set -x
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_${TAG:-final}.json 2> gpurun_out/bench_${TAG:-final}.err; tail -c 3000 gpurun_out/bench_${TAG:-final}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG:-final}.json 2> gpurun_out/bench_ref_${TAG:-final}.err; tail -c 1500 gpurun_out/bench_ref_${TAG:-final}.json
timeout 300 python tools/bench_lookup.py --json gpurun_out/lookup_sweep_${TAG:-final}.json 2>&1 | tail -6
timeout 300 python bench.py --steps 2 --warmup 3 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG:-final}.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 28 -c 7 -o gpurun_out/prof_step_${TAG:-final} -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
