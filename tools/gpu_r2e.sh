set -x
timeout 600 python tools/bench_sustained.py --json gpurun_out/r2e_sustained.json 2>&1 | tail -6
timeout 600 python tools/bench_sustained.py --detections 8 --rounds 1 2>&1 | tail -3
# ncu of one long-batch contraction launch: tensor pipe, L2, DRAM, issue
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_gemm_kernel -s 1 -c 1 -o gpurun_out/r2e_gemm_batch -f python tools/bench_sustained.py --detections 8 --seconds 0.05 --rounds 1 > gpurun_out/r2e_ncu.log 2>&1; tail -3 gpurun_out/r2e_ncu.log
