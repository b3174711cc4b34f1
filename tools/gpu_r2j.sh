set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_tma -s 1 -c 1 -o gpurun_out/r2j_lookup_tma -f python tools/bench_lookup.py --once --radii 4 --layouts tiled > gpurun_out/r2j_ncu.log 2>&1; tail -2 gpurun_out/r2j_ncu.log
