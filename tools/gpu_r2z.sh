set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python tools/bench_lookup.py --radii 3 4 5 6 7 8 --json gpurun_out/r2z_lookup_sweep.json 2>&1 | tail -13
timeout 300 python tools/bench_lookup.py --once --radii 4 8 > /dev/null 2>&1 && timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:corr_lookup --csv --log-file gpurun_out/r2z_lookup_traffic.csv python tools/bench_lookup.py --once --radii 4 8 > gpurun_out/ncu_lookup.log 2>&1
timeout 300 python tools/bench_lookup.py --once --radii 4 --layouts tiled > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_banded -s 1 -c 1 -o gpurun_out/prof_lookup_tiled4_r2z -f python tools/bench_lookup.py --once --radii 4 --layouts tiled > gpurun_out/ncu_lk.log 2>&1; tail -2 gpurun_out/ncu_lk.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-config-blocks > gpurun_out/bench_r2z.json 2>/dev/null; tail -c 300 gpurun_out/bench_r2z.json
