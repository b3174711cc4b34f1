set -x
timeout 800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python tools/bench_lookup.py --json gpurun_out/r2i_lookup_sweep.json 2>&1 | tail -12
timeout 300 python tools/bench_lookup.py --once --radii 4 8 > /dev/null 2>&1 && timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:corr_lookup --csv --log-file gpurun_out/r2i_lookup_traffic.csv python tools/bench_lookup.py --once --radii 4 8 > gpurun_out/r2i_ncu.log 2>&1; tail -2 gpurun_out/r2i_ncu.log
