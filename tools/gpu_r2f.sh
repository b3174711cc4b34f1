set -x
timeout 800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 600 python tools/bench_sustained.py --json gpurun_out/r2f_sustained.json 2>&1 | tail -3
timeout 600 python tools/bench_sustained.py --detections 8 --rounds 1 2>&1 | tail -2
timeout 600 python bench.py --steps 100 --warmup 10 --no-config-blocks --no-lookup-roofline --no-cpu-baseline > gpurun_out/r2f_bench.json 2>gpurun_out/r2f_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['warm_bank']['value'])"
timeout 900 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:match_gemm_kernel -s 1 -c 1 python tools/bench_sustained.py --detections 8 --seconds 0.05 --rounds 1 2>&1 | grep -E "dram__|gpu__time|tensor|alu|inst_exec" 
