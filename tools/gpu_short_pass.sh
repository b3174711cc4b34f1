set -x
T=r2i
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 400 gpurun_out/bench_${T}.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${T}.json 2> gpurun_out/bench_ref_${T}.err; tail -c 300 gpurun_out/bench_ref_${T}.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 3 --no-config-blocks --no-lookup-roofline --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
