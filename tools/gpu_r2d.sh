set -x
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo rc=$?; tail -c 6000 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; tail -c 1500 gpurun_out/r2d_bench_ref.json
