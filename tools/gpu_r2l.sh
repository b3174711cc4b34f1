set -x
timeout 600 python -m pytest tests/test_gpu_two_ranks.py -q -m gpu -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err; echo rc=$?; tail -c 3500 gpurun_out/r2l_bench_2gpu.json; tail -5 gpurun_out/r2l_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2l_bench_ref_2gpu.json 2>gpurun_out/r2l_bench_ref_2gpu.err; echo rc=$?; tail -c 600 gpurun_out/r2l_bench_ref_2gpu.json
