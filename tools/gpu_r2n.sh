set -x
N=${NGPU:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n_bench_${N}gpu.json 2> gpurun_out/r2n_bench_${N}gpu.err; echo rc=$?; tail -c 1200 gpurun_out/r2n_bench_${N}gpu.json; tail -3 gpurun_out/r2n_bench_${N}gpu.err
