set -x
timeout 800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 600 python tools/bench_vs_torch_eager.py --json gpurun_out/r2k_vs_torch_eager.json 2>&1 | tail -14
