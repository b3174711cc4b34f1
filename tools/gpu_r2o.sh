set -x
timeout 800 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/bench_lookup.py --json gpurun_out/r2o_lookup_sweep.json 2>&1 | tail -11
timeout 300 python tools/bench_lookup.py --batch 64 --levels 3 --radii 2 4 2>&1 | tail -5
