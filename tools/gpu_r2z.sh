set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "lookup or nonfinite or tiled or flow" 2>&1 | tail -3
timeout 300 python tools/bench_lookup.py --radii 3 4 5 6 7 8 2>&1 | tail -13
