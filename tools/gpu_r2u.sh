set -x
for jb in 5 6 7; do PICOPOSE_LOOKUP_JB=$jb timeout 300 python tools/bench_lookup.py --radii 4 2>&1 | tail -2 | sed "s/^/JB=$jb /"; done
timeout 300 python tools/bench_lookup.py --radii 3 4 5 6 7 8 --json gpurun_out/r2u_lookup_sweep.json 2>&1 | tail -12
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
