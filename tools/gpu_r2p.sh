set -x
timeout 800 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 10 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new N=1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['warm_bank']['value'])"
(cd _ab_old && timeout 600 python bench.py --steps 100 --warmup 10 --no-lookup-roofline --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old N=1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['warm_bank']['value'])")
timeout 600 python bench.py --steps 100 --warmup 10 --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new N=1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['warm_bank']['value'])"
timeout 600 python tools/bench_match_shapes.py 2>&1 | tail -12 | cut -c1-250
timeout 300 python tools/bench_sustained.py --rounds 1 2>&1 | tail -1 | cut -c1-400
