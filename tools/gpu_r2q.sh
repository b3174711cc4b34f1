set -x
B="--steps 100 --warmup 10 --no-lookup-roofline --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],4), "gemm", round(d["roofline"]["kernel_ms"],4), "warm", round(d["warm_bank"]["value"]))'
for r in 1 2; do
(cd _ab_old && timeout 600 python bench.py $B 2>/dev/null | python -c "$P" old)
timeout 600 python bench.py $B --no-config-blocks 2>/dev/null | python -c "$P" new
PICOPOSE_B200_EXACT_KEYS=1 timeout 600 python bench.py $B --no-config-blocks 2>/dev/null | python -c "$P" new_exact_keys
PICOPOSE_B200_CLUSTER=1 timeout 600 python bench.py $B --no-config-blocks 2>/dev/null | python -c "$P" new_cl1
done
(cd _ab_old && PICOPOSE_B200_CLUSTER=1 timeout 600 python bench.py $B 2>/dev/null | python -c "$P" old_cl1)
