set -x
N=${NGPU:-8}
run() { # dir tag extra
  (cd $1 && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $4 bench.py --gpus $N --steps 50 --warmup 5 --no-lookup-roofline --no-cpu-baseline $3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$2', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'host_e2e_ms', round(d['e2e']['host_enqueue_ms_per_step'],3), 'gemm_ms', round(d['roofline']['kernel_ms'],4))")
}
run _ab_old old "" 29531
run . new "--no-config-blocks" 29532
run _ab_old old "" 29533
run . new "--no-config-blocks" 29534
