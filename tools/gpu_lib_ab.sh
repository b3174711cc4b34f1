#!/bin/bash
# Same-box A/B of two builds of the library: the in-tree one against $ALT (another libpicopose_b200.so, e.g. built from
# the same sources with an extra -D into picopose_b200/lib_alt/; loaded through PICOPOSE_B200_LIB), GPU tests on $ALT
# first, then the short bench alternating.
#   gpurun --timeout 600 -- 'ALT=$PWD/picopose_b200/lib_alt/libpicopose_b200.so bash tools/gpu_lib_ab.sh'
cd "$(dirname "$0")/.."
ALT=${ALT:?path of the other build}
PICOPOSE_B200_LIB=$ALT timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
rm -f gpurun_out/lib_ab.jsonl
for v in base alt base alt base alt; do
  if [ $v = alt ]; then export PICOPOSE_B200_LIB=$ALT; else unset PICOPOSE_B200_LIB; fi
  timeout 200 python bench.py --no-config-blocks --no-lookup-roofline --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r={'lib':'$v','ms_per_step':d['ms_per_step'],'value':d['value'],'e2e':d['e2e']['value'],'gemm_ms':d['roofline']['kernel_ms'],'warm_bank':d['warm_bank']['value'],'clocks':d['clocks']}
print(json.dumps(r))" | tee -a gpurun_out/lib_ab.jsonl
done
