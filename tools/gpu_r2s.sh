set -x
B="--steps 4 --warmup 3 --no-lookup-roofline --no-cpu-baseline"
for d in ${DIRS:-_ab_old _ab_mid _ab_mida .}; do
  X=""; if [ "$d" = "." ] || [ -f $d/.has_blocks ]; then X="--no-config-blocks"; fi
  (cd $d && timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:match_gemm_kernel -c 8 python bench.py $B $X 2>&1 | grep -E "gpu__time|cycles_elapsed" | awk '{print $NF}' | paste - - | tail -4 | sed "s|^|$d |")
done
