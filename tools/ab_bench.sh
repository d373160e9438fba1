#!/bin/bash
# A/B two builds in the same gpurun call (same box, same clocks):  tools/ab_bench.sh libA.so libB.so [reps]
A=$1; B=$2; R=${3:-2}
for i in $(seq $R); do
  for L in $A $B; do
    PF_LIB=$L timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['roofline']['kernels']
print('$L', round(d['ms_per_step'], 2), {n: round(v['ms_per_step'], 2) for n, v in k.items()})"
  done
done
