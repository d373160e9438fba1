#!/bin/bash
# A/B several builds in the same gpurun call (same box, same clocks):
#   [WORKLOAD=256x20x200] tools/ab_bench.sh REPS libA.so libB.so [libC.so ...]
R=$1; shift
for i in $(seq $R); do
  for L in "$@"; do
    PF_LIB=$L timeout 120 python bench.py --workload ${WORKLOAD:-200x1000} --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['roofline']['kernels']
print('$L', round(d['ms_per_step'], 2), {n: round(v['ms_per_step'], 2) for n, v in k.items()}, 'parity', d.get('checks', {}).get('parity_vs_oracle_fixture_max_rel'), 'MHz', d.get('clocks', {}).get('sm_mhz'))"
  done
done
