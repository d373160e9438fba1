#!/bin/bash
# Run under gpurun: full ncu capture (with source) of one tcgen05 FFN launch.
set -u
TAG=${1:-ffn}
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:'k_colapply_ffn' \
    -s 2 -c 1 -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/prof_${TAG}_source.csv 2>/dev/null
ls -la $OUT | tail -5
