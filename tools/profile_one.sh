#!/bin/bash
# Run under gpurun: full ncu capture (with source) of one launch of one kernel.   tools/profile_one.sh <tag> <kernel regex> [skip]
set -u
TAG=$1; K=$2; SKIP=${3:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o $OUT/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
rm -f $OUT/prof.ncu-rep
ls -la $OUT
