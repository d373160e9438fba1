#!/bin/bash
# Run under gpurun: launch list of one bench step + full ncu capture of the dominant kernels.
#   tools/profile_gpu.sh [tag]
# Outputs land in gpurun_out/ (scratch); copy the summaries you want kept into profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
# (2) full capture of the FFN kernel and the two streaming kernels (one launch each, block 2)
ncu --set full --clock-control none --import-source on -k regex:'k_colapply_ffn_tc|k_row_attn|k_col_partial' \
    -s 6 -c 3 -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ls -la $OUT
