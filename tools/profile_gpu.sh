#!/bin/bash
# Run under gpurun: launch list of one bench step + full ncu capture of the dominant kernels.
#   tools/profile_gpu.sh [tag]
# Outputs land in gpurun_out/ (scratch); tools/ncu_summary.py turns them into the text
# summaries committed under profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 93 -c 62 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
# (2) full capture of one launch of each hot kernel (block 1 of the first timed step)
ncu --set full --clock-control none --import-source on -k regex:'k_colapply_ffn|k_row_attn|k_col_partial' \
    -s 12 -c 3 -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv -k regex:k_colapply_ffn > $OUT/prof_${TAG}_source.csv 2>/dev/null
PF_WS_PROF=1 python tools/ws_role_timing.py > $OUT/role_timing_$TAG.txt 2>&1
ls -la $OUT | tail -8
