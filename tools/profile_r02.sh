#!/bin/bash
# Run under gpurun (one GPU): round-2 evidence for profiles/.
#   (1) launch list of two forwards (25 launches each: per block row / column summaries / exchange+finalize / FFN,
#       + k_head_reduce), cold-cache and serialised: compare SHARES with bench.py's CUDA-event timing, not absolutes
#   (2) ncu --set full of one launch of each hot kernel, with source
#   (3) the default bench line
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 75 -c 50 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_colapply_ffn_ws|k_row_attn_ws|k_col_partial_ws|k_row_attn_combo' \
    -s 36 -c 4 -o $OUT/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
for k in k_colapply_ffn_ws k_row_attn_ws k_col_partial_ws k_row_attn_combo; do
  ncu -i $OUT/prof.ncu-rep --page source --csv -k regex:$k > $OUT/prof_source_$k.csv 2>/dev/null
done
PF_WS_PROF=1 timeout 300 python tools/ws_role_timing.py > $OUT/role_timing.txt 2>&1
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
ls -la $OUT | tail -12
