#!/bin/bash
# One gpurun call = several measurements; everything lands in gpurun_out/ (merged back).  usage: tools/gpu_call.sh <tag> <steps...>
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for step in "$@"; do
  case $step in
    probe)   timeout 120 tools/probe/umma_mn_probe > $out/probe.txt 2>&1; echo "probe rc=$?" >> $out/probe.txt ;;
    tcattn)  timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_attention or stage_taps or synthetic_shapes" -s > $out/pytest_tcattn.txt 2>&1; echo "rc=$?" >> $out/pytest_tcattn.txt ;;
    full)    timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size_values" -s > $out/pytest_full.txt 2>&1; echo "rc=$?" >> $out/pytest_full.txt ;;
    gputests) timeout 2400 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.txt 2>&1; echo "rc=$?" >> $out/pytest_gpu.txt ;;
    bench_cc) PF_COL_IMPL=cc timeout 600 python bench.py --steps 5 --no-cpu-baseline > $out/bench_cc.json 2> $out/bench_cc.err ;;
    bench_tc) timeout 240 python bench.py --steps 5 --no-cpu-baseline > $out/bench_tc.json 2> $out/bench_tc.err ;;
    bench_rowtma) PF_ROW_IMPL=tma timeout 240 python bench.py --steps 5 --no-cpu-baseline > $out/bench_rowtma.json 2> $out/bench_rowtma.err ;;
    bench_tc1) PF_COL_IMPL=tc1 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $out/bench_tc1.json 2> $out/bench_tc1.err ;;
    ncu_col) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_col_partial_ws -s 3 -c 1 -o $out/prof_col python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_col.log 2>&1 ;;
    ncu_row) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_row_attn_ws -s 3 -c 1 -o $out/prof_row python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_row.log 2>&1 ;;
    ncu_ffn) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_colapply_ffn_ws -s 3 -c 1 -o $out/prof_ffn python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_ffn.log 2>&1 ;;
    bench)   timeout 900 python bench.py > $out/bench.json 2> $out/bench.err ;;
    diag)    timeout 600 python tools/diag_variants.py > $out/diag.txt 2>&1 ;;
    diag_rows) timeout 1200 python tools/diag_rows.py > $out/diag_rows.txt 2>&1 ;;
    sanit_col) PF_COL_IMPL=tc PF_ROW_IMPL=tma timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python tools/diag_variants.py one 100 500 1 tc tma > $out/sanit_col.txt 2>&1 ;;
    coredump) ( export CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_ENABLE_LIGHTWEIGHT_COREDUMP=1 CUDA_COREDUMP_FILE=/tmp/pf_core_%p CUDA_COREDUMP_SHOW_PROGRESS=0;
               timeout 600 python tools/diag_variants.py one ${DIAG_SHAPE:-100 500 1} tc tma > $out/coredump_run.txt 2>&1 );
             for c in /tmp/pf_core_*; do [ -f "$c" ] && timeout 300 cuda-gdb -batch -ex "target cudacore $c" -ex "info cuda kernels" -ex "info cuda lanes" -ex "bt" -ex "x/6i \$pc" > $out/coredump_gdb.txt 2>&1 && break; done ;;
    gdbrun)  timeout 900 cuda-gdb -batch -ex run -ex "info cuda kernels" -ex bt -ex "x/6i \$pc" --args python tools/diag_variants.py one ${DIAG_SHAPE:-100 500 1} tc tma > $out/gdbrun.txt 2>&1 ;;
    head)    timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_head or stage_taps or synthetic_shapes or tcgen05_ffn or forward_idx_equals" -s > $out/pytest_head.txt 2>&1; echo "rc=$?" >> $out/pytest_head.txt ;;
    lean)    PF_LIB=build/libpf_lean.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_head or stage_taps or synthetic_shapes or forward_idx_equals" -s > $out/pytest_lean.txt 2>&1; echo "rc=$?" >> $out/pytest_lean.txt ;;
    ab)      tools/ab_bench.sh ${AB_REPS:-2} phyloformer_b200/libpf_sm100.so $AB_LIBS > $out/ab.txt 2>&1 ;;
    sharded) timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -s > $out/pytest_sharded.txt 2>&1; echo "rc=$?" >> $out/pytest_sharded.txt ;;
    ab_exch) ( tools/ab_bench.sh 2 phyloformer_b200/libpf_sm100.so; echo split; PF_EXCH_IMPL=split tools/ab_bench.sh 2 phyloformer_b200/libpf_sm100.so; echo 256x20x200; WORKLOAD=256x20x200 tools/ab_bench.sh 2 phyloformer_b200/libpf_sm100.so; echo split; WORKLOAD=256x20x200 PF_EXCH_IMPL=split tools/ab_bench.sh 2 phyloformer_b200/libpf_sm100.so ) > $out/ab_exch.txt 2>&1 ;;
    ab_env)  ( for i in 1 2; do tools/ab_bench.sh 1 phyloformer_b200/libpf_sm100.so; echo "$AB_ENV"; env $AB_ENV tools/ab_bench.sh 1 phyloformer_b200/libpf_sm100.so; done; if [ -n "$AB_WORKLOAD2" ]; then echo $AB_WORKLOAD2; WORKLOAD=$AB_WORKLOAD2 tools/ab_bench.sh 1 phyloformer_b200/libpf_sm100.so; echo "$AB_ENV"; WORKLOAD=$AB_WORKLOAD2 env $AB_ENV tools/ab_bench.sh 1 phyloformer_b200/libpf_sm100.so; fi ) > $out/ab_env.txt 2>&1 ;;
    sanit)   timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_smoke.py > $out/sanitizer.txt 2>&1; echo "rc=$?" >> $out/sanitizer.txt ;;
    quick)   timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_attention or fused_head or stage_taps or cuda_graph or cli_end or reference_cases or block0_combo or duplicate or native_library" -s > $out/pytest_quick.txt 2>&1; echo "rc=$?" >> $out/pytest_quick.txt ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "rc=$?" >> $out/smoke.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
tail -n 5 $out/*.txt 2>/dev/null
for f in $out/bench*.json; do [ -f "$f" ] && python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[1], "ms/step %.2f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], {a:round(b["ms_per_step"],2) for a,b in k.items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
