#!/bin/bash
# Multi-GPU measurements in one gpurun --gpus N call.  usage: tools/multi_gpu_call.sh <tag> <N> <steps...>
tag=$1; N=$2; shift 2
out=gpurun_out/$tag; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29540
for step in "$@"; do
  port=$((port + 1))
  case $step in
    check)        timeout 600 $TR --master-port $port tools/check_sharded.py > $out/check_fused.txt 2>&1; echo "rc=$?" >> $out/check_fused.txt ;;
    check_split)  PF_EXCH_IMPL=split timeout 600 $TR --master-port $port tools/check_sharded.py > $out/check_split.txt 2>&1; echo "rc=$?" >> $out/check_split.txt ;;
    bench)        timeout 600 $TR --master-port $port bench.py --gpus $N --steps 10 > $out/bench_n$N.json 2> $out/bench_n$N.err ;;
    bench_split)  PF_EXCH_IMPL=split timeout 600 $TR --master-port $port bench.py --gpus $N --steps 10 > $out/bench_split_n$N.json 2> $out/bench_split_n$N.err ;;
    bench_nccl)   timeout 600 $TR --master-port $port bench.py --gpus $N --steps 10 --exchange nccl > $out/bench_nccl_n$N.json 2> $out/bench_nccl_n$N.err ;;
    bench_500)    timeout 900 $TR --master-port $port bench.py --gpus $N --steps 5 --workload 500x500 > $out/bench_500x500_n$N.json 2> $out/bench_500x500_n$N.err ;;
    bench_cfg4)   timeout 600 $TR --master-port $port bench.py --gpus $N --steps 5 --workload 256x20x200 > $out/bench_256x20x200_n$N.json 2> $out/bench_256x20x200_n$N.err ;;
    pytest)       timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -s > $out/pytest_sharded.txt 2>&1; echo "rc=$?" >> $out/pytest_sharded.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
tail -n 3 $out/*.txt 2>/dev/null
for f in $out/bench*.json; do [ -f "$f" ] && python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[1], "ms/step %.3f"%d["ms_per_step"], "value %.3e"%d["value"], "e2e %.3f"%d["e2e"]["ms_per_step"], {a:round(b["ms_per_step"],3) for a,b in k.items()}, d.get("checks"), d.get("msas_100x500"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
