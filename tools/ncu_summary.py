#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump and (optionally) its `--page source --csv`
companion: headline counters, stall mix, and per-role instruction/sample shares.

    python tools/ncu_summary.py gpurun_out/prof_X_raw.csv [gpurun_out/prof_X_source.csv]
"""
import collections
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[ix["Kernel Name"]][:70])
        for k in KEYS:
            if k in ix:
                print(f"  {k} = {r[ix[k]]} {units[ix[k]]}")
        st = {h: float(r[i]) for h, i in ix.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
        for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:22s} {v:.2f}")
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr, data = rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
        samp = sum(int(r[ix["# Samples"]]) for r in data)
        print(f"source page: {tot} warp instructions, {samp} samples, {len(data)} SASS lines")
        op, ops = collections.Counter(), collections.Counter()
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            o = m.group(2).split(".")[0] if m else "?"
            op[o] += int(r[ix["Instructions Executed"]])
            ops[o] += int(r[ix["# Samples"]])
        for o, c in op.most_common(24):
            print(f"  {o:10s} {100 * c / tot:6.2f}% inst {100 * ops[o] / samp:6.2f}% samples")
        # hottest SASS lines by samples
        hot = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:25]
        for r in hot:
            print(f"  {int(r[ix['# Samples']]):7d}  {r[ix['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
