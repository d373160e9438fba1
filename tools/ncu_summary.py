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
        # the source page holds one section per captured kernel: "Kernel Name" row, header row, SASS rows
        secs, cur = [], None
        for r in csv.reader(open(sys.argv[2])):
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1], "hdr": None, "data": []}
                secs.append(cur)
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = r
            elif cur is not None and r:
                cur["data"].append(r)
        seen = set()
        for sec in secs:
            if sec["name"] in seen or not sec["data"]:
                continue
            seen.add(sec["name"])
            source_section(sec)


def source_section(sec):
    ix = {h: i for i, h in enumerate(sec["hdr"])}
    data = sec["data"]
    ne = lambda r: int(r[ix["Instructions Executed"]] or 0)   # noqa: E731
    ns = lambda r: int(r[ix["# Samples"]] or 0)               # noqa: E731
    tot, samp = sum(map(ne, data)), sum(map(ns, data))
    print(f"== source page: {sec['name'][:60]}")
    print(f"  {tot} warp instructions, {samp} samples, {len(data)} SASS lines")
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += ne(r)
        ops[o] += ns(r)
    for o, c in op.most_common(24):
        print(f"  {o:10s} {100 * c / max(tot, 1):6.2f}% inst {100 * ops[o] / max(samp, 1):6.2f}% samples")
    # mbarrier polling: a try_wait loop is the run of SASS lines that executed exactly as often as
    # its TRYWAIT; everything beyond one pass per entry (the line before the loop) is spinning.
    # NOTE: this page comes from an instrumented replay pass whose timing differs from the
    # real run; compare the total with smsp__inst_executed.sum of the raw page.
    spin = 0
    for k, r in enumerate(data):
        if "TRYWAIT" in r[ix["Source"]] and k > 0:
            body = 0
            while k + body < len(data) and ne(data[k + body]) == ne(r):
                body += 1
            spin += max(0, ne(r) - ne(data[k - 1])) * body
    print(f"  mbarrier polling (extra try_wait iterations): {spin} instructions = {100 * spin / max(tot, 1):.1f}% of this pass"
          f" ({tot - spin} without them)")
    # warp-specialised kernels: split the SASS at the role boundaries (first STTM.x16 = producer,
    # UTCBAR = MMA issuer, LDTM = epilogue) and report each role's share
    marks = []
    for k, r in enumerate(data):
        srcl = r[ix["Source"]]
        role = "producer" if "STTM.x16" in srcl else "mma" if ("UTCBAR" in srcl or "UTCHMMA" in srcl or "UTCMMA" in srcl) \
            else "epilogue" if "LDTM" in srcl else None
        if role and ne(r) > 1000:
            marks.append((k, role))
    if marks and len({m[1] for m in marks}) == 3:
        first = {}
        for k, role in marks:
            first.setdefault(role, k)
        order = sorted(first.items(), key=lambda kv: kv[1])
        bounds = []
        for n, (role, k) in enumerate(order):
            lo = 0 if n == 0 else (last_of(marks, order[n - 1][0]) + first[role]) // 2
            hi = len(data) if n == len(order) - 1 else (last_of(marks, role) + order[n + 1][1]) // 2
            bounds.append((role, lo, hi))
        for role, lo, hi in bounds:
            ri, rs = sum(map(ne, data[lo:hi])), sum(map(ns, data[lo:hi]))
            print(f"  role {role:9s} SASS[{lo}:{hi}]  {100 * ri / max(tot, 1):5.1f}% inst  {100 * rs / max(samp, 1):5.1f}% samples")
    hot = sorted(data, key=lambda r: -ns(r))[:25]
    for r in hot:
        print(f"  {ns(r):7d}  {r[ix['Source']].strip()[:90]}")


def last_of(marks, role):
    return max(k for k, r in marks if r == role)


if __name__ == "__main__":
    main()
