#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page hot spots) into text: python profiles/ncu_summary.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name[:100])
    for h, u, v in zip(hdr, units, r):
        if h in want or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            print(f"  {h:95s} {v:>16s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
try:
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
except StopIteration:
    sys.exit(0)
hdr = rows[hi]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:                      # first kernel of the report only
    if len(r) != len(hdr) or not r[ia].isdigit():
        break
    data.append(r)
tot = sum(int(r[ia]) for r in data) or 1
tots = sum(int(r[isamp]) for r in data) or 1
print(f"-- source page: {tot} warp instructions, {tots} samples; top {top} by samples")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
    print(f"  {r[isrc].strip()[:72]:72s} {100 * int(r[ia]) / tot:5.2f}% inst {100 * int(r[isamp]) / tots:5.2f}% samples")
ops = {}
for r in data:
    op = r[isrc].strip().split()
    op = [t for t in op if not t.startswith("@")]
    if not op: continue
    o = op[0].split(".")[0]
    ops[o] = ops.get(o, 0) + int(r[ia])
print("-- executed instruction mix:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
