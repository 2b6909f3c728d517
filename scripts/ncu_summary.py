#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (first kernel in the report) into text: key raw metrics plus a
warp-stall breakdown of the hottest loop vs the rest (source page).  Usage: ncu_summary.py rep.ncu-rep [loop_min_exec_ratio]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum"]
print("== %s" % rep)
for k in keys:
    if k in m:
        print("%-62s %s %s" % (k, m[k][0], m[k][1]))
def num(k):
    return float(m[k][0].replace(",", "")) if k in m else None
r, w = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
if r is not None:
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tb = r * scale.get(m["dram__bytes_read.sum"][1], 1) + w * scale.get(m["dram__bytes_write.sum"][1], 1)
    print("%-62s %.0f bytes" % ("dram traffic per launch (read + write)", tb))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {n: hdr.index(n) for n in names}
ex = [int(r[iex]) for r in data]
base = sorted(e for e in ex if e > 0)
med = base[len(base) // 2] if base else 1
tot = sum(int(r[isamp]) for r in data)
def agg(sel, label):
    c = collections.Counter(); s = 0; e = 0
    for r in sel:
        for n in names:
            c[n] += int(r[idx[n]])
        s += int(r[isamp]); e += int(r[iex])
    top = ", ".join("%s %.0f%%" % (n.replace("stall_", ""), 100.0 * v / max(1, s)) for n, v in c.most_common(5))
    print("%-28s static SASS %6d  executed %12d  stall samples %7d (%.1f%%)  [%s]" % (label, len(sel), e, s, 100.0 * s / max(1, tot), top))
print("-- source page: instructions executed >= %gx the median count = the hot loop" % ratio)
agg([r for r in data if int(r[iex]) >= ratio * med], "hot loop")
agg([r for r in data if 0 < int(r[iex]) < ratio * med], "once-per-step code")
agg([r for r in data if int(r[iex]) == 0], "not executed")
