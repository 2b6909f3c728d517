#!/usr/bin/env python
"""Condense an `ncu --set full` capture of ONE launch of a bench workload's dominant kernel into the small JSON that bench.py's roofline
block reads (profiles/r02_<workload>_ncu.json), keyed by the SASS hash of the kernel in the library the capture ran on.
Usage (on the GPU box, right after the capture):  ncu_to_json.py <kuka|mobile> <rep.ncu-rep> <out.json> [capture command ...]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import bench
from srl_sim._abi import CUDA_LIBRARY_PATH

workload, rep, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def num(k, unit_scale=False):
    if k not in m:
        return None
    v = float(m[k][0].replace(",", ""))
    return v * scale.get(m[k][1], 1.0) if unit_scale else v


doc = {
    "workload": workload,
    "kernel": m["Kernel Name"][0],
    "sass_sha16": bench.kernel_sass_sha16(CUDA_LIBRARY_PATH, workload),
    "library": os.path.relpath(CUDA_LIBRARY_PATH, ROOT),
    "grid": m.get("launch__grid_size", ("", ""))[0], "block": m.get("launch__block_size", ("", ""))[0],
    "registers_per_thread": num("launch__registers_per_thread"),
    "duration_ms": num("gpu__time_duration.sum", True),
    "sm_clock_ghz": num("sm__cycles_elapsed.avg.per_second"),
    "warp_inst_per_launch": num("smsp__inst_executed.sum"),
    "threads_per_warp_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "dram_bytes_per_launch": (num("dram__bytes_read.sum", True) or 0) + (num("dram__bytes_write.sum", True) or 0),
    "dram_read_bytes": num("dram__bytes_read.sum", True), "dram_write_bytes": num("dram__bytes_write.sum", True),
    "local_ld_inst": num("smsp__sass_inst_executed_op_local_ld.sum"), "local_st_inst": num("smsp__sass_inst_executed_op_local_st.sum"),
    "capture": " ".join(sys.argv[4:]) or None,
    "note": "one launch, ncu --set full --clock-control none; per-launch times under ncu are cold-cache and serialised -- bench.py uses only the "
            "instruction / byte COUNTS of this capture, with its own live launch time",
}
# share of the warp instructions executed in the hottest loop (the PGS sweep for Kuka): instructions executed >= 100x the median count
try:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    shdr = srows[1]; iex = shdr.index("Instructions Executed")
    ex = [int(r[iex]) for r in srows[2:] if len(r) > iex and r[iex].isdigit()]
    pos = sorted(e for e in ex if e > 0)
    med = pos[len(pos) // 2]
    doc["hot_loop_inst_share"] = sum(e for e in ex if e >= 100 * med) / float(sum(ex))
except Exception as exn:
    doc["hot_loop_inst_share"] = None
with open(out, "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps(doc))
