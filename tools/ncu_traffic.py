#!/usr/bin/env python3
"""Summarises an `ncu --set full` report of one extraction step into the JSON bench.py reads for roofline.traffic:
per kernel (summed over its launches in the capture) DRAM bytes, duration, executed warp instructions, and the mean
integer-ALU pipe / issue utilisation. Usage: tools/ncu_traffic.py report.ncu-rep out.json "<how it was captured>"."""
import csv, json, subprocess, sys, collections

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}

def scaled(r, name):
    v = float(r[col[name]].replace(",", "") or 0)
    u = units[col[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "s": 1e3, "ms": 1, "us": 1e-3, "ns": 1e-6}.get(u, 1)

k = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    e = k.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_ms": 0.0, "inst": 0.0, "_alu": 0.0, "_issue": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum")
    e["time_ms"] += scaled(r, "gpu__time_duration.sum")
    e["inst"] += scaled(r, "smsp__inst_executed.sum")
    e["_alu"] += scaled(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
    e["_issue"] += scaled(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
for e in k.values():
    e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
    e["alu_pipe_pct"] = e.pop("_alu") / e["launches"]
    e["issue_active_pct"] = e.pop("_issue") / e["launches"]
json.dump({"source": note, "kernels": k}, open(out, "w"), indent=1)
for n, e in k.items():
    print("%-26s x%-2d %8.3f ms %8.1f MB dram %8.1f Minst  alu %4.1f%% issue %4.1f%%" % (n, e["launches"], e["time_ms"], e["dram_bytes"] / 1e6, e["inst"] / 1e6, e["alu_pipe_pct"], e["issue_active_pct"]))
