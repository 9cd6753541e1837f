#!/usr/bin/env python3
"""Summarises an `ncu --set full` report of the matcher kernels into JSON: per kernel (first launch in the capture) duration,
executed warp instructions, grid size and the utilisation of the execution pipes that matter for Hamming matching
(XU = POPC, ALU = LOP3 / compares, LSU = shared / global accesses, FMA = IMAD) plus the issue-slot utilisation.
Usage: tools/ncu_match_pipes.py report.ncu-rep [more.ncu-rep ...] out.json "<how it was captured>"."""
import collections
import csv
import json
import subprocess
import sys

reps, out, note = sys.argv[1:-2], sys.argv[-2], sys.argv[-1]
PIPES = {"xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active"}
k = collections.OrderedDict()
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, name):
        try:
            return float(r[col[name]].replace(",", ""))
        except Exception:
            return None
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        t = num(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1e-3)
        k[name] = {"time_us": t, "warp_inst": num(r, "smsp__inst_executed.sum"), "grid": num(r, "launch__grid_size"), "block": num(r, "launch__block_size"),
                   "pipe_pct": {p: num(r, m) for p, m in PIPES.items() if m in col}, "report": rep.split("/")[-1]}
json.dump({"source": note, "kernels": k}, open(out, "w"), indent=1)
for n, e in k.items():
    print("%-26s %8.1f us %10.0f inst  %s" % (n, e["time_us"], e["warp_inst"], {p: round(v, 1) for p, v in e["pipe_pct"].items() if v is not None}))
