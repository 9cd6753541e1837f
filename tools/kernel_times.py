"""Development helper: per-kernel durations of the last call in an ncu launch list (gpu__time_duration.sum csv).
Usage: tools/kernel_times.py launches.csv calls"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi, bi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
names = [(r[ki].split('(')[0][:40], r[gi], r[bi], float(r[vi].replace(',', '')) / 1000) for r in rows[1:]]
per = len(names) // int(sys.argv[2])
for x in names[-per:]:
    print("%-42s %-14s %-14s %8.2f us" % x)
