"""Development probe: wall time of the blocking single-frame C call (median of 200)."""
import sys, time
sys.path[:0] = ['coeb-slam_b200/python']
import numpy as np, ctypes as C, coeb_b200 as cb
from coeb_b200 import synth
ex = cb.Extractor()
gray = synth.make_frame(0)
boxes, tm, blur = synth.make_dynamic(0)
for _ in range(8):
    ex.extract(gray, boxes, tm, blur)
ts = []
for _ in range(200):
    t = time.perf_counter(); ex.extract(gray, boxes, tm, blur); ts.append(time.perf_counter() - t)
print("median %.1f us  p10 %.1f  p90 %.1f" % (1e6 * np.median(ts), 1e6 * np.percentile(ts, 10), 1e6 * np.percentile(ts, 90)))
