"""Development probe: wall time of the blocking single-frame C call (median of 300), ctypes arguments marshalled up front."""
import sys, time
sys.path[:0] = ['coeb-slam_b200/python']
import numpy as np, ctypes as C, coeb_b200 as cb
from coeb_b200 import synth
ex = cb.Extractor()
gray = synth.make_frame(0)
boxes, tm, blur = synth.make_dynamic(0)
boxes, tm, blur = np.ascontiguousarray(boxes, np.float32), np.ascontiguousarray(tm, np.float32), np.ascontiguousarray(blur, np.int32)
for _ in range(8):
    ex.extract(gray, boxes, tm, blur)
cap = ex.default_cap(640, 480)
kps, desc, n = np.empty(cap, cb.KP_DTYPE), np.empty((cap, 32), np.uint8), C.c_int()
P = lambda a: a.ctypes.data_as(C.c_void_p)
args = (ex.h, P(gray), 640, 480, 640, P(boxes), len(boxes), P(tm), len(tm), P(blur), len(blur), P(kps), P(desc), cap, C.byref(n))
fn = cb.lib().coeb_extract
ts = []
for _ in range(300):
    t = time.perf_counter(); fn(*args); ts.append(time.perf_counter() - t)
print("median %.1f us  p10 %.1f  p90 %.1f  (%d keypoints)" % (1e6 * np.median(ts), 1e6 * np.percentile(ts, 10), 1e6 * np.percentile(ts, 90), n.value))
