"""Development probe: a few single-frame calls (the third and later ones replay the CUDA graph); with COEB_PIPE_TRACE=1 the library
prints the call's timeline, under ncu the launch list shows the single-frame kernels."""
import sys
sys.path[:0] = ['coeb-slam_b200/python']
import numpy as np, coeb_b200 as cb
from coeb_b200 import synth
ex = cb.Extractor()
gray = synth.make_frame(0)
boxes, tm, blur = synth.make_dynamic(0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    k, d = ex.extract(gray, boxes, tm, blur)
print(len(k))
