#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle diagnostic for one frame; prints where the first divergence is.
usage: gpu_debug_extract.py [seed] [w h] [nfeatures]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "coeb-slam_b200", "python")]
import numpy as np  # noqa: E402

import coeb_b200 as gpu  # noqa: E402
import orc  # noqa: E402
from coeb_b200 import synth  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
gray = synth.make_frame(seed, w, h)
boxes, tm, blur = synth.make_dynamic(seed, w, h, force_area=(seed % 8 == 3))
g, c = gpu.Extractor(nfeatures=nf), orc.Extractor(nfeatures=nf)
kb, db = c.extract(gray, boxes, tm, blur)
try:
    kg, dg = g.extract(gray, boxes, tm, blur)
except Exception as e:  # keep going: the stage dumps below still say where it broke
    print("gpu extract raised:", e)
    kg, dg = np.empty(0, gpu.KP_DTYPE), np.empty((0, 32), np.uint8)
print("seed", seed, "size", w, h, "counts cpu/gpu", len(kb), len(kg), "dyn", c.dyn_info(), g.dyn_info())
for l in range(8):
    pc, pg = c.level_image(l), g.level_image(l)
    bc, bg = c.level_image(l, blurred=True), g.level_image(l, blurred=True)
    cc, gc = c.level_candidates(l).astype(np.int64), g.level_candidates(l)
    cs, gs = set(map(tuple, cc)), set(map(tuple, gc))
    ck, gk = c.level_keypoints(l), g.level_keys(l)
    same_sel = len(ck) == len(gk) and np.array_equal(ck["x"], gk[:, 0]) and np.array_equal(ck["y"], gk[:, 1])
    same_ang = same_sel and np.array_equal(ck["angle"], gk[:, 3])
    print("L%d pyr %s (maxdiff %d) blur %s cand %d/%d (only cpu %d, only gpu %d) sel %d/%d same=%s angle=%s" % (
        l, np.array_equal(pc, pg), int(np.abs(pc.astype(int) - pg).max()), bc is None or np.array_equal(bc, bg), len(cc),
        len(gc), len(cs - gs), len(gs - cs), len(ck), len(gk), same_sel, same_ang))
    if cs != gs:
        print("   only cpu:", sorted(cs - gs)[:5], " only gpu:", sorted(gs - cs)[:5])
    if not same_sel and len(ck) and len(gk):
        n = min(len(ck), len(gk))
        bad = [i for i in range(n) if ck["x"][i] != gk[i, 0] or ck["y"][i] != gk[i, 1]]
        print("   first selection mismatch at", bad[:3], "cpu", [(ck["x"][i], ck["y"][i]) for i in bad[:3]], "gpu",
              [(gk[i, 0], gk[i, 1]) for i in bad[:3]], "set-equal:",
              set(zip(ck["x"].tolist(), ck["y"].tolist())) == set(map(tuple, gk[:, :2].tolist())))
    if same_sel and not same_ang:
        d = np.abs(ck["angle"] - gk[:, 3])
        print("   angle max diff", d.max(), "at", int(d.argmax()), ck["angle"][d.argmax()], gk[d.argmax(), 3])
if len(kb) == len(kg):
    print("final kp equal:", kb.tobytes() == kg.tobytes(), "desc bit diff:", int(np.unpackbits(db ^ dg).sum()), "of",
          db.size * 8)
