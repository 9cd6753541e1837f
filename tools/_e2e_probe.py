import sys, time, numpy as np
sys.path.insert(0, 'coeb-slam_b200/python')
import coeb_b200 as cb
from coeb_b200 import synth
import torch
cap = 1100
for B in (64, 128, 256, 512):
    batch = synth.make_batch(B, base_seed=1000, unique=16)
    pin = {}; own = []
    for k in ("gray", "boxes", "nbox", "tm", "ntm", "blur"):
        a, o = cb.pinned_array(batch[k].shape, batch[k].dtype); a[...] = batch[k]; pin[k] = a; own.append(o)
    o_kps, p1 = cb.pinned_array((B, cap), cb.KP_DTYPE); o_desc, p2 = cb.pinned_array((B, cap, 32), np.uint8)
    o_cnt, p3 = cb.pinned_array((B,), np.int32); o_st, p4 = cb.pinned_array((B,), np.int32)
    ex = cb.Extractor(1000, 1.2, 8, 20, 7)
    f = lambda: ex.extract_batch_host(pin["gray"], pin["boxes"], pin["nbox"], pin["tm"], pin["ntm"], pin["blur"], cap=cap, out=(o_kps, o_desc, o_cnt, o_st))
    for _ in range(4): f()
    ts = []
    for _ in range(10):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    print("B=%d  median %.3f ms  -> %.0f frames/s" % (B, 1e3 * np.median(ts), B / np.median(ts)))
