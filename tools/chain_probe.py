"""Development probe: the tracking-thread chain on one frame (coeb_extract -> coeb_frame_from_extractor -> coeb_search_local_points),
a few times; with a tools/build_trace.sh library and COEB_KERNEL_TRACE=1 the matcher kernels' timeline is printed per call."""
import sys
sys.path[:0] = ['coeb-slam_b200/python']
import numpy as np, ctypes as C, coeb_b200 as cb
from coeb_b200 import synth
W, H = 640, 480
ex, m = cb.Extractor(), cb.Matcher()
gray = synth.make_frame(1)
kps, desc = ex.extract(gray)
scale = ex.tables()["scale"]
cam = cb.Camera(535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
Tcw, Ow = synth.make_pose(0)
lm, skip, has_obs = synth.make_local_map(kps, desc, scale, Tcw, seed=0)
dev_map = m.local_map(lm)
depth = synth.make_depth(0, W, H)
state = np.full(len(kps), -1, np.int32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    k, d = ex.extract(gray)
    f = m.frame_from_extractor(ex, cam, n=len(k), depth=depth, depth_factor=1.0 / 5000.0)[0]
    out = m.search_local_points(f, dev_map, skip, has_obs, Tcw, Ow, 3.0, 0.8, state, want_proj=False)
    f.close()
print("ok")
