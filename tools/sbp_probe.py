"""Development probe: the flat SearchByProjection call (coeb_match_projection, 5000 map points uploaded per call) a few times; with a
tools/build_trace.sh library and COEB_KERNEL_TRACE=1 the matcher kernels' timeline is printed per call."""
import sys
sys.path[:0] = ['coeb-slam_b200/python']
import numpy as np, coeb_b200 as cb
from coeb_b200 import synth
ex, m = cb.Extractor(), cb.Matcher()
gray = synth.make_frame(1)
kps, desc = ex.extract(gray)
scale = ex.tables()["scale"]
cam = cb.Camera(535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
mp, uright = synth.make_map_points(kps, desc, scale, seed=30)
f = m.frame(kps, desc, cam, scale, uright=uright)
state = np.full(len(kps), -1, np.int32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    n, km = m.match_projection(f, mp, 3.0, 0.8, state)
print("ok", n)
