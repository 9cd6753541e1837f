#!/usr/bin/env python3
"""Generates tests/golden/motion_seed{0,3}.npz: the intermediates of Frame::ProcessMovingObject (reference src/Frame.cc:311-393)
as the REAL OpenCV 4.13.0 computes them (oracle/pmo.py calls cv2 with the reference's arguments) on two seeded synthetic frame
pairs. The frames are regenerated from the seed at test time (numpy only); a CRC of the pair is stored with the outputs.
Run in the dev container (needs cv2)."""
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "coeb-slam_b200", "python")]
import pmo  # noqa: E402
from coeb_b200 import synth  # noqa: E402

for seed in (0, 3):
    prev, cur, boxes = synth.make_motion_pair(seed)
    r = pmo.process_moving_object(prev, cur)
    corners = pmo.good_features(prev)
    _, dist = pmo.epipolar_outliers(r["F"], r["prepoint"], r["nextpoint"], r["state"])
    np.savez_compressed(os.path.join(HERE, "motion_seed%d.npz" % seed), cv2_version=np.array(cv2.__version__),
                        crc=np.array(zlib.crc32(prev.tobytes() + cur.tobytes())), boxes=boxes, corners=corners, prepoint=r["prepoint"],
                        nextpoint=r["nextpoint"], state=r["state"], F=r["F"], tm=r["tm"], tm_index=r["tm_index"], dist=dist)
    print(seed, len(corners), int((r["state"] != 0).sum()), len(r["tm"]))
