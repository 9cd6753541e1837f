#!/usr/bin/env python3
"""Generates the committed golden vectors under tests/golden/ (run in the dev container: needs cv2 4.13.0).

Two kinds of goldens:
  * cv2_primitives.npz  -- outputs of the REAL OpenCV primitives the reference calls (cv2.resize INTER_LINEAR,
    cv2.GaussianBlur 7x7 sigma 2, cv2.FastFeatureDetector on ROIs, cv2.fastAtan2) on small seeded inputs;
    they pin the oracle's integer models without needing cv2 at test time.
  * extract_seedN.npz   -- full extractor outputs (keypoints + descriptors + per-level candidate counts) produced by
    oracle A (tests/oracle_cv2.py: the reference's control flow driving cv2), on seeded synthetic 640x480 frames
    with injected person boxes. Inputs are regenerated from the seed (numpy only); a CRC of the input is stored.
  * match_seed0.npz     -- matcher outputs of oracle B on seeded inputs (no third-party arithmetic involved).
"""
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "coeb-slam_b200", "python"), os.path.join(ROOT, "tests")]
import orc  # noqa: E402
from coeb_b200 import synth  # noqa: E402
from conftest import load_pattern  # noqa: E402
from oracle_cv2 import ExtractorA  # noqa: E402


def primitives():
    rng = np.random.default_rng(1234)
    out = {"cv2_version": np.array(cv2.__version__)}
    img = rng.integers(0, 256, size=(96, 128), dtype=np.uint8)
    tex = synth.make_frame(77)[100:196, 200:328].copy()
    out["img_noise"], out["img_tex"] = img, tex
    for name, src in (("noise", img), ("tex", tex)):
        for (dw, dh) in ((107, 80), (89, 67), (150, 100)):
            out["resize_%s_%dx%d" % (name, dw, dh)] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        out["blur_%s" % name] = cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        for th in (20, 7, 30, 10):
            kps = cv2.FastFeatureDetector_create(th, True).detect(src[10:50, 20:58])
            out["fast_%s_th%d" % (name, th)] = np.array([(k.pt[0], k.pt[1], k.response) for k in kps], np.int32).reshape(-1, 3)
    yx = rng.integers(-30000, 30000, size=(512, 2)).astype(np.float32)
    yx[:6] = [(0, 0), (0, -5), (-3, 0), (5, 0), (1, 1), (-7, 7)]
    out["atan2_in"] = yx
    out["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), **out)


def extraction():
    pat = load_pattern()
    for seed in (0, 3):
        gray = synth.make_frame(seed)
        boxes, tm, blur = synth.make_dynamic(seed, force_area=(seed % 8 == 3))
        ka, da, st, dyn = ExtractorA().extract(gray, boxes, tm, blur, pattern=pat, want_stages=True)
        kps = np.array(ka, dtype=orc.KP_DTYPE)
        np.savez_compressed(os.path.join(HERE, "extract_seed%d.npz" % seed), input_crc=np.uint32(zlib.crc32(gray.tobytes())),
                            boxes=boxes, tm=tm, blur=blur, kps=kps, desc=da, area_flag=np.bool_(dyn["area_flag"]),
                            rects=np.array(dyn["rects"], np.int32).reshape(-1, 4),
                            n_candidates=np.array([len(c) for c in st["candidates"]], np.int32),
                            level_crc=np.array([zlib.crc32(p.tobytes()) for p in st["pyramid"]], np.uint32))


def matching():
    ex = orc.Extractor()
    kps, desc = ex.extract(synth.make_frame(100))
    scale = ex.tables()["scale"]
    cam = orc.Camera(535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
    mp, uright = synth.make_map_points(kps, desc, scale, seed=30)
    f = orc.Frame(kps, desc, cam, scale, uright)
    state = np.random.default_rng(5).choice([-1, -1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
    n2, km2 = orc.match_projection(f, mp, 3.0, 0.8, state)
    last, Tc, Tl = synth.make_last_frame(kps, desc, seed=15)
    f0 = orc.Frame(kps, desc, cam, scale, None)
    n3, km3 = orc.match_lastframe(f0, last, Tc, Tl, 15.0, False, True, np.full(len(kps), -1, np.int32))
    q, t = synth.make_knn_sets(300, 4000, seed=1)
    idx, d1, d2, _ = orc.knn2(q, t, 0.7)
    np.savez_compressed(os.path.join(HERE, "match_seed0.npz"), kps_crc=np.uint32(zlib.crc32(kps.tobytes())), n_proj=np.int32(n2), kp_match_proj=km2,
                        n_last=np.int32(n3), kp_match_last=km3, knn_idx=idx, knn_d1=d1, knn_d2=d2)


def frame_tail():
    """Real-OpenCV outputs for the arithmetic of the Frame constructor tail / isInFrustum: undistortPoints with the TUM1
    calibration, cv::gemm on 3x3 . 3x1 + 3x1 (mRcw * P + mtcw) and cv::norm of a 3-vector."""
    import cv2
    rng = np.random.default_rng(9)
    n = 2000
    pts = np.zeros((n, 2), np.float32)   # same draw order as tests/test_frame_tail_cpu.py::_random_keys(2000, 9)
    pts[:, 0] = rng.uniform(0, 639, n)
    pts[:, 1] = rng.uniform(0, 479, n)
    K = np.array([[517.306408, 0, 318.643040], [0, 516.469215, 255.313989], [0, 0, 1]], np.float32)
    D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
    und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    rng = np.random.default_rng(10)
    m = 500
    R = rng.standard_normal((m, 3, 3)).astype(np.float32)
    P = (rng.standard_normal((m, 3)) * 5).astype(np.float32)
    t = rng.standard_normal((m, 3)).astype(np.float32)
    out = np.stack([cv2.gemm(R[i], P[i].reshape(3, 1), 1.0, t[i].reshape(3, 1), 1.0).reshape(3) for i in range(m)])
    v = (rng.standard_normal((m, 3)) * 3).astype(np.float32)
    nrm = np.array([np.float32(cv2.norm(v[i].reshape(3, 1))) for i in range(m)], np.float32)
    np.savez_compressed(os.path.join(HERE, "frame_tail.npz"), pts=pts, undist=und, gemm_R=R, gemm_P=P, gemm_t=t, gemm_out=out,
                        norm_v=v, norm_out=nrm)


if __name__ == "__main__":
    primitives()
    extraction()
    matching()
    frame_tail()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
