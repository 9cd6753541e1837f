"""CPU: the oracle's restatement of the Frame constructor tail and of the visibility test in front of SearchByProjection
(oracle/coeb_oracle_frame.hpp), pinned against OpenCV 4.13 where the arithmetic is OpenCV's and against the C library
where it is libm's; plus a numpy restatement of Frame::isInFrustum as an independent check of the control flow."""
import ctypes as C
import os

import numpy as np
import pytest

import orc
from coeb_b200 import synth

TUM1 = dict(cam=(517.306408, 516.469215, 318.643040, 255.313989, 40.0, 40.0 / 517.306408, 0.0, 640.0, 0.0, 480.0),
            dist=(0.262383, -0.953104, -0.005358, 0.002628, 1.163314))
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _random_keys(n, seed, w=640, h=480):
    rng = np.random.default_rng(seed)
    kps = np.zeros(n, orc.KP_DTYPE)
    kps["x"] = rng.uniform(0, w - 1, n)
    kps["y"] = rng.uniform(0, h - 1, n)
    kps["octave"] = rng.integers(0, 8, n)
    kps["angle"] = rng.uniform(0, 360, n)
    return kps


def test_logf_restatement_equals_libm():
    """MapPoint::PredictScale calls the float log of the C library: the restatement (glibc 2.39 algorithm) must agree
    with the libm this test runs on for every sampled normal float, and densely over the ratios that occur (0.01 .. 100)."""
    fn = orc.lib().orc_logf_mismatches
    fn.restype = C.c_long
    assert fn(C.c_uint32(0x00800000), C.c_uint32(0x7F800000), C.c_uint32(1009)) == 0
    lo, hi = np.float32(0.01).view(np.uint32), np.float32(100.0).view(np.uint32)
    assert fn(C.c_uint32(int(lo)), C.c_uint32(int(hi)), C.c_uint32(13)) == 0


def test_undistort_matches_cv2_bit_exact():
    cv2 = pytest.importorskip("cv2")
    cam = orc.Camera(*TUM1["cam"])
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], np.float32)
    for dist in (TUM1["dist"], (-0.2, 0.05, 0.001, -0.0005, 0.0), (0.0, 0.5, 0.0, 0.0, 0.0)):
        kps = _random_keys(4000, 5)
        un = orc.undistort_keypoints(kps, cam, dist)
        if dist[0] == 0.0:   # `if (mDistCoef.at<float>(0) == 0.0) mvKeysUn = mvKeys` (src/Frame.cc:581-585)
            assert un.tobytes() == kps.tobytes()
            continue
        ref = cv2.undistortPoints(np.stack([kps["x"], kps["y"]], 1).reshape(-1, 1, 2), K, np.array(dist, np.float32), None, K).reshape(-1, 2)
        assert np.array_equal(un["x"], ref[:, 0]) and np.array_equal(un["y"], ref[:, 1])
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(un[f], kps[f])


def test_undistort_golden_without_cv2():
    g = np.load(os.path.join(G, "frame_tail.npz"))
    cam = orc.Camera(*TUM1["cam"])
    kps = _random_keys(2000, 9)
    assert np.array_equal(np.stack([kps["x"], kps["y"]], 1), g["pts"])
    un = orc.undistort_keypoints(kps, cam, TUM1["dist"])
    assert np.array_equal(np.stack([un["x"], un["y"]], 1), g["undist"])
    # gemm / norm pins: the fp32 left-to-right projection and the double-accumulated norm
    R, P, t = g["gemm_R"], g["gemm_P"], g["gemm_t"]
    for i in range(len(R)):
        a = np.empty(3, np.float32)
        for r in range(3):
            s = np.float32(R[i, r, 0] * P[i, 0])
            s = np.float32(s + np.float32(R[i, r, 1] * P[i, 1]))
            s = np.float32(s + np.float32(R[i, r, 2] * P[i, 2]))
            a[r] = np.float32(s + t[i, r])
        assert np.array_equal(a, g["gemm_out"][i])
    nrm = np.sqrt((g["norm_v"].astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    assert np.array_equal(nrm, g["norm_out"])


def test_depth_gather_and_convert_to():
    cv2 = pytest.importorskip("cv2")
    cam = orc.Camera(*TUM1["cam"])
    kps = _random_keys(3000, 6)
    un = orc.undistort_keypoints(kps, cam, TUM1["dist"])
    d16 = synth.make_depth(3)
    f = np.float32(1.0) / np.float32(5000.0)   # mDepthMapFactor = 1.0f / DepthMapFactor (src/Tracking.cc:113-116)
    ur, dp = orc.stereo_from_rgbd(kps, un, d16, cam.bf, f)
    dm = cv2.multiply(d16, float(f), dtype=cv2.CV_32F)   # one fp32 rounding per pixel, as convertTo(CV_32F, f)
    ur2, dp2 = orc.stereo_from_rgbd(kps, un, dm, cam.bf)
    assert np.array_equal(ur, ur2) and np.array_equal(dp, dp2)
    rows, cols = kps["y"].astype(np.int32), kps["x"].astype(np.int32)
    d = dm[rows, cols]
    assert np.array_equal(dp, np.where(d > 0, d, np.float32(-1)))
    with np.errstate(divide="ignore"):
        assert np.array_equal(ur, np.where(d > 0, un["x"] - np.float32(cam.bf) / d, np.float32(-1)).astype(np.float32))
    assert 0.5 < (dp > 0).mean() < 0.95
    ur0, dp0 = orc.stereo_from_rgbd(kps, un, None, cam.bf)
    assert (ur0 == -1).all() and (dp0 == -1).all()


def _frustum_numpy(lm, skip, Tcw, Ow, cam, cos_limit, scale):
    """Frame::isInFrustum (src/Frame.cc:445-501) written independently with numpy float32/float64 scalars."""
    f32 = np.float32
    n = len(lm["min_dist"])
    out = np.zeros((n, 5), np.float32)
    inv = np.zeros(n, np.uint8)
    logs = f32(np.log(np.float64(scale[1])))   # only used through the integer level below
    T = Tcw.reshape(3, 4).astype(np.float32)
    for i in range(n):
        if skip[i]:
            continue
        P = lm["xyz"][i]
        pc = [f32(f32(f32(f32(T[r, 0] * P[0]) + f32(T[r, 1] * P[1])) + f32(T[r, 2] * P[2])) + T[r, 3]) for r in range(3)]
        if pc[2] < 0:
            continue
        invz = f32(1.0) / pc[2]
        u = f32(f32(f32(f32(cam.fx) * pc[0]) * invz) + f32(cam.cx))
        v = f32(f32(f32(f32(cam.fy) * pc[1]) * invz) + f32(cam.cy))
        if u < cam.min_x or u > cam.max_x or v < cam.min_y or v > cam.max_y:
            continue
        po = (P - Ow).astype(np.float32)
        dist = f32(np.sqrt((po.astype(np.float64) ** 2).sum()))
        if dist < f32(0.8) * lm["min_dist"][i] or dist > f32(1.2) * lm["max_dist"][i]:
            continue
        vc = f32((po.astype(np.float64) * lm["normal"][i].astype(np.float64)).sum() / np.float64(dist))
        if vc < cos_limit:
            continue
        ratio = lm["max_dist"][i] / dist
        lvl = int(np.clip(np.ceil(np.log(np.float64(ratio)) / np.float64(logs)), 0, len(scale) - 1))
        inv[i] = 1
        out[i] = (u, v, f32(u - f32(f32(cam.bf) * invz)), vc, lvl)
    return inv, out


def test_search_local_points_oracle_against_numpy_and_flat_matcher():
    ex = orc.Extractor()
    kps, desc = ex.extract(synth.make_frame(100))
    scale = ex.tables()["scale"]
    cam_args = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
    cam = orc.Camera(*cam_args)
    Tcw, Ow = synth.make_pose(2)
    lm, skip, has_obs = synth.make_local_map(kps, desc, scale, Tcw, seed=2, n_map=3000, n_true=600)
    rng = np.random.default_rng(8)
    uright = np.where(rng.random(len(kps)) < 0.3, kps["x"] - np.float32(40.0) / rng.uniform(0.5, 5.0, len(kps)).astype(np.float32),
                      np.float32(-1)).astype(np.float32)
    F = orc.Frame(kps, desc, cam, scale, uright)
    state = np.full(len(kps), -1, np.int32)
    state[rng.random(len(kps)) < 0.1] = -2
    nm, kpm, in_view, proj = orc.search_local_points(F, lm, skip, has_obs, Tcw, Ow, 3.0, 0.8, state)
    inv2, proj2 = _frustum_numpy(lm, skip, Tcw, Ow, cam, np.float32(0.5), scale)
    assert np.array_equal(in_view, inv2)
    # every rejection branch is exercised and a healthy share survives
    assert 0.2 < in_view.mean() < 0.8 and nm > 200
    v = in_view.astype(bool)
    assert np.array_equal(proj[v][:, :4], proj2[v][:, :4])
    # the numpy level uses a double log: it may differ from the float one only when log(ratio)/log(s) sits on an integer
    assert (proj[v][:, 4] != proj2[v][:, 4]).mean() < 1e-3
    # the fused call == the visibility fields handed to the flat SearchByProjection oracle
    mp = dict(track_in_view=in_view, bad=np.zeros_like(in_view), has_obs=has_obs, proj_x=proj[:, 0], proj_y=proj[:, 1], proj_xr=proj[:, 2],
              view_cos=proj[:, 3], level=proj[:, 4].astype(np.int32), desc=lm["desc"])
    nm2, kpm2 = orc.match_projection(F, mp, 3.0, 0.8, state)
    assert nm2 == nm and np.array_equal(kpm, kpm2)


def test_relocalisation_oracle_against_python_transcription():
    """src/ORBmatcher.cc:1473-1600 transcribed with numpy scalars; windows through the (separately pinned) grid query."""
    f32 = np.float32
    ex = orc.Extractor()
    kps, desc = ex.extract(synth.make_frame(210))
    scale = ex.tables()["scale"]
    cam_args = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
    cam = orc.Camera(*cam_args)
    F = orc.Frame(kps, desc, cam, scale, None)
    Tcw, Ow = synth.make_pose(5)
    lm, skip, _ = synth.make_local_map(kps, desc, scale, Tcw, seed=5, n_map=1200, n_true=600)
    rng = np.random.default_rng(5)
    valid = (1 - skip).astype(np.uint8)
    angle = rng.uniform(0, 360, len(skip)).astype(np.float32)
    state = rng.choice([-1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
    kf = dict(valid=valid, xyz=lm["xyz"], min_dist=lm["min_dist"], max_dist=lm["max_dist"], angle=angle, desc=lm["desc"])
    th, orb_dist = f32(10.0), 100
    n, km = orc.match_reloc(F, kf, Tcw, Ow, th, orb_dist, True, state)

    T = Tcw.reshape(3, 4).astype(np.float32)
    logs = np.log(np.float64(scale[1]))
    km2 = state.copy()
    hist = [[] for _ in range(30)]
    n2 = 0
    for i in range(len(valid)):
        if not valid[i]:
            continue
        P = lm["xyz"][i]
        pc = [f32(f32(f32(f32(T[r, 0] * P[0]) + f32(T[r, 1] * P[1])) + f32(T[r, 2] * P[2])) + T[r, 3]) for r in range(3)]
        with np.errstate(divide="ignore"):
            invz = f32(np.float64(1.0) / np.float64(pc[2]))
        u = f32(f32(f32(f32(cam.fx) * pc[0]) * invz) + f32(cam.cx))
        v = f32(f32(f32(f32(cam.fy) * pc[1]) * invz) + f32(cam.cy))
        if u < cam.min_x or u > cam.max_x or v < cam.min_y or v > cam.max_y:   # no test on the sign of z in this overload
            continue
        po = (P - Ow).astype(np.float32)
        d3 = f32(np.sqrt((po.astype(np.float64) ** 2).sum()))
        if d3 < f32(0.8) * lm["min_dist"][i] or d3 > f32(1.2) * lm["max_dist"][i]:
            continue
        lvl = int(np.clip(np.ceil(np.log(np.float64(lm["max_dist"][i] / d3)) / logs), 0, len(scale) - 1))
        win = F.features_in_area(u, v, f32(th * scale[lvl]), lvl - 1, lvl + 1)
        best, bi = 256, -1
        for i2 in win:
            if km2[i2] != -1:
                continue
            dist = int(np.unpackbits(lm["desc"][i] ^ desc[i2]).sum())
            if dist < best:
                best, bi = dist, int(i2)
        if best <= orb_dist:
            km2[bi] = i
            n2 += 1
            rot = f32(angle[i]) - f32(kps["angle"][bi])
            if rot < 0:
                rot = f32(rot + f32(360.0))
            b = int(np.floor(float(f32(rot * f32(1.0 / 30))) + 0.5))
            hist[0 if b == 30 else b].append(bi)
    sizes = [len(h) for h in hist]
    order = sorted(range(30), key=lambda k: (-sizes[k], k))
    keep = [order[0]] if sizes[order[0]] > 0 else []
    if keep and sizes[order[1]] > 0 and not f32(sizes[order[1]]) < f32(0.1) * f32(sizes[order[0]]):
        keep.append(order[1])
        if sizes[order[2]] > 0 and not f32(sizes[order[2]]) < f32(0.1) * f32(sizes[order[0]]):
            keep.append(order[2])
    for b in range(30):
        if b not in keep:
            for k in hist[b]:
                km2[k] = -1
                n2 -= 1
    assert n == n2 and np.array_equal(km, km2) and n > 20
