"""ctypes binding of the CPU oracle (oracle/libcoeb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU arms.
The product package (coeb-slam_b200/) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcoeb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

MAX_BOXES = 32


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class DynInfo(C.Structure):
    _fields_ = [("area_flag", C.c_int32), ("n_dynamic", C.c_int32), ("rect", (C.c_int32 * 4) * MAX_BOXES),
                ("area", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("b", C.c_float), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float),
                ("max_y", C.c_float)]


def build(force=False):
    """Compile the oracle if the shared library is missing (g++ is in the image)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_cv_round_f.argtypes = [C.c_float]
        _lib.orc_extractor_create.restype = C.c_void_p
        _lib.orc_frame_create.restype = C.c_void_p
        _lib.orc_extract_batch_mt.restype = C.c_double
        _lib.orc_knn2_mt.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def resize_linear(src, dw, dh):
    src = _u8(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def gaussian7(src):
    src = _u8(src)
    dst = np.empty_like(src)
    lib().orc_gaussian7(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def rgb_to_gray(img, bgr=False):
    img = _u8(img)
    h, w, c = img.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_rgb_to_gray(_p(img), w, h, img.strides[0], c, int(bgr), _p(out), w)
    return out


def blur_flags(gray, boxes):
    gray = _u8(gray)
    boxes = _f32(boxes).reshape(-1, 4)
    flags = np.zeros(len(boxes), np.int32)
    means = np.zeros(len(boxes), np.float64)
    lib().orc_blur_flags(_p(gray), gray.shape[1], gray.shape[0], gray.strides[0], _p(boxes), len(boxes), _p(flags), _p(means))
    return flags, means


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def fast_roi(img, threshold, cap=8192):
    """FAST-9/16 + NMS on a (possibly non-contiguous-row) uint8 view. Returns (n,3) int32 x,y,score."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast_roi(C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0],
                           int(threshold), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def octree(cand_xyr, minX, maxX, minY, maxY, N):
    c = _f32(cand_xyr).reshape(-1, 3)
    out = np.empty(max(4 * N + 64, len(c) + 1), np.int32)
    n = lib().orc_octree(_p(c), len(c), minX, maxX, minY, maxY, N, _p(out), len(out))
    assert n <= len(out)
    return out[:n].copy()


def octree_tie_stats(reset=False):
    """(octree calls, careful-phase rounds, rounds with an order tie, rounds with a cut tie) on this thread; see coeb_oracle.hpp."""
    out = (C.c_long * 4)()
    lib().orc_octree_tie_stats(out, int(reset))
    return dict(calls=out[0], careful_rounds=out[1], order_tie_rounds=out[2], cut_tie_rounds=out[3])


def hamming256(a, b):
    return int(lib().orc_hamming256(_p(_u8(a)), _p(_u8(b))))


class Extractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.params = OrbParams(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.h = C.c_void_p(lib().orc_extractor_create(C.byref(self.params)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        per, umax = np.empty(n, np.int32), np.empty(16, np.int32)
        lib().orc_extractor_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(per), _p(umax))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax)

    def extract(self, gray, boxes=None, tm=None, blur_flag=None, cap=20000):
        gray = _u8(gray)
        h, w = gray.shape
        boxes = _f32(boxes if boxes is not None else np.zeros((0, 4))).reshape(-1, 4)
        tm = _f32(tm if tm is not None else np.zeros((0, 2))).reshape(-1, 2)
        blur = _i32(blur_flag if blur_flag is not None else np.zeros(len(boxes)))
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        st = lib().orc_extract(self.h, _p(gray), w, h, gray.strides[0], _p(boxes), len(boxes), _p(tm), len(tm),
                               _p(blur), len(blur), _p(kps), _p(desc), cap, C.byref(n))
        if st != 0:
            raise RuntimeError("oracle extract failed: status %d" % st)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def dyn_info(self):
        d = DynInfo()
        lib().orc_dyn_info(self.h, C.byref(d))
        rects = np.array([[d.rect[i][k] for k in range(4)] for i in range(min(d.n_dynamic, MAX_BOXES))],
                         np.int32).reshape(-1, 4)
        return dict(area_flag=bool(d.area_flag), n_dynamic=d.n_dynamic, rects=rects, area=d.area)

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        lib().orc_level_size(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_image(self, level, blurred=False):
        w, h = self.level_size(level)
        img = np.zeros((h, w), np.uint8)
        n = lib().orc_level_image(self.h, level, 1 if blurred else 0, _p(img))
        return img if n else None

    def level_candidates(self, level, cap=200000):
        out = np.empty((cap, 3), np.float32)
        n = lib().orc_level_candidates(self.h, level, _p(out), cap)
        assert n <= cap
        return out[:n].copy()

    def level_keypoints(self, level, cap=20000):
        out = np.empty(cap, KP_DTYPE)
        n = lib().orc_level_keypoints(self.h, level, _p(out), cap)
        assert n <= cap
        return out[:n].copy()

    def stage_times(self):
        t = (C.c_double * 7)()
        fr = C.c_long()
        lib().orc_stage_times(self.h, t, C.byref(fr))
        names = ["pyramid", "fast", "octree", "angle", "blur", "desc", "total"]
        return {k: t[i] for i, k in enumerate(names)}, fr.value


def extract_batch_mt(params, gray, boxes, nbox, tm, ntm, blur, nthreads, cap=4096, want_outputs=False):
    """gray [B,h,w] u8; boxes [B,max_box,4] f32; nbox [B]; tm [B,max_tm,2]; ntm [B]; blur [B,max_box] i32."""
    gray = _u8(gray)
    B, h, w = gray.shape
    boxes, tm, blur = _f32(boxes), _f32(tm), _i32(blur)
    nbox, ntm = _i32(nbox), _i32(ntm)
    counts = np.zeros(B, np.int32)
    kps = np.empty((B, cap), KP_DTYPE) if want_outputs else None
    desc = np.empty((B, cap, 32), np.uint8) if want_outputs else None
    secs = lib().orc_extract_batch_mt(C.byref(params), B, _p(gray), w, h, _p(boxes), _p(nbox), boxes.shape[1],
                                      _p(tm), _p(ntm), tm.shape[1], _p(blur), int(nthreads), _p(counts),
                                      _p(kps), _p(desc), cap)
    return secs, counts, kps, desc


def hardware_threads():
    return int(lib().orc_hardware_threads())


class Frame:
    """Flattened ORB_SLAM2::Frame view for the matcher oracles (keypoints must be the undistorted ones)."""

    def __init__(self, kps, desc, cam, scale, uright=None):
        self.kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        self.desc = _u8(desc).reshape(-1, 32)
        self.n = len(self.kps)
        self.cam = cam
        self.scale = _f32(scale)
        self.uright = None if uright is None else _f32(uright)
        self.h = C.c_void_p(lib().orc_frame_create(_p(self.kps), _p(self.desc), self.n, _p(self.uright),
                                                   C.byref(cam), _p(self.scale), len(self.scale)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_frame_destroy(self.h)
            self.h = None

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.empty(self.n + 1, np.int32)
        n = lib().orc_features_in_area(self.h, C.c_float(x), C.c_float(y), C.c_float(r), min_level, max_level,
                                       _p(out), len(out))
        return out[:n].copy()

    def grid_cell(self, ix, iy):
        out = np.empty(self.n + 1, np.int32)
        n = lib().orc_grid_cell(self.h, ix, iy, _p(out), len(out))
        return out[:n].copy()


def match_projection(frame, mp, th, nnratio, kp_match):
    """mp: dict of arrays (track_in_view, bad, has_obs, proj_x, proj_y, proj_xr, level, view_cos, desc)."""
    kp_match = _i32(kp_match).copy()
    a = dict(track_in_view=_u8(mp["track_in_view"]), bad=_u8(mp["bad"]), has_obs=_u8(mp["has_obs"]),
             proj_x=_f32(mp["proj_x"]), proj_y=_f32(mp["proj_y"]), proj_xr=_f32(mp["proj_xr"]),
             level=_i32(mp["level"]), view_cos=_f32(mp["view_cos"]), desc=_u8(mp["desc"]))
    n = lib().orc_match_projection(frame.h, len(a["proj_x"]), _p(a["track_in_view"]), _p(a["bad"]),
                                   _p(a["has_obs"]), _p(a["proj_x"]), _p(a["proj_y"]), _p(a["proj_xr"]),
                                   _p(a["level"]), _p(a["view_cos"]), _p(a["desc"]), C.c_float(th),
                                   C.c_float(nnratio), _p(kp_match))
    return n, kp_match


def match_lastframe(cur, last, Tcw_cur, Tcw_last, th, mono, check_ori, kp_match):
    """last: dict of arrays (valid, has_obs, xyz, octave, angle, desc). Tcw_*: 3x4 row-major float32."""
    kp_match = _i32(kp_match).copy()
    a = dict(valid=_u8(last["valid"]), has_obs=_u8(last["has_obs"]), xyz=_f32(last["xyz"]),
             octave=_i32(last["octave"]), angle=_f32(last["angle"]), desc=_u8(last["desc"]))
    tc, tl = _f32(Tcw_cur).reshape(12), _f32(Tcw_last).reshape(12)
    n = lib().orc_match_lastframe(cur.h, len(a["valid"]), _p(a["valid"]), _p(a["has_obs"]), _p(a["xyz"]),
                                  _p(a["octave"]), _p(a["angle"]), _p(a["desc"]), _p(tc), _p(tl), C.c_float(th),
                                  int(mono), int(check_ori), _p(kp_match))
    return n, kp_match


def match_init(f1, f2, prev_matched, window, nnratio, check_ori=True):
    prev = _f32(prev_matched).reshape(-1, 2).copy()
    m12 = np.empty(f1.n, np.int32)
    n = lib().orc_match_init(f1.h, f2.h, _p(prev), _p(m12), int(window), C.c_float(nnratio), int(check_ori))
    return n, m12, prev


def stereo_match(exL, exR, kpsL, descL, kpsR, descR, bf, b):
    kpsL = np.ascontiguousarray(kpsL, dtype=KP_DTYPE)
    kpsR = np.ascontiguousarray(kpsR, dtype=KP_DTYPE)
    descL, descR = _u8(descL), _u8(descR)
    ur = np.empty(len(kpsL), np.float32)
    dp = np.empty(len(kpsL), np.float32)
    n = lib().orc_stereo_match(exL.h, exR.h, len(kpsL), _p(kpsL), _p(descL), len(kpsR), _p(kpsR), _p(descR),
                               C.c_float(bf), C.c_float(b), _p(ur), _p(dp))
    return n, ur, dp


def knn2(q, t, nnratio, nthreads=1):
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    idx = np.empty(len(q), np.int32)
    d1 = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.int32)
    secs = lib().orc_knn2_mt(len(q), _p(q), len(t), _p(t), C.c_float(nnratio), _p(idx), _p(d1), _p(d2),
                             int(nthreads))
    return idx, d1, d2, secs


def undistort_keypoints(kps, cam, dist5):
    """Frame::UndistortKeyPoints: dist5 = (k1, k2, p1, p2, k3) or None."""
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    out = np.empty_like(kps)
    d = None if dist5 is None else _f32(dist5).reshape(5)
    lib().orc_undistort_keypoints(_p(kps), len(kps), C.byref(cam), _p(d), _p(out))
    return out


def stereo_from_rgbd(kps, kps_un, depth, mbf, factor=1.0):
    """Frame::ComputeStereoFromRGBD: depth is a float32 map, or a uint16 map scaled by `factor` (convertTo)."""
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    kps_un = np.ascontiguousarray(kps_un, dtype=KP_DTYPE)
    n = len(kps)
    ur, dp = np.empty(n, np.float32), np.empty(n, np.float32)
    if depth is None:
        kind, dptr, stride = 0, None, 0
    else:
        depth = np.ascontiguousarray(depth)
        assert depth.dtype in (np.float32, np.uint16)
        kind, dptr, stride = (1 if depth.dtype == np.float32 else 2), _p(depth), depth.strides[0]
    lib().orc_stereo_from_rgbd(_p(kps), _p(kps_un), n, dptr, kind, stride, C.c_float(factor), C.c_float(mbf), _p(ur), _p(dp))
    return ur, dp


def search_local_points(frame, lm, skip, has_obs, Tcw, Ow, th, nnratio, kp_match, cos_limit=0.5):
    """Tracking::SearchLocalPoints (isInFrustum + SearchByProjection). lm: dict of arrays (xyz, normal, min_dist,
    max_dist, desc). Returns (nmatches, kp_match, in_view, proj[n,5] = u, v, ur, viewCos, level)."""
    kp_match = _i32(kp_match).copy()
    a = dict(xyz=_f32(lm["xyz"]), normal=_f32(lm["normal"]), min_dist=_f32(lm["min_dist"]), max_dist=_f32(lm["max_dist"]),
             desc=_u8(lm["desc"]))
    n = len(a["min_dist"])
    skip, has_obs = _u8(skip), _u8(has_obs)
    tc, ow = _f32(Tcw).reshape(12), _f32(Ow).reshape(3)
    in_view = np.zeros(n, np.uint8)
    proj = np.zeros((n, 5), np.float32)
    nm = lib().orc_search_local_points(frame.h, n, _p(a["xyz"]), _p(a["normal"]), _p(a["min_dist"]), _p(a["max_dist"]),
                                       _p(a["desc"]), _p(skip), _p(has_obs), _p(tc), _p(ow), C.c_float(cos_limit),
                                       C.c_float(th), C.c_float(nnratio), _p(kp_match), _p(in_view), _p(proj))
    return nm, kp_match, in_view, proj


def match_bow(f1, f2, valid1, valid2, fv1, fv2, nnratio, check_ori=True, strict_low=False):
    """ORBmatcher::SearchByBoW. fv = (node[nn], start[nn+1], items) CSR of the DBoW2 FeatureVector. Returns (n, match12)."""
    valid1 = _u8(valid1)
    valid2 = None if valid2 is None else _u8(valid2)
    n1, s1, i1 = (_i32(a) for a in fv1)
    n2, s2, i2 = (_i32(a) for a in fv2)
    m12 = np.empty(f1.n, np.int32)
    n = lib().orc_match_bow(f1.h, f2.h, _p(valid1), _p(valid2), len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2), _p(i2),
                            C.c_float(nnratio), int(check_ori), int(strict_low), _p(m12))
    return n, m12


def match_triangulation(f1, f2, free1, free2, fv1, fv2, F12, epipole, only_stereo=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation. Returns (n, match12)."""
    free1, free2 = _u8(free1), _u8(free2)
    n1, s1, i1 = (_i32(a) for a in fv1)
    n2, s2, i2 = (_i32(a) for a in fv2)
    F = _f32(F12).reshape(9)
    m12 = np.empty(f1.n, np.int32)
    n = lib().orc_match_triangulation(f1.h, f2.h, _p(free1), _p(free2), len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2), _p(i2),
                                      _p(F), C.c_float(epipole[0]), C.c_float(epipole[1]), int(only_stereo), int(check_ori), _p(m12))
    return n, m12


def match_reloc(cur, kf, Tcw, Ow, th, orb_dist, check_ori, kp_match):
    """Relocalisation SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist). kf: dict of arrays (valid, xyz,
    min_dist, max_dist, angle, desc). Returns (nmatches, kp_match)."""
    kp_match = _i32(kp_match).copy()
    a = dict(valid=_u8(kf["valid"]), xyz=_f32(kf["xyz"]), min_dist=_f32(kf["min_dist"]), max_dist=_f32(kf["max_dist"]),
             angle=_f32(kf["angle"]), desc=_u8(kf["desc"]))
    tc, ow = _f32(Tcw).reshape(12), _f32(Ow).reshape(3)
    n = lib().orc_match_reloc(cur.h, len(a["valid"]), _p(a["valid"]), _p(a["xyz"]), _p(a["min_dist"]), _p(a["max_dist"]),
                              _p(a["angle"]), _p(a["desc"]), _p(tc), _p(ow), C.c_float(th), int(orb_dist), int(check_ori), _p(kp_match))
    return n, kp_match


def fuse_search(frame, lm, valid, Tcw, Ow, th, chi2_tests=True):
    """Search half of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th). Returns (nFused, best_idx[n])."""
    a = dict(xyz=_f32(lm["xyz"]), normal=_f32(lm["normal"]), min_dist=_f32(lm["min_dist"]), max_dist=_f32(lm["max_dist"]), desc=_u8(lm["desc"]))
    n = len(a["min_dist"])
    valid = _u8(valid)
    tc, ow = _f32(Tcw).reshape(12), _f32(Ow).reshape(3)
    best = np.empty(n, np.int32)
    nf = lib().orc_fuse_search(frame.h, n, _p(a["xyz"]), _p(a["normal"]), _p(a["min_dist"]), _p(a["max_dist"]), _p(a["desc"]), _p(valid),
                               _p(tc), _p(ow), C.c_float(th), int(chi2_tests), _p(best))
    return nf, best
