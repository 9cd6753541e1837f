"""ctypes binding of oracle/_ref: the reference's OWN sources compiled unchanged against the test-only OpenCV shim
(oracle/ref_shim/, recipe: `make -C oracle _ref`).

TEST INFRASTRUCTURE ONLY: imported by tests/ (oracle B == _ref) and by bench.py's reference arm / cpu_baseline leg.
The product package (coeb-slam_b200/) must never import this module.

Variants (see oracle/Makefile): "ref" = reference flags, glibc heap; "nofma" = the same with -ffp-contract=off;
"mono" = reference flags + monotonic heap (isolates the heap-address tie-break of src/ORBextractor.cc:691).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from orc import KP_DTYPE, OrbParams, _f32, _i32, _p, _u8

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_NAMES = {"ref": "libcoeb_ref.so", "nofma": "libcoeb_ref_nofma.so", "mono": "libcoeb_ref_mono.so"}
REFERENCE_TREE = "/root/reference"

_libs = {}


def available(variant="ref"):
    """True if the prebuilt library exists, or the reference tree is here to build it from."""
    return os.path.exists(os.path.join(_DIR, _NAMES[variant])) or os.path.isdir(os.path.join(REFERENCE_TREE, "src"))


def build(force=False):
    """Compile oracle/_ref from the sources under /root/reference (dev container only; the GPU box uses the prebuilt files)."""
    if not os.path.isdir(os.path.join(REFERENCE_TREE, "src")):
        return False
    subprocess.check_call(["make", "-C", _HERE, "_ref"] + (["-B"] if force else []))
    return True


def lib(variant="ref"):
    if variant not in _libs:
        path = os.path.join(_DIR, _NAMES[variant])
        if not os.path.exists(path):
            if not build():
                raise RuntimeError("oracle/_ref/%s is missing and /root/reference is not here to build it" % _NAMES[variant])
        L = C.CDLL(path)
        L.ref_extractor_create.restype = C.c_void_p
        L.ref_build_info.restype = C.c_char_p
        L.ref_extract_batch_mt.restype = C.c_double
        L.ref_heap_fallbacks.restype = C.c_long
        _libs[variant] = L
    return _libs[variant]


class Extractor:
    """ORB_SLAM2::ORBextractor of the reference, same call shape as orc.Extractor / coeb_b200.Extractor."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, variant="ref"):
        self.L = lib(variant)
        self.params = OrbParams(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.h = C.c_void_p(self.L.ref_extractor_create(C.byref(self.params)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        self.L.ref_extractor_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2)

    def extract(self, gray, boxes=None, tm=None, blur_flag=None, cap=20000):
        gray = _u8(gray)
        h, w = gray.shape
        boxes = _f32(boxes if boxes is not None else np.zeros((0, 4))).reshape(-1, 4)
        tm = _f32(tm if tm is not None else np.zeros((0, 2))).reshape(-1, 2)
        blur = _i32(blur_flag if blur_flag is not None else np.zeros(len(boxes)))
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        st = self.L.ref_extract(self.h, _p(gray), w, h, gray.strides[0], _p(boxes), len(boxes), _p(tm), len(tm),
                                _p(blur), len(blur), _p(kps), _p(desc), cap, C.byref(n))
        if st != 0:
            raise RuntimeError("reference extract refused: status %d" % st)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        if self.L.ref_level_size(self.h, level, C.byref(w), C.byref(h)) != 0:
            return None
        return w.value, h.value

    def level_image(self, level, border=0):
        w, h = self.level_size(level)
        img = np.zeros((h + 2 * border, w + 2 * border), np.uint8)
        n = self.L.ref_level_image(self.h, level, border, _p(img))
        return img if n else None

    def heap_fallbacks(self):
        return int(self.L.ref_heap_fallbacks())


def extract_batch_mt(params, gray, boxes, nbox, tm, ntm, blur, nthreads, cap=4096, want_outputs=False, variant="ref"):
    """Same packing and return value as orc.extract_batch_mt, run by the reference's own extractor (one instance per thread)."""
    gray = _u8(gray)
    B, h, w = gray.shape
    boxes, tm, blur = _f32(boxes), _f32(tm), _i32(blur)
    nbox, ntm = _i32(nbox), _i32(ntm)
    counts = np.zeros(B, np.int32)
    kps = np.empty((B, cap), KP_DTYPE) if want_outputs else None
    desc = np.empty((B, cap, 32), np.uint8) if want_outputs else None
    secs = lib(variant).ref_extract_batch_mt(C.byref(params), B, _p(gray), w, h, _p(boxes), _p(nbox), boxes.shape[1], _p(tm), _p(ntm),
                                             tm.shape[1], _p(blur), int(nthreads), cap, _p(counts), _p(kps), _p(desc))
    return secs, counts, kps, desc


# ---- the reference's Frame / MapPoint / ORBmatcher (oracle/ref_shim/ref_match_c.cpp); same call shapes as orc.py ----
class Frame:
    """A real ORB_SLAM2::Frame filled with undistorted keypoints, descriptors, uRight, camera statics and its grid."""

    def __init__(self, kps, desc, cam, scale, uright=None, variant="ref"):
        self.L = lib(variant)
        self.L.ref_frame_create.restype = C.c_void_p
        self.kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        self.desc = _u8(desc).reshape(-1, 32)
        self.n = len(self.kps)
        self.scale = _f32(scale)
        self.uright = None if uright is None else _f32(uright)
        self.h = C.c_void_p(self.L.ref_frame_create(_p(self.kps), _p(self.desc), self.n, _p(self.uright), C.byref(cam), _p(self.scale),
                                                    len(self.scale)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_frame_destroy(self.h)
            self.h = None

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.empty(self.n + 1, np.int32)
        n = self.L.ref_features_in_area(self.h, C.c_float(x), C.c_float(y), C.c_float(r), min_level, max_level, _p(out), len(out))
        return out[:n].copy()

    def grid_cell(self, ix, iy):
        out = np.empty(self.n + 1, np.int32)
        n = self.L.ref_grid_cell(self.h, ix, iy, _p(out), len(out))
        return out[:n].copy()


def hamming256(a, b, variant="ref"):
    return int(lib(variant).ref_hamming256(_p(_u8(a)), _p(_u8(b))))


def match_projection(frame, mp, th, nnratio, kp_match):
    kp_match = _i32(kp_match).copy()
    a = dict(track_in_view=_u8(mp["track_in_view"]), bad=_u8(mp["bad"]), has_obs=_u8(mp["has_obs"]), proj_x=_f32(mp["proj_x"]),
             proj_y=_f32(mp["proj_y"]), proj_xr=_f32(mp["proj_xr"]), level=_i32(mp["level"]), view_cos=_f32(mp["view_cos"]), desc=_u8(mp["desc"]))
    n = frame.L.ref_match_projection(frame.h, len(a["proj_x"]), _p(a["track_in_view"]), _p(a["bad"]), _p(a["has_obs"]), _p(a["proj_x"]),
                                     _p(a["proj_y"]), _p(a["proj_xr"]), _p(a["level"]), _p(a["view_cos"]), _p(a["desc"]), C.c_float(th),
                                     C.c_float(nnratio), _p(kp_match))
    return n, kp_match


def match_lastframe(cur, last, Tcw_cur, Tcw_last, th, mono, check_ori, kp_match):
    kp_match = _i32(kp_match).copy()
    a = dict(valid=_u8(last["valid"]), has_obs=_u8(last["has_obs"]), xyz=_f32(last["xyz"]), octave=_i32(last["octave"]),
             angle=_f32(last["angle"]), desc=_u8(last["desc"]))
    tc, tl = _f32(Tcw_cur).reshape(12), _f32(Tcw_last).reshape(12)
    n = cur.L.ref_match_lastframe(cur.h, len(a["valid"]), _p(a["valid"]), _p(a["has_obs"]), _p(a["xyz"]), _p(a["octave"]), _p(a["angle"]),
                                  _p(a["desc"]), _p(tc), _p(tl), C.c_float(th), int(mono), int(check_ori), _p(kp_match))
    return n, kp_match


def match_init(f1, f2, prev_matched, window, nnratio, check_ori=True):
    prev = _f32(prev_matched).reshape(-1, 2).copy()
    m12 = np.empty(f1.n, np.int32)
    n = f1.L.ref_match_init(f1.h, f2.h, _p(prev), _p(m12), int(window), C.c_float(nnratio), int(check_ori))
    return n, m12, prev


def stereo_match(exL, exR, kpsL, descL, kpsR, descR, bf, b):
    kpsL = np.ascontiguousarray(kpsL, dtype=KP_DTYPE)
    kpsR = np.ascontiguousarray(kpsR, dtype=KP_DTYPE)
    descL, descR = _u8(descL), _u8(descR)
    ur = np.empty(len(kpsL), np.float32)
    dp = np.empty(len(kpsL), np.float32)
    n = exL.L.ref_stereo_match(exL.h, exR.h, len(kpsL), _p(kpsL), _p(descL), len(kpsR), _p(kpsR), _p(descR), C.c_float(bf), C.c_float(b),
                               _p(ur), _p(dp))
    return n, ur, dp


def search_local_points(frame, lm, skip, has_obs, Tcw, th, nnratio, kp_match, cos_limit=0.5):
    """Returns (nmatches, kp_match, in_view, proj[n,5], Ow) -- Ow is the camera centre Frame::UpdatePoseMatrices computed."""
    kp_match = _i32(kp_match).copy()
    a = dict(xyz=_f32(lm["xyz"]), normal=_f32(lm["normal"]), min_dist=_f32(lm["min_dist"]), max_dist=_f32(lm["max_dist"]), desc=_u8(lm["desc"]))
    n = len(a["min_dist"])
    skip, has_obs = _u8(skip), _u8(has_obs)
    tc = _f32(Tcw).reshape(12)
    ow = np.zeros(3, np.float32)
    in_view = np.zeros(n, np.uint8)
    proj = np.zeros((n, 5), np.float32)
    nm = frame.L.ref_search_local_points(frame.h, n, _p(a["xyz"]), _p(a["normal"]), _p(a["min_dist"]), _p(a["max_dist"]), _p(a["desc"]),
                                         _p(skip), _p(has_obs), _p(tc), _p(ow), C.c_float(cos_limit), C.c_float(th), C.c_float(nnratio),
                                         _p(kp_match), _p(in_view), _p(proj))
    return nm, kp_match, in_view, proj, ow


def undistort_keypoints(kps, cam, dist5, variant="ref"):
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    out = np.empty_like(kps)
    d = None if dist5 is None else _f32(dist5).reshape(5)
    lib(variant).ref_undistort_keypoints(_p(kps), len(kps), C.byref(cam), _p(d), _p(out))
    return out


def stereo_from_rgbd(kps, kps_un, depth, mbf, factor=1.0, variant="ref"):
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    kps_un = np.ascontiguousarray(kps_un, dtype=KP_DTYPE)
    n = len(kps)
    ur, dp = np.empty(n, np.float32), np.empty(n, np.float32)
    depth = np.ascontiguousarray(depth)
    assert depth.dtype in (np.float32, np.uint16)
    kind = 1 if depth.dtype == np.float32 else 2
    lib(variant).ref_stereo_from_rgbd(_p(kps), _p(kps_un), n, _p(depth), kind, depth.shape[1], depth.shape[0], depth.strides[0],
                                      C.c_float(factor), C.c_float(mbf), _p(ur), _p(dp))
    return ur, dp


def match_bow(f1, f2, valid1, valid2, fv1, fv2, nnratio, check_ori=True, strict_low=False):
    valid1 = _u8(valid1)
    valid2 = None if valid2 is None else _u8(valid2)
    n1, s1, i1 = (_i32(a) for a in fv1)
    n2, s2, i2 = (_i32(a) for a in fv2)
    m12 = np.empty(f1.n, np.int32)
    n = f1.L.ref_match_bow(f1.h, f2.h, _p(valid1), _p(valid2), len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2), _p(i2),
                           C.c_float(nnratio), int(check_ori), int(strict_low), _p(m12))
    return n, m12


def match_triangulation(f1, f2, free1, free2, fv1, fv2, F12, t2w, only_stereo=False, check_ori=True):
    """Keyframe 1 at the origin, keyframe 2 at [I | t2w]. Returns (n, match12, (ex, ey)) with the epipole the reference's
    expression gives for these poses."""
    free1, free2 = _u8(free1), _u8(free2)
    n1, s1, i1 = (_i32(a) for a in fv1)
    n2, s2, i2 = (_i32(a) for a in fv2)
    F, t = _f32(F12).reshape(9), _f32(t2w).reshape(3)
    m12 = np.empty(f1.n, np.int32)
    ex, ey = C.c_float(), C.c_float()
    n = f1.L.ref_match_triangulation(f1.h, f2.h, _p(free1), _p(free2), len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2), _p(i2),
                                     _p(F), _p(t), int(only_stereo), int(check_ori), _p(m12), C.byref(ex), C.byref(ey))
    return n, m12, (ex.value, ey.value)


def match_reloc(cur, kf, Tcw, th, orb_dist, check_ori, kp_match):
    """Returns (nmatches, kp_match, Ow)."""
    kp_match = _i32(kp_match).copy()
    a = dict(valid=_u8(kf["valid"]), xyz=_f32(kf["xyz"]), min_dist=_f32(kf["min_dist"]), max_dist=_f32(kf["max_dist"]),
             angle=_f32(kf["angle"]), desc=_u8(kf["desc"]))
    tc = _f32(Tcw).reshape(12)
    ow = np.zeros(3, np.float32)
    n = cur.L.ref_match_reloc(cur.h, len(a["valid"]), _p(a["valid"]), _p(a["xyz"]), _p(a["min_dist"]), _p(a["max_dist"]), _p(a["angle"]),
                              _p(a["desc"]), _p(tc), _p(ow), C.c_float(th), int(orb_dist), int(check_ori), _p(kp_match))
    return n, kp_match, ow


def blur_flags(gray, boxes, variant="ref"):
    gray = _u8(gray)
    boxes = _f32(boxes).reshape(-1, 4)
    flags = np.zeros(len(boxes), np.int32)
    means = np.zeros(len(boxes), np.float64)
    lib(variant).ref_blur_flags(_p(gray), gray.shape[1], gray.shape[0], gray.strides[0], _p(boxes), len(boxes), _p(flags), _p(means))
    return flags, means
