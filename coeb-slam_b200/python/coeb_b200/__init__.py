"""coeb_b200 -- thin ctypes binding of libcoeb_frontend.so (the C ABI in include/coeb_frontend.h).

This is test/bench plumbing around the C boundary; the reference-facing host API is the C++ drop-in
in include/ORBextractor.h / include/ORBmatcher.h. Nothing here computes: every call goes to the
hand-written sm_100a kernels, and loading fails loudly if the library is missing. There is no CPU
fallback and this package never imports the oracle.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.normpath(os.path.join(_HERE, "..", ".."))
LIB_PATH = os.environ.get("COEB_B200_LIB") or os.path.join(PKG_ROOT, "lib", "libcoeb_frontend.so")   # override: development builds only

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
MAX_BOXES = 32
OK, ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_CAPACITY, ERR_BAD_BOX, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6


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


class DepthImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("kind", C.c_int32), ("stride_bytes", C.c_int32), ("factor", C.c_float),
                ("on_device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


class CoebError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("coeb status %d: %s" % (status, msg))
        self.status = status


_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `make -C coeb-slam_b200` (or __graft_entry__.build())" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.coeb_last_error.restype = C.c_char_p
        _lib.coeb_version.restype = C.c_char_p
    return _lib


def _check(st):
    if st != 0:
        raise CoebError(st, lib().coeb_last_error().decode())


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw device pointer


def device_count():
    return int(lib().coeb_device_count())


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Extractor:
    """ORB_SLAM2::ORBextractor over the C ABI."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, device=0):
        self.params = OrbParams(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        h = C.c_void_p()
        _check(lib().coeb_extractor_create(C.byref(self.params), int(device), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().coeb_extractor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_cap(self, w=640, h=480):
        """Upper bound of the keypoints one w x h frame can yield (coeb_extractor_max_keypoints), rounded up to a multiple of 8."""
        n = C.c_int()
        _check(lib().coeb_extractor_max_keypoints(self.h, int(w), int(h), C.byref(n)))
        return (n.value + 7) & ~7

    def set_stream(self, stream_ptr):
        _check(lib().coeb_extractor_set_stream(self.h, C.c_void_p(int(stream_ptr) if stream_ptr else 0)))

    def reserve(self, w, h, max_batch):
        _check(lib().coeb_extractor_reserve(self.h, int(w), int(h), int(max_batch)))

    def launches_per_call(self):
        return int(lib().coeb_extractor_launches_per_call(self.h))

    STAGES = ("classify", "pyramid", "blur", "fast", "select", "describe")

    def set_profiling(self, on=True):
        _check(lib().coeb_extractor_set_profiling(self.h, int(bool(on))))

    def stage_ms(self):
        ms = (C.c_float * 6)()
        _check(lib().coeb_extractor_stage_ms(self.h, ms))
        return {k: float(ms[i]) for i, k in enumerate(self.STAGES)}

    def tables(self):
        n = self.nlevels
        nl = C.c_int()
        sc, isc, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        per = np.empty(n, np.int32)
        _check(lib().coeb_extractor_tables(self.h, C.byref(nl), _p(sc), _p(isc), _p(s2), _p(is2), _p(per)))
        return dict(nlevels=nl.value, scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, per_level=per)

    def extract(self, gray, boxes=None, tm=None, blur_flag=None, cap=None):
        """Single frame, host buffers (ORBextractor::operator())."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape
        boxes = _c(boxes if boxes is not None else np.zeros((0, 4)), np.float32).reshape(-1, 4)
        tm = _c(tm if tm is not None else np.zeros((0, 2)), np.float32).reshape(-1, 2)
        blur = _c(blur_flag if blur_flag is not None else np.zeros(len(boxes)), np.int32)
        cap = cap or self.default_cap(w, h)
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(lib().coeb_extract(self.h, _p(gray), w, h, gray.strides[0], _p(boxes), len(boxes), _p(tm), len(tm),
                                  _p(blur), len(blur), _p(kps), _p(desc), cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch_host(self, gray, boxes=None, nbox=None, tm=None, ntm=None, blur=None, cap=None, out=None,
                           check=True):
        """gray [B,h,w] u8 (host). Returns (kps [B,cap], desc [B,cap,32], counts [B], status [B])."""
        assert gray.dtype == np.uint8 and gray.ndim == 3 and gray.flags.c_contiguous
        B, h, w = gray.shape
        cap = cap or self.default_cap(w, h)
        if out is None:
            out = (np.empty((B, cap), KP_DTYPE), np.empty((B, cap, 32), np.uint8), np.empty(B, np.int32),
                   np.empty(B, np.int32))
        kps, desc, counts, status = out
        max_box = boxes.shape[1] if boxes is not None else 0
        max_tm = tm.shape[1] if tm is not None else 0
        st = lib().coeb_extract_batch_host(self.h, B, _p(gray), w, h, w, C.c_size_t(w * h), _p(boxes), _p(nbox), max_box,
                                           _p(tm), _p(ntm), max_tm, _p(blur), _p(kps), _p(desc), _p(counts), _p(status),
                                           cap)
        if check:
            _check(st)
        return kps, desc, counts, status

    def extract_batch_device(self, B, d_gray, w, h, stride, frame_stride, d_boxes, d_nbox, max_box, d_tm, d_ntm, max_tm,
                             d_blur, d_kps, d_desc, d_counts, d_status, cap):
        """All pointers are raw device addresses (ints). Enqueues on the extractor's stream."""
        _check(lib().coeb_extract_batch_device(self.h, int(B), _p(d_gray), int(w), int(h), int(stride),
                                               C.c_size_t(frame_stride), _p(d_boxes), _p(d_nbox), int(max_box), _p(d_tm),
                                               _p(d_ntm), int(max_tm), _p(d_blur), _p(d_kps), _p(d_desc), _p(d_counts),
                                               _p(d_status), int(cap)))

    def dyn_info(self, frame=0):
        d = DynInfo()
        _check(lib().coeb_extractor_dyn_info(self.h, frame, C.byref(d)))
        rects = np.array([[d.rect[i][k] for k in range(4)] for i in range(min(d.n_dynamic, MAX_BOXES))],
                         np.int32).reshape(-1, 4)
        return dict(area_flag=bool(d.area_flag), n_dynamic=d.n_dynamic, rects=rects, area=d.area)

    def level_image(self, level, frame=0, blurred=False):
        w, h, pitch = C.c_int(), C.c_int(), C.c_int()
        ptr = C.c_void_p()
        _check(lib().coeb_pyramid_level(self.h, frame, level, int(blurred), C.byref(ptr), C.byref(w), C.byref(h),
                                        C.byref(pitch)))
        img = np.empty((h.value, w.value), np.uint8)
        _check(lib().coeb_pyramid_level_copy(self.h, frame, level, int(blurred), _p(img)))
        return img

    def level_candidates(self, level, frame=0, cap=400000):
        """FAST candidates (x, y, response) as an (n,3) int array sorted into the reference's vector order
        is NOT guaranteed; callers sort."""
        out = np.empty(cap, np.uint32)
        n = C.c_int()
        _check(lib().coeb_debug_candidates(self.h, frame, level, _p(out), cap, C.byref(n)))
        p = out[:n.value]
        return np.stack([p & 0xFFF, (p >> 12) & 0xFFF, p >> 24], axis=1).astype(np.int64)

    def level_keys(self, level, frame=0, cap=20000):
        out = np.empty((cap, 4), np.float32)
        n = C.c_int()
        _check(lib().coeb_debug_level_keys(self.h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()


def host_alloc(nbytes):
    p = C.c_void_p()
    _check(lib().coeb_host_alloc(C.byref(p), C.c_size_t(nbytes)))
    return p


def host_free(p):
    _check(lib().coeb_host_free(p))


def pinned_array(shape, dtype):
    """numpy array backed by cudaHostAlloc'ed (pinned) memory; keep the returned owner alive."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = host_alloc(max(n, 1))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p


# ---------------------------------------------------------------------------------------------------
# Frame grid + ORBmatcher over the C ABI
# ---------------------------------------------------------------------------------------------------
def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Matcher:
    """Device context (stream + scratch) of the ORBmatcher kernels."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().coeb_matcher_create(int(device), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().coeb_matcher_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        _check(lib().coeb_matcher_set_stream(self.h, C.c_void_p(int(stream_ptr) if stream_ptr else 0)))

    def hamming(self, a, b):
        a, b = _u8(a).reshape(-1, 32), _u8(b).reshape(-1, 32)
        out = np.empty(len(a), np.int32)
        _check(lib().coeb_hamming256_batch(self.h, _p(a), _p(b), len(a), _p(out)))
        return out

    def frame(self, kps, desc, cam, scale, uright=None):
        return Frame(self, kps, desc, cam, scale, uright)

    def frame_from_extractor(self, ex, cam, n=-1, frame_index=0, dist5=None, depth=None, depth_factor=1.0, download=True):
        """Frame constructor tail on the device (UndistortKeyPoints + ComputeStereoFromRGBD + grid) from the extractor's
        resident output. depth: None, a float32 map or a uint16 map (scaled by depth_factor), host numpy arrays.
        Returns (frame, keys_un, uright, depth) (the arrays are None when download is False)."""
        return ExtractedFrame.build(self, ex, cam, n, frame_index, dist5, depth, depth_factor, download)

    def local_map(self, lm):
        return LocalMap(self, lm)

    def search_local_points(self, frame, local_map, skip, has_obs, Tcw, Ow, th, nnratio, kp_match, cos_limit=0.5, want_proj=True):
        kp_match = _i32(kp_match).copy()
        skip, has_obs = _u8(skip), _u8(has_obs)
        tc, ow = _f32(Tcw).reshape(12), _f32(Ow).reshape(3)
        in_view = np.zeros(local_map.n, np.uint8)
        proj = np.zeros((local_map.n, 5), np.float32) if want_proj else None
        n = C.c_int()
        _check(lib().coeb_search_local_points(self.h, frame.h, local_map.h, _p(skip), _p(has_obs), _p(tc), _p(ow),
                                              C.c_float(cos_limit), C.c_float(th), C.c_float(nnratio), _p(kp_match),
                                              _p(in_view), _p(proj), C.byref(n)))
        return n.value, kp_match, in_view, proj

    def match_bow(self, f1, f2, valid1, valid2, fv1, fv2, nnratio, check_ori=True, strict_low=False):
        """ORBmatcher::SearchByBoW; fv = (node, start, items) CSR of a DBoW2 FeatureVector. Returns (n, match12)."""
        valid1 = _u8(valid1)
        valid2 = None if valid2 is None else _u8(valid2)
        n1, s1, i1 = (_i32(a) for a in fv1)
        n2, s2, i2 = (_i32(a) for a in fv2)
        m12 = np.empty(f1.n, np.int32)
        n = C.c_int()
        _check(lib().coeb_match_bow(self.h, f1.h, f2.h, _p(valid1), _p(valid2), len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2),
                                    _p(s2), _p(i2), C.c_float(nnratio), int(check_ori), int(strict_low), _p(m12), C.byref(n)))
        return n.value, m12

    def match_triangulation(self, f1, f2, free1, free2, fv1, fv2, F12, epipole, only_stereo=False, check_ori=True):
        """ORBmatcher::SearchForTriangulation. Returns (n, match12)."""
        free1, free2 = _u8(free1), _u8(free2)
        n1, s1, i1 = (_i32(a) for a in fv1)
        n2, s2, i2 = (_i32(a) for a in fv2)
        F, e = _f32(F12).reshape(9), _f32(epipole).reshape(2)
        m12 = np.empty(f1.n, np.int32)
        n = C.c_int()
        _check(lib().coeb_match_triangulation(self.h, f1.h, f2.h, _p(free1), _p(free2), len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2),
                                              _p(s2), _p(i2), _p(F), _p(e), int(only_stereo), int(check_ori), _p(m12), C.byref(n)))
        return n.value, m12

    def match_reloc(self, cur, kf, Tcw, Ow, th, orb_dist, check_ori, kp_match):
        """Relocalisation SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist). Returns (n, kp_match)."""
        kp_match = _i32(kp_match).copy()
        a = dict(valid=_u8(kf["valid"]), xyz=_f32(kf["xyz"]), min_dist=_f32(kf["min_dist"]), max_dist=_f32(kf["max_dist"]),
                 angle=_f32(kf["angle"]), desc=_u8(kf["desc"]))
        tc, ow = _f32(Tcw).reshape(12), _f32(Ow).reshape(3)
        n = C.c_int()
        _check(lib().coeb_match_reloc(self.h, cur.h, len(a["valid"]), _p(a["valid"]), _p(a["xyz"]), _p(a["min_dist"]), _p(a["max_dist"]),
                                      _p(a["angle"]), _p(a["desc"]), _p(tc), _p(ow), C.c_float(th), int(orb_dist), int(check_ori),
                                      _p(kp_match), C.byref(n)))
        return n.value, kp_match

    def fuse_search(self, frame, local_map, valid, Tcw, Ow, th, chi2_tests=True):
        """Search half of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th). Returns (nFused, best_idx[n])."""
        valid = _u8(valid)
        tc, ow = _f32(Tcw).reshape(12), _f32(Ow).reshape(3)
        best = np.empty(local_map.n, np.int32)
        n = C.c_int()
        _check(lib().coeb_fuse_search(self.h, frame.h, local_map.h, _p(valid), _p(tc), _p(ow), C.c_float(th), int(chi2_tests), _p(best), C.byref(n)))
        return n.value, best

    def match_projection(self, frame, mp, th, nnratio, kp_match):
        kp_match = _i32(kp_match).copy()
        a = dict(track_in_view=_u8(mp["track_in_view"]), bad=_u8(mp["bad"]), has_obs=_u8(mp["has_obs"]),
                 proj_x=_f32(mp["proj_x"]), proj_y=_f32(mp["proj_y"]), proj_xr=_f32(mp["proj_xr"]),
                 level=_i32(mp["level"]), view_cos=_f32(mp["view_cos"]), desc=_u8(mp["desc"]))
        n = C.c_int()
        _check(lib().coeb_match_projection(self.h, frame.h, len(a["proj_x"]), _p(a["track_in_view"]), _p(a["bad"]),
                                           _p(a["has_obs"]), _p(a["proj_x"]), _p(a["proj_y"]), _p(a["proj_xr"]),
                                           _p(a["level"]), _p(a["view_cos"]), _p(a["desc"]), C.c_float(th),
                                           C.c_float(nnratio), _p(kp_match), C.byref(n)))
        return n.value, kp_match

    def match_lastframe(self, cur, last, Tcw_cur, Tcw_last, th, mono, check_ori, kp_match):
        kp_match = _i32(kp_match).copy()
        a = dict(valid=_u8(last["valid"]), has_obs=_u8(last["has_obs"]), xyz=_f32(last["xyz"]),
                 octave=_i32(last["octave"]), angle=_f32(last["angle"]), desc=_u8(last["desc"]))
        tc, tl = _f32(Tcw_cur).reshape(12), _f32(Tcw_last).reshape(12)
        n = C.c_int()
        _check(lib().coeb_match_lastframe(self.h, cur.h, len(a["valid"]), _p(a["valid"]), _p(a["has_obs"]), _p(a["xyz"]),
                                          _p(a["octave"]), _p(a["angle"]), _p(a["desc"]), _p(tc), _p(tl), C.c_float(th),
                                          int(mono), int(check_ori), _p(kp_match), C.byref(n)))
        return n.value, kp_match

    def match_init(self, f1, f2, prev_matched, window, nnratio, check_ori=True):
        prev = _f32(prev_matched).reshape(-1, 2).copy()
        m12 = np.empty(f1.n, np.int32)
        n = C.c_int()
        _check(lib().coeb_match_init(self.h, f1.h, f2.h, _p(prev), _p(m12), int(window), C.c_float(nnratio),
                                     int(check_ori), C.byref(n)))
        return n.value, m12, prev

    def stereo_match(self, exL, exR, kpsL, descL, kpsR, descR, bf, b):
        kpsL = np.ascontiguousarray(kpsL, dtype=KP_DTYPE)
        kpsR = np.ascontiguousarray(kpsR, dtype=KP_DTYPE)
        descL, descR = _u8(descL), _u8(descR)
        ur = np.empty(len(kpsL), np.float32)
        dp = np.empty(len(kpsL), np.float32)
        n = C.c_int()
        _check(lib().coeb_stereo_match(self.h, exL.h, exR.h, len(kpsL), _p(kpsL), _p(descL), len(kpsR), _p(kpsR),
                                       _p(descR), C.c_float(bf), C.c_float(b), _p(ur), _p(dp), C.byref(n)))
        return n.value, ur, dp

    def stereo_match_frames(self, exL, frame_left, exR, frame_right, kpsL, descL, kpsR, descR, bf, b):
        """ComputeStereoMatches on the pyramids of (exL, frame_left) and (exR, frame_right): e.g. frames 0 and 1 of one two-frame call."""
        kpsL = np.ascontiguousarray(kpsL, dtype=KP_DTYPE)
        kpsR = np.ascontiguousarray(kpsR, dtype=KP_DTYPE)
        descL, descR = _u8(descL), _u8(descR)
        ur = np.empty(len(kpsL), np.float32)
        dp = np.empty(len(kpsL), np.float32)
        n = C.c_int()
        _check(lib().coeb_stereo_match_frames(self.h, exL.h, int(frame_left), exR.h, int(frame_right), len(kpsL), _p(kpsL), _p(descL), len(kpsR), _p(kpsR),
                                              _p(descR), C.c_float(bf), C.c_float(b), _p(ur), _p(dp), C.byref(n)))
        return n.value, ur, dp

    def knn2(self, q, t, nnratio):
        q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
        idx, d1, d2 = (np.empty(len(q), np.int32) for _ in range(3))
        n = C.c_int()
        _check(lib().coeb_knn2(self.h, _p(q), len(q), _p(t), len(t), C.c_float(nnratio), _p(idx), _p(d1), _p(d2),
                               C.byref(n)))
        return idx, d1, d2, n.value

    def knn2_device(self, d_q, nq, d_t, nt, nnratio, d_idx, d_d1, d_d2):
        _check(lib().coeb_knn2_device(self.h, _p(d_q), int(nq), _p(d_t), int(nt), C.c_float(nnratio), _p(d_idx),
                                      _p(d_d1), _p(d_d2)))


class Frame:
    """Device-resident view of ORB_SLAM2::Frame for the matchers (undistorted keypoints + 64x48 grid)."""

    def __init__(self, matcher, kps, desc, cam, scale, uright=None):
        self.kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
        self.desc = _u8(desc).reshape(-1, 32)
        self.n = len(self.kps)
        self.scale = _f32(scale)
        self.uright = None if uright is None else _f32(uright)
        self.matcher = matcher
        h = C.c_void_p()
        _check(lib().coeb_frame_create(matcher.h, _p(self.kps), _p(self.desc), self.n, _p(self.uright), C.byref(cam),
                                       _p(self.scale), len(self.scale), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().coeb_frame_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.empty(self.n + 1, np.int32)
        n = C.c_int()
        _check(lib().coeb_frame_features_in_area(self.h, C.c_float(x), C.c_float(y), C.c_float(r), int(min_level),
                                                 int(max_level), _p(out), len(out), C.byref(n)))
        return out[:n.value].copy()


class ExtractedFrame(Frame):
    """A Frame built on the device from an Extractor's last call (coeb_frame_from_extractor)."""

    def __init__(self):   # use build()
        self.h = None

    @classmethod
    def build(cls, matcher, ex, cam, n, frame_index, dist5, depth, depth_factor, download):
        f = cls()
        f.matcher = matcher
        d5 = None if dist5 is None else _f32(dist5).reshape(5)
        dimg = None
        if depth is not None:
            depth = np.ascontiguousarray(depth)   # the call is blocking, so this temporary outlives the upload
            assert depth.dtype in (np.float32, np.uint16) and depth.ndim == 2
            dimg = DepthImage(depth.ctypes.data, 1 if depth.dtype == np.float32 else 2, depth.strides[0], float(depth_factor), 0,
                              depth.shape[1], depth.shape[0])
        cap = ex.default_cap() if n < 0 else max(n, 1)
        keys_un = np.empty(cap, KP_DTYPE) if download else None
        ur = np.empty(cap, np.float32) if download else None
        dp = np.empty(cap, np.float32) if download else None
        h = C.c_void_p()
        nn = C.c_int()
        _check(lib().coeb_frame_from_extractor(matcher.h, ex.h, int(frame_index), int(n), C.byref(cam), _p(d5),
                                               C.byref(dimg) if dimg is not None else None, _p(keys_un), _p(ur), _p(dp),
                                               C.byref(nn), C.byref(h)))
        f.h = h
        f.n = nn.value
        if download:
            keys_un, ur, dp = keys_un[:f.n].copy(), ur[:f.n].copy(), dp[:f.n].copy()
        return f, keys_un, ur, dp


class LocalMap:
    """Tracking::mvpLocalMapPoints resident on the device. lm: dict of arrays (xyz, normal, min_dist, max_dist, desc)."""

    def __init__(self, matcher, lm):
        a = dict(xyz=_f32(lm["xyz"]).reshape(-1, 3), normal=_f32(lm["normal"]).reshape(-1, 3), min_dist=_f32(lm["min_dist"]),
                 max_dist=_f32(lm["max_dist"]), desc=_u8(lm["desc"]).reshape(-1, 32))
        self.n = len(a["min_dist"])
        self.matcher = matcher
        h = C.c_void_p()
        _check(lib().coeb_local_map_create(matcher.h, self.n, _p(a["xyz"]), _p(a["normal"]), _p(a["min_dist"]), _p(a["max_dist"]),
                                           _p(a["desc"]), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().coeb_local_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
