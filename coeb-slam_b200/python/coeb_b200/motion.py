"""ctypes binding of the Frame::ProcessMovingObject entry points of libcoeb_frontend.so (include/coeb_frontend.h,
coeb-slam_b200/csrc/motion.cu). Test / bench plumbing only; nothing here computes."""
import ctypes as C

import numpy as np

from . import _check, _p, lib

TRACE_POINTS = 1000


class MotionTrace(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("n_tracked", C.c_int32), ("n_inliers", C.c_int32), ("have_F", C.c_int32),
                ("pre_xy", C.c_float * (2 * TRACE_POINTS)), ("next_xy", C.c_float * (2 * TRACE_POINTS)), ("state", C.c_uint8 * TRACE_POINTS),
                ("F", C.c_double * 9)]


def _gray(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


class Motion:
    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().coeb_motion_create(int(device), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().coeb_motion_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def good_features(self, gray, max_corners=1000, quality=0.01, min_distance=8.0, harris_k=0.04):
        gray = _gray(gray)
        out = np.empty((max(max_corners, 1), 2), np.float32)
        n = C.c_int()
        _check(lib().coeb_motion_good_features(self.h, _p(gray), gray.shape[1], gray.shape[0], gray.strides[0], int(max_corners), C.c_double(quality),
                                               C.c_double(min_distance), C.c_double(harris_k), _p(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def corner_subpix(self, gray, pts, half_win=10, max_iters=20, eps=0.03):
        gray = _gray(gray)
        q = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2).copy()
        _check(lib().coeb_motion_corner_subpix(self.h, _p(gray), gray.shape[1], gray.shape[0], gray.strides[0], _p(q), len(q), int(half_win), int(max_iters),
                                               C.c_double(eps)))
        return q

    def lk(self, prev, cur, pts, win=22, max_level=5, max_iters=20, eps=0.01, min_eig=1e-4):
        prev, cur = _gray(prev), _gray(cur)
        assert prev.shape == cur.shape
        p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
        nxt = np.zeros_like(p)
        st = np.zeros(len(p), np.uint8)
        _check(lib().coeb_motion_lk(self.h, _p(prev), _p(cur), prev.shape[1], prev.shape[0], prev.strides[0], _p(p), len(p), int(win), int(max_level),
                                    int(max_iters), C.c_double(eps), C.c_double(min_eig), _p(nxt), _p(st)))
        return nxt, st

    def epipolar_outliers(self, pre, nxt, state, F, limit=1.0):
        pre = np.ascontiguousarray(pre, dtype=np.float32).reshape(-1, 2)
        nxt = np.ascontiguousarray(nxt, dtype=np.float32).reshape(-1, 2)
        state = np.ascontiguousarray(state, dtype=np.uint8)
        F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
        mv = np.zeros(len(pre), np.uint8)
        dist = np.zeros(len(pre), np.float64)
        _check(lib().coeb_epipolar_outliers(self.h, _p(pre), _p(nxt), _p(state), len(pre), _p(F), C.c_double(limit), _p(mv), _p(dist)))
        return mv, dist

    def process(self, prev, cur, cap=1000):
        """Returns (T_M [n,2], trace dict)."""
        prev, cur = _gray(prev), _gray(cur)
        assert prev.shape == cur.shape
        tm = np.zeros((cap, 2), np.float32)
        n = C.c_int()
        tr = MotionTrace()
        _check(lib().coeb_process_moving_object(self.h, _p(prev), _p(cur), prev.shape[1], prev.shape[0], prev.strides[0], _p(tm), cap, C.byref(n), C.byref(tr)))
        k = min(tr.n_points, TRACE_POINTS)
        trace = dict(n_points=tr.n_points, n_tracked=tr.n_tracked, n_inliers=tr.n_inliers,
                     prepoint=np.array(tr.pre_xy[:2 * k], np.float32).reshape(-1, 2), nextpoint=np.array(tr.next_xy[:2 * k], np.float32).reshape(-1, 2),
                     state=np.array(tr.state[:k], np.uint8), F=np.array(tr.F[:], np.float64).reshape(3, 3) if tr.have_F else None)
        return tm[:n.value].copy(), trace

    def process_next(self, cur, cap=1000):
        """Sequence form (coeb_process_moving_object_next): the previous frame is the current frame of the last process / process_next
        call, resident on the device. Returns (T_M [n,2], trace dict); the first call of a sequence returns no points."""
        cur = _gray(cur)
        tm = np.zeros((cap, 2), np.float32)
        n = C.c_int()
        tr = MotionTrace()
        _check(lib().coeb_process_moving_object_next(self.h, _p(cur), cur.shape[1], cur.shape[0], cur.strides[0], _p(tm), cap, C.byref(n), C.byref(tr)))
        k = min(tr.n_points, TRACE_POINTS)
        trace = dict(n_points=tr.n_points, n_tracked=tr.n_tracked, n_inliers=tr.n_inliers,
                     prepoint=np.array(tr.pre_xy[:2 * k], np.float32).reshape(-1, 2), nextpoint=np.array(tr.next_xy[:2 * k], np.float32).reshape(-1, 2),
                     state=np.array(tr.state[:k], np.uint8), F=np.array(tr.F[:], np.float64).reshape(3, 3) if tr.have_F else None)
        return tm[:n.value].copy(), trace

    def prepared_process_next(self, frames, cap=1000):
        """The sequence call with its ctypes arguments built once: returns a zero-argument callable that feeds the frames round robin
        (each call pairs the previous call's frame with the next one) and returns the number of T_M points."""
        frames = [_gray(f) for f in frames]
        tm = np.zeros((cap, 2), np.float32)
        n = C.c_int()
        fn = lib().coeb_process_moving_object_next
        args = [(self.h, _p(f), f.shape[1], f.shape[0], f.strides[0], _p(tm), cap, C.byref(n), None) for f in frames]
        state = {"i": 0}

        def call():
            _check(fn(*args[state["i"]]))
            state["i"] = (state["i"] + 1) % len(args)
            return n.value
        call.keep = (frames, tm, n)
        return call

    def prepared_process(self, prev, cur, cap=1000):
        """The C call with its ctypes arguments built once (what a C++ caller pays): returns a zero-argument callable that runs
        coeb_process_moving_object on the two frames and returns the number of T_M points (results stay in the prepared buffers)."""
        prev, cur = _gray(prev), _gray(cur)
        assert prev.shape == cur.shape
        tm = np.zeros((cap, 2), np.float32)
        n = C.c_int()
        fn = lib().coeb_process_moving_object
        a = (self.h, _p(prev), _p(cur), prev.shape[1], prev.shape[0], prev.strides[0], _p(tm), cap, C.byref(n), None)
        keep = (prev, cur, tm, n)

        def call():
            _check(fn(*a))
            return n.value
        call.keep = keep
        return call


def fundamental_ransac(p1, p2, threshold=0.1, confidence=0.99, max_iters=1000, seed=12345):
    """Host code of the library (no device needed). Returns (F [3,3], inlier mask)."""
    p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 2)
    F = np.zeros(9, np.float64)
    mask = np.zeros(len(p1), np.uint8)
    n = C.c_int()
    _check(lib().coeb_fundamental_ransac(_p(p1), _p(p2), len(p1), C.c_double(threshold), C.c_double(confidence), int(max_iters), C.c_uint(seed), _p(F), _p(mask),
                                         C.byref(n)))
    return F.reshape(3, 3), mask
