"""oracle/pmo.py -- CPU oracle of Frame::ProcessMovingObject (reference src/Frame.cc:311-393), SURVEY.md section 8(f) row 1.

TEST INFRASTRUCTURE ONLY (see oracle/coeb_oracle.hpp). The arithmetic of this function lives entirely in OpenCV, a third-party
dependency that is not under /root/reference (required `OpenCV 3` else `>= 2.4.3`, CMakeLists.txt:34-40; pinned here to the
OpenCV 4.13.0 of the image's `cv2` module): goodFeaturesToTrack (Harris, k = 0.04, quality 0.01, min distance 8, 1000 corners),
cornerSubPix (window 10, 20 iterations, eps 0.03), calcOpticalFlowPyrLK (window 22x22, 5 levels, 20 iterations, eps 0.01),
findFundamentalMat (RANSAC 0.1 / 0.99). The oracle therefore CALLS those functions, in the reference's order and with the
reference's arguments, and restates only what the reference itself wrote around them: the 5-px border test and the 3x3 SAD test
(:337-364) and the epipolar-distance test (:372-385).

Parity for this row is by tolerance, not bit-exact: OpenCV's float paths are runtime-dispatched SIMD with
FMA contraction, its box filter carries a running column sum, and its RANSAC draws from its own RNG.
  * corners: the same integer pixel set up to response near-ties;
  * sub-pixel positions and tracked positions: within 0.01 px and 0.05 px on the same input points;
  * T_M membership: identical for the same fundamental matrix except within 1e-3 px of the 1-px epipolar threshold.
"""
import numpy as np

LIMIT_EDGE_CORNER = 5       # include/Frame.h: limit_edge_corner
LIMIT_OF_CHECK = 2120.0     # include/Frame.h: limit_of_check
LIMIT_DIS_EPI = 1.0         # include/Frame.h: limit_dis_epi (the code compares against the literal 1, :381)


def good_features(gray_prev):
    import cv2
    p = cv2.goodFeaturesToTrack(gray_prev, 1000, 0.01, 8, None, None, 3, True, 0.04)     # :333
    return np.zeros((0, 2), np.float32) if p is None else p.reshape(-1, 2).astype(np.float32)


def corner_subpix(gray_prev, pts):
    import cv2
    if len(pts) == 0:
        return pts.copy()
    q = pts.reshape(-1, 1, 2).astype(np.float32).copy()
    cv2.cornerSubPix(gray_prev, q, (10, 10), (-1, -1), (cv2.TERM_CRITERIA_MAX_ITER | cv2.TERM_CRITERIA_EPS, 20, 0.03))   # :334
    return q.reshape(-1, 2)


def lk_flow(gray_prev, gray_cur, pts):
    import cv2
    if len(pts) == 0:
        return pts.copy(), np.zeros(0, np.uint8)
    nxt, st, _ = cv2.calcOpticalFlowPyrLK(gray_prev, gray_cur, pts.reshape(-1, 1, 2).astype(np.float32), None, winSize=(22, 22), maxLevel=5,
                                          criteria=(cv2.TERM_CRITERIA_MAX_ITER | cv2.TERM_CRITERIA_EPS, 20, 0.01))                 # :335
    return nxt.reshape(-1, 2), st.reshape(-1).astype(np.uint8)


def border_and_sad_check(gray_prev, gray_cur, prepoint, nextpoint, state):
    """src/Frame.cc:336-364: returns the updated state (uint8)."""
    h, w = gray_cur.shape
    state = state.copy()
    e = LIMIT_EDGE_CORNER
    for i in range(len(state)):
        if state[i] == 0:
            continue
        x1, y1 = int(prepoint[i, 0]), int(prepoint[i, 1])      # float -> int truncation
        x2, y2 = int(nextpoint[i, 0]), int(nextpoint[i, 1])
        if x1 < e or x1 >= w - e or x2 < e or x2 >= w - e or y1 < e or y1 >= h - e or y2 < e or y2 >= h - e:
            state[i] = 0
            continue
        a = gray_prev[y1 - 1:y1 + 2, x1 - 1:x1 + 2].astype(np.int32)
        b = gray_cur[y2 - 1:y2 + 2, x2 - 1:x2 + 2].astype(np.int32)
        if float(np.abs(a - b).sum()) > LIMIT_OF_CHECK:
            state[i] = 0
    return state


def epipolar_outliers(F, prepoint, nextpoint, state):
    """src/Frame.cc:372-385 in double: indices i with state[i] != 0 whose epipolar distance exceeds 1, and the distances."""
    F = np.asarray(F, np.float64).reshape(3, 3)
    idx, dist = [], np.full(len(state), -1.0)
    for i in range(len(state)):
        if state[i] == 0:
            continue
        px, py = float(prepoint[i, 0]), float(prepoint[i, 1])
        A = F[0, 0] * px + F[0, 1] * py + F[0, 2]
        B = F[1, 0] * px + F[1, 1] * py + F[1, 2]
        C = F[2, 0] * px + F[2, 1] * py + F[2, 2]
        dd = abs(A * float(nextpoint[i, 0]) + B * float(nextpoint[i, 1]) + C) / np.sqrt(A * A + B * B)
        dist[i] = dd
        if dd <= 1:
            continue
        idx.append(i)
    return np.array(idx, np.int32), dist


def process_moving_object(gray_prev, gray_cur):
    """The whole function. Returns a dict with every intermediate (prepoint, nextpoint, state, F, T_M)."""
    import cv2
    pre = corner_subpix(gray_prev, good_features(gray_prev))
    nxt, st = lk_flow(gray_prev, gray_cur, pre)
    st = border_and_sad_check(gray_prev, gray_cur, pre, nxt, st)
    keep = st != 0
    F = None
    tm = np.zeros((0, 2), np.float32)
    out_idx = np.zeros(0, np.int32)
    if keep.sum() >= 8:
        F, _ = cv2.findFundamentalMat(pre[keep], nxt[keep], cv2.FM_RANSAC, 0.1, 0.99)      # :370
    if F is not None and F.shape == (3, 3):
        out_idx, _ = epipolar_outliers(F, pre, nxt, st)
        tm = nxt[out_idx]
    return dict(prepoint=pre, nextpoint=nxt, state=st, F=F, tm=tm, tm_index=out_idx)
