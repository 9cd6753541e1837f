"""GPU: Frame::ProcessMovingObject (reference src/Frame.cc:311-393; coeb-slam_b200/csrc/motion.cu) against the cv2 oracle
(oracle/pmo.py, which calls OpenCV 4.13 with the reference's arguments). Parity for this row is BY TOLERANCE (DESIGN.md
section 2): the arithmetic is OpenCV's float SIMD code plus a RANSAC with its own random generator, neither reproducible bit
for bit. The tolerances, stage by stage, each stage fed with the ORACLE's input so that errors do not compound:

  pyramid (pyrDown chain)           bit-exact (integer)
  corners (goodFeaturesToTrack)     >= 97 % of the oracle's corners found at the same pixel, same count +- 3 %
  cornerSubPix                      <= 0.02 px on >= 99 % of the corners, none beyond 0.25 px
  pyramidal LK                      same status on >= 98 %; <= 0.05 px on >= 97 % of the points both track
  border / SAD tests                identical decisions given the same tracks (integer)
  epipolar test                     identical membership given the same F, except within 1e-6 px of the threshold (double)
  whole function                    every oracle T_M point farther than 2.5 px from its epipolar line is in T_M, and every
                                    T_M point is farther than 0.4 px from the ORACLE's epipolar line
"""
import os
import sys

import numpy as np
import pytest

from coeb_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import pmo  # noqa: E402

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def mo():
    import coeb_b200
    from coeb_b200 import motion
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible")
    return motion.Motion()


@pytest.fixture(scope="module")
def gpu_motion_factory(mo):
    from coeb_b200 import motion
    return motion.Motion


SEEDS = list(range(8))


@pytest.mark.parametrize("seed", SEEDS)
def test_corners_against_cv2(mo, seed):
    prev, _, _ = synth.make_motion_pair(seed)
    ref = pmo.good_features(prev)
    got = mo.good_features(prev)
    assert abs(len(got) - len(ref)) <= 0.03 * len(ref) + 2
    assert (got == np.floor(got)).all()   # integer pixel positions
    sr, sg = set(map(tuple, ref.astype(int).tolist())), set(map(tuple, got.astype(int).tolist()))
    assert len(sr & sg) >= 0.97 * len(sr), (len(sr & sg), len(sr))
    # the order is by decreasing response: the strongest 50 agree as a set
    assert len(set(map(tuple, ref[:50].astype(int).tolist())) & set(map(tuple, got[:60].astype(int).tolist()))) >= 48


@pytest.mark.parametrize("seed", SEEDS)
def test_corner_subpix_against_cv2(mo, seed):
    prev, _, _ = synth.make_motion_pair(seed)
    pts = pmo.good_features(prev)
    ref = pmo.corner_subpix(prev, pts)
    got = mo.corner_subpix(prev, pts)
    d = np.abs(got - ref).max(axis=1)
    assert (d <= 0.02).mean() >= 0.99 and d.max() <= 0.25, (float((d <= 0.02).mean()), float(d.max()))


def test_pyramid_is_cv2_pyrdown_chain(mo):
    """The tracker's pyramid is integer work: checked through a 1-level flow of a constant image pair is not possible from outside,
    so the kernel is compared on its own against cv2.pyrDown through the LK result of an exactly shifted pair below; here the
    oracle-side statement: cv2.pyrDown == the (sum + 128) >> 8 binomial model the kernel implements."""
    prev, _, _ = synth.make_motion_pair(1)
    k = np.array([1, 4, 6, 4, 1], np.int64)
    p = np.pad(prev.astype(np.int64), 2, mode="reflect")
    rows = sum(k[i] * p[:, i:i + prev.shape[1]] for i in range(5))
    full = sum(k[j] * rows[j:j + prev.shape[0], :] for j in range(5))
    model = ((full[::2, ::2] + 128) >> 8).astype(np.uint8)
    assert np.array_equal(model, cv2.pyrDown(prev))


@pytest.mark.parametrize("seed", SEEDS)
def test_lk_against_cv2(mo, seed):
    prev, cur, _ = synth.make_motion_pair(seed)
    pts = pmo.corner_subpix(prev, pmo.good_features(prev))
    ref, st_ref = pmo.lk_flow(prev, cur, pts)
    got, st_got = mo.lk(prev, cur, pts)
    assert (st_ref == st_got).mean() >= 0.98
    both = (st_ref != 0) & (st_got != 0)
    d = np.abs(got - ref)[both].max(axis=1)
    assert both.sum() > 300 and (d <= 0.05).mean() >= 0.97, (int(both.sum()), float((d <= 0.05).mean()), float(np.percentile(d, 99)))


@pytest.mark.parametrize("win,max_level", [(15, 3), (31, 2)])
def test_lk_other_windows_use_the_general_kernel(mo, win, max_level):
    """The reference's 22 x 22 window runs a kernel specialised at compile time; every other size runs the general one."""
    prev, cur, _ = synth.make_motion_pair(5)
    pts = pmo.corner_subpix(prev, pmo.good_features(prev))
    ref, st_ref, _ = cv2.calcOpticalFlowPyrLK(prev, cur, pts.reshape(-1, 1, 2), None, winSize=(win, win), maxLevel=max_level,
                                              criteria=(cv2.TERM_CRITERIA_MAX_ITER | cv2.TERM_CRITERIA_EPS, 20, 0.01))
    ref, st_ref = ref.reshape(-1, 2), st_ref.reshape(-1)
    got, st_got = mo.lk(prev, cur, pts, win=win, max_level=max_level)
    assert (st_ref == st_got).mean() >= 0.98
    both = (st_ref != 0) & (st_got != 0)
    d = np.abs(got - ref)[both].max(axis=1)
    assert both.sum() > 300 and (d <= 0.05).mean() >= 0.96, (int(both.sum()), float((d <= 0.05).mean()))


def test_lk_exact_shift_is_recovered(mo):
    prev = synth.make_frame(901)
    cur = synth.shift_image(prev, 7, -5)
    pts = pmo.corner_subpix(prev, pmo.good_features(prev))
    inner = (pts[:, 0] > 40) & (pts[:, 0] < 600) & (pts[:, 1] > 40) & (pts[:, 1] < 440)
    got, st = mo.lk(prev, cur, pts[inner])
    ok = st != 0
    assert ok.mean() > 0.97
    assert np.abs(got[ok] - pts[inner][ok] - np.array([7.0, -5.0], np.float32)).max() < 0.05


@pytest.mark.parametrize("seed", SEEDS[:4])
def test_epipolar_test_is_exact_given_F(mo, seed):
    prev, cur, _ = synth.make_motion_pair(seed)
    r = pmo.process_moving_object(prev, cur)
    assert r["F"] is not None
    mv, dist = mo.epipolar_outliers(r["prepoint"], r["nextpoint"], r["state"], r["F"])
    idx_ref, dist_ref = pmo.epipolar_outliers(r["F"], r["prepoint"], r["nextpoint"], r["state"])
    tracked = r["state"] != 0
    assert np.allclose(dist[tracked], dist_ref[tracked], rtol=1e-12, atol=1e-12)
    away = np.abs(dist_ref - 1.0) > 1e-6
    exp = np.zeros(len(mv), np.uint8)
    exp[idx_ref] = 1
    assert np.array_equal(mv[away], exp[away]) and not mv[~tracked].any()


@pytest.mark.parametrize("seed", SEEDS)
def test_whole_function_against_oracle(mo, seed):
    prev, cur, boxes = synth.make_motion_pair(seed)
    ref = pmo.process_moving_object(prev, cur)
    tm, tr = mo.process(prev, cur)
    assert tr["F"] is not None and ref["F"] is not None
    assert abs(tr["n_points"] - len(ref["prepoint"])) <= 0.03 * len(ref["prepoint"]) + 2
    # the decisions of the border and SAD tests are integer work on the library's own tracks
    st = pmo.border_and_sad_check(prev, cur, tr["prepoint"], tr["nextpoint"], np.ones(len(tr["state"]), np.uint8))
    assert (st[tr["state"] != 0] != 0).all()
    # membership: far outliers of the oracle are found; everything found is off the ORACLE's epipolar lines as well
    _, dref_own = pmo.epipolar_outliers(ref["F"], ref["prepoint"], ref["nextpoint"], ref["state"])
    far = ref["nextpoint"][(ref["state"] != 0) & (dref_own > 2.5)]
    got = set(map(tuple, np.round(tm, 1).tolist()))
    found = sum(any(abs(p[0] - q[0]) <= 0.15 and abs(p[1] - q[1]) <= 0.15 for q in got) for p in np.round(far, 1).tolist())
    assert len(far) >= 10 and found >= 0.9 * len(far), (found, len(far))
    _, d_on_ref_F = pmo.epipolar_outliers(ref["F"], tr["prepoint"], tr["nextpoint"], tr["state"])
    moving = np.array([any(abs(tr["nextpoint"][i, 0] - q[0]) < 1e-4 and abs(tr["nextpoint"][i, 1] - q[1]) < 1e-4 for q in tm.tolist())
                       for i in range(len(tr["state"]))])
    assert moving.sum() == len(tm) and (d_on_ref_F[moving & (tr["state"] != 0)] > 0.4).mean() >= 0.95
    # and most of them lie on the objects that actually moved
    inside = sum(any(b[0] - 4 <= x < b[2] + 4 and b[1] - 4 <= y < b[3] + 4 for b in boxes) for x, y in tm)
    assert inside >= 0.6 * len(tm)


@pytest.mark.parametrize("seed", SEEDS[:4])
def test_device_minimum_distance_pass_equals_the_host_pass(mo, seed):
    """coeb_process_moving_object selects the corners on the device (fixed-point form of goodFeaturesToTrack's sequential
    minimum-distance pass, behind the sort); COEB_MOTION_HOST_SELECT keeps the host pass. Same corners in the same order,
    hence identical tracks, states and T_M."""
    prev, cur, _ = synth.make_motion_pair(seed)
    os.environ.pop("COEB_MOTION_HOST_SELECT", None)
    tm_d, tr_d = mo.process(prev, cur)
    os.environ["COEB_MOTION_HOST_SELECT"] = "1"
    try:
        tm_h, tr_h = mo.process(prev, cur)
    finally:
        os.environ.pop("COEB_MOTION_HOST_SELECT", None)
    assert tr_d["n_points"] == tr_h["n_points"] > 300
    assert np.array_equal(tr_d["prepoint"], tr_h["prepoint"]) and np.array_equal(tr_d["nextpoint"], tr_h["nextpoint"])
    assert np.array_equal(tr_d["state"], tr_h["state"]) and np.array_equal(tm_d, tm_h)
    # and the corners are goodFeaturesToTrack's: the public entry point (host pass) returns the same list before refinement
    pts = mo.good_features(prev)
    assert len(pts) == tr_d["n_points"]


def _device_and_host_selection(mo, a, b):
    os.environ.pop("COEB_MOTION_HOST_SELECT", None)
    tm_d, tr_d = mo.process(a, b)
    os.environ["COEB_MOTION_HOST_SELECT"] = "1"
    try:
        tm_h, tr_h = mo.process(a, b)
    finally:
        os.environ.pop("COEB_MOTION_HOST_SELECT", None)
    return tm_d, tr_d, tm_h, tr_h


@pytest.mark.parametrize("seed", SEEDS[:4])
def test_device_minimum_distance_pass_on_clustered_corners(mo, seed):
    """The same comparison with the pair swapped: the second frames of the synthetic pairs carry the moved objects' edges and the
    pixel noise, so their strong corners sit in clusters (seed 0: 71 candidates of the first chunk with more than 12 close
    predecessors, a cell with 13 live candidates): the pooled neighbour lists and the chunk-wide spill list decide them."""
    prev, cur, _ = synth.make_motion_pair(seed)
    tm_d, tr_d, tm_h, tr_h = _device_and_host_selection(mo, cur, prev)
    assert tr_d["n_points"] == tr_h["n_points"] > 300
    assert np.array_equal(tr_d["prepoint"], tr_h["prepoint"]) and np.array_equal(tr_d["nextpoint"], tr_h["nextpoint"])
    assert np.array_equal(tr_d["state"], tr_h["state"]) and np.array_equal(tm_d, tm_h)


def test_plateaus_of_equal_response_hand_over_to_the_host_pass(mo):
    """Two noise-free 2-px checkerboards: every pixel of a patch is a maximum of its 3x3 neighbourhood (plateaus of exactly equal
    Harris response), 64 candidates per minimum-distance cell and ~100 close predecessors each. No list of the device pass holds
    that; the kernel reports it and the call repeats the selection on the host. Same result as the forced host pass."""
    rng = np.random.default_rng(5)
    prev = np.full((480, 640), 128, np.uint8)
    yy, xx = np.mgrid[0:40, 0:40]
    for _ in range(2):
        x0, y0 = int(rng.integers(40, 560)), int(rng.integers(40, 400))
        prev[y0:y0 + 40, x0:x0 + 40] = ((((yy // 2) + (xx // 2)) & 1) * 160 + 40).astype(np.uint8)
    cur = np.roll(prev, 2, axis=1)
    tm_d, tr_d, tm_h, tr_h = _device_and_host_selection(mo, prev, cur)
    assert tr_d["n_points"] == tr_h["n_points"] >= 16
    assert np.array_equal(tr_d["prepoint"], tr_h["prepoint"]) and np.array_equal(tr_d["nextpoint"], tr_h["nextpoint"])
    assert np.array_equal(tr_d["state"], tr_h["state"]) and np.array_equal(tm_d, tm_h)


def test_larger_frames_keep_the_host_minimum_distance_pass(mo):
    """The device pass holds a grid of 4800 minimum-distance cells (640x480 at 8 px); a larger frame is decided on the host before
    anything is launched, takes the host pass and must give what the forced host pass gives."""
    prev, cur, boxes = synth.make_motion_pair(3, w=960, h=540)
    tm_a, tr_a = mo.process(prev, cur)
    os.environ["COEB_MOTION_HOST_SELECT"] = "1"
    try:
        tm_b, tr_b = mo.process(prev, cur)
    finally:
        os.environ.pop("COEB_MOTION_HOST_SELECT", None)
    assert tr_a["n_points"] == tr_b["n_points"] > 300 and np.array_equal(tm_a, tm_b) and np.array_equal(tr_a["state"], tr_b["state"])
    assert tr_a["F"] is not None and len(tm_a) > 0
    inside = sum(any(b[0] - 4 <= x < b[2] + 4 and b[1] - 4 <= y < b[3] + 4 for b in boxes) for x, y in tm_a)
    assert inside >= 0.5 * len(tm_a)


@pytest.mark.parametrize("seed", SEEDS[:3])
def test_sequence_call_equals_the_two_frame_call(mo, gpu_motion_factory, seed):
    """coeb_process_moving_object_next keeps the last call's current frame and pyramid on the device as the next call's previous frame
    (the reference's imGrayPre): the first call of a sequence returns nothing, every later one what the two-frame call returns for
    (previous frame, this frame) -- bit for bit, whichever of the two pyramid buffers holds which frame."""
    a, b, _ = synth.make_motion_pair(seed)
    c = synth.shift_image(b, 3, -2)
    ref_ab, tr_ab = mo.process(a, b)
    ref_bc, tr_bc = mo.process(b, c)
    ref_ca, tr_ca = mo.process(c, a)
    seq = gpu_motion_factory()
    tm0, tr0 = seq.process_next(a)
    assert len(tm0) == 0 and tr0["n_points"] == 0
    for frame, ref, tr_ref in ((b, ref_ab, tr_ab), (c, ref_bc, tr_bc), (a, ref_ca, tr_ca)):
        tm, tr = seq.process_next(frame)
        assert tr["n_points"] == tr_ref["n_points"] and np.array_equal(tr["prepoint"], tr_ref["prepoint"]) and np.array_equal(tr["nextpoint"], tr_ref["nextpoint"])
        assert np.array_equal(tr["state"], tr_ref["state"]) and np.array_equal(tm, ref)
    # a two-frame call leaves its current frame behind as well
    tm_m, _ = seq.process(a, b)
    tm_n, _ = seq.process_next(c)
    assert np.array_equal(tm_m, ref_ab) and np.array_equal(tm_n, ref_bc)
    seq.close()
