"""GPU parity of the steps either side of the extractor and the matchers ("next" rows of SURVEY.md section 8f): the
Frame constructor tail built on the device from the extractor's resident output (UndistortKeyPoints,
ComputeStereoFromRGBD, AssignFeaturesToGrid) and Tracking::SearchLocalPoints (isInFrustum + SearchByProjection) against
a device-resident local map. Everything is compared bit for bit with the CPU oracle, floats included."""
import numpy as np
import pytest

import orc
from coeb_b200 import synth

pytestmark = pytest.mark.gpu

TUM1_CAM = (517.306408, 516.469215, 318.643040, 255.313989, 40.0, 40.0 / 517.306408, 0.0, 640.0, 0.0, 480.0)
TUM1_DIST = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)
TUM3_CAM = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
DEPTH_FACTOR = np.float32(1.0) / np.float32(5000.0)


@pytest.fixture(scope="module")
def gpu():
    import coeb_b200
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible")
    return coeb_b200


def _extract_both(gpu, seed, dynamic=True):
    gray = synth.make_frame(seed)
    boxes, tm, blur = synth.make_dynamic(seed) if dynamic else (None, None, None)
    g, c = gpu.Extractor(), orc.Extractor()
    kg, dg = g.extract(gray, boxes, tm, blur)
    kc, dc = c.extract(gray, boxes, tm, blur)
    assert kg.tobytes() == kc.tobytes() and np.array_equal(dg, dc)
    return g, c, kc, dc


@pytest.mark.parametrize("cam_args,dist,depth_kind", [(TUM1_CAM, TUM1_DIST, "u16"), (TUM3_CAM, None, "f32"), (TUM1_CAM, TUM1_DIST, None),
                                                     (TUM3_CAM, (0.0, 0.1, 0.0, 0.0, 0.0), "u16")])
def test_frame_from_extractor_matches_oracle(gpu, cam_args, dist, depth_kind):
    g, c, kps, desc = _extract_both(gpu, 21)
    scale = c.tables()["scale"]
    d16 = synth.make_depth(21)
    depth = {None: None, "u16": d16, "f32": (d16.astype(np.float32) * DEPTH_FACTOR).astype(np.float32)}[depth_kind]
    m = gpu.Matcher()
    fg, keys_un, ur, dp = m.frame_from_extractor(g, gpu.Camera(*cam_args), n=len(kps), dist5=dist, depth=depth, depth_factor=DEPTH_FACTOR)
    ref_un = orc.undistort_keypoints(kps, orc.Camera(*cam_args), dist)
    ref_ur, ref_dp = orc.stereo_from_rgbd(kps, ref_un, depth, cam_args[4], DEPTH_FACTOR)
    assert fg.n == len(kps)
    assert keys_un.tobytes() == ref_un.tobytes()
    assert ur.tobytes() == ref_ur.tobytes() and dp.tobytes() == ref_dp.tobytes()
    if dist is not None and dist[0] != 0.0:
        assert np.abs(keys_un["x"] - kps["x"]).max() > 0.5   # the distortion really moved points
    else:
        assert keys_un.tobytes() == kps.tobytes()
    if depth is not None:
        assert 0.4 < (dp > 0).mean() < 0.98
    # the device-built grid answers window queries like the oracle's Frame over the same undistorted keypoints
    fc = orc.Frame(ref_un, desc, orc.Camera(*cam_args), scale, ref_ur if depth is not None else None)
    rng = np.random.default_rng(3)
    for _ in range(150):
        x, y, r = rng.uniform(-20, 660), rng.uniform(-20, 500), rng.uniform(1, 60)
        lo = int(rng.integers(-1, 5))
        hi = lo + int(rng.integers(0, 3)) if lo >= 0 else -1
        assert np.array_equal(fg.features_in_area(x, y, r, lo, hi), fc.features_in_area(x, y, r, lo, hi))
    # ... and feeds the matchers: frame-to-frame SearchByProjection against the oracle on the same inputs
    last, Tc, Tl = synth.make_last_frame(ref_un, desc, seed=4, fx=cam_args[0], fy=cam_args[1], cx=cam_args[2], cy=cam_args[3])
    state = np.full(len(kps), -1, np.int32)
    ng, kg = m.match_lastframe(fg, last, Tc, Tl, 15.0, False, True, state)
    nc, kc = orc.match_lastframe(fc, last, Tc, Tl, 15.0, False, True, state)
    assert ng == nc and np.array_equal(kg, kc) and ng > 100


def test_frame_from_extractor_reads_the_count_from_the_device(gpu):
    g, c, kps, desc = _extract_both(gpu, 5, dynamic=False)
    m = gpu.Matcher()
    fg, keys_un, ur, dp = m.frame_from_extractor(g, gpu.Camera(*TUM3_CAM), n=-1)
    assert fg.n == len(kps) and keys_un.tobytes() == kps.tobytes() and (ur == -1).all() and (dp == -1).all()
    # a frame of a batch call: frame 2 of 3
    b = synth.make_batch(3, base_seed=40)
    kb, db, counts, status = g.extract_batch_host(b["gray"], b["boxes"], b["nbox"], b["tm"], b["ntm"], b["blur"])
    f2, k2, _, _ = m.frame_from_extractor(g, gpu.Camera(*TUM3_CAM), n=int(counts[2]), frame_index=2)
    assert k2.tobytes() == kb[2, :counts[2]].tobytes()
    with pytest.raises(gpu.CoebError):
        m.frame_from_extractor(g, gpu.Camera(*TUM3_CAM), n=10 ** 6)


@pytest.mark.parametrize("th,ratio,seed", [(3.0, 0.8, 2), (1.0, 0.8, 3), (5.0, 0.8, 4), (3.0, 0.6, 5)])
def test_search_local_points_matches_oracle(gpu, th, ratio, seed):
    g, c, kps, desc = _extract_both(gpu, 100 + seed)
    scale = c.tables()["scale"]
    d16 = synth.make_depth(seed)
    m = gpu.Matcher()
    fg, keys_un, ur, dp = m.frame_from_extractor(g, gpu.Camera(*TUM3_CAM), n=len(kps), depth=d16, depth_factor=DEPTH_FACTOR)
    fc = orc.Frame(keys_un, desc, orc.Camera(*TUM3_CAM), scale, ur)
    Tcw, Ow = synth.make_pose(seed)
    lm, skip, has_obs = synth.make_local_map(keys_un, desc, scale, Tcw, seed=seed)
    rng = np.random.default_rng(seed)
    state = rng.choice([-1, -1, -1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
    dev_map = m.local_map(lm)
    ng, kg, vg, pg = m.search_local_points(fg, dev_map, skip, has_obs, Tcw, Ow, th, ratio, state)
    nc, kc, vc, pc = orc.search_local_points(fc, lm, skip, has_obs, Tcw, Ow, th, ratio, state)
    assert np.array_equal(vg, vc), "isInFrustum decisions differ"
    assert 0.2 < vc.mean() < 0.8
    assert pg.tobytes() == pc.tobytes(), "projected fields (u, v, uR, viewCos, level) differ"
    assert ng == nc and np.array_equal(kg, kc)
    assert nc > 150
    # a second call on the same resident map with another pose (the per-frame use)
    Tcw2, Ow2 = synth.make_pose(seed + 50)
    ng, kg, vg, pg = m.search_local_points(fg, dev_map, skip, has_obs, Tcw2, Ow2, th, ratio, state)
    nc, kc, vc, pc = orc.search_local_points(fc, lm, skip, has_obs, Tcw2, Ow2, th, ratio, state)
    assert np.array_equal(vg, vc) and pg.tobytes() == pc.tobytes() and ng == nc and np.array_equal(kg, kc)


def test_search_local_points_edge_cases(gpu):
    g, c, kps, desc = _extract_both(gpu, 7, dynamic=False)
    scale = c.tables()["scale"]
    m = gpu.Matcher()
    fg = m.frame(kps, desc, gpu.Camera(*TUM3_CAM), scale, None)
    fc = orc.Frame(kps, desc, orc.Camera(*TUM3_CAM), scale, None)
    Tcw, Ow = synth.make_pose(1)
    lm, skip, has_obs = synth.make_local_map(kps, desc, scale, Tcw, seed=1, n_map=2000, n_true=700)
    state = np.full(len(kps), -1, np.int32)
    dev_map = m.local_map(lm)
    # very wide windows overflow the 32-entry candidate lists: the window-walking fallback must give the same answer
    ng, kg, vg, pg = m.search_local_points(fg, dev_map, skip, has_obs, Tcw, Ow, 40.0, 0.8, state)
    nc, kc, vc, pc = orc.search_local_points(fc, lm, skip, has_obs, Tcw, Ow, 40.0, 0.8, state)
    assert ng == nc and np.array_equal(kg, kc) and np.array_equal(vg, vc)
    # everything skipped: nothing in view, nothing matched, the keypoint state comes back untouched
    ng, kg, vg, _ = m.search_local_points(fg, dev_map, np.ones_like(skip), has_obs, Tcw, Ow, 3.0, 0.8, state)
    assert ng == 0 and vg.sum() == 0 and np.array_equal(kg, state)
    # a stricter viewing-cosine limit only removes points
    _, _, v9, _ = m.search_local_points(fg, dev_map, skip, has_obs, Tcw, Ow, 3.0, 0.8, state, cos_limit=0.9)
    _, _, c9, _ = orc.search_local_points(fc, lm, skip, has_obs, Tcw, Ow, 3.0, 0.8, state, cos_limit=0.9)
    assert np.array_equal(v9, c9) and v9.sum() < vc.sum() and not (v9 & ~vc).any()
    # empty map
    empty = {k: a[:0] for k, a in lm.items()}
    ng, kg, vg, _ = m.search_local_points(fg, m.local_map(empty), skip[:0], has_obs[:0], Tcw, Ow, 3.0, 0.8, state)
    assert ng == 0 and len(vg) == 0 and np.array_equal(kg, state)


@pytest.mark.parametrize("th,orb_dist,ori,seed", [(10.0, 100, True, 1), (3.0, 64, True, 2), (10.0, 100, False, 3), (60.0, 100, True, 4)])
def test_relocalisation_search_by_projection_matches_oracle(gpu, th, orb_dist, ori, seed):
    """SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) of Tracking::Relocalization: every non-null entry of
    mvpMapPoints blocks, points behind the camera are NOT rejected (the reference has no such test), the level comes from
    PredictScale. th = 60 makes windows wide enough to overflow the 64-entry candidate lists (window-walking fallback)."""
    g, c, kps, desc = _extract_both(gpu, 200 + seed, dynamic=False)
    scale = c.tables()["scale"]
    m = gpu.Matcher()
    fg = m.frame(kps, desc, gpu.Camera(*TUM3_CAM), scale, None)
    fc = orc.Frame(kps, desc, orc.Camera(*TUM3_CAM), scale, None)
    Tcw, Ow = synth.make_pose(seed)
    lm, skip, _ = synth.make_local_map(kps, desc, scale, Tcw, seed=seed, n_map=1500, n_true=700)
    rng = np.random.default_rng(seed)
    kf = dict(valid=(1 - skip).astype(np.uint8), xyz=lm["xyz"], min_dist=lm["min_dist"], max_dist=lm["max_dist"],
              angle=rng.uniform(0, 360, len(skip)).astype(np.float32), desc=lm["desc"])
    state = rng.choice([-1, -1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
    ng, kg = m.match_reloc(fg, kf, Tcw, Ow, th, orb_dist, ori, state)
    nc, kc = orc.match_reloc(fc, kf, Tcw, Ow, th, orb_dist, ori, state)
    assert ng == nc and np.array_equal(kg, kc)
    assert nc > (5 if ori else 100)
    assert np.array_equal(kg[state != -1], state[state != -1]), "occupied entries must come back untouched"
    got = kg[kg >= 0]
    assert len(np.unique(got)) == len(got), "a keyframe map point was assigned twice"


@pytest.mark.parametrize("th,seed,stereo,chi2", [(3.0, 1, True, True), (3.0, 2, False, True), (1.0, 3, True, True), (25.0, 4, True, True),
                                                 (4.0, 5, True, False), (10.0, 6, False, False)])
def test_fuse_search_matches_oracle(gpu, th, seed, stereo, chi2):
    """Search half of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th): the keypoint every map point would be fused into."""
    g, c, kps, desc = _extract_both(gpu, 300 + seed, dynamic=False)
    scale = c.tables()["scale"]
    rng = np.random.default_rng(seed)
    uright = np.where(rng.random(len(kps)) < 0.5, kps["x"] - np.float32(40.0) / rng.uniform(0.5, 5.0, len(kps)).astype(np.float32),
                      np.float32(-1)).astype(np.float32) if stereo else None
    m = gpu.Matcher()
    fg = m.frame(kps, desc, gpu.Camera(*TUM3_CAM), scale, uright)
    fc = orc.Frame(kps, desc, orc.Camera(*TUM3_CAM), scale, uright)
    Tcw, Ow = synth.make_pose(seed)
    lm, skip, _ = synth.make_local_map(kps, desc, scale, Tcw, seed=seed, n_map=3000, n_true=min(900, len(kps)))
    valid = (1 - skip).astype(np.uint8)
    ng, bg = m.fuse_search(fg, m.local_map(lm), valid, Tcw, Ow, th, chi2)
    nc, bc = orc.fuse_search(fc, lm, valid, Tcw, Ow, th, chi2)
    assert ng == nc and np.array_equal(bg, bc)
    assert nc > 50 and (bc[valid == 0] == -1).all()
