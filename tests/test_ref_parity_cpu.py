"""CPU: oracle B (oracle/coeb_oracle*.hpp) against oracle/_ref -- the reference's OWN sources
(/root/reference/src/{ORBextractor,ORBmatcher,Frame,MapPoint,KeyFrame,Map}.cc) compiled unchanged against the test-only
OpenCV shim of oracle/ref_shim/. This is the pin that ties the oracle (and through it the CUDA path, tests/test_*_gpu.py)
to the reference itself rather than to a reading of it.

What is compared, all bit for bit:
  * extraction: keypoints (all seven fields), descriptors, every pyramid level, on the normal, area_flag and blur-flag
    paths, other nFeatures, a KITTI-shaped frame -- against the MONOTONIC-HEAP build of the reference (see below);
  * Frame grid + GetFeaturesInArea, SearchByProjection (local map / last frame), SearchForInitialization,
    ComputeStereoMatches, isInFrustum + PredictScale + SearchByProjection in Tracking::SearchLocalPoints order,
    UndistortKeyPoints, ComputeStereoFromRGBD, DescriptorDistance -- against the plain build.

The octree's heap-address tie-break (src/ORBextractor.cc:691 sorts pair<int, ExtractorNode*>): with glibc's allocator the
reference's own output depends on the allocation history of the process -- test_reference_is_not_reproducible_under_glibc_heap
shows the SAME frame giving different keypoints from the SAME extractor instance. Oracle B and the CUDA path resolve the
tie as "created later first", which is what strictly increasing addresses give; the monotonic-heap build of the
reference (operator new inside the library never reuses an address) is that reference, and must equal oracle B everywhere.
"""
import os
import sys

import numpy as np
import pytest

import orc
from coeb_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import ref  # noqa: E402

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref is not built and /root/reference is not present")

CAM_ARGS = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)


def _dyn(seed):
    return synth.make_dynamic(seed, force_area=(seed % 8 == 3))


def _same(a, b):
    return len(a[0]) == len(b[0]) and a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])


# --------------------------------------------------------------------------------------------------------------------
# extraction
# --------------------------------------------------------------------------------------------------------------------
def test_tables_equal_reference_constructor():
    for nf, sf, nl in ((1000, 1.2, 8), (2000, 1.2, 8), (500, 1.5, 5)):
        r, o = ref.Extractor(nf, sf, nl).tables(), orc.Extractor(nf, sf, nl).tables()
        for k in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
            assert r[k].tobytes() == o[k].tobytes(), k


def test_extraction_equals_reference_monotonic_heap_40_frames():
    R, O = ref.Extractor(variant="mono"), orc.Extractor()
    n_area = n_box = 0
    for seed in range(40):
        gray = synth.make_frame(seed)
        boxes, tm, blur = _dyn(seed)
        kr, ko = R.extract(gray, boxes, tm, blur), O.extract(gray, boxes, tm, blur)
        assert _same(kr, ko), seed
        n_area += bool(O.dyn_info()["area_flag"])
        n_box += O.dyn_info()["n_dynamic"] > 0
        for lvl in range(8):
            assert np.array_equal(R.level_image(lvl), O.level_image(lvl)), (seed, lvl)
    assert n_area >= 4 and n_box >= 15     # the area_flag (30/10, x0.7, cull before the octree) and the per-box paths ran
    assert R.heap_fallbacks() == 0         # every allocation of the reference stayed inside the monotonic arena


def test_extraction_blur_flag_layer2_path():
    """A box with one T_M point inside and blur_flag = 1 is dynamic through 'layer 2' only (src/ORBextractor.cc:1168-1184)."""
    R, O = ref.Extractor(variant="mono"), orc.Extractor()
    gray = synth.make_frame(77)
    boxes = np.array([[100, 60, 380, 420], [400, 100, 600, 400]], np.float32)
    tm = np.array([[150.5, 200.25], [500.0, 300.0]], np.float32)
    for blur in ([1, 0], [0, 1], [1, 1], [0, 0], [1]):
        a, b = R.extract(gray, boxes, tm, np.array(blur, np.int32)), O.extract(gray, boxes, tm, np.array(blur, np.int32))
        assert _same(a, b), blur
    assert O.dyn_info()["n_dynamic"] == 1   # the last call: only box 0 has a flag


@pytest.mark.parametrize("nf,w,h", [(500, 640, 480), (2000, 640, 480), (2000, 1241, 376), (300, 320, 240)])
def test_extraction_other_sizes_and_quotas(nf, w, h):
    R, O = ref.Extractor(nfeatures=nf, variant="mono"), orc.Extractor(nfeatures=nf)
    for seed in (500, 501):
        gray = synth.make_frame(seed, w, h)
        assert _same(R.extract(gray), O.extract(gray)), seed


def test_reference_border_is_reflect101_of_the_level():
    """The 19-px frame the reference builds around every level (src/ORBextractor.cc:1347-1364) is a pure function of
    the level (BORDER_REFLECT_101), so not materialising it loses nothing."""
    R = ref.Extractor()
    R.extract(synth.make_frame(5))
    for lvl in (0, 3, 7):
        inner, framed = R.level_image(lvl), R.level_image(lvl, border=19)
        assert np.array_equal(framed, np.pad(inner, 19, mode="reflect"))


def test_reference_is_not_reproducible_under_glibc_heap():
    """The same frame through the same reference extractor, before and after other frames: with glibc's allocator the
    keypoint lists differ (heap-address tie-break), while pyramid, count and the bulk of the keypoints agree."""
    R = ref.Extractor(variant="ref")
    frames = [synth.make_frame(s) for s in range(6)]
    first = [R.extract(f) for f in frames]
    again = [R.extract(f) for f in frames]
    differing = sum(a[0].tobytes() != b[0].tobytes() for a, b in zip(first, again))
    assert differing >= 1
    for a, b in zip(first, again):
        sa = set(zip(a[0]["x"].tolist(), a[0]["y"].tolist(), a[0]["octave"].tolist()))
        sb = set(zip(b[0]["x"].tolist(), b[0]["y"].tolist(), b[0]["octave"].tolist()))
        assert len(sa & sb) >= 0.95 * max(len(sa), len(sb))


def test_glibc_heap_build_differs_from_oracle_only_by_tie_break():
    """Plain reference build (glibc heap, FMA contraction allowed) against oracle B: the keypoints both select carry
    bit-identical angle, response, size and descriptor; what differs is which node of an equal-size run was expanded.
    Prints the numbers DESIGN.md quotes."""
    R, O = ref.Extractor(variant="ref"), orc.Extractor()
    orc.octree_tie_stats(reset=True)
    common = only = 0
    for seed in range(24):
        gray = synth.make_frame(seed)
        boxes, tm, blur = _dyn(seed)
        (kr, dr), (ko, do) = R.extract(gray, boxes, tm, blur), O.extract(gray, boxes, tm, blur)
        ir = {(x, y, o): i for i, (x, y, o) in enumerate(zip(kr["x"].tolist(), kr["y"].tolist(), kr["octave"].tolist()))}
        io = {(x, y, o): i for i, (x, y, o) in enumerate(zip(ko["x"].tolist(), ko["y"].tolist(), ko["octave"].tolist()))}
        both = sorted(set(ir) & set(io))
        a, b = np.array([ir[k] for k in both]), np.array([io[k] for k in both])
        assert kr[a].tobytes() == ko[b].tobytes()          # same seven fields, bit for bit
        assert np.array_equal(dr[a], do[b])                # same 256 descriptor bits
        common += len(both)
        only += len(set(ir) ^ set(io))
        assert abs(len(kr) - len(ko)) <= 0.02 * len(ko)
    ties = orc.octree_tie_stats()
    print("glibc-heap reference vs oracle B: %d common keypoints, %d in the symmetric difference (%.2f%%); octree calls %d, "
          "careful-phase rounds %d, with an order tie %d, with a cut-off tie %d"
          % (common, only, 100.0 * only / (2 * common + only), ties["calls"], ties["careful_rounds"], ties["order_tie_rounds"],
             ties["cut_tie_rounds"]))
    assert only <= 0.03 * common
    assert ties["cut_tie_rounds"] > 0


# --------------------------------------------------------------------------------------------------------------------
# Frame grid, matchers, stereo, frame tail -- the reference's own classes
# --------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tum():
    ex = orc.Extractor()
    kps, desc = ex.extract(synth.make_frame(100))
    return dict(kps=kps, desc=desc, scale=ex.tables()["scale"])


# The matchers run against two builds of the reference: "nofma" (-ffp-contract=off, the arithmetic oracle B and the CUDA
# path pin) and "ref" (the reference's own flags, where gcc contracts e.g. `fx * PcX * invz + cx`, src/Frame.cc:464-465 and
# src/ORBmatcher.cc:1370-1371, into a fused multiply-add). Decisions and match indices must be identical against BOTH;
# projected floats are bit-identical against "nofma" and within one rounding against "ref".
@pytest.fixture(scope="module", params=["nofma", "ref"])
def variant(request):
    return request.param


def _pair(t, kps=None, desc=None, uright=None, scale=None, variant="ref"):
    kps = t["kps"] if kps is None else kps
    desc = t["desc"] if desc is None else desc
    scale = t["scale"] if scale is None else scale
    cam = orc.Camera(*CAM_ARGS)
    return orc.Frame(kps, desc, cam, scale, uright), ref.Frame(kps, desc, cam, scale, uright, variant=variant)


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=(512, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(512, 32), dtype=np.uint8)
    b[:4] = a[:4]
    b[4:8] = ~a[4:8]
    for i in range(512):
        assert ref.hamming256(a[i], b[i]) == orc.hamming256(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())


def test_grid_and_window_queries(tum, variant):
    fo, fr = _pair(tum, variant=variant)
    for ix in range(0, 64, 7):
        for iy in range(0, 48, 5):
            assert np.array_equal(fo.grid_cell(ix, iy), fr.grid_cell(ix, iy))
    rng = np.random.default_rng(1)
    total = 0
    for _ in range(400):
        x, y = float(rng.uniform(-30, 670)), float(rng.uniform(-30, 510))
        r = float(rng.choice([3.0, 7.5, 15.0, 36.0, 100.0]))
        lv = [(-1, -1), (0, 0), (2, 3), (1, -1), (0, 4)][int(rng.integers(0, 5))]
        a, b = fo.features_in_area(x, y, r, *lv), fr.features_in_area(x, y, r, *lv)
        assert np.array_equal(a, b), (x, y, r, lv)
        total += len(a)
    assert total > 1000


@pytest.mark.parametrize("th,ratio", [(3.0, 0.8), (1.0, 0.8), (5.0, 0.6)])
def test_search_by_projection_local_map(tum, th, ratio, variant):
    mp, uright = synth.make_map_points(tum["kps"], tum["desc"], tum["scale"], seed=int(th * 10))
    fo, fr = _pair(tum, uright=uright, variant=variant)
    state = np.random.default_rng(5).choice([-1, -1, -1, -1, -2, -3], size=len(tum["kps"])).astype(np.int32)
    (n_o, k_o), (n_r, k_r) = orc.match_projection(fo, mp, th, ratio, state), ref.match_projection(fr, mp, th, ratio, state)
    assert n_o == n_r and np.array_equal(k_o, k_r) and n_o > 100


def test_search_by_projection_contention(tum, variant):
    kps, desc = tum["kps"][:60], tum["desc"][:60]
    rng = np.random.default_rng(7)
    n_map = 3000
    src = rng.integers(0, len(kps), n_map)
    mp = dict(track_in_view=np.ones(n_map, np.uint8), bad=(rng.random(n_map) < 0.05).astype(np.uint8),
              has_obs=(rng.random(n_map) < 0.9).astype(np.uint8),
              proj_x=(kps["x"][src] + rng.normal(0, 6, n_map)).astype(np.float32),
              proj_y=(kps["y"][src] + rng.normal(0, 6, n_map)).astype(np.float32), proj_xr=np.zeros(n_map, np.float32),
              level=np.minimum(kps["octave"][src] + 1, 7).astype(np.int32), view_cos=rng.uniform(0.99, 1.0, n_map).astype(np.float32),
              desc=synth.flip_bits(desc[src], rng, 20))
    fo, fr = _pair(tum, kps=kps, desc=desc, variant=variant)
    state = np.full(len(kps), -1, np.int32)
    a, b = orc.match_projection(fo, mp, 3.0, 0.8, state), ref.match_projection(fr, mp, 3.0, 0.8, state)
    assert (a[0], a[1].tolist()) == (b[0], b[1].tolist())


@pytest.mark.parametrize("th,mono,ori,shift", [(15.0, False, True, None), (30.0, True, True, None), (7.0, False, False, None),
                                               (15.0, False, True, (0.0, 0.0, -0.5)), (15.0, False, True, (0.0, 0.0, 0.5))])
def test_search_by_projection_last_frame(tum, th, mono, ori, shift, variant):
    kw = {} if shift is None else dict(shift=shift)
    last, Tc, Tl = synth.make_last_frame(tum["kps"], tum["desc"], seed=int(th), **kw)
    uright = np.where(np.random.default_rng(3).random(len(tum["kps"])) < 0.3, tum["kps"]["x"] - 10.0, -1.0).astype(np.float32)
    fo, fr = _pair(tum, uright=uright, variant=variant)
    state = np.full(len(tum["kps"]), -1, np.int32)
    a, b = orc.match_lastframe(fo, last, Tc, Tl, th, mono, ori, state), ref.match_lastframe(fr, last, Tc, Tl, th, mono, ori, state)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 50


@pytest.mark.parametrize("dx,dy,nf,window", [(20, 10, 2000, 100), (-20, -10, 2000, 100), (3, 1, 1000, 100), (20, 10, 2000, 400)])
def test_search_for_initialization(dx, dy, nf, window, variant):
    ex = orc.Extractor(nfeatures=nf)
    g1 = synth.make_frame(200)
    k1, d1 = ex.extract(g1)
    k2, d2 = ex.extract(synth.shift_image(g1, dx, dy))
    scale = ex.tables()["scale"]
    t = dict(kps=k1, desc=d1, scale=scale)
    f1o, f1r = _pair(t, variant=variant)
    f2o, f2r = _pair(t, kps=k2, desc=d2, variant=variant)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    a, b = orc.match_init(f1o, f2o, prev, window, 0.9, True), ref.match_init(f1r, f2r, prev, window, 0.9, True)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[0] > 20


def test_stereo_matches_kitti_shape(variant):
    w, h, nf = 1241, 376, 2000
    left = synth.make_frame(300, w, h)
    right = synth.make_stereo_right(left, seed=300)
    ol, orr = orc.Extractor(nfeatures=nf), orc.Extractor(nfeatures=nf)
    rl, rr = ref.Extractor(nfeatures=nf, variant="mono"), ref.Extractor(nfeatures=nf, variant="mono")
    kl, dl = ol.extract(left)
    kr, dr = orr.extract(right)
    assert _same(rl.extract(left), (kl, dl)) and _same(rr.extract(right), (kr, dr))
    bf, b = 386.1448, 386.1448 / 718.856
    n_o, ur_o, dp_o = orc.stereo_match(ol, orr, kl, dl, kr, dr, bf, b)
    n_r, ur_r, dp_r = ref.stereo_match(rl, rr, kl, dl, kr, dr, bf, b)
    assert n_o == n_r and n_o > 100
    assert ur_o.tobytes() == ur_r.tobytes() and dp_o.tobytes() == dp_r.tobytes()
    # the plain builds hold the same pyramids (they are integer work), so Frame::ComputeStereoMatches can run on them too
    pl, pr = ref.Extractor(nfeatures=nf, variant=variant), ref.Extractor(nfeatures=nf, variant=variant)
    pl.extract(left)
    pr.extract(right)
    n_p, ur_p, dp_p = ref.stereo_match(pl, pr, kl, dl, kr, dr, bf, b)
    same_match = (ur_p > 0) == (ur_o > 0)
    assert n_p == n_o and same_match.all()
    if variant == "nofma":
        assert ur_o.tobytes() == ur_p.tobytes() and dp_o.tobytes() == dp_p.tobytes()
    else:   # the sub-pixel parabola and `mbf / disparity` may be contracted
        assert np.allclose(ur_o, ur_p, rtol=1e-6, atol=1e-4) and np.allclose(dp_o, dp_p, rtol=1e-5)


@pytest.mark.parametrize("seed,th", [(2, 3.0), (4, 1.0), (6, 5.0)])
def test_search_local_points(tum, seed, th, variant):
    """Frame::isInFrustum + MapPoint::PredictScale + SearchByProjection, in Tracking::SearchLocalPoints order."""
    kps, desc, scale = tum["kps"], tum["desc"], tum["scale"]
    Tcw, Ow = synth.make_pose(seed)
    lm, skip, has_obs = synth.make_local_map(kps, desc, scale, Tcw, seed=seed, n_map=3000, n_true=600)
    rng = np.random.default_rng(8)
    uright = np.where(rng.random(len(kps)) < 0.3, kps["x"] - np.float32(40.0) / rng.uniform(0.5, 5.0, len(kps)).astype(np.float32),
                      np.float32(-1)).astype(np.float32)
    fo, fr = _pair(tum, uright=uright, variant=variant)
    state = rng.choice([-1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
    n_r, k_r, v_r, p_r, ow_r = ref.search_local_points(fr, lm, skip, has_obs, Tcw, th, 0.8, state)
    assert np.allclose(ow_r, Ow, atol=1e-6)
    # Ow is an INPUT of the C ABI: the drop-in adapter hands over the reference's own mOw (include/ORBmatcher.h)
    n_o, k_o, v_o, p_o = orc.search_local_points(fo, lm, skip, has_obs, Tcw, ow_r, th, 0.8, state)
    assert np.array_equal(v_o, v_r) and v_o.sum() > 500
    # u, v, uR, viewCos, predicted level of every visible point
    if variant == "nofma":
        assert p_o.tobytes() == p_r.tobytes()
    else:   # the fused `fx * PcX * invz + cx` of the reference's own flags rounds once instead of twice
        assert np.array_equal(p_o[:, 3:], p_r[:, 3:])                # viewCos and the predicted level do not go through it
        # one rounding of a product of magnitude <= 1024 px (fx * PcX * invz before cx is added): 2^-13 px
        assert (np.abs(p_o[:, :3] - p_r[:, :3]) <= 2 * np.spacing(np.float32(1024.0))).all()
        assert ((p_o[:, :3] != p_r[:, :3]).any(axis=1)).mean() < 0.5
    assert n_o == n_r and np.array_equal(k_o, k_r) and n_o > 100


def test_undistort_and_rgbd_depth():
    tum1 = dict(cam=(517.306408, 516.469215, 318.643040, 255.313989, 40.0, 40.0 / 517.306408, 0.0, 640.0, 0.0, 480.0),
                dist=(0.262383, -0.953104, -0.005358, 0.002628, 1.163314))
    cam = orc.Camera(*tum1["cam"])
    rng = np.random.default_rng(5)
    kps = np.zeros(3000, orc.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(0, 639, 3000), rng.uniform(0, 479, 3000)
    kps["octave"], kps["angle"] = rng.integers(0, 8, 3000), rng.uniform(0, 360, 3000)
    for dist in (tum1["dist"], (-0.2, 0.05, 0.001, -0.0005, 0.0), (0.0, 0.5, 0.0, 0.0, 0.0)):
        uo, ur = orc.undistort_keypoints(kps, cam, dist), ref.undistort_keypoints(kps, cam, dist)
        assert uo.tobytes() == ur.tobytes()
    un = orc.undistort_keypoints(kps, cam, tum1["dist"])
    d16 = synth.make_depth(3)
    f = np.float32(1.0) / np.float32(5000.0)
    a, b = orc.stereo_from_rgbd(kps, un, d16, cam.bf, f), ref.stereo_from_rgbd(kps, un, d16, cam.bf, f)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()
    dm = (d16.astype(np.float32) * f).astype(np.float32)
    a, b = orc.stereo_from_rgbd(kps, un, dm, cam.bf), ref.stereo_from_rgbd(kps, un, dm, cam.bf)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes() and (a[1] > 0).mean() > 0.5


# --------------------------------------------------------------------------------------------------------------------
# the "next" rows of SURVEY.md section 8(f) that go through KeyFrame objects, and the blur flag
# --------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def two_views():
    ex = orc.Extractor(nfeatures=600)
    g1 = synth.make_frame(301)
    k1, d1 = ex.extract(g1)
    k2, d2 = ex.extract(synth.shift_image(g1, -5, 3))
    return dict(k1=k1, d1=d1, k2=k2, d2=d2, scale=ex.tables()["scale"])


@pytest.mark.parametrize("ratio,ori,strict,use_valid2,n_nodes", [(0.7, True, False, False, 60), (0.9, True, True, True, 60),
                                                                 (0.9, False, False, False, 5), (0.75, True, True, False, 20)])
def test_search_by_bow_both_overloads(two_views, ratio, ori, strict, use_valid2, n_nodes, variant):
    v = two_views
    t1, t2 = dict(kps=v["k1"], desc=v["d1"], scale=v["scale"]), dict(kps=v["k2"], desc=v["d2"], scale=v["scale"])
    f1o, f1r = _pair(t1, variant=variant)
    f2o, f2r = _pair(t2, variant=variant)
    rng = np.random.default_rng(3)
    valid1 = (rng.random(len(v["k1"])) < 0.85).astype(np.uint8)
    valid2 = (rng.random(len(v["k2"])) < 0.9).astype(np.uint8) if use_valid2 else None
    fv1, fv2 = synth.make_feature_vector(v["d1"], n_nodes, seed=4), synth.make_feature_vector(v["d2"], n_nodes, seed=4)
    a = orc.match_bow(f1o, f2o, valid1, valid2, fv1, fv2, ratio, ori, strict)
    b = ref.match_bow(f1r, f2r, valid1, valid2, fv1, fv2, ratio, ori, strict)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 30


@pytest.mark.parametrize("only_stereo,t2w", [(False, (-0.0375, -0.0883, 1.0)), (True, (1.5, 1.5, 0.001)), (False, (0.2, 0.1, 0.5))])
def test_search_for_triangulation(two_views, only_stereo, t2w, variant):
    v = two_views
    rng = np.random.default_rng(6)
    ur1 = np.where(rng.random(len(v["k1"])) < 0.6, v["k1"]["x"] - np.float32(5), np.float32(-1)).astype(np.float32)
    ur2 = np.where(rng.random(len(v["k2"])) < 0.6, v["k2"]["x"] - np.float32(5), np.float32(-1)).astype(np.float32)
    t1, t2 = dict(kps=v["k1"], desc=v["d1"], scale=v["scale"]), dict(kps=v["k2"], desc=v["d2"], scale=v["scale"])
    f1o, f1r = _pair(t1, uright=ur1, variant=variant)
    f2o, f2r = _pair(t2, uright=ur2, variant=variant)
    free1, free2 = (rng.random(len(v["k1"])) < 0.8).astype(np.uint8), (rng.random(len(v["k2"])) < 0.8).astype(np.uint8)
    fv1, fv2 = synth.make_feature_vector(v["d1"], 30, seed=7), synth.make_feature_vector(v["d2"], 30, seed=7)
    F12 = np.array([[0, 0, 3.0], [0, 0, 5.0], [-3.0, -5.0, 0]], np.float32)   # pure shift (-5, 3)
    n_r, m_r, epi = ref.match_triangulation(f1r, f2r, free1, free2, fv1, fv2, F12, t2w, only_stereo, True)
    # the epipole as the callers of the C ABI compute it: fp32, left to right, no FMA (src/ORBmatcher.cc:663-670)
    f32 = np.float32
    invz = f32(1.0) / f32(t2w[2])
    ex = f32(f32(f32(f32(CAM_ARGS[0]) * f32(t2w[0])) * invz) + f32(CAM_ARGS[2]))
    ey = f32(f32(f32(f32(CAM_ARGS[1]) * f32(t2w[1])) * invz) + f32(CAM_ARGS[3]))
    if variant == "nofma":
        assert (f32(epi[0]), f32(epi[1])) == (ex, ey)
    n_o, m_o = orc.match_triangulation(f1o, f2o, free1, free2, fv1, fv2, F12, (ex, ey), only_stereo, True)
    assert n_o == n_r and np.array_equal(m_o, m_r) and n_o > 10


@pytest.mark.parametrize("th,orb_dist,ori,seed", [(10.0, 100, True, 1), (3.0, 64, True, 2), (60.0, 100, False, 3)])
def test_relocalisation_search_by_projection(th, orb_dist, ori, seed, variant):
    ex = orc.Extractor()
    kps, desc = ex.extract(synth.make_frame(200 + seed))
    scale = ex.tables()["scale"]
    fo, fr = _pair(dict(kps=kps, desc=desc, scale=scale), variant=variant)
    Tcw, Ow = synth.make_pose(seed)
    lm, skip, _ = synth.make_local_map(kps, desc, scale, Tcw, seed=seed, n_map=1500, n_true=700)
    rng = np.random.default_rng(seed)
    kf = dict(valid=(1 - skip).astype(np.uint8), xyz=lm["xyz"], min_dist=lm["min_dist"], max_dist=lm["max_dist"],
              angle=rng.uniform(0, 360, len(skip)).astype(np.float32), desc=lm["desc"])
    state = rng.choice([-1, -1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
    n_r, k_r, ow_r = ref.match_reloc(fr, kf, Tcw, th, orb_dist, ori, state)
    assert np.allclose(ow_r, Ow, atol=1e-6)
    # Ow is an INPUT of the C ABI: the drop-in adapter hands over the reference's own mOw (include/ORBmatcher.h)
    n_o, k_o = orc.match_reloc(fo, kf, Tcw, ow_r, th, orb_dist, ori, state)
    assert n_o == n_r and np.array_equal(k_o, k_r) and n_o > (5 if ori else 100)


def test_blur_flags_of_the_frame_constructor():
    """Frame::detect_laplacian per box and the 4.2 threshold (src/Frame.cc:171-202, 905-913)."""
    rng = np.random.default_rng(11)
    for seed in range(6):
        gray = synth.make_frame(seed).copy()
        if seed % 2:   # a smooth region so that some boxes fall below the threshold
            gray[100:400, 150:500] = (gray[100:400, 150:500].astype(np.int32) // 16 * 16 + 8).astype(np.uint8)
            gray[120:380, 170:480] = 128
        boxes = []
        for _ in range(4):
            x0, y0 = int(rng.integers(0, 400)), int(rng.integers(0, 250))
            boxes.append([x0, y0, x0 + int(rng.integers(40, 230)), y0 + int(rng.integers(40, 220))])
        boxes = np.array(boxes, np.float32)
        (fo, mo), (fr, mr) = orc.blur_flags(gray, boxes), ref.blur_flags(gray, boxes)
        assert np.array_equal(fo, fr) and mo.tobytes() == mr.tobytes()
