"""GPU parity of the keypoint grid, the Hamming matchers, the stereo matcher and the brute-force k=2 matcher
against the CPU oracle: match indices, counts and Hamming-derived decisions must be bit-exact; stereo
uRight/depth are compared as exact floats (same fp32 operation order, no FMA)."""
import numpy as np
import pytest

import orc
from coeb_b200 import synth

pytestmark = pytest.mark.gpu

NLEVELS = 8


@pytest.fixture(scope="module")
def gpu():
    import coeb_b200
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible")
    return coeb_b200


@pytest.fixture(scope="module")
def tum(gpu):
    """One extracted 640x480 frame (CPU oracle output == GPU output, see test_extract_gpu) + camera."""
    ex = orc.Extractor()
    gray = synth.make_frame(100)
    kps, desc = ex.extract(gray)
    scale = ex.tables()["scale"]
    cam_args = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
    return dict(gray=gray, kps=kps, desc=desc, scale=scale, cam_args=cam_args)


def _frames(gpu, t, kps=None, desc=None, uright=None):
    kps = t["kps"] if kps is None else kps
    desc = t["desc"] if desc is None else desc
    m = gpu.Matcher()
    fg = m.frame(kps, desc, gpu.Camera(*t["cam_args"]), t["scale"], uright)
    fc = orc.Frame(kps, desc, orc.Camera(*t["cam_args"]), t["scale"], uright)
    return m, fg, fc


def test_hamming_batch(gpu):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=(4096, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(4096, 32), dtype=np.uint8)
    b[:10] = a[:10]
    b[10:20] = ~a[10:20]
    got = gpu.Matcher().hamming(a, b)
    ref = np.unpackbits(a ^ b, axis=1).sum(axis=1)
    assert np.array_equal(got, ref)
    assert got[:10].max() == 0 and got[10:20].min() == 256
    assert all(orc.hamming256(a[i], b[i]) == ref[i] for i in range(0, 4096, 97))


def test_grid_window_queries_match_traversal_order(gpu, tum):
    m, fg, fc = _frames(gpu, tum)
    rng = np.random.default_rng(1)
    total = 0
    for _ in range(300):
        x, y = float(rng.uniform(-30, 670)), float(rng.uniform(-30, 510))
        r = float(rng.choice([3.0, 7.5, 15.0, 36.0, 100.0]))
        lv = [(-1, -1), (0, 0), (2, 3), (1, -1), (0, 4)][int(rng.integers(0, 5))]
        a, b = fg.features_in_area(x, y, r, *lv), fc.features_in_area(x, y, r, *lv)
        assert np.array_equal(a, b), (x, y, r, lv)
        total += len(a)
    assert total > 1000


@pytest.mark.parametrize("th,ratio", [(3.0, 0.8), (1.0, 0.8), (5.0, 0.6)])
def test_search_by_projection_local_map_5k(gpu, tum, th, ratio):
    mp, uright = synth.make_map_points(tum["kps"], tum["desc"], tum["scale"], seed=int(th * 10))
    m, fg, fc = _frames(gpu, tum, uright=uright)
    rng = np.random.default_rng(5)
    state = rng.choice([-1, -1, -1, -1, -2, -3], size=len(tum["kps"])).astype(np.int32)  # claims left by earlier matching
    n_c, km_c = orc.match_projection(fc, mp, th, ratio, state)
    n_g, km_g = m.match_projection(fg, mp, th, ratio, state)
    assert n_c == n_g
    assert np.array_equal(km_c, km_g)
    assert n_c > 100


def test_search_by_projection_contention(gpu, tum):
    # many map points competing for few keypoints: long claim chains for the fixed-point iteration
    kps, desc = tum["kps"][:60], tum["desc"][:60]
    rng = np.random.default_rng(7)
    n_map = 3000
    src = rng.integers(0, len(kps), n_map)
    mp = dict(track_in_view=np.ones(n_map, np.uint8), bad=np.zeros(n_map, np.uint8), has_obs=(rng.random(n_map) < 0.9).astype(np.uint8),
              proj_x=(kps["x"][src] + rng.normal(0, 6, n_map)).astype(np.float32),
              proj_y=(kps["y"][src] + rng.normal(0, 6, n_map)).astype(np.float32),
              proj_xr=np.zeros(n_map, np.float32), level=np.minimum(kps["octave"][src] + 1, 7).astype(np.int32),
              view_cos=np.full(n_map, 0.95, np.float32), desc=synth.flip_bits(desc[src], rng, 20))
    m, fg, fc = _frames(gpu, tum, kps=kps, desc=desc)
    state = np.full(len(kps), -1, np.int32)
    n_c, km_c = orc.match_projection(fc, mp, 3.0, 0.8, state)
    n_g, km_g = m.match_projection(fg, mp, 3.0, 0.8, state)
    assert (n_c, km_c.tolist()) == (n_g, km_g.tolist())


@pytest.mark.parametrize("th,mono,ori", [(15.0, False, True), (30.0, True, True), (7.0, False, False)])
def test_search_by_projection_last_frame(gpu, tum, th, mono, ori):
    last, Tc, Tl = synth.make_last_frame(tum["kps"], tum["desc"], seed=int(th))
    rng = np.random.default_rng(3)
    uright = np.where(rng.random(len(tum["kps"])) < 0.3, tum["kps"]["x"] - 10.0, -1.0).astype(np.float32)
    m, fg, fc = _frames(gpu, tum, uright=uright)
    state = np.full(len(tum["kps"]), -1, np.int32)  # Tracking clears mvpMapPoints before the call (Tracking.cc:943)
    n_c, km_c = orc.match_lastframe(fc, last, Tc, Tl, th, mono, ori, state)
    n_g, km_g = m.match_lastframe(fg, last, Tc, Tl, th, mono, ori, state)
    assert n_c == n_g
    assert np.array_equal(km_c, km_g)
    assert n_c > 50


@pytest.mark.parametrize("forward", [True, False])
def test_last_frame_forward_backward_level_windows(gpu, tum, forward):
    last, Tc, Tl = synth.make_last_frame(tum["kps"], tum["desc"], seed=9, shift=(0.0, 0.0, -0.5 if forward else 0.5))
    m, fg, fc = _frames(gpu, tum)
    state = np.full(len(tum["kps"]), -1, np.int32)
    n_c, km_c = orc.match_lastframe(fc, last, Tc, Tl, 15.0, False, True, state)
    n_g, km_g = m.match_lastframe(fg, last, Tc, Tl, 15.0, False, True, state)
    assert (n_c, km_c.tolist()) == (n_g, km_g.tolist())


@pytest.mark.parametrize("dx,dy,nf", [(20, 10, 2000), (-20, -10, 2000), (3, 1, 1000)])
def test_search_for_initialization(gpu, dx, dy, nf):
    ex = orc.Extractor(nfeatures=nf)
    g1 = synth.make_frame(200)
    g2 = synth.shift_image(g1, dx, dy)
    k1, d1 = ex.extract(g1)
    k2, d2 = ex.extract(g2)
    scale = ex.tables()["scale"]
    cam_args = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
    m = gpu.Matcher()
    f1g, f2g = m.frame(k1, d1, gpu.Camera(*cam_args), scale), m.frame(k2, d2, gpu.Camera(*cam_args), scale)
    f1c, f2c = orc.Frame(k1, d1, orc.Camera(*cam_args), scale), orc.Frame(k2, d2, orc.Camera(*cam_args), scale)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)  # vbPrevMatched starts as the F1 keypoints
    n_c, m_c, p_c = orc.match_init(f1c, f2c, prev, 100, 0.9, True)
    n_g, m_g, p_g = m.match_init(f1g, f2g, prev, 100, 0.9, True)
    assert n_c == n_g
    assert np.array_equal(m_c, m_g)
    assert np.array_equal(p_c, p_g)
    assert n_c > 20


def test_stereo_matches_kitti_shape(gpu):
    w, h, nf = 1241, 376, 2000
    left = synth.make_frame(300, w, h)
    right = synth.make_stereo_right(left, seed=300)
    gl, gr = gpu.Extractor(nfeatures=nf), gpu.Extractor(nfeatures=nf)
    cl, cr = orc.Extractor(nfeatures=nf), orc.Extractor(nfeatures=nf)
    kl, dl = gl.extract(left)
    kr, dr = gr.extract(right)
    kl2, dl2 = cl.extract(left)
    kr2, dr2 = cr.extract(right)
    assert kl.tobytes() == kl2.tobytes() and kr.tobytes() == kr2.tobytes() and np.array_equal(dl, dl2)
    bf, b = 386.1448, 386.1448 / 718.856
    n_c, ur_c, dp_c = orc.stereo_match(cl, cr, kl, dl, kr, dr, bf, b)
    n_g, ur_g, dp_g = gpu.Matcher().stereo_match(gl, gr, kl, dl, kr, dr, bf, b)
    assert n_c == n_g
    assert np.array_equal(ur_c, ur_g)
    assert np.array_equal(dp_c, dp_g)
    assert n_c > 100
    # the pair as ONE two-frame call of one extractor, matched as frames 0 and 1 of it: same keypoints, same matches
    g2 = gpu.Extractor(nfeatures=nf)
    k2, d2, c2, s2 = g2.extract_batch_host(np.ascontiguousarray(np.stack([left, right])))
    assert (s2 == 0).all() and c2[0] == len(kl) and c2[1] == len(kr)
    assert k2[0, :c2[0]].tobytes() == kl.tobytes() and k2[1, :c2[1]].tobytes() == kr.tobytes()
    assert np.array_equal(d2[0, :c2[0]], dl) and np.array_equal(d2[1, :c2[1]], dr)
    n_2, ur_2, dp_2 = gpu.Matcher().stereo_match_frames(g2, 0, g2, 1, kl, dl, kr, dr, bf, b)
    assert n_2 == n_c and np.array_equal(ur_2, ur_c) and np.array_equal(dp_2, dp_c)


def test_knn2_small_exact_and_ties(gpu):
    q, t = synth.make_knn_sets(600, 7000, seed=1)
    t[100] = t[5000] = q[3]  # exact duplicates: first index must win, second distance equals the first
    t[200] = q[4]
    m = gpu.Matcher()
    for ratio in (0.6, 0.9):
        ig, d1g, d2g, ng = m.knn2(q, t, ratio)
        ic, d1c, d2c, _ = orc.knn2(q, t, ratio)
        assert np.array_equal(d1g, d1c) and np.array_equal(d2g, d2c) and np.array_equal(ig, ic)
        assert ng == int((ic >= 0).sum())
    assert d1g[3] == 0 and d2g[3] == 0 and ig[3] == -1  # ratio test rejects a tie at distance 0
    assert ig[4] == 200


def test_knn2_full_size_properties_and_sample(gpu):
    q, t = synth.make_knn_sets(4000, 100000, seed=2)
    m = gpu.Matcher()
    ig, d1g, d2g, ng = m.knn2(q, t, 0.7)
    assert (d1g <= d2g).all() and (d1g >= 0).all() and (d2g <= 256).all()
    acc = ig >= 0
    assert ng == int(acc.sum()) and ng > 500
    # accepted matches are real: distance to the reported neighbour equals d1
    sel = np.flatnonzero(acc)[:512]
    assert np.array_equal(np.unpackbits(q[sel] ^ t[ig[sel]], axis=1).sum(axis=1), d1g[sel])
    # exact comparison with the oracle on a 256-query sample against the whole train set
    ic, d1c, d2c, _ = orc.knn2(q[:256], t, 0.7, nthreads=orc.hardware_threads())
    assert np.array_equal(ig[:256], ic) and np.array_equal(d1g[:256], d1c) and np.array_equal(d2g[:256], d2c)


def test_wide_windows_take_the_overflow_fallback(gpu, tum):
    """Windows with more candidates than the per-query list capacity (32 / 64 / 256) must fall back to the
    window-walking kernels and still match the oracle exactly."""
    mp, uright = synth.make_map_points(tum["kps"], tum["desc"], tum["scale"], seed=77, n_map=600, n_true=300)
    m, fg, fc = _frames(gpu, tum, uright=uright)
    state = np.full(len(tum["kps"]), -1, np.int32)
    n_c, km_c = orc.match_projection(fc, mp, 40.0, 0.8, state)
    n_g, km_g = m.match_projection(fg, mp, 40.0, 0.8, state)
    assert (n_c, km_c.tolist()) == (n_g, km_g.tolist()) and n_c > 20
    last, Tc, Tl = synth.make_last_frame(tum["kps"], tum["desc"], seed=78)
    n_c, km_c = orc.match_lastframe(fc, last, Tc, Tl, 120.0, False, True, state)
    n_g, km_g = m.match_lastframe(fg, last, Tc, Tl, 120.0, False, True, state)
    assert (n_c, km_c.tolist()) == (n_g, km_g.tolist()) and n_c > 20
    ex = orc.Extractor(nfeatures=2000)
    g1 = synth.make_frame(210)
    k1, d1 = ex.extract(g1)
    k2, d2 = ex.extract(synth.shift_image(g1, 8, 4))
    cam_args = tum["cam_args"]
    f1g, f2g = m.frame(k1, d1, gpu.Camera(*cam_args), tum["scale"]), m.frame(k2, d2, gpu.Camera(*cam_args), tum["scale"])
    f1c, f2c = orc.Frame(k1, d1, orc.Camera(*cam_args), tum["scale"]), orc.Frame(k2, d2, orc.Camera(*cam_args), tum["scale"])
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    n_c, m_c, p_c = orc.match_init(f1c, f2c, prev, 400, 0.9, True)
    n_g, m_g, p_g = m.match_init(f1g, f2g, prev, 400, 0.9, True)
    assert n_c == n_g and np.array_equal(m_c, m_g) and np.array_equal(p_c, p_g)


@pytest.mark.parametrize("n_map,th", [(5000, 1.0), (5000, 3.0), (9000, 1.0)])
def test_dense_frames_and_large_maps(gpu, tum, n_map, th):
    """The latency form of the resolve kernel keeps its claim tables in shared memory up to 4096 keypoints and caches up to ~7000
    queries; a denser frame (tables in global memory, long candidate lists) and a larger map (general kernel) must equal the
    oracle as well."""
    ex = orc.Extractor()
    parts = [ex.extract(synth.make_frame(300 + s)) for s in range(6)]
    kps, desc = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    assert len(kps) > 4200
    mp, uright = synth.make_map_points(kps, desc, tum["scale"], seed=n_map, n_map=n_map, n_true=2500)
    m, fg, fc = _frames(gpu, tum, kps=kps, desc=desc, uright=uright)
    rng = np.random.default_rng(n_map)
    state = rng.choice([-1, -1, -1, -2], size=len(kps)).astype(np.int32)
    n_c, km_c = orc.match_projection(fc, mp, th, 0.8, state)
    n_g, km_g = m.match_projection(fg, mp, th, 0.8, state)
    assert n_c == n_g and np.array_equal(km_c, km_g) and n_c > 300
    last, Tc, Tl = synth.make_last_frame(kps, desc, seed=n_map + 1)
    n_c, km_c = orc.match_lastframe(fc, last, Tc, Tl, 7.0 * th, False, True, state)
    n_g, km_g = m.match_lastframe(fg, last, Tc, Tl, 7.0 * th, False, True, state)
    assert n_c == n_g and np.array_equal(km_c, km_g) and n_c > 300
