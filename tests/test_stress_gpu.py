"""GPU: randomised differential runs of every matcher entry point against the oracle over many seeds (ties, contention,
list overflow and empty windows turn up by chance here rather than by construction). Bit-exact like the targeted tests."""
import numpy as np
import pytest

import orc
from coeb_b200 import synth

pytestmark = pytest.mark.gpu
CAM = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)


@pytest.fixture(scope="module")
def gpu():
    import coeb_b200
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible")
    return coeb_b200


@pytest.fixture(scope="module")
def frames():
    ex = orc.Extractor()
    out = []
    for seed in (400, 401, 402):
        g = synth.make_frame(seed)
        k, d = ex.extract(g)
        k2, d2 = ex.extract(synth.shift_image(g, 9 - 6 * (seed % 3), 4 - 3 * (seed % 2)))
        out.append(dict(k=k, d=d, k2=k2, d2=d2))
    return out, ex.tables()["scale"]


def test_randomised_matchers_against_oracle(gpu, frames):
    fr, scale = frames
    m = gpu.Matcher()
    checked = 0
    for s in range(18):
        f = fr[s % 3]
        rng = np.random.default_rng(1000 + s)
        kps, desc = f["k"], f["d"]
        uright = np.where(rng.random(len(kps)) < 0.3, kps["x"] - np.float32(40.0) / rng.uniform(0.5, 5.0, len(kps)).astype(np.float32),
                          np.float32(-1)).astype(np.float32) if s % 2 else None
        fg, fc = m.frame(kps, desc, gpu.Camera(*CAM), scale, uright), orc.Frame(kps, desc, orc.Camera(*CAM), scale, uright)
        state = rng.choice([-1, -1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
        th = float(rng.choice([1.0, 3.0, 5.0, 12.0]))
        ratio = float(rng.choice([0.6, 0.8, 0.9]))
        # M2, flat
        nt = int(rng.integers(100, 800))
        mp, _ = synth.make_map_points(kps, desc, scale, seed=s, n_map=nt + int(rng.integers(100, 4000)), n_true=nt)
        a, b = m.match_projection(fg, mp, th, ratio, state), orc.match_projection(fc, mp, th, ratio, state)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), ("M2", s)
        # SearchLocalPoints
        Tcw, Ow = synth.make_pose(s)
        nt = min(int(rng.integers(100, 800)), len(kps))
        lm, skip, obs = synth.make_local_map(kps, desc, scale, Tcw, seed=s, n_map=nt + int(rng.integers(100, 3500)), n_true=nt)
        a = m.search_local_points(fg, m.local_map(lm), skip, obs, Tcw, Ow, th, ratio, state)
        b = orc.search_local_points(fc, lm, skip, obs, Tcw, Ow, th, ratio, state)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3].tobytes() == b[3].tobytes(), ("local", s)
        # M3
        last, Tc, Tl = synth.make_last_frame(kps, desc, seed=s)
        mono, ori, th3 = bool(s % 3 == 0), bool(s % 4), float(rng.choice([7.0, 15.0, 30.0]))
        a, b = m.match_lastframe(fg, last, Tc, Tl, th3, mono, ori, state), orc.match_lastframe(fc, last, Tc, Tl, th3, mono, ori, state)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), ("M3", s)
        # relocalisation
        kf = dict(valid=(1 - skip).astype(np.uint8), xyz=lm["xyz"], min_dist=lm["min_dist"], max_dist=lm["max_dist"],
                  angle=rng.uniform(0, 360, len(skip)).astype(np.float32), desc=lm["desc"])
        od = int(rng.choice([64, 100]))
        a, b = m.match_reloc(fg, kf, Tcw, Ow, th3, od, ori, state), orc.match_reloc(fc, kf, Tcw, Ow, th3, od, ori, state)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), ("reloc", s)
        # M4 + BoW + triangulation on the shifted pair
        k2, d2 = f["k2"], f["d2"]
        f2g, f2c = m.frame(k2, d2, gpu.Camera(*CAM), scale, None), orc.Frame(k2, d2, orc.Camera(*CAM), scale, None)
        prev = np.stack([kps["x"], kps["y"]], axis=1).astype(np.float32) + rng.normal(0, 2, (len(kps), 2)).astype(np.float32)
        win = int(rng.choice([20, 50, 100]))
        a, b = m.match_init(fg, f2g, prev, win, 0.9, ori), orc.match_init(fc, f2c, prev, win, 0.9, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2].tobytes() == b[2].tobytes(), ("M4", s)
        nn = int(rng.choice([1, 8, 60, 150]))
        fv1, fv2 = synth.make_feature_vector(desc, nn, seed=s), synth.make_feature_vector(d2, nn, seed=s)
        v1 = (rng.random(len(kps)) < 0.8).astype(np.uint8)
        v2 = (rng.random(len(k2)) < 0.8).astype(np.uint8)
        strict = bool(s % 2)
        a = m.match_bow(fg, f2g, v1, v2 if strict else None, fv1, fv2, ratio, ori, strict)
        b = orc.match_bow(fc, f2c, v1, v2 if strict else None, fv1, fv2, ratio, ori, strict)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), ("bow", s)
        F12 = rng.normal(0, 1, (3, 3)).astype(np.float32) * np.float32(1e-2)
        ep = (float(rng.uniform(0, 640)), float(rng.uniform(0, 480)))
        a = m.match_triangulation(fg, f2g, v1, v2, fv1, fv2, F12, ep, False, ori)
        b = orc.match_triangulation(fc, f2c, v1, v2, fv1, fv2, F12, ep, False, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), ("tri", s)
        checked += 1
    assert checked == 18


def test_randomised_extraction_parameters_against_oracle(gpu):
    """Odd sizes and parameter mixes: every keypoint and descriptor byte against the oracle."""
    rng = np.random.default_rng(77)
    for t in range(6):
        w, h = int(rng.integers(200, 900)), int(rng.integers(160, 600))
        h = min(h, w)   # landscape: the octree starts from round(width / height) roots (src/ORBextractor.cc:494), 0 for tall images
        nf = int(rng.integers(200, 2500))
        nl = int(rng.integers(2, 9))
        sf = float(rng.choice([1.15, 1.2, 1.25, 1.4]))
        if min(w, h) / sf ** (nl - 1) < 100:   # the reference needs at least one 30 px cell row and column on its coarsest level
            nl = 3
        gray = synth.make_frame(600 + t, w, h)
        boxes, tm, blur = synth.make_dynamic(600 + t, w, h, force_area=(t == 2))
        g, c = gpu.Extractor(nf, sf, nl, 20, 7), orc.Extractor(nf, sf, nl, 20, 7)
        kg, dg = g.extract(gray, boxes, tm, blur, cap=nf + 600)
        kc, dc = c.extract(gray, boxes, tm, blur)
        assert kg.tobytes() == kc.tobytes() and np.array_equal(dg, dc), (w, h, nf, nl, sf)
        g.close()
