"""GPU parity of ORBmatcher::SearchByBoW (both overloads) against the CPU oracle: match indices and counts bit-exact. The
DBoW2 FeatureVectors are inputs (CSR); tests build them with a stand-in vocabulary (nearest of 100 random centres)."""
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
def pair():
    """A keyframe and a later frame of the same scene: the second image is the first shifted by a few pixels."""
    ex = orc.Extractor()
    g1 = synth.make_frame(300)
    g2 = synth.shift_image(g1, 6, -4)
    k1, d1 = ex.extract(g1)
    k2, d2 = ex.extract(g2)
    return dict(k1=k1, d1=d1, k2=k2, d2=d2, scale=ex.tables()["scale"])


def _both(gpu, p):
    m = gpu.Matcher()
    fg1, fg2 = m.frame(p["k1"], p["d1"], gpu.Camera(*CAM), p["scale"]), m.frame(p["k2"], p["d2"], gpu.Camera(*CAM), p["scale"])
    fc1, fc2 = orc.Frame(p["k1"], p["d1"], orc.Camera(*CAM), p["scale"]), orc.Frame(p["k2"], p["d2"], orc.Camera(*CAM), p["scale"])
    return m, fg1, fg2, fc1, fc2


@pytest.mark.parametrize("ratio,ori,strict,use_valid2,n_nodes", [(0.7, True, False, False, 100), (0.9, True, False, False, 100), (0.75, True, True, True, 100),
                                                                 (0.9, False, False, False, 100), (0.9, True, False, False, 7), (0.8, True, True, True, 1)])
def test_search_by_bow_matches_oracle(gpu, pair, ratio, ori, strict, use_valid2, n_nodes):
    m, fg1, fg2, fc1, fc2 = _both(gpu, pair)
    rng = np.random.default_rng(11)
    valid1 = (rng.random(fc1.n) < 0.85).astype(np.uint8)
    valid2 = (rng.random(fc2.n) < 0.9).astype(np.uint8) if use_valid2 else None
    fv1 = synth.make_feature_vector(pair["d1"], n_nodes, seed=1)
    fv2 = synth.make_feature_vector(pair["d2"], n_nodes, seed=1)
    ng, mg = m.match_bow(fg1, fg2, valid1, valid2, fv1, fv2, ratio, ori, strict)
    nc, mc = orc.match_bow(fc1, fc2, valid1, valid2, fv1, fv2, ratio, ori, strict)
    assert ng == nc and np.array_equal(mg, mc)
    assert nc > (20 if n_nodes > 1 else 5)
    got = mg[mg >= 0]
    assert len(np.unique(got)) == len(got), "a feature of the searched frame was claimed twice"
    assert (valid1[mg >= 0] == 1).all()
    if valid2 is not None:
        assert (valid2[got] == 1).all()


def test_search_by_bow_disjoint_and_partial_vocabularies(gpu, pair):
    m, fg1, fg2, fc1, fc2 = _both(gpu, pair)
    valid1 = np.ones(fc1.n, np.uint8)
    n1, s1, i1 = synth.make_feature_vector(pair["d1"], 50, seed=2)
    n2, s2, i2 = synth.make_feature_vector(pair["d2"], 50, seed=2)
    # drop every third node of the second vector (lower_bound skips), then make the id sets disjoint
    keep = np.arange(len(n2)) % 3 != 0
    items2, start2 = [], [0]
    for k in np.nonzero(keep)[0]:
        items2.extend(i2[s2[k]:s2[k + 1]].tolist())
        start2.append(len(items2))
    fv2 = (n2[keep], np.array(start2, np.int32), np.array(items2, np.int32))
    ng, mg = m.match_bow(fg1, fg2, valid1, None, (n1, s1, i1), fv2, 0.9)
    nc, mc = orc.match_bow(fc1, fc2, valid1, None, (n1, s1, i1), fv2, 0.9)
    assert ng == nc and np.array_equal(mg, mc) and nc > 10
    ng, mg = m.match_bow(fg1, fg2, valid1, None, (n1, s1, i1), (fv2[0] + 1, fv2[1], fv2[2]), 0.9)
    assert ng == 0 and (mg == -1).all()
    # malformed vectors are rejected, not read
    with pytest.raises(gpu.CoebError):
        m.match_bow(fg1, fg2, valid1, None, (n1[::-1].copy(), s1, i1), fv2, 0.9)
    dup = i1.copy()
    dup[1] = dup[0]
    with pytest.raises(gpu.CoebError):
        m.match_bow(fg1, fg2, valid1, None, (n1, s1, dup), fv2, 0.9)


def _shift_fundamental(dx, dy):
    """F12 of two views related by a pure image shift (dx, dy): the epipolar line of x1 is the line through x1 along the shift."""
    return np.array([[0, 0, dy], [0, 0, -dx], [-dy, dx, 0]], np.float32)


@pytest.mark.parametrize("only_stereo,ori,epipole,with_uright", [(False, True, (1e7, -1e7), False), (False, True, (320.0, 240.0), True),
                                                                 (True, True, (320.0, 240.0), True), (False, False, (100.0, 50.0), False)])
def test_search_for_triangulation_matches_oracle(gpu, pair, only_stereo, ori, epipole, with_uright):
    rng = np.random.default_rng(21)
    ur1 = ur2 = None
    if with_uright:
        ur1 = np.where(rng.random(len(pair["k1"])) < 0.6, pair["k1"]["x"] - np.float32(8), np.float32(-1)).astype(np.float32)
        ur2 = np.where(rng.random(len(pair["k2"])) < 0.6, pair["k2"]["x"] - np.float32(8), np.float32(-1)).astype(np.float32)
    m = gpu.Matcher()
    fg1 = m.frame(pair["k1"], pair["d1"], gpu.Camera(*CAM), pair["scale"], ur1)
    fg2 = m.frame(pair["k2"], pair["d2"], gpu.Camera(*CAM), pair["scale"], ur2)
    fc1 = orc.Frame(pair["k1"], pair["d1"], orc.Camera(*CAM), pair["scale"], ur1)
    fc2 = orc.Frame(pair["k2"], pair["d2"], orc.Camera(*CAM), pair["scale"], ur2)
    free1 = (rng.random(fc1.n) < 0.7).astype(np.uint8)
    free2 = (rng.random(fc2.n) < 0.7).astype(np.uint8)
    fv1 = synth.make_feature_vector(pair["d1"], 40, seed=5)
    fv2 = synth.make_feature_vector(pair["d2"], 40, seed=5)
    F12 = _shift_fundamental(6.0, -4.0)
    ng, mg = m.match_triangulation(fg1, fg2, free1, free2, fv1, fv2, F12, epipole, only_stereo, ori)
    nc, mc = orc.match_triangulation(fc1, fc2, free1, free2, fv1, fv2, F12, epipole, only_stereo, ori)
    assert ng == nc and np.array_equal(mg, mc)
    assert nc > 20
    ok = mg >= 0
    assert free1[ok].all() and free2[mg[ok]].all()
    if only_stereo:
        assert (ur1[ok] >= 0).all() and (ur2[mg[ok]] >= 0).all()
    # duplicated descriptors on the searched side: equal distances, the LAST candidate in list order must win on both sides
    k2d = np.concatenate([pair["k2"], pair["k2"]])
    d2d = np.concatenate([pair["d2"], pair["d2"]])
    fg2d = m.frame(k2d, d2d, gpu.Camera(*CAM), pair["scale"], None if ur2 is None else np.concatenate([ur2, ur2]))
    fc2d = orc.Frame(k2d, d2d, orc.Camera(*CAM), pair["scale"], None if ur2 is None else np.concatenate([ur2, ur2]))
    fv2d = synth.make_feature_vector(d2d, 40, seed=5)
    free2d = np.concatenate([free2, free2])
    ng, mg = m.match_triangulation(fg1, fg2d, free1, free2d, fv1, fv2d, F12, epipole, only_stereo, ori)
    nc, mc = orc.match_triangulation(fc1, fc2d, free1, free2d, fv1, fv2d, F12, epipole, only_stereo, ori)
    assert ng == nc and np.array_equal(mg, mc)
    assert (mc[mc >= 0] >= len(pair["k2"])).all(), "ties must resolve to the later duplicate"
