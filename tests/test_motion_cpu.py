"""CPU: the ProcessMovingObject oracle (oracle/pmo.py = OpenCV 4.13 called with the reference's arguments) against the committed
golden intermediates, the reference-written parts of the function (border / SAD tests, epipolar test) against independent
numpy statements, and the library's host-side fundamental-matrix estimator (no device needed)."""
import os
import sys
import zlib

import numpy as np
import pytest

from coeb_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pmo  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("seed", [0, 3])
def test_synthetic_pairs_are_reproducible(seed):
    g = np.load(os.path.join(G, "motion_seed%d.npz" % seed))
    prev, cur, boxes = synth.make_motion_pair(seed)
    assert zlib.crc32(prev.tobytes() + cur.tobytes()) == int(g["crc"]) and np.array_equal(boxes, g["boxes"])


@pytest.mark.parametrize("seed", [0, 3])
def test_oracle_reproduces_the_golden_intermediates(seed):
    """Guards against a different OpenCV build behind `cv2`: corners, sub-pixel corners, tracks and states must be the stored ones
    (same version: bit for bit; the fundamental matrix comes from OpenCV's seeded RANSAC and is reproducible as well)."""
    cv2 = pytest.importorskip("cv2")
    g = np.load(os.path.join(G, "motion_seed%d.npz" % seed))
    if cv2.__version__ != str(g["cv2_version"]):
        pytest.skip("goldens were made with OpenCV %s" % g["cv2_version"])
    prev, cur, _ = synth.make_motion_pair(seed)
    assert np.array_equal(pmo.good_features(prev), g["corners"])
    r = pmo.process_moving_object(prev, cur)
    assert np.allclose(r["prepoint"], g["prepoint"], atol=1e-3) and np.allclose(r["nextpoint"], g["nextpoint"], atol=2e-2)
    assert (r["state"] == g["state"]).mean() > 0.995
    assert abs(len(r["tm"]) - len(g["tm"])) <= 3


@pytest.mark.parametrize("seed", [0, 3])
def test_reference_written_parts_against_numpy(seed):
    """src/Frame.cc:336-364 and :372-385 restated independently on the golden tracks."""
    g = np.load(os.path.join(G, "motion_seed%d.npz" % seed))
    prev, cur, _ = synth.make_motion_pair(seed)
    pre, nxt, st = g["prepoint"], g["nextpoint"], g["state"]
    # epipolar distances and membership
    F = g["F"]
    h1 = np.concatenate([pre.astype(np.float64), np.ones((len(pre), 1))], axis=1)
    l = h1 @ F.T
    d = np.abs(l[:, 0] * nxt[:, 0] + l[:, 1] * nxt[:, 1] + l[:, 2]) / np.hypot(l[:, 0], l[:, 1])
    idx, dist = pmo.epipolar_outliers(F, pre, nxt, st)
    assert np.allclose(dist[st != 0], d[st != 0], rtol=1e-12) and np.array_equal(idx, g["tm_index"])
    assert np.array_equal(idx, np.nonzero((st != 0) & ~(d <= 1))[0])
    # border / SAD tests: the stored state already passed them, so re-applying them changes nothing; a track pushed to the image
    # edge or onto a very different patch is dropped
    assert np.array_equal(pmo.border_and_sad_check(prev, cur, pre, nxt, st), st)
    bad = nxt.copy()
    bad[:5, 0] = 2.0
    st2 = pmo.border_and_sad_check(prev, cur, pre, bad, np.ones(len(st), np.uint8))
    assert not st2[:5].any()
    white = np.full_like(cur, 255)
    black = np.zeros_like(prev)
    st3 = pmo.border_and_sad_check(black, white, pre, nxt, np.ones(len(st), np.uint8))
    inner = (pre[:, 0] >= 5) & (pre[:, 0] < 635) & (pre[:, 1] >= 5) & (pre[:, 1] < 475) & (nxt[:, 0] >= 5) & (nxt[:, 0] < 635) & (nxt[:, 1] >= 5) & (nxt[:, 1] < 475)
    assert not st3[inner].any()   # 9 * 255 = 2295 > 2120


def test_fundamental_ransac_recovers_a_known_geometry():
    """Host code of the product library (coeb_fundamental_ransac): exact correspondences of a known two-view geometry plus 30 %
    gross outliers; no device is needed for this entry point."""
    from coeb_b200 import motion
    rng = np.random.default_rng(4)
    X = np.concatenate([rng.uniform(-2, 2, (400, 2)), rng.uniform(4, 9, (400, 1))], axis=1)
    K = np.array([[520.0, 0, 320], [0, 520, 240], [0, 0, 1]])
    ang = 0.03
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.2, -0.05, 0.1])
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:]
    x2 = (K @ (R @ X.T + t[:, None])).T
    x2 = x2[:, :2] / x2[:, 2:]
    bad = rng.random(400) < 0.3
    x2[bad] += rng.uniform(-40, 40, (int(bad.sum()), 2))
    F, mask = motion.fundamental_ransac(x1.astype(np.float32), x2.astype(np.float32), 0.1, 0.99, 2000, 7)
    assert mask[~bad].mean() > 0.95 and mask[bad].mean() < 0.05
    h1 = np.concatenate([x1, np.ones((400, 1))], axis=1)
    h2 = np.concatenate([x2, np.ones((400, 1))], axis=1)
    l2 = (F @ h1.T).T
    d = np.abs((h2 * l2).sum(axis=1)) / np.hypot(l2[:, 0], l2[:, 1])
    assert d[~bad].max() < 0.05 and abs(np.linalg.det(F)) < 1e-9 * np.abs(F).max() ** 3 + 1e-12
    # the same estimate on the golden tracks agrees with OpenCV's on which points are off their epipolar lines
    g = np.load(os.path.join(G, "motion_seed0.npz"))
    keep = g["state"] != 0
    F2, _ = motion.fundamental_ransac(g["prepoint"][keep], g["nextpoint"][keep], 0.1, 0.99, 1000, 12345)
    _, d_own = pmo.epipolar_outliers(F2, g["prepoint"], g["nextpoint"], g["state"])
    far = keep & (g["dist"] > 2.5)
    assert far.sum() > 20 and (d_own[far] > 1).mean() > 0.9 and (d_own[keep & (g["dist"] < 0.3)] <= 1).mean() > 0.95


def test_fundamental_ransac_degenerate_and_translation_inputs():
    """Host code of the product library: fewer than eight points and identical points fail loudly (no model, no crash); a pure
    sideways translation (epipoles at infinity) is recovered like any other geometry."""
    from coeb_b200 import motion
    import coeb_b200
    rng = np.random.default_rng(11)
    with pytest.raises(coeb_b200.CoebError):
        motion.fundamental_ransac(rng.uniform(0, 600, (7, 2)).astype(np.float32), rng.uniform(0, 600, (7, 2)).astype(np.float32))
    same = np.tile(np.array([[100.0, 50.0]], np.float32), (40, 1))
    with pytest.raises(coeb_b200.CoebError):
        motion.fundamental_ransac(same, same)
    X = np.concatenate([rng.uniform(-3, 3, (300, 2)), rng.uniform(4, 12, (300, 1))], axis=1)
    K = np.array([[520.0, 0, 320], [0, 520, 240], [0, 0, 1]])
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:]
    x2 = (K @ (X + np.array([0.3, 0.0, 0.0])).T).T
    x2 = x2[:, :2] / x2[:, 2:]
    bad = rng.random(300) < 0.25
    x2[bad] += rng.uniform(-30, 30, (int(bad.sum()), 2))
    F, mask = motion.fundamental_ransac(x1.astype(np.float32), x2.astype(np.float32), 0.1, 0.99, 2000, 3)
    assert mask[~bad].mean() > 0.95 and mask[bad].mean() < 0.1
    h1 = np.concatenate([x1, np.ones((300, 1))], axis=1)
    h2 = np.concatenate([x2, np.ones((300, 1))], axis=1)
    l2 = (F @ h1.T).T
    d = np.abs((h2 * l2).sum(axis=1)) / np.hypot(l2[:, 0], l2[:, 1])
    assert d[~bad].max() < 0.05


def test_fundamental_ransac_vector_clones_agree():
    """The hypothesis scoring of coeb_fundamental_ransac has one clone per vector width (picked once per process from the CPU's flags;
    COEB_MOTION_NO_AVX512 keeps the 256-bit one where the 512-bit one would run): the same expression element by element, so the
    consensus set and F must be the same to the last bit; so must F with the final fit's normal matrix accumulated as whole rows
    (COEB_MOTION_SCALAR_FIT keeps the scalar loop). One process per variant, on the golden tracks."""
    import subprocess
    code = ("import sys, zlib, numpy as np; sys.path.insert(0, %r); from coeb_b200 import motion; g = np.load(%r); keep = g['state'] != 0; "
            "F, m = motion.fundamental_ransac(g['prepoint'][keep], g['nextpoint'][keep], 0.1, 0.99, 1000, 12345); "
            "print(zlib.crc32(F.tobytes()), zlib.crc32(m.tobytes()), int(m.sum()))") % (os.path.join(ROOT, "coeb-slam_b200", "python"), os.path.join(G, "motion_seed0.npz"))
    outs = []
    for extra in ({}, {"COEB_MOTION_NO_AVX512": "1"}, {"COEB_MOTION_SCALAR_FIT": "1"}):
        env = {k: v for k, v in os.environ.items() if k not in ("COEB_MOTION_NO_AVX512", "COEB_MOTION_SCALAR_FIT")}
        env.update(extra)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split())
    # (the third process accumulates the final fit's normal matrix with the scalar upper-triangle loop instead of whole 4-wide rows)
    assert outs[0] == outs[1] == outs[2] and int(outs[0][2]) > 100
