"""GPU: the CUDA path against the reference ITSELF (oracle/_ref: the reference's sources compiled unchanged against the test-only
OpenCV shim, prebuilt in the dev container and shipped with the snapshot), without the oracle restatement in between.
Extraction against the monotonic-heap build (the reference's octree tie order follows heap addresses);
matchers against the plain build."""
import os
import sys

import numpy as np
import pytest

import orc
from coeb_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import ref  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available("mono"), reason="oracle/_ref was not shipped with this snapshot")]
CAM_ARGS = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)


@pytest.fixture(scope="module")
def gpu():
    import coeb_b200
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible")
    return coeb_b200


def test_extraction_equals_the_reference_on_32_frames(gpu):
    g, r = gpu.Extractor(), ref.Extractor(variant="mono")
    n_area = 0
    for seed in range(32):
        gray = synth.make_frame(seed)
        boxes, tm, blur = synth.make_dynamic(seed, force_area=(seed % 8 == 3))
        kg, dg = g.extract(gray, boxes, tm, blur)
        kr, dr = r.extract(gray, boxes, tm, blur)
        assert len(kg) == len(kr) and kg.tobytes() == kr.tobytes() and np.array_equal(dg, dr), seed
        n_area += bool(g.dyn_info()["area_flag"])
        for lvl in (0, 1, 4, 7):
            assert np.array_equal(g.level_image(lvl), r.level_image(lvl)), (seed, lvl)
    assert n_area >= 4


def test_batched_extraction_equals_the_reference(gpu):
    batch = synth.make_batch(48, base_seed=7000, unique=16)
    g = gpu.Extractor()
    kps, desc, counts, status = g.extract_batch_host(batch["gray"], batch["boxes"], batch["nbox"], batch["tm"], batch["ntm"], batch["blur"])
    assert (status == 0).all()
    r = ref.Extractor(variant="mono")
    for i in range(48):
        nb, nt = int(batch["nbox"][i]), int(batch["ntm"][i])
        kr, dr = r.extract(batch["gray"][i], batch["boxes"][i, :nb], batch["tm"][i, :nt], batch["blur"][i, :nb])
        assert counts[i] == len(kr) and kps[i, :counts[i]].tobytes() == kr.tobytes() and np.array_equal(desc[i, :counts[i]], dr), i


def test_matchers_equal_the_reference(gpu):
    ex = orc.Extractor()
    kps, desc = ex.extract(synth.make_frame(100))
    scale = ex.tables()["scale"]
    m = gpu.Matcher()
    rng = np.random.default_rng(5)
    # SearchByProjection against a 5000-point map
    mp, uright = synth.make_map_points(kps, desc, scale, seed=30)
    fg, fr = m.frame(kps, desc, gpu.Camera(*CAM_ARGS), scale, uright), ref.Frame(kps, desc, orc.Camera(*CAM_ARGS), scale, uright, variant="nofma")
    state = rng.choice([-1, -1, -1, -1, -2, -3], size=len(kps)).astype(np.int32)
    a, b = m.match_projection(fg, mp, 3.0, 0.8, state), ref.match_projection(fr, mp, 3.0, 0.8, state)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 100
    # frame to frame
    last, Tc, Tl = synth.make_last_frame(kps, desc, seed=15)
    state = np.full(len(kps), -1, np.int32)
    a, b = m.match_lastframe(fg, last, Tc, Tl, 15.0, False, True, state), ref.match_lastframe(fr, last, Tc, Tl, 15.0, False, True, state)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 50
    # window queries
    for _ in range(100):
        x, y, rr = float(rng.uniform(-30, 670)), float(rng.uniform(-30, 510)), float(rng.choice([7.5, 36.0, 100.0]))
        assert np.array_equal(fg.features_in_area(x, y, rr, 0, 4), fr.features_in_area(x, y, rr, 0, 4))
    # initialisation
    ex2 = orc.Extractor(nfeatures=2000)
    g1 = synth.make_frame(200)
    k1, d1 = ex2.extract(g1)
    k2, d2 = ex2.extract(synth.shift_image(g1, 20, 10))
    sc = ex2.tables()["scale"]
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    a = m.match_init(m.frame(k1, d1, gpu.Camera(*CAM_ARGS), sc), m.frame(k2, d2, gpu.Camera(*CAM_ARGS), sc), prev, 100, 0.9, True)
    b = ref.match_init(ref.Frame(k1, d1, orc.Camera(*CAM_ARGS), sc), ref.Frame(k2, d2, orc.Camera(*CAM_ARGS), sc), prev, 100, 0.9, True)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_stereo_equals_the_reference(gpu):
    w, h, nf = 1241, 376, 2000
    left = synth.make_frame(300, w, h)
    right = synth.make_stereo_right(left, seed=300)
    gl, gr = gpu.Extractor(nfeatures=nf), gpu.Extractor(nfeatures=nf)
    rl, rr = ref.Extractor(nfeatures=nf, variant="mono"), ref.Extractor(nfeatures=nf, variant="mono")
    kl, dl = gl.extract(left)
    kr, dr = gr.extract(right)
    kl2, dl2 = rl.extract(left)
    kr2, dr2 = rr.extract(right)
    assert kl.tobytes() == kl2.tobytes() and kr.tobytes() == kr2.tobytes() and np.array_equal(dl, dl2) and np.array_equal(dr, dr2)
    bf, b = 386.1448, 386.1448 / 718.856
    n_g, ur_g, dp_g = gpu.Matcher().stereo_match(gl, gr, kl, dl, kr, dr, bf, b)
    n_r, ur_r, dp_r = ref.stereo_match(rl, rr, kl, dl, kr, dr, bf, b)
    assert n_g == n_r and n_g > 100 and ur_g.tobytes() == ur_r.tobytes() and dp_g.tobytes() == dp_r.tobytes()
