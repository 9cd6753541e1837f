"""Pins oracle B (C++ restatement) against the real OpenCV 4.13.0 primitives and against oracle A
(Python transcription driving cv2). CPU only. Skipped if cv2 is not importable."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

import orc  # noqa: E402
from coeb_b200 import synth  # noqa: E402
from oracle_cv2 import ExtractorA  # noqa: E402


def test_cv2_version_is_the_pinned_one():
    assert cv2.__version__.startswith("4."), cv2.__version__


@pytest.mark.parametrize("seed", [0, 1])
def test_resize_chain_bit_exact(seed):
    rng = np.random.default_rng(seed)
    for (w, h) in [(640, 480), (1241, 376), (321, 203)]:
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8) if seed == 0 else synth.make_frame(seed, w, h)
        ex = orc.Extractor()
        inv = ex.tables()["inv_scale"]
        prev = img
        for l in range(1, 8):
            dw = int(np.rint(np.float32(w) * inv[l]))
            dh = int(np.rint(np.float32(h) * inv[l]))
            ref = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
            got = orc.resize_linear(prev, dw, dh)
            assert np.array_equal(ref, got), (w, h, l, np.abs(ref.astype(int) - got).max())
            prev = ref


def test_resize_upscale_and_odd_ratios():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(37, 53), dtype=np.uint8)
    for (dw, dh) in [(53, 37), (80, 50), (17, 9), (52, 36), (106, 74), (30, 37)]:
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        if (dw, dh) == (img.shape[1] // 2, img.shape[0] // 2):
            continue  # OpenCV switches exact 2x decimation to INTER_AREA
        got = orc.resize_linear(img, dw, dh)
        assert np.array_equal(ref, got), (dw, dh)


@pytest.mark.parametrize("shape", [(480, 640), (333, 444), (134, 179), (9, 11), (7, 300)])
def test_gaussian_bit_exact(shape):
    rng = np.random.default_rng(shape[0])
    for img in (rng.integers(0, 256, size=shape, dtype=np.uint8),
                synth.make_frame(3, max(shape[1], 64), max(shape[0], 64))[:shape[0], :shape[1]].copy()):
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        got = orc.gaussian7(img)
        assert np.array_equal(ref, got), np.abs(ref.astype(int) - got).max()


def test_fast_atan2_bit_exact():
    rng = np.random.default_rng(11)
    ys = rng.integers(-40000, 40000, size=20000)
    xs = rng.integers(-40000, 40000, size=20000)
    pairs = list(zip(ys, xs)) + [(0, 0), (0, -5), (-3, 0), (5, 0), (1, 1), (-1, -1), (7, -7)]
    for y, x in pairs:
        ref = np.float32(cv2.fastAtan2(float(y), float(x)))
        got = np.float32(orc.fast_atan2(float(y), float(x)))
        assert ref == got, (y, x, ref, got)


def test_fast_roi_matches_cv2_including_order_and_response():
    frame = synth.make_frame(2)
    rng = np.random.default_rng(2)
    checked = 0
    for _ in range(150):
        w, h = int(rng.integers(7, 45)), int(rng.integers(7, 45))
        x0, y0 = int(rng.integers(0, 640 - w)), int(rng.integers(0, 480 - h))
        roi = frame[y0:y0 + h, x0:x0 + w]
        for th in (20, 7, 30, 10):
            ref = cv2.FastFeatureDetector_create(th, True).detect(roi)
            got = orc.fast_roi(roi, th)
            assert len(ref) == len(got), (w, h, th)
            for kp, g in zip(ref, got):
                assert (int(kp.pt[0]), int(kp.pt[1]), int(kp.response)) == tuple(int(v) for v in g)
                assert kp.size == 7.0 and kp.angle == -1.0 and kp.octave == 0 and kp.class_id == -1
            checked += len(ref)
    assert checked > 500


def test_fast_random_noise_rois():
    rng = np.random.default_rng(9)
    for _ in range(40):
        roi = rng.integers(0, 256, size=(int(rng.integers(8, 40)), int(rng.integers(8, 40))), dtype=np.uint8)
        for th in (7, 20):
            ref = cv2.FastFeatureDetector_create(th, True).detect(roi)
            got = orc.fast_roi(roi, th)
            assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in ref] == [tuple(int(v) for v in g) for g in got]


@pytest.mark.parametrize("seed,dynamic", [(0, False), (3, True), (4, True)])
def test_full_extraction_oracle_b_equals_oracle_a(seed, dynamic, orb_pattern):
    gray = synth.make_frame(seed)
    if dynamic:
        boxes, tm, blur = synth.make_dynamic(seed, force_area=(seed % 8 == 3))
        if seed == 4:
            boxes, tm, blur = synth.make_dynamic(seed, nbox=2)
            blur[:] = 1
    else:
        boxes, tm, blur = np.zeros((0, 4), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.int32)
    exb = orc.Extractor()
    kb, db = exb.extract(gray, boxes, tm, blur)
    exa = ExtractorA()
    ka, da, st, dyn = exa.extract(gray, boxes, tm, blur, pattern=orb_pattern, want_stages=True)
    info = exb.dyn_info()
    assert info["area_flag"] == dyn["area_flag"]
    assert [tuple(r) for r in info["rects"]] == dyn["rects"]
    for l in range(8):
        assert np.array_equal(st["pyramid"][l], exb.level_image(l)), l
        ca = np.array(st["candidates"][l], np.float32).reshape(-1, 3)
        assert np.array_equal(ca, exb.level_candidates(l)), l
        if st["blurred"][l] is not None:
            assert np.array_equal(st["blurred"][l], exb.level_image(l, blurred=True)), l
    assert len(ka) == len(kb)
    for a, b in zip(ka, kb):
        assert (a[0], a[1], a[2], a[3], a[4], a[5], a[6]) == (b["x"], b["y"], b["size"], b["angle"], b["response"],
                                                              b["octave"], b["class_id"])
    assert np.array_equal(da, db)
    if dynamic:
        assert info["n_dynamic"] >= 1


@pytest.mark.parametrize("code,ch,bgr", [(cv2.COLOR_RGB2GRAY, 3, False), (cv2.COLOR_BGR2GRAY, 3, True), (cv2.COLOR_RGBA2GRAY, 4, False),
                                         (cv2.COLOR_BGRA2GRAY, 4, True)])
def test_rgb_to_gray_bit_exact(code, ch, bgr):
    rng = np.random.default_rng(ch)
    img = rng.integers(0, 256, size=(123, 217, ch), dtype=np.uint8)
    img[:4, :4] = [[0] * ch, [255] * ch, [1] * ch, [254] * ch]
    assert np.array_equal(orc.rgb_to_gray(img, bgr), cv2.cvtColor(img, code))


def test_laplacian_blur_flag_model():
    gray = synth.make_frame(12)
    gray[50:200, 300:520] = 90   # flat box -> mean 0 -> blurred
    boxes = np.array([[300, 50, 520, 200], [20, 30, 200, 400], [5, 5, 6, 6], [0, 0, 640, 480], [100.7, 80.2, 300.9, 333.3]], np.float32)
    flags, means = orc.blur_flags(gray, boxes)
    for b, f, m in zip(boxes, flags, means):
        x0, y0, bw, bh = int(b[0]), int(b[1]), int(b[2] - b[0]), int(b[3] - b[1])
        roi = gray[y0:y0 + bh, x0:x0 + bw].copy()
        ref = cv2.mean(np.abs(cv2.Laplacian(roi, cv2.CV_16U)))[0]
        assert m == ref and f == int(ref < 4.2)
    assert flags[0] == 1 and flags[1] == 0
