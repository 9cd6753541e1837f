"""GPU parity: the CUDA extractor (through the C ABI) against the CPU oracle, stage by stage and on the
final keypoints/descriptors. Bit-exact for everything integer; angles must be bit-identical floats
(same fp32 polynomial, no FMA), which is stricter than the 1e-3 rad the north star allows."""
import numpy as np
import pytest

import orc
from coeb_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import coeb_b200
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests need the CUDA library to run for real")
    return coeb_b200


def _inputs(seed, w=640, h=480, dynamic=True):
    gray = synth.make_frame(seed, w, h)
    if dynamic:
        boxes, tm, blur = synth.make_dynamic(seed, w, h, force_area=(seed % 8 == 3))
    else:
        boxes, tm, blur = np.zeros((0, 4), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.int32)
    return gray, boxes, tm, blur


def compare_frame(gpu_ex, cpu_ex, gray, boxes, tm, blur, stages=True, frame=0, kps=None, desc=None):
    kb, db = cpu_ex.extract(gray, boxes, tm, blur)
    if kps is None:
        kps, desc = gpu_ex.extract(gray, boxes, tm, blur)
    if stages:
        info_c, info_g = cpu_ex.dyn_info(), gpu_ex.dyn_info(frame)
        assert info_c["area_flag"] == info_g["area_flag"]
        assert np.array_equal(info_c["rects"], info_g["rects"])
        assert info_c["area"] == info_g["area"]
        for l in range(cpu_ex.nlevels):
            assert np.array_equal(cpu_ex.level_image(l), gpu_ex.level_image(l, frame)), "pyramid level %d" % l
            cb = cpu_ex.level_image(l, blurred=True)
            if cb is not None:
                assert np.array_equal(cb, gpu_ex.level_image(l, frame, blurred=True)), "blurred level %d" % l
            cc = cpu_ex.level_candidates(l).astype(np.int64)
            gc = gpu_ex.level_candidates(l, frame)
            cs = cc[np.lexsort((cc[:, 0], cc[:, 1]))] if len(cc) else cc.reshape(0, 3)
            gs = gc[np.lexsort((gc[:, 0], gc[:, 1]))] if len(gc) else gc.reshape(0, 3)
            assert len(cs) == len(gs), "level %d: %d vs %d FAST candidates" % (l, len(cs), len(gs))
            assert np.array_equal(cs, gs), "FAST candidate set, level %d" % l
            ck = cpu_ex.level_keypoints(l)
            gk = gpu_ex.level_keys(l, frame)
            assert len(ck) == len(gk), "level %d: %d vs %d selected keys" % (l, len(ck), len(gk))
            assert np.array_equal(ck["x"], gk[:, 0]) and np.array_equal(ck["y"], gk[:, 1]), \
                "octree selection/order, level %d" % l
            assert np.array_equal(ck["response"], gk[:, 2])
            assert np.array_equal(ck["angle"], gk[:, 3]), "IC angle, level %d (max diff %g deg)" % (
                l, np.abs(ck["angle"] - gk[:, 3]).max())
    assert len(kb) == len(kps)
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(kb[f], kps[f]), f
    bits = np.unpackbits(np.bitwise_xor(db, desc)).sum()
    assert bits == 0, "descriptor bit agreement %.6f" % (1 - bits / max(db.size * 8, 1))
    return len(kb)


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 11])
def test_single_frame_640x480_bit_exact(gpu, seed):
    gray, boxes, tm, blur = _inputs(seed)
    n = compare_frame(gpu.Extractor(), orc.Extractor(), gray, boxes, tm, blur)
    assert n > 300


def test_tables_match_oracle(gpu):
    for nf in (1000, 1500, 2000, 4000):
        g, c = gpu.Extractor(nfeatures=nf).tables(), orc.Extractor(nfeatures=nf).tables()
        for k in ("scale", "inv_scale", "sigma2", "inv_sigma2", "per_level"):
            assert np.array_equal(g[k], c[k]), (nf, k)


def test_no_boxes_equals_classic_operator(gpu):
    gray = synth.make_frame(21)
    compare_frame(gpu.Extractor(), orc.Extractor(), gray, None, None, None)


def test_blur_flag_layer_and_short_blur_array(gpu):
    # layer 2: few T_M points inside a box, blur flag set -> dynamic; blur_flag shorter than the box list
    gray = synth.make_frame(5)
    boxes = np.array([[100, 60, 300, 400], [350, 100, 500, 380]], np.float32)
    tm = np.array([[150.5, 100.2], [400.7, 200.9], [20.0, 20.0]], np.float32)
    g, c = gpu.Extractor(), orc.Extractor()
    compare_frame(g, c, gray, boxes, tm, np.array([1], np.int32))
    assert g.dyn_info()["n_dynamic"] == 1
    compare_frame(g, c, gray, boxes, tm, np.array([1, 1], np.int32))
    assert g.dyn_info()["n_dynamic"] == 2
    compare_frame(g, c, gray, boxes, tm, np.array([0, 0], np.int32))
    assert g.dyn_info()["n_dynamic"] == 0


def test_nfeatures_bump_1500_2000(gpu):
    # COEB's frame-loss logic re-creates the extractor with nFeatures += 500 (src/Tracking.cc:434-465)
    gray, boxes, tm, blur = _inputs(6)
    for nf in (1500, 2000):
        compare_frame(gpu.Extractor(nfeatures=nf), orc.Extractor(nfeatures=nf), gray, boxes, tm, blur)


def test_flat_and_noise_images(gpu):
    g, c = gpu.Extractor(), orc.Extractor()
    flat = np.full((480, 640), 127, np.uint8)
    k, d = g.extract(flat)
    assert len(k) == 0 and len(c.extract(flat)[0]) == 0
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, size=(480, 640), dtype=np.uint8)
    compare_frame(g, c, noise, None, None, None)


def test_clustered_features_deep_octree(gpu):
    # all texture in one corner: deep quadtree, many single-key nodes, careful phase with ties
    rng = np.random.default_rng(8)
    img = np.full((480, 640), 90, np.uint8)
    img[40:160, 60:220] = rng.integers(0, 256, size=(120, 160), dtype=np.uint8)
    img[300:330, 500:560] = synth.make_frame(9)[300:330, 500:560]
    compare_frame(gpu.Extractor(), orc.Extractor(), img, None, None, None)


def test_bad_box_is_rejected(gpu):
    gray = synth.make_frame(1)
    g = gpu.Extractor()
    with pytest.raises(gpu.CoebError) as e:
        g.extract(gray, np.array([[600, 400, 700, 500]], np.float32), np.array([[610., 410.]], np.float32), [1])
    assert e.value.status == gpu.ERR_BAD_BOX


@pytest.mark.parametrize("w,h,nf", [(1241, 376, 2000), (1920, 1080, 4000), (752, 480, 1200), (333, 257, 500)])
def test_other_resolutions(gpu, w, h, nf):
    gray = synth.make_frame(31, w, h)
    compare_frame(gpu.Extractor(nfeatures=nf), orc.Extractor(nfeatures=nf), gray, None, None, None)


def test_batch_host_matches_per_frame_oracle(gpu):
    B = 12
    batch = synth.make_batch(B, base_seed=40)
    g = gpu.Extractor()
    kps, desc, counts, status = g.extract_batch_host(batch["gray"], batch["boxes"], batch["nbox"], batch["tm"],
                                                     batch["ntm"], batch["blur"])
    assert (status == 0).all()
    c = orc.Extractor()
    flags = 0
    for i in range(B):
        nb, nt = batch["nbox"][i], batch["ntm"][i]
        compare_frame(g, c, batch["gray"][i], batch["boxes"][i, :nb], batch["tm"][i, :nt], batch["blur"][i, :nb],
                      stages=True, frame=i, kps=kps[i, :counts[i]], desc=desc[i, :counts[i]])
        flags += c.dyn_info()["area_flag"]
    assert flags >= 1  # the area_flag path (30/10 thresholds, x0.7 quota, pre-octree cull) was exercised


def test_repeat_call_is_deterministic(gpu):
    gray, boxes, tm, blur = _inputs(2)
    g = gpu.Extractor()
    k1, d1 = g.extract(gray, boxes, tm, blur)
    for _ in range(3):
        k2, d2 = g.extract(gray, boxes, tm, blur)
        assert k1.tobytes() == k2.tobytes() and d1.tobytes() == d2.tobytes()


@pytest.mark.parametrize("nlevels,sf,nf", [(4, 1.5, 800), (6, 1.2, 600), (8, 1.1, 1000), (1, 1.2, 300), (3, 2.5, 300), (3, 2.0, 500)])
def test_other_pyramid_parameters(gpu, nlevels, sf, nf):
    gray = synth.make_frame(33)
    g = gpu.Extractor(nfeatures=nf, scale_factor=sf, nlevels=nlevels)
    c = orc.Extractor(nfeatures=nf, scale_factor=sf, nlevels=nlevels)
    kb, db = c.extract(gray)
    kg, dg = g.extract(gray)
    assert len(kb) == len(kg) and kb.tobytes() == kg.tobytes() and np.array_equal(db, dg)


def test_one_handle_many_shapes_and_capacities(gpu):
    # geometry is rebuilt (and the cached CUDA graphs dropped) when the image size changes; a too-small output capacity
    # is reported, not truncated
    g, c = gpu.Extractor(), orc.Extractor()
    for (w, h) in [(640, 480), (400, 300), (640, 480), (752, 480), (400, 300)]:
        gray = synth.make_frame(w + h, w, h)
        for _ in range(3):   # 3 calls per shape: eager, eager, graph replay
            kg, dg = g.extract(gray)
        kb, db = c.extract(gray)
        assert kb.tobytes() == kg.tobytes() and np.array_equal(db, dg), (w, h)
    gray = synth.make_frame(1)
    with pytest.raises(gpu.CoebError) as e:
        g.extract(gray, cap=100)
    assert e.value.status == gpu.ERR_CAPACITY
    kg, dg = g.extract(gray, cap=5000)
    kb, db = c.extract(gray)
    assert kb.tobytes() == kg.tobytes()


def test_large_host_batch_takes_the_pipelined_path(gpu):
    B = 80   # > 64 frames: 32-frame sub-batches over three streams
    batch = synth.make_batch(B, base_seed=500, unique=10)
    g = gpu.Extractor()
    kps, desc, counts, status = g.extract_batch_host(batch["gray"], batch["boxes"], batch["nbox"], batch["tm"], batch["ntm"], batch["blur"])
    assert (status == 0).all()
    c = orc.Extractor()
    for i in list(range(0, B, 7)) + [31, 32, 63, 64, B - 1]:
        nb, nt = batch["nbox"][i], batch["ntm"][i]
        kb, db = c.extract(batch["gray"][i], batch["boxes"][i, :nb], batch["tm"][i, :nt], batch["blur"][i, :nb])
        assert counts[i] == len(kb) and kps[i, :counts[i]].tobytes() == kb.tobytes() and np.array_equal(desc[i, :counts[i]], db), i


def test_host_batch_of_260_frames_uses_64_frame_sub_batches_with_a_ragged_tail(gpu):
    B = 260   # 4 x 64 + 4: copy stream, three compute streams, download stream
    batch = synth.make_batch(B, base_seed=700, unique=6)
    g = gpu.Extractor()
    kps, desc, counts, status = g.extract_batch_host(batch["gray"], batch["boxes"], batch["nbox"], batch["tm"], batch["ntm"], batch["blur"])
    assert (status == 0).all()
    # same frames through single-frame calls of another handle: the sub-batch boundaries must not show
    g1 = gpu.Extractor()
    c = orc.Extractor()
    for i in (0, 63, 64, 127, 128, 191, 192, 255, 256, 259):
        nb, nt = batch["nbox"][i], batch["ntm"][i]
        args = (batch["gray"][i], batch["boxes"][i, :nb], batch["tm"][i, :nt], batch["blur"][i, :nb])
        k1, d1 = g1.extract(*args)
        kb, db = c.extract(*args)
        assert counts[i] == len(kb) and kps[i, :counts[i]].tobytes() == kb.tobytes() and np.array_equal(desc[i, :counts[i]], db), i
        assert k1.tobytes() == kb.tobytes() and np.array_equal(d1, db)
    # and a second call on the same handle replays the captured graphs
    kps2, desc2, counts2, _ = g.extract_batch_host(batch["gray"], batch["boxes"], batch["nbox"], batch["tm"], batch["ntm"], batch["blur"])
    assert np.array_equal(counts, counts2)
    for i in range(0, B, 13):
        assert kps[i, :counts[i]].tobytes() == kps2[i, :counts[i]].tobytes() and np.array_equal(desc[i, :counts[i]], desc2[i, :counts[i]])


def test_plain_load_twins_of_the_tma_kernels(gpu):
    """COEB_TMA=0 selects the vector-load staging the TMA kernels fall back to (unaligned views, missing driver entry point):
    the switch is read once per process, so the check runs in a child process and compares with the oracle there."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path[:0] = %r\n"
        "import coeb_b200 as cb, orc\n"
        "from coeb_b200 import synth\n"
        "for seed in (0, 3):\n"
        "    gray = synth.make_frame(seed)\n"
        "    boxes, tm, blur = synth.make_dynamic(seed, force_area=(seed == 3))\n"
        "    kg, dg = cb.Extractor().extract(gray, boxes, tm, blur)\n"
        "    kc, dc = orc.Extractor().extract(gray, boxes, tm, blur)\n"
        "    assert kg.tobytes() == kc.tobytes() and np.array_equal(dg, dc), seed\n"
        "b = synth.make_batch(70, base_seed=900, unique=5)\n"
        "ex = cb.Extractor()\n"
        "k, d, c, s = ex.extract_batch_host(b['gray'], b['boxes'], b['nbox'], b['tm'], b['ntm'], b['blur'])\n"
        "o = orc.Extractor()\n"
        "for i in (0, 33, 69):\n"
        "    nb, nt = b['nbox'][i], b['ntm'][i]\n"
        "    kb, db = o.extract(b['gray'][i], b['boxes'][i, :nb], b['tm'][i, :nt], b['blur'][i, :nb])\n"
        "    assert c[i] == len(kb) and k[i, :c[i]].tobytes() == kb.tobytes() and np.array_equal(d[i, :c[i]], db), i\n"
        "print('plain-load path ok')\n") % ([p for p in sys.path if p],)
    env = dict(os.environ, COEB_TMA="0")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "plain-load path ok" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_create_destroy_cycles_do_not_leak_device_memory(gpu):
    """COEB's frame-loss logic re-creates the extractor (`new ORBextractor(nFeatures + 500, ...)`, src/Tracking.cc:434-465) and the
    drop-in matchers create and destroy device frames per call: handles must give their device memory back."""
    import torch
    gray = synth.make_frame(5)
    kps0 = None

    def cycle():
        nonlocal kps0
        ex = gpu.Extractor(1500)
        k, d = ex.extract(gray)
        if kps0 is None:
            kps0 = k.tobytes()
        assert k.tobytes() == kps0
        m = gpu.Matcher()
        f = m.frame(k, d, gpu.Camera(535.4, 539.2, 320.1, 247.6, 40.0, 0.0747, 0.0, 640.0, 0.0, 480.0), ex.tables()["scale"])
        f.close()
        m.close()
        ex.close()
    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(40):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 8 << 20, "device memory shrank by %.1f MB over 40 create/destroy cycles" % ((free0 - free1) / 2 ** 20)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(640, 480), (1241, 376), (333, 257), (1920, 1080)])
def test_small_batch_path_equals_the_batch_path(gpu, w, h):
    """Up to four frames take the latency path (level 0 end to end on a side stream, the whole pyramid in ONE launch whose CTAs
    compute a region of every level in shared memory, per-cell fallback, results mirrored into mapped pinned memory); larger
    batches take the resize chain. Both must produce the same pyramid, bit for bit, and the same keypoints and descriptors."""
    ex = gpu.Extractor(1000, 1.2, 8, 20, 7)
    frames = np.stack([synth.make_frame(40 + i, w, h) for i in range(6)])
    one_k, one_d = ex.extract(frames[0])
    small = ex.launches_per_call()
    one_levels = [ex.level_image(l) for l in range(8)]
    one_blur = [ex.level_image(l, blurred=True) for l in range(8)]
    kps, desc, counts, status = ex.extract_batch_host(frames)
    assert ex.launches_per_call() == 1 + 7 + 1 + 3 + 1 + 1
    assert small == 10, "the single-frame call did not take the one-launch pyramid"
    assert (status == 0).all() and counts[0] == len(one_k)
    for l in range(8):
        assert np.array_equal(one_levels[l], ex.level_image(l, frame=0)), "pyramid level %d" % l
        assert np.array_equal(one_blur[l], ex.level_image(l, frame=0, blurred=True)), "blurred level %d" % l
    assert kps[0, :counts[0]].tobytes() == one_k.tobytes() and np.array_equal(desc[0, :counts[0]], one_d)
    ex.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [2, 3, 4])
def test_batches_of_two_to_four_frames_take_the_latency_path(gpu, B):
    """2-4 frames run like a single one (side-stream level 0, one-launch pyramid with one grid row per frame, per-cell fallback):
    every stage of every frame against the oracle, dynamic boxes and the area_flag thresholds included."""
    batch = synth.make_batch(B, base_seed=3)   # seed 3 takes the area_flag path
    g = gpu.Extractor()
    kps, desc, counts, status = g.extract_batch_host(batch["gray"], batch["boxes"], batch["nbox"], batch["tm"], batch["ntm"], batch["blur"])
    assert (status == 0).all() and g.launches_per_call() == 10
    c = orc.Extractor()
    flags = 0
    for i in range(B):
        nb, nt = batch["nbox"][i], batch["ntm"][i]
        compare_frame(g, c, batch["gray"][i], batch["boxes"][i, :nb], batch["tm"][i, :nt], batch["blur"][i, :nb],
                      stages=True, frame=i, kps=kps[i, :counts[i]], desc=desc[i, :counts[i]])
        flags += c.dyn_info()["area_flag"]
    assert flags >= 1
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("switches", [("COEB_NO_L0_OVERLAP",), ("COEB_NO_PYR_REGIONS",), ("COEB_NO_MIRROR", "COEB_NO_STAGE")])
def test_single_frame_fast_paths_can_be_switched_off(gpu, switches):
    """The latency path of a single-frame call is made of independent shortcuts (level 0 on a side stream, one-launch pyramid,
    results mirrored into mapped memory, staged upload); each has a development switch that restores the plain path. Same results
    either way (child process: some switches are read once)."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path[:0] = %r\n"
        "import coeb_b200 as cb, orc\n"
        "from coeb_b200 import synth\n"
        "ex, o = cb.Extractor(), orc.Extractor()\n"
        "for seed in (0, 3, 5, 0):\n"
        "    gray = synth.make_frame(seed)\n"
        "    boxes, tm, blur = synth.make_dynamic(seed, force_area=(seed == 3))\n"
        "    kg, dg = ex.extract(gray, boxes, tm, blur)\n"
        "    kc, dc = o.extract(gray, boxes, tm, blur)\n"
        "    assert kg.tobytes() == kc.tobytes() and np.array_equal(dg, dc), seed\n"
        "print('switched path ok')\n") % ([p for p in sys.path if p],)
    env = dict(os.environ, **{k: "1" for k in switches})
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "switched path ok" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("w,stride", [(640, 704), (600, 640), (333, 400)])
def test_single_frame_with_a_row_stride(gpu, w, stride):
    """A cv::Mat ROI or an aligned allocation hands over rows that are `stride` bytes apart; the single-frame call stages them
    into the arena pitch on the host. The buffer ends right after the last row's pixels (no trailing padding to read)."""
    import ctypes as C
    h = 257 if w == 333 else 480
    img = synth.make_frame(77, w, h)
    flat = np.zeros(stride * (h - 1) + w, np.uint8)   # exactly as long as the last pixel
    rows = np.lib.stride_tricks.as_strided(flat, shape=(h, w), strides=(stride, 1))
    rows[:] = img
    ex = gpu.Extractor(800, 1.2, 8, 20, 7)
    cap = ex.default_cap(w, h)
    kps, desc, n = np.empty(cap, gpu.KP_DTYPE), np.empty((cap, 32), np.uint8), C.c_int()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for _ in range(3):   # eager calls, then the replayed graph
        st = gpu.lib().coeb_extract(ex.h, P(flat), w, h, stride, None, 0, None, 0, None, 0, P(kps), P(desc), cap, C.byref(n))
        assert st == 0
        kc, dc = orc.Extractor(800, 1.2, 8, 20, 7).extract(img)
        assert n.value == len(kc) and kps[:n.value].tobytes() == kc.tobytes() and np.array_equal(desc[:n.value], dc)
    ex.close()
