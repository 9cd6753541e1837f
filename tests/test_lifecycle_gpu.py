"""GPU: handle lifetimes across threads (ORB-SLAM hands Frames / KeyFrames with device twins from Tracking to the mapping
threads and destroys them there), the keypoint bound the library reports, and the error paths of the single-frame call."""
import ctypes as C
import threading

import numpy as np
import pytest

import orc
from coeb_b200 import synth

pytestmark = pytest.mark.gpu

CAM_ARGS = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)


@pytest.fixture(scope="module")
def gpu():
    import coeb_b200
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible")
    return coeb_b200


def test_frames_destroyed_on_other_threads_and_after_their_matcher(gpu):
    ex = orc.Extractor()
    kps, desc = ex.extract(synth.make_frame(7))
    scale = ex.tables()["scale"]
    cam = gpu.Camera(*CAM_ARGS)
    m = gpu.Matcher()
    ref = m.frame(kps, desc, cam, scale).features_in_area(320.0, 240.0, 50.0)
    errors = []

    def churn(seed):   # frames are created by the owner thread and destroyed by four others, concurrently with new creations
        try:
            for _ in range(40):
                f = handoff[seed].pop() if handoff[seed] else None
                if f is not None:
                    f.close()
        except Exception as e:   # noqa: BLE001
            errors.append(e)
    for rounds in range(5):
        handoff = [[m.frame(kps, desc, cam, scale) for _ in range(10)] for _ in range(4)]
        ts = [threading.Thread(target=churn, args=(i,)) for i in range(4)]
        for t in ts:
            t.start()
        live = [m.frame(kps, desc, cam, scale) for _ in range(10)]   # pops the pool while the others push to it
        for t in ts:
            t.join()
        for f in live:
            assert np.array_equal(f.features_in_area(320.0, 240.0, 50.0), ref)
            f.close()
    assert not errors
    # a frame that outlives its matcher: destroying it is fine, using it is a loud error
    m2 = gpu.Matcher()
    orphan = m2.frame(kps, desc, cam, scale)
    m2.close()
    with pytest.raises(RuntimeError):
        orphan.features_in_area(320.0, 240.0, 50.0)
    orphan.close()
    assert np.array_equal(m.frame(kps, desc, cam, scale).features_in_area(320.0, 240.0, 50.0), ref)


@pytest.mark.parametrize("nf,w,h", [(1000, 640, 480), (40, 1241, 376), (16, 640, 480), (4000, 1920, 1080)])
def test_max_keypoints_is_a_bound_and_small_quotas_succeed(gpu, nf, w, h):
    """The first octree round splits every root (4 x nIni nodes per level) whatever the quota, so nfeatures + a margin is
    not a bound; coeb_extractor_max_keypoints is, and the default capacity of the wrappers comes from it."""
    g, c = gpu.Extractor(nfeatures=nf), orc.Extractor(nfeatures=nf)
    gray = synth.make_frame(21, w, h)
    bound = C.c_int()
    assert gpu.lib().coeb_extractor_max_keypoints(g.h, w, h, C.byref(bound)) == 0
    kg, dg = g.extract(gray)
    kc, dc = c.extract(gray)
    assert kg.tobytes() == kc.tobytes() and np.array_equal(dg, dc)
    assert len(kg) <= bound.value <= g.default_cap(w, h)


def test_single_frame_error_paths(gpu):
    L = gpu.lib()
    g = gpu.Extractor()
    gray = synth.make_frame(3)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    kps, desc, n = np.zeros(2000, gpu.KP_DTYPE), np.zeros((2000, 32), np.uint8), C.c_int(-5)
    # a positive count with a null array is refused, not dereferenced
    assert L.coeb_extract(g.h, P(gray), 640, 480, 640, None, 2, None, 0, None, 0, P(kps), P(desc), 2000, C.byref(n)) == -1
    assert L.coeb_extract(g.h, P(gray), 640, 480, 640, None, 0, None, 3, None, 0, P(kps), P(desc), 2000, C.byref(n)) == -1
    # capacity too small: the count is reported, the caller's arrays are left alone
    assert L.coeb_extract(g.h, P(gray), 640, 480, 640, None, 0, None, 0, None, 0, P(kps), P(desc), 2000, C.byref(n)) == 0
    full = n.value
    kps2, desc2 = np.full(100, 0, gpu.KP_DTYPE), np.full((100, 32), 0xAB, np.uint8)
    st = L.coeb_extract(g.h, P(gray), 640, 480, 640, None, 0, None, 0, None, 0, P(kps2), P(desc2), 100, C.byref(n))
    assert st == -4 and n.value == full
    assert (desc2 == 0xAB).all() and not kps2["x"].any()
    # and the handle still works afterwards
    assert L.coeb_extract(g.h, P(gray), 640, 480, 640, None, 0, None, 0, None, 0, P(kps), P(desc), 2000, C.byref(n)) == 0 and n.value == full


def test_growing_batches_do_not_replay_stale_graphs(gpu):
    """Batch sizes that force the arenas and staging buffers to be reallocated between graph-replayed calls."""
    g, c = gpu.Extractor(), orc.Extractor()
    batch = synth.make_batch(40, base_seed=50, unique=8)
    ref = [c.extract(batch["gray"][i], batch["boxes"][i, :batch["nbox"][i]], batch["tm"][i, :batch["ntm"][i]], batch["blur"][i, :batch["nbox"][i]])
           for i in range(40)]
    for B in (1, 1, 1, 4, 4, 4, 1, 1, 40, 40, 40, 4, 4, 1):
        sub = {k: np.ascontiguousarray(v[:B]) for k, v in batch.items()}
        kps, desc, counts, status = g.extract_batch_host(sub["gray"], sub["boxes"], sub["nbox"], sub["tm"], sub["ntm"], sub["blur"])
        assert (status == 0).all()
        for i in range(B):
            assert kps[i, :counts[i]].tobytes() == ref[i][0].tobytes() and np.array_equal(desc[i, :counts[i]], ref[i][1]), (B, i)
