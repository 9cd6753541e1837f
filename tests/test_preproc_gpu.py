"""GPU parity of the producers in front of the extractor ("next" rows of SURVEY.md section 8f): RGB->gray conversion and
the per-box Laplacian blur flag, plus the fully device-resident chain RGB -> gray -> blur flags -> extraction and the
unaligned device-input path of the extractor."""
import numpy as np
import pytest

import orc
from coeb_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import coeb_b200
    if coeb_b200.device_count() < 1:
        pytest.fail("no sm_100 device visible")
    return coeb_b200


@pytest.mark.parametrize("shape,bgr", [((480, 640, 3), False), ((480, 640, 3), True), ((97, 131, 4), False), ((33, 47, 4), True)])
def test_rgb_to_gray_bit_exact(gpu, shape, bgr):
    from coeb_b200 import preproc
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    ex = gpu.Extractor()
    assert np.array_equal(preproc.rgb_to_gray(ex, img, bgr), orc.rgb_to_gray(img, bgr))


def test_blur_flags_match_oracle_and_flag_both_ways(gpu):
    from coeb_b200 import preproc
    gray = synth.make_frame(5)
    gray[100:300, 50:250] = 120          # a flat region: Laplacian mean 0 -> blurred
    gray[310:400, 300:500] = synth.make_frame(6)[310:400, 300:500] // 8 + 100   # faint texture
    boxes = np.array([[60, 110, 240, 290], [300, 310, 500, 400], [400, 20, 630, 300], [0, 0, 640, 480], [10, 10, 11, 11]], np.float32)
    ex = gpu.Extractor()
    fg, mg = preproc.blur_flags(ex, gray, boxes)
    fc, mc = orc.blur_flags(gray, boxes)
    assert np.array_equal(fg, fc) and np.array_equal(mg, mc)
    assert fg[0] == 1 and fg[2] == 0 and mg[0] == 0.0


def test_device_resident_chain_rgb_to_keypoints(gpu):
    """RGB frames + boxes in HBM -> gray -> blur flags -> extraction, all enqueued on one stream, against the oracle."""
    import torch
    from coeb_b200 import preproc
    B, W, H = 4, 640, 480
    rng = np.random.default_rng(2)
    batch = synth.make_batch(B, base_seed=60)
    rgb = np.stack([batch["gray"]] * 3, axis=-1).astype(np.int16) + rng.integers(-12, 13, size=(B, H, W, 3))
    rgb = np.clip(rgb, 0, 255).astype(np.uint8)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    ex = gpu.Extractor()
    ex.set_stream(stream.cuda_stream)
    cap = ex.default_cap()
    with torch.cuda.stream(stream):
        d_rgb = torch.from_numpy(rgb).to(dev)
        d_gray = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        d_boxes, d_nbox = torch.from_numpy(batch["boxes"]).to(dev), torch.from_numpy(batch["nbox"]).to(dev)
        d_tm, d_ntm = torch.from_numpy(batch["tm"]).to(dev), torch.from_numpy(batch["ntm"]).to(dev)
        d_flags = torch.zeros((B, synth.MAX_BOX), dtype=torch.int32, device=dev)
        d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
        d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
        d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
        d_st = torch.zeros(B, dtype=torch.int32, device=dev)
        preproc.rgb_to_gray_batch_device(ex, B, d_rgb.data_ptr(), W, H, W * 3, W * H * 3, 3, False, d_gray.data_ptr(), W, W * H)
        preproc.blur_flags_batch_device(ex, B, d_gray.data_ptr(), W, H, W, W * H, d_boxes.data_ptr(), d_nbox.data_ptr(), synth.MAX_BOX,
                                        d_flags.data_ptr())
        ex.extract_batch_device(B, d_gray.data_ptr(), W, H, W, W * H, d_boxes.data_ptr(), d_nbox.data_ptr(), synth.MAX_BOX, d_tm.data_ptr(),
                                d_ntm.data_ptr(), synth.MAX_TM, d_flags.data_ptr(), d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(),
                                d_st.data_ptr(), cap)
    stream.synchronize()
    assert (d_st.cpu().numpy() == 0).all()
    cnt = d_cnt.cpu().numpy()
    kps = d_kps.cpu().numpy().view(gpu.KP_DTYPE).reshape(B, cap)
    desc = d_desc.cpu().numpy()
    c = orc.Extractor()
    for i in range(B):
        gray = orc.rgb_to_gray(rgb[i])
        nb, nt = batch["nbox"][i], batch["ntm"][i]
        flags, _ = orc.blur_flags(gray, batch["boxes"][i, :nb])
        assert np.array_equal(flags, d_flags[i, :nb].cpu().numpy())
        kc, dc = c.extract(gray, batch["boxes"][i, :nb], batch["tm"][i, :nt], flags)
        assert cnt[i] == len(kc) and kps[i, :cnt[i]].tobytes() == kc.tobytes() and np.array_equal(desc[i, :cnt[i]], dc)


def test_unaligned_device_input_is_staged(gpu):
    """A tightly packed 1241-px-wide device buffer is not 16-byte aligned per row: the extractor must copy it into its
    aligned arena and still match the oracle."""
    import torch
    W, H, nf = 1241, 376, 2000
    gray = synth.make_frame(70, W, H)
    dev = torch.device("cuda:0")
    ex = gpu.Extractor(nfeatures=nf)
    cap = ex.default_cap()
    d_gray = torch.from_numpy(np.stack([gray, gray[::-1].copy()])).to(dev)
    d_kps = torch.empty((2, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((2, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(2, dtype=torch.int32, device=dev)
    d_st = torch.zeros(2, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ex.extract_batch_device(2, d_gray.data_ptr(), W, H, W, W * H, None, None, 0, None, None, 0, None, d_kps.data_ptr(), d_desc.data_ptr(),
                            d_cnt.data_ptr(), d_st.data_ptr(), cap)
    ex.dyn_info(0)  # blocks on the extractor's stream
    cnt = d_cnt.cpu().numpy()
    kps = d_kps.cpu().numpy().view(gpu.KP_DTYPE).reshape(2, cap)
    c = orc.Extractor(nfeatures=nf)
    for i, g in enumerate((gray, gray[::-1].copy())):
        kc, dc = c.extract(g)
        assert cnt[i] == len(kc) and kps[i, :cnt[i]].tobytes() == kc.tobytes()
        assert np.array_equal(d_desc[i, :cnt[i]].cpu().numpy(), dc)
