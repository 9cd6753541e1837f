"""Seeded synthetic inputs for the COEB front end.

numpy only, so the same seed gives the same bytes in the dev container and on the GPU box
(golden fixtures store an input checksum, not the input).

Frame generator ("TUM-shaped"): low-frequency background + random axis-aligned / rotated rectangles
and checker patches (contrast 20..120) + low-contrast zones (contrast 8..16 on a flat background, so
that cells there yield nothing at iniThFAST=20 and must fall back to minThFAST=7) + sigma=2 noise.
"""
import numpy as np

MAX_BOX = 4
MAX_TM = 64


def _upsample_bilinear(small, h, w):
    sh, sw = small.shape
    ys = (np.arange(h) + 0.5) * sh / h - 0.5
    xs = (np.arange(w) + 0.5) * sw / w - 0.5
    y0 = np.clip(np.floor(ys).astype(np.int64), 0, sh - 2)
    x0 = np.clip(np.floor(xs).astype(np.int64), 0, sw - 2)
    fy = np.clip(ys - y0, 0, 1)[:, None]
    fx = np.clip(xs - x0, 0, 1)[None, :]
    a = small[y0][:, x0]
    b = small[y0][:, x0 + 1]
    c = small[y0 + 1][:, x0]
    d = small[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def _add_shape(img, rng, lo, hi, region=None):
    h, w = img.shape
    if region is None:
        rx0, ry0, rx1, ry1 = 0, 0, w, h
    else:
        rx0, ry0, rx1, ry1 = region
    sw = int(rng.integers(6, 60))
    sh = int(rng.integers(6, 60))
    cx = float(rng.uniform(rx0, rx1))
    cy = float(rng.uniform(ry0, ry1))
    contrast = float(rng.uniform(lo, hi)) * (1 if rng.random() < 0.5 else -1)
    kind = rng.random()
    ang = float(rng.uniform(0, np.pi)) if kind > 0.4 else 0.0
    rad = int(np.ceil(0.5 * np.hypot(sw, sh))) + 1
    x0, x1 = max(int(cx) - rad, rx0), min(int(cx) + rad + 1, rx1)
    y0, y1 = max(int(cy) - rad, ry0), min(int(cy) + rad + 1, ry1)
    if x1 <= x0 or y1 <= y0:
        return
    yy, xx = np.mgrid[y0:y1, x0:x1]
    dx, dy = xx - cx, yy - cy
    ca, sa = np.cos(ang), np.sin(ang)
    u = dx * ca + dy * sa
    v = -dx * sa + dy * ca
    inside = (np.abs(u) <= sw / 2) & (np.abs(v) <= sh / 2)
    if kind > 0.8:  # checker patch
        cell = float(rng.integers(4, 10))
        chk = ((np.floor(u / cell) + np.floor(v / cell)) % 2) * 2 - 1
        img[y0:y1, x0:x1] += inside * chk * contrast
    else:
        img[y0:y1, x0:x1] += inside * contrast


def make_frame(seed, w=640, h=480):
    """Returns a uint8 [h, w] grayscale frame."""
    rng = np.random.default_rng(int(seed))
    small = rng.uniform(0, 255, size=(max(h // 16, 2), max(w // 16, 2)))
    img = np.ascontiguousarray(_upsample_bilinear(small, h, w))
    img = 0.6 * img + 50.0
    nshape = int(rng.integers(200, 401) * (w * h) / (640 * 480))
    for _ in range(nshape):
        _add_shape(img, rng, 20, 120)
    # low-contrast zones: flat background, faint shapes only (forces the minThFAST fallback)
    nz = int(rng.integers(2, 5))
    for _ in range(nz):
        zw, zh = int(rng.integers(w // 8, w // 4)), int(rng.integers(h // 8, h // 3))
        zx, zy = int(rng.integers(0, w - zw)), int(rng.integers(0, h - zh))
        img[zy:zy + zh, zx:zx + zw] = float(rng.uniform(60, 190))
        for _ in range(int(rng.integers(0, 12))):
            _add_shape(img, rng, 9, 17, region=(zx, zy, zx + zw, zy + zh))
    img += rng.normal(0.0, 2.0, size=img.shape)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def make_dynamic(seed, w=640, h=480, force_area=False, nbox=None):
    """Person boxes, moving points T_M and blur flags for one frame.

    Returns boxes [n,4] float32 (integer-valued xmin,ymin,xmax,ymax), tm [m,2] float32, blur [n] int32.
    force_area: two large boxes densely covered by T_M so that the summed dynamic area exceeds
    200000 px^2 (area_flag path: thresholds 30/10, quota x0.7, cull before the octree).
    """
    rng = np.random.default_rng(int(seed) + 7_000_003)
    sx, sy = w / 640.0, h / 480.0
    if force_area:
        n = 2
    elif nbox is not None:
        n = nbox
    else:
        n = int(rng.integers(0, 4))
    boxes = np.zeros((n, 4), np.float32)
    for b in range(n):
        if force_area:
            bw, bh = int(rng.integers(280, 301) * sx), int(rng.integers(400, 421) * sy)
        else:
            bw, bh = int(rng.integers(80, 301) * sx), int(rng.integers(150, 421) * sy)
        x0 = int(rng.integers(0, w - bw))
        y0 = int(rng.integers(0, h - bh))
        boxes[b] = (x0, y0, x0 + bw, y0 + bh)
    m = int(rng.integers(40, 61)) if force_area else int(rng.integers(0, 61))
    tm = np.zeros((m, 2), np.float32)
    for t in range(m):
        if n > 0 and rng.random() < (0.95 if force_area else 0.7):
            b = int(rng.integers(0, n))
            tm[t] = (rng.uniform(boxes[b, 0], boxes[b, 2]), rng.uniform(boxes[b, 1], boxes[b, 3]))
        else:
            tm[t] = (rng.uniform(5, w - 5), rng.uniform(5, h - 5))
    tm[:, 0] = np.clip(tm[:, 0], 5, w - 5.001)
    tm[:, 1] = np.clip(tm[:, 1], 5, h - 5.001)
    blur = (rng.random(n) < 0.3).astype(np.int32)
    return boxes, tm, blur


def make_batch(n, base_seed=0, w=640, h=480, with_dynamic=True, unique=None):
    """Packed batch: gray [n,h,w] u8, boxes [n,MAX_BOX,4] f32, nbox [n] i32, tm [n,MAX_TM,2] f32,
    ntm [n] i32, blur [n,MAX_BOX] i32. Every 8th frame (seed % 8 == 3) takes the area_flag path.

    unique: generate only this many distinct images and derive the rest by circular shifts of them
    (keeps start-up short for large batches); dynamic inputs are always per-frame."""
    gray = np.empty((n, h, w), np.uint8)
    boxes = np.zeros((n, MAX_BOX, 4), np.float32)
    nbox = np.zeros(n, np.int32)
    tm = np.zeros((n, MAX_TM, 2), np.float32)
    ntm = np.zeros(n, np.int32)
    blur = np.zeros((n, MAX_BOX), np.int32)
    nuniq = n if unique is None else min(unique, n)
    base = [make_frame(base_seed + i, w, h) for i in range(nuniq)]
    for i in range(n):
        if i < nuniq:
            gray[i] = base[i]
        else:
            k = i // nuniq
            gray[i] = np.roll(base[i % nuniq], (7 * k, 13 * k), axis=(0, 1))
        if with_dynamic:
            seed = base_seed + i
            b, t, f = make_dynamic(seed, w, h, force_area=(seed % 8 == 3))
            nbox[i], ntm[i] = len(b), len(t)
            boxes[i, :len(b)] = b
            tm[i, :len(t)] = t
            blur[i, :len(b)] = f
    return dict(gray=gray, boxes=boxes, nbox=nbox, tm=tm, ntm=ntm, blur=blur)


def flip_bits(desc, rng, max_flips):
    """Copy of uint8 descriptors [n,32] with 0..max_flips random bit flips per row."""
    out = desc.copy()
    for i in range(len(out)):
        k = int(rng.integers(0, max_flips + 1))
        if k:
            pos = rng.choice(256, size=k, replace=False)
            for p in pos:
                out[i, p >> 3] ^= np.uint8(1 << (p & 7))
    return out
