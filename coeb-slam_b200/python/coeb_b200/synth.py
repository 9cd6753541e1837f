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


# ---------------------------------------------------------------------------------------------------
# Matching workloads
# ---------------------------------------------------------------------------------------------------
def make_map_points(kps, desc, scale, seed=0, n_map=5000, n_true=800, w=640, h=480, bf=40.0):
    """Config 3: `n_true` true correspondences (a frame descriptor with 0..40 bit flips, projected at the
    keypoint + N(0, 3 px), predicted level = octave or octave + 1) and distractors with random descriptors
    at uniform positions. Returns the SoA dict of coeb_match_projection plus a uRight array for the frame."""
    rng = np.random.default_rng(seed + 91)
    n = len(kps)
    nlevels = len(scale)
    n_true = min(n_true, n)
    src = rng.choice(n, size=n_true, replace=False)
    d = np.empty((n_map, 32), np.uint8)
    d[:n_true] = flip_bits(desc[src], rng, 40)
    d[n_true:] = rng.integers(0, 256, size=(n_map - n_true, 32), dtype=np.uint8)
    px = np.empty(n_map, np.float32)
    py = np.empty(n_map, np.float32)
    px[:n_true] = kps["x"][src] + rng.normal(0, 3, n_true)
    py[:n_true] = kps["y"][src] + rng.normal(0, 3, n_true)
    px[n_true:] = rng.uniform(0, w, n_map - n_true)
    py[n_true:] = rng.uniform(0, h, n_map - n_true)
    level = np.empty(n_map, np.int32)
    level[:n_true] = np.minimum(kps["octave"][src] + rng.integers(0, 2, n_true), nlevels - 1)
    level[n_true:] = rng.integers(0, nlevels, n_map - n_true)
    perm = rng.permutation(n_map)
    depth = rng.uniform(0.5, 5.0, n_map).astype(np.float32)
    mp = dict(track_in_view=(rng.random(n_map) < 0.95).astype(np.uint8), bad=(rng.random(n_map) < 0.02).astype(np.uint8),
              has_obs=(rng.random(n_map) < 0.97).astype(np.uint8), proj_x=px, proj_y=py,
              proj_xr=(px - np.float32(bf) / depth).astype(np.float32), level=level,
              view_cos=rng.uniform(0.9, 1.0, n_map).astype(np.float32), desc=d)
    mp = {k: np.ascontiguousarray(v[perm]) for k, v in mp.items()}
    uright = np.where(rng.random(n) < 0.3, kps["x"] - np.float32(bf) / rng.uniform(0.5, 5.0, n).astype(np.float32),
                      np.float32(-1)).astype(np.float32)
    return mp, uright


def make_last_frame(kps, desc, seed=0, fx=535.4, fy=539.2, cx=320.1, cy=247.6, shift=(0.02, -0.01, 0.03)):
    """Config 3 (frame to frame): every keypoint of a 'last frame' gets a 3-D point that projects, under the
    current pose, near a keypoint of the current frame. Returns (last SoA dict, Tcw_cur, Tcw_last)."""
    rng = np.random.default_rng(seed + 17)
    n = len(kps)
    z = rng.uniform(0.8, 6.0, n).astype(np.float32)
    # small rotation about y plus a translation
    a = 0.01
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    t = np.array(shift, np.float32)
    Tc = np.concatenate([R, t[:, None]], axis=1).astype(np.float32)
    Tl = np.concatenate([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)], axis=1)
    # camera-frame points that project onto the keypoints (+ noise), moved back to the world frame
    u = kps["x"] + rng.normal(0, 2.0, n).astype(np.float32)
    v = kps["y"] + rng.normal(0, 2.0, n).astype(np.float32)
    pc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], axis=1).astype(np.float32)
    pw = ((pc - t) @ R).astype(np.float32)  # R^T (pc - t)
    last = dict(valid=(rng.random(n) < 0.8).astype(np.uint8), has_obs=(rng.random(n) < 0.97).astype(np.uint8), xyz=pw,
                octave=np.clip(kps["octave"] + rng.integers(-1, 2, n), 0, 7).astype(np.int32), angle=(kps["angle"] + rng.normal(0, 4.0, n)).astype(np.float32) % 360,
                desc=flip_bits(desc, rng, 30))
    return last, Tc, Tl


def shift_image(gray, dx, dy):
    """Second view for the initialisation matcher: the frame shifted by (dx, dy), edges replicated."""
    h, w = gray.shape
    ys = np.clip(np.arange(h) - dy, 0, h - 1)
    xs = np.clip(np.arange(w) - dx, 0, w - 1)
    return np.ascontiguousarray(gray[ys][:, xs])


def make_stereo_right(left, seed=0, dmin=2, dmax=80, band=24):
    """Config 4: right image = left image with a per-row-band horizontal disparity in [dmin, dmax] px."""
    rng = np.random.default_rng(seed + 33)
    h, w = left.shape
    right = np.empty_like(left)
    for y0 in range(0, h, band):
        d = int(rng.integers(dmin, dmax + 1))
        xs = np.clip(np.arange(w) + d, 0, w - 1)
        right[y0:y0 + band] = left[y0:y0 + band][:, xs]
    return right


def make_knn_sets(nq=4000, nt=100000, seed=0):
    """Config 5: random query descriptors; 10% of the train set are noisy copies of queries."""
    rng = np.random.default_rng(seed + 55)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    ncopy = nt // 10
    src = rng.integers(0, nq, ncopy)
    pos = rng.choice(nt, size=ncopy, replace=False)
    noisy = q[src].copy()
    flips = rng.integers(0, 60, ncopy)
    for i in range(ncopy):
        k = int(flips[i])
        if k:
            p = rng.choice(256, size=k, replace=False)
            np.bitwise_xor.at(noisy[i], p >> 3, (1 << (p & 7)).astype(np.uint8))
    t[pos] = noisy
    return q, t


def make_pose(seed=0):
    """A small camera motion: Tcw (3x4 row-major, float32) and the camera centre Ow = -Rcw^T tcw (float32)."""
    rng = np.random.default_rng(seed + 301)
    ax, ay, az = rng.uniform(-0.05, 0.05, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    R = (Rz @ Ry @ Rx).astype(np.float32)
    t = rng.uniform(-0.2, 0.2, 3).astype(np.float32)
    Tcw = np.concatenate([R, t[:, None]], axis=1).astype(np.float32)
    Ow = (-(R.T.astype(np.float32) @ t)).astype(np.float32)
    return Tcw, Ow


def make_local_map(kps, desc, scale, Tcw, seed=0, n_map=5000, n_true=800, fx=535.4, fy=539.2, cx=320.1, cy=247.6):
    """Config 3 with the visibility test in front (Tracking::SearchLocalPoints): world points, normals and scale-invariance
    distances of a local map. `n_true` points project (under Tcw) within a few pixels of a frame keypoint, carry its
    descriptor with 0..40 bit flips and a distance range that predicts the keypoint's octave or the next one; the rest are
    distractors spread so that every rejection of Frame::isInFrustum fires (behind the camera, outside the image, outside
    the distance range, grazing view). Returns (local-map SoA dict, skip, has_obs)."""
    rng = np.random.default_rng(seed + 57)
    n = len(kps)
    nlevels = len(scale)
    sf = float(scale[1]) if nlevels > 1 else 1.2
    n_true = min(n_true, n)
    R, t = Tcw[:, :3].astype(np.float64), Tcw[:, 3].astype(np.float64)
    Ow = -R.T @ t
    src = rng.choice(n, size=n_true, replace=False)
    u = np.empty(n_map)
    v = np.empty(n_map)
    u[:n_true] = kps["x"][src] + rng.normal(0, 3, n_true)
    v[:n_true] = kps["y"][src] + rng.normal(0, 3, n_true)
    u[n_true:] = rng.uniform(-120, 760, n_map - n_true)    # some outside the image
    v[n_true:] = rng.uniform(-90, 570, n_map - n_true)
    z = rng.uniform(0.6, 6.0, n_map)
    z[n_true:][rng.random(n_map - n_true) < 0.1] *= -1      # behind the camera
    pc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], axis=1)
    pw = (pc - t) @ R                                       # R^T (pc - t)
    po = pw - Ow
    dist = np.linalg.norm(po, axis=1)
    # normals: towards the camera, tilted; distractors include grazing / back-facing ones
    tilt = rng.normal(0, 0.25, (n_map, 3))
    tilt[n_true:] = rng.normal(0, 1.2, (n_map - n_true, 3))
    nrm = po / dist[:, None] + tilt
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    level = np.empty(n_map, np.int64)
    level[:n_true] = np.minimum(kps["octave"][src] + rng.integers(0, 2, n_true), nlevels - 1)
    level[n_true:] = rng.integers(0, nlevels, n_map - n_true)
    max_dist = dist * sf ** (level - rng.uniform(0.05, 0.95, n_map))     # PredictScale -> level
    far = rng.random(n_map) < 0.05
    far[:n_true] = False
    max_dist[far] *= rng.choice([0.05, 40.0], size=int(far.sum()))       # outside [0.8 min, 1.2 max] either way
    min_dist = max_dist / float(scale[-1])
    d = np.empty((n_map, 32), np.uint8)
    d[:n_true] = flip_bits(desc[src], rng, 40)
    d[n_true:] = rng.integers(0, 256, size=(n_map - n_true, 32), dtype=np.uint8)
    perm = rng.permutation(n_map)
    lm = dict(xyz=pw.astype(np.float32)[perm], normal=nrm.astype(np.float32)[perm], min_dist=min_dist.astype(np.float32)[perm],
              max_dist=max_dist.astype(np.float32)[perm], desc=d[perm])
    lm = {k: np.ascontiguousarray(a) for k, a in lm.items()}
    skip = (rng.random(n_map) < 0.04).astype(np.uint8)
    has_obs = (rng.random(n_map) < 0.97).astype(np.uint8)
    return lm, skip, has_obs


def make_depth(seed=0, w=640, h=480, factor=5000.0):
    """TUM-style raw depth map: uint16, metres x `factor`, smooth 0.5-5 m surface with ~15 % invalid (0) pixels."""
    rng = np.random.default_rng(seed + 411)
    base = _upsample_bilinear(rng.uniform(0.5, 5.0, (h // 32 + 2, w // 32 + 2)), h, w)
    d = np.clip(base * factor, 1, 65535).astype(np.uint16)
    holes = _upsample_bilinear(rng.random((h // 16 + 2, w // 16 + 2)), h, w) < 0.3
    d[holes] = 0
    return np.ascontiguousarray(d)


def make_feature_vector(desc, n_nodes=100, seed=0):
    """A stand-in for DBoW2's FeatureVector (vocabulary->transform(..., levelsup=4)): every descriptor is assigned to the
    nearest of `n_nodes` fixed random 256-bit centres (Hamming), which is what a vocabulary-tree node at that level is.
    Returns the CSR triple (node ids ascending, start, items in ascending feature index = DBoW2's push order)."""
    rng = np.random.default_rng(seed + 733)   # the 'vocabulary' depends on the seed only
    centres = rng.integers(0, 256, size=(n_nodes, 32), dtype=np.uint8)
    d = np.unpackbits(desc[:, None, :] ^ centres[None, :, :], axis=2).sum(axis=2)
    node_of = d.argmin(axis=1)
    nodes, start, items = [], [0], []
    for k in range(n_nodes):
        idx = np.nonzero(node_of == k)[0]
        if len(idx) == 0:
            continue   # a FeatureVector only holds the nodes that occur
        nodes.append(7 * k + 3)   # arbitrary ascending ids
        items.extend(idx.tolist())
        start.append(len(items))
    return np.array(nodes, np.int32), np.array(start, np.int32), np.array(items, np.int32)


def warp_affine(gray, A, t):
    """Bilinear resampling of `gray` under x_src = A @ x_dst + t (2x2 matrix, 2-vector), edges replicated, rounded to uint8."""
    h, w = gray.shape
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = A[0][0] * xs + A[0][1] * ys + t[0]
    sy = A[1][0] * xs + A[1][1] * ys + t[1]
    sx = np.clip(sx, 0, w - 1.001)
    sy = np.clip(sy, 0, h - 1.001)
    x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
    fx, fy = sx - x0, sy - y0
    g = gray.astype(np.float64)
    v = (g[y0, x0] * (1 - fx) * (1 - fy) + g[y0, x0 + 1] * fx * (1 - fy) + g[y0 + 1, x0] * (1 - fx) * fy + g[y0 + 1, x0 + 1] * fx * fy)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def warp_flow(gray, fx, fy):
    """cur(x, y) = gray(x - fx(x, y), y - fy(x, y)), bilinear, edges replicated, rounded to uint8."""
    h, w = gray.shape
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = np.clip(xs - fx, 0, w - 1.001)
    sy = np.clip(ys - fy, 0, h - 1.001)
    x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
    ax, ay = sx - x0, sy - y0
    g = gray.astype(np.float64)
    v = g[y0, x0] * (1 - ax) * (1 - ay) + g[y0, x0 + 1] * ax * (1 - ay) + g[y0 + 1, x0] * (1 - ax) * ay + g[y0 + 1, x0 + 1] * ax * ay
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def make_motion_pair(seed, w=640, h=480, moving=True):
    """Two consecutive frames for Frame::ProcessMovingObject: a translating camera in front of a scene with smoothly varying
    depth (motion parallax, so that the epipolar geometry is well determined), one or two rectangular 'objects' that move on
    their own against it, and pixel noise that differs between the frames. Returns (prev, cur, boxes[n, 4] of the moving
    objects in the current frame)."""
    rng = np.random.default_rng(seed + 9001)
    prev = make_frame(seed + 500, w, h)
    depth = 2.0 + 3.0 * _upsample_bilinear(rng.random((5, 6)), h, w)             # metres
    f, cx, cy = 520.0, w / 2.0, h / 2.0
    t = np.array([rng.uniform(0.015, 0.03) * rng.choice([-1, 1]), rng.uniform(-0.008, 0.008), rng.uniform(-0.03, 0.03)])
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    flow_x = (f * t[0] - (xs - cx) * t[2]) / depth
    flow_y = (f * t[1] - (ys - cy) * t[2]) / depth
    cur = warp_flow(prev, flow_x, flow_y)
    boxes = []
    if moving:
        for _ in range(int(rng.integers(1, 3))):
            bw, bh = int(rng.integers(90, 200)), int(rng.integers(120, 260))
            x0, y0 = int(rng.integers(20, w - bw - 20)), int(rng.integers(20, h - bh - 20))
            d = np.array([rng.uniform(-3, 3), rng.uniform(6, 11) * rng.choice([-1, 1])])   # mostly across the epipolar lines
            obj = warp_flow(prev, np.full((h, w), d[0]), np.full((h, w), d[1]))
            cur[y0:y0 + bh, x0:x0 + bw] = obj[y0:y0 + bh, x0:x0 + bw]
            boxes.append([x0, y0, x0 + bw, y0 + bh])
    noise = rng.normal(0, 1.5, cur.shape)
    cur = np.clip(np.rint(cur.astype(np.float64) + noise), 0, 255).astype(np.uint8)
    return prev, cur, np.array(boxes, np.float32).reshape(-1, 4)
