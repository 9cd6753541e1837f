"""CPU: the oracle's SearchByBoW restatement against an independent pure-Python transcription of the reference loops
(src/ORBmatcher.cc:158-288, :522-655) that keeps the FeatureVector as a {node: [indices]} map like DBoW2 does."""
import numpy as np
import pytest

import orc
from coeb_b200 import synth

CAM = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
TH_LOW, HISTO = 50, 30


def _bow_python(k1, d1, k2, d2, valid1, valid2, fv1, fv2, ratio, check_ori, strict):
    to_map = lambda fv: {int(fv[0][k]): [int(i) for i in fv[2][fv[1][k]:fv[1][k + 1]]] for k in range(len(fv[0]))}
    m1, m2 = to_map(fv1), to_map(fv2)
    match12 = np.full(len(k1), -1, np.int32)
    matched2 = np.zeros(len(k2), bool)
    hist = [[] for _ in range(HISTO)]
    n = 0
    for node in sorted(set(m1) & set(m2)):   # the lower_bound merge visits exactly the common keys in ascending order
        for i1 in m1[node]:
            if not valid1[i1]:
                continue
            b1, b2, best = 256, 256, -1
            for i2 in m2[node]:
                if matched2[i2] or (valid2 is not None and not valid2[i2]):
                    continue
                dist = int(np.unpackbits(d1[i1] ^ d2[i2]).sum())
                if dist < b1:
                    b2, b1, best = b1, dist, i2
                elif dist < b2:
                    b2 = dist
            if (b1 < TH_LOW if strict else b1 <= TH_LOW) and np.float32(b1) < np.float32(ratio) * np.float32(b2):
                match12[i1] = best
                matched2[best] = True
                if check_ori:
                    rot = np.float32(k1["angle"][i1]) - np.float32(k2["angle"][best])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360.0))
                    v = float(np.float32(rot * np.float32(1.0 / HISTO)))
                    b = int(np.floor(v + 0.5))   # round(): half away from zero, v >= 0
                    hist[0 if b == HISTO else b].append(i1)
                n += 1
    if check_ori:
        sizes = [len(h) for h in hist]
        order = sorted(range(HISTO), key=lambda i: (-sizes[i], i))   # ComputeThreeMaxima: strict '>' keeps the lower bin on ties
        keep = [order[0]] if sizes[order[0]] > 0 else []
        if len(keep) and sizes[order[1]] > 0 and not np.float32(sizes[order[1]]) < np.float32(0.1) * np.float32(sizes[order[0]]):
            keep.append(order[1])
            if sizes[order[2]] > 0 and not np.float32(sizes[order[2]]) < np.float32(0.1) * np.float32(sizes[order[0]]):
                keep.append(order[2])
        for b in range(HISTO):
            if b not in keep:
                for i1 in hist[b]:
                    match12[i1] = -1
                    n -= 1
    return n, match12


@pytest.mark.parametrize("ratio,ori,strict,use_valid2,n_nodes", [(0.7, True, False, False, 60), (0.9, True, True, True, 60), (0.9, False, False, False, 5)])
def test_bow_oracle_against_python_transcription(ratio, ori, strict, use_valid2, n_nodes):
    ex = orc.Extractor(nfeatures=400)
    g1 = synth.make_frame(301)
    k1, d1 = ex.extract(g1)
    k2, d2 = ex.extract(synth.shift_image(g1, -5, 3))
    scale = ex.tables()["scale"]
    f1, f2 = orc.Frame(k1, d1, orc.Camera(*CAM), scale), orc.Frame(k2, d2, orc.Camera(*CAM), scale)
    rng = np.random.default_rng(3)
    valid1 = (rng.random(len(k1)) < 0.85).astype(np.uint8)
    valid2 = (rng.random(len(k2)) < 0.9).astype(np.uint8) if use_valid2 else None
    fv1, fv2 = synth.make_feature_vector(d1, n_nodes, seed=4), synth.make_feature_vector(d2, n_nodes, seed=4)
    n, m12 = orc.match_bow(f1, f2, valid1, valid2, fv1, fv2, ratio, ori, strict)
    n_py, m_py = _bow_python(k1, d1, k2, d2, valid1, valid2, fv1, fv2, ratio, ori, strict)
    assert n == n_py and np.array_equal(m12, m_py)
    assert n > 30


def _tri_python(k1, d1, ur1, k2, d2, ur2, free1, free2, fv1, fv2, F12, ex, ey, scale, only_stereo, check_ori):
    """src/ORBmatcher.cc:657-824 with numpy float32 scalars (every product and sum rounded to float, like the reference)."""
    f32 = np.float32
    to_map = lambda fv: {int(fv[0][k]): [int(i) for i in fv[2][fv[1][k]:fv[1][k + 1]]] for k in range(len(fv[0]))}
    m1, m2 = to_map(fv1), to_map(fv2)
    F = np.asarray(F12, np.float32).reshape(3, 3)
    match12 = np.full(len(k1), -1, np.int32)
    hist = [[] for _ in range(HISTO)]
    n = 0
    for node in sorted(set(m1) & set(m2)):
        for i1 in m1[node]:
            if not free1[i1]:
                continue
            s1 = ur1 is not None and ur1[i1] >= 0
            if only_stereo and not s1:
                continue
            x1, y1 = f32(k1["x"][i1]), f32(k1["y"][i1])
            best, bi = TH_LOW, -1
            for i2 in m2[node]:
                if not free2[i2]:
                    continue
                s2 = ur2 is not None and ur2[i2] >= 0
                if only_stereo and not s2:
                    continue
                dist = int(np.unpackbits(d1[i1] ^ d2[i2]).sum())
                if dist > TH_LOW or dist > best:
                    continue
                x2, y2, sc = f32(k2["x"][i2]), f32(k2["y"][i2]), f32(scale[k2["octave"][i2]])
                if not s1 and not s2:
                    dx, dy = f32(f32(ex) - x2), f32(f32(ey) - y2)
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(f32(100) * sc):
                        continue
                a = f32(f32(f32(x1 * F[0, 0]) + f32(y1 * F[1, 0])) + F[2, 0])
                b = f32(f32(f32(x1 * F[0, 1]) + f32(y1 * F[1, 1])) + F[2, 1])
                c = f32(f32(f32(x1 * F[0, 2]) + f32(y1 * F[1, 2])) + F[2, 2])
                num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
                den = f32(f32(a * a) + f32(b * b))
                if den == 0:
                    continue
                if np.float64(f32(f32(num * num) / den)) < 3.84 * np.float64(f32(sc * sc)):
                    best, bi = dist, i2
            if bi >= 0:
                match12[i1] = bi
                n += 1
                if check_ori:
                    rot = f32(k1["angle"][i1]) - f32(k2["angle"][bi])
                    if rot < 0:
                        rot = f32(rot + f32(360.0))
                    b_ = int(np.floor(float(f32(rot * f32(1.0 / HISTO))) + 0.5))
                    hist[0 if b_ == HISTO else b_].append(i1)
    if check_ori:
        sizes = [len(h) for h in hist]
        order = sorted(range(HISTO), key=lambda i: (-sizes[i], i))
        keep = [order[0]] if sizes[order[0]] > 0 else []
        if len(keep) and sizes[order[1]] > 0 and not f32(sizes[order[1]]) < f32(0.1) * f32(sizes[order[0]]):
            keep.append(order[1])
            if sizes[order[2]] > 0 and not f32(sizes[order[2]]) < f32(0.1) * f32(sizes[order[0]]):
                keep.append(order[2])
        for b_ in range(HISTO):
            if b_ not in keep:
                for i1 in hist[b_]:
                    match12[i1] = -1
                    n -= 1
    return n, match12


@pytest.mark.parametrize("only_stereo,epipole", [(False, (300.0, 200.0)), (True, (1e6, 1e6))])
def test_triangulation_oracle_against_python_transcription(only_stereo, epipole):
    ex = orc.Extractor(nfeatures=400)
    g1 = synth.make_frame(302)
    k1, d1 = ex.extract(g1)
    k2, d2 = ex.extract(synth.shift_image(g1, 5, 2))
    scale = ex.tables()["scale"]
    rng = np.random.default_rng(6)
    ur1 = np.where(rng.random(len(k1)) < 0.6, k1["x"] - np.float32(5), np.float32(-1)).astype(np.float32)
    ur2 = np.where(rng.random(len(k2)) < 0.6, k2["x"] - np.float32(5), np.float32(-1)).astype(np.float32)
    f1, f2 = orc.Frame(k1, d1, orc.Camera(*CAM), scale, ur1), orc.Frame(k2, d2, orc.Camera(*CAM), scale, ur2)
    free1, free2 = (rng.random(len(k1)) < 0.8).astype(np.uint8), (rng.random(len(k2)) < 0.8).astype(np.uint8)
    fv1, fv2 = synth.make_feature_vector(d1, 30, seed=7), synth.make_feature_vector(d2, 30, seed=7)
    F12 = np.array([[0, 0, 2.0], [0, 0, -5.0], [-2.0, 5.0, 0]], np.float32)   # pure shift (5, 2)
    n, m12 = orc.match_triangulation(f1, f2, free1, free2, fv1, fv2, F12, epipole, only_stereo, True)
    n_py, m_py = _tri_python(k1, d1, ur1, k2, d2, ur2, free1, free2, fv1, fv2, F12, epipole[0], epipole[1], scale, only_stereo, True)
    assert n == n_py and np.array_equal(m12, m_py)
    assert n > 15
