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
