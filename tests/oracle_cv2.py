"""Oracle A: an independent Python transcription of the reference extractor's control flow that calls
the REAL OpenCV primitives (cv2 4.13.0): cv2.resize, cv2.FastFeatureDetector on ROI views,
cv2.fastAtan2, cv2.GaussianBlur. It validates oracle B (oracle/, C++ restatement with integer models
of those primitives) stage by stage. Test infrastructure only.

Reference lines followed: src/ORBextractor.cc:80-156 (IC_Angle, descriptor), :418-477 (ctor),
:489-769 (octree), :771-904 (cell loop), :1088-1342 (operator()), :1344-1450 (pyramid, culling).
"""
import math

import cv2
import numpy as np

f32 = np.float32
EDGE = 19
PATCH = 31
HALF = 15


def cv_round(v):
    return int(np.rint(v))


class ExtractorA:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        sf = float(f32(scale_factor))  # double member initialised from a float
        self.scale = [f32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(f32(float(self.scale[-1]) * sf))
        self.inv_scale = [f32(1.0) / s for s in self.scale]
        factor = f32(1.0 / sf)
        nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
        self.per_level = []
        s = 0
        for _ in range(nlevels - 1):
            self.per_level.append(cv_round(nd))
            s += self.per_level[-1]
            nd = f32(nd * factor)
        self.per_level.append(max(nfeatures - s, 0))
        self.umax = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]

    # ---- pyramid (:1344-1367) ----
    def pyramid(self, gray):
        h, w = gray.shape
        pyr = [gray.copy()]
        for l in range(1, self.nlevels):
            s = self.inv_scale[l]
            sz = (cv_round(f32(w) * s), cv_round(f32(h) * s))
            pyr.append(cv2.resize(pyr[-1], sz, interpolation=cv2.INTER_LINEAR))
        return pyr

    # ---- cell loop (:793-850) ----
    def cells(self, im, th_ini, th_min):
        h, w = im.shape
        minBX = minBY = EDGE - 3
        maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
        width, height = f32(maxBX - minBX), f32(maxBY - minBY)
        nCols, nRows = int(width / f32(30)), int(height / f32(30))
        wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
        det_ini = cv2.FastFeatureDetector_create(th_ini, True)
        det_min = cv2.FastFeatureDetector_create(th_min, True)
        out = []
        for i in range(nRows):
            iniY = minBY + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                roi = im[iniY:maxY, iniX:maxX]
                kps = det_ini.detect(roi)
                if len(kps) == 0:
                    kps = det_min.detect(roi)
                for kp in kps:
                    out.append((f32(kp.pt[0]) + f32(j * wCell), f32(kp.pt[1]) + f32(i * hCell), f32(kp.response)))
        return out

    # ---- octree (:489-769), same documented tie-break as oracle B ----
    @staticmethod
    def octree(cands, minX, maxX, minY, maxY, N):
        if not cands:
            return []
        nIni = int(math.floor(float(f32(maxX - minX) / f32(maxY - minY)) + 0.5))  # C round(), positive arg
        hX = f32(maxX - minX) / f32(nIni)
        seq = [0]

        def mk(x0, x1, y0, y1, keys):
            n = dict(x0=x0, x1=x1, y0=y0, y1=y1, keys=keys, nomore=(len(keys) == 1), seq=seq[0])
            seq[0] += 1
            return n

        nodes = []  # python list used as the std::list; index 0 is the front
        roots = [mk(int(hX * f32(i)), int(hX * f32(i + 1)), 0, maxY - minY, []) for i in range(nIni)]
        for k, c in enumerate(cands):
            roots[int(f32(c[0]) / hX)]["keys"].append(k)
        for r in roots:
            if len(r["keys"]) == 1:
                r["nomore"] = True
            if r["keys"]:
                nodes.append(r)

        def divide(p):
            halfX = int(math.ceil(f32(p["x1"] - p["x0"]) / 2))
            halfY = int(math.ceil(f32(p["y1"] - p["y0"]) / 2))
            xm, ym = p["x0"] + halfX, p["y0"] + halfY
            ks = [[], [], [], []]
            for k in p["keys"]:
                x, y = cands[k][0], cands[k][1]
                if x < xm:
                    ks[0 if y < ym else 2].append(k)
                else:
                    ks[1 if y < ym else 3].append(k)
            geo = [(p["x0"], xm, p["y0"], ym), (xm, p["x1"], p["y0"], ym), (p["x0"], xm, ym, p["y1"]),
                   (xm, p["x1"], ym, p["y1"])]
            return [(g, k) for g, k in zip(geo, ks)]

        finish = False
        while not finish:
            prev = len(nodes)
            expandable = []
            nToExpand = 0
            old = list(nodes)
            for nd in old:
                if nd["nomore"]:
                    continue
                for g, k in divide(nd):
                    if k:
                        c = mk(g[0], g[1], g[2], g[3], k)
                        nodes.insert(0, c)
                        if len(k) > 1:
                            nToExpand += 1
                            expandable.append(c)
                nodes.remove(nd)
            if len(nodes) >= N or len(nodes) == prev:
                finish = True
            elif len(nodes) + nToExpand * 3 > N:
                while not finish:
                    prev = len(nodes)
                    pv = sorted(expandable, key=lambda n: (len(n["keys"]), n["seq"]))
                    expandable = []
                    for nd in reversed(pv):
                        for g, k in divide(nd):
                            if k:
                                c = mk(g[0], g[1], g[2], g[3], k)
                                nodes.insert(0, c)
                                if len(k) > 1:
                                    expandable.append(c)
                        nodes.remove(nd)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev:
                        finish = True
        res = []
        for nd in nodes:
            best = nd["keys"][0]
            for k in nd["keys"][1:]:
                if cands[k][2] > cands[best][2]:
                    best = k
            res.append(best)
        return res

    # ---- IC_Angle (:80-107) ----
    def ic_angle(self, im, x, y):
        cx = cv_round(x)
        cy = cv_round(y)
        m01 = m10 = 0
        row = im[cy].astype(np.int64)
        for u in range(-HALF, HALF + 1):
            m10 += u * int(row[cx + u])
        for v in range(1, HALF + 1):
            d = self.umax[v]
            plus = im[cy + v, cx - d:cx + d + 1].astype(np.int64)
            minus = im[cy - v, cx - d:cx + d + 1].astype(np.int64)
            us = np.arange(-d, d + 1)
            m01 += v * int((plus - minus).sum())
            m10 += int((us * (plus + minus)).sum())
        return f32(cv2.fastAtan2(float(m01), float(m10)))

    # ---- descriptor (:109-156) ----
    @staticmethod
    def descriptor(blurred, x, y, angle_deg, pattern):
        factorPI = f32(np.pi / 180.0)
        ang = f32(angle_deg) * factorPI
        a = f32(math.cos(float(ang)))
        b = f32(math.sin(float(ang)))
        cx, cy = cv_round(x), cv_round(y)
        px = pattern[:, 0].astype(np.float32)
        py = pattern[:, 1].astype(np.float32)
        yy = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
        xx = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
        vals = blurred[cy + yy, cx + xx].astype(np.int64)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)

    # ---- operator() (:1088-1342) ----
    def classify(self, w, h, boxes, tm, blur):
        rects, area = [], f32(0)
        for b, bx in enumerate(boxes):
            xmin, ymin, xmax, ymax = (f32(v) for v in bx)
            rx, ry, rw, rh = int(xmin), int(ymin), int(xmax - xmin), int(ymax - ymin)
            area_box = f32((xmax - xmin) * (ymax - ymin))
            cnt, mark = 0, False
            for t in tm:
                tx, ty = int(t[0]), int(t[1])
                if rx <= tx < rx + rw and ry <= ty < ry + rh:
                    cnt += 1
                if f32(cnt * 10000) > area_box:
                    mark = True
                    break
            if not mark and b < len(blur) and blur[b] == 1 and cnt > 0:
                mark = True
            if mark:
                area = f32(area + area_box)
                rects.append((int(xmin), int(ymin), int(xmax), int(ymax)))
        return rects, bool(area > 200000)

    def moving(self, rects, x, y, level, w0, h0):
        s = self.scale[level] if level else f32(1)
        sx, sy = f32(x) * s, f32(y) * s
        if sx >= w0 - 1:
            sx = f32(w0 - 1)
        if sy >= h0 - 1:
            sy = f32(h0 - 1)
        ix, iy = int(sx), int(sy)
        return any(r[0] <= ix < r[2] and r[1] <= iy < r[3] for r in rects)

    def extract(self, gray, boxes=(), tm=(), blur=(), pattern=None, want_stages=False):
        h0, w0 = gray.shape
        rects, area_flag = self.classify(w0, h0, boxes, tm, blur)
        pyr = self.pyramid(gray)
        th_ini, th_min = (30, 10) if area_flag else (20, 7)
        stages = dict(pyramid=pyr, candidates=[], level_keys=[], blurred=[])
        all_keys = []
        for l in range(self.nlevels):
            im = pyr[l]
            h, w = im.shape
            cands = self.cells(im, th_ini, th_min)
            if area_flag:
                cands = [c for c in cands if not self.moving(rects, c[0], c[1], l, w0, h0)]
            stages["candidates"].append(cands)
            N = self.per_level[l]
            if area_flag:
                N = int(int(N) * 0.7)
            sel = self.octree(cands, 16, w - 16, 16, h - 16, N)
            size = f32(int(f32(PATCH) * self.scale[l]))
            keys = []
            for k in sel:
                x, y = f32(cands[k][0] + f32(16)), f32(cands[k][1] + f32(16))
                keys.append([x, y, size, f32(-1), cands[k][2], l])
            all_keys.append(keys)
        for l in range(self.nlevels):
            for k in all_keys[l]:
                k[3] = self.ic_angle(pyr[l], k[0], k[1])
        if not area_flag:
            for l in range(min(self.nlevels, 8)):
                all_keys[l] = [k for k in all_keys[l] if not self.moving(rects, k[0], k[1], l, w0, h0)]
        stages["level_keys"] = all_keys
        out_k, out_d = [], []
        for l in range(self.nlevels):
            if not all_keys[l]:
                stages["blurred"].append(None)
                continue
            bl = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            stages["blurred"].append(bl)
            for k in all_keys[l]:
                out_d.append(self.descriptor(bl, k[0], k[1], k[3], pattern))
                s = self.scale[l]
                x, y = (k[0] * s, k[1] * s) if l else (k[0], k[1])
                out_k.append((f32(x), f32(y), k[2], k[3], k[4], l, -1))
        desc = np.array(out_d, np.uint8).reshape(-1, 32)
        if want_stages:
            return out_k, desc, stages, dict(rects=rects, area_flag=area_flag)
        return out_k, desc
