// coeb_oracle_frame.hpp -- CPU restatement of the per-point steps either side of the extractor and the
// matchers: the tail of the RGB-D Frame constructor (UndistortKeyPoints, ComputeStereoFromRGBD) and the
// visibility test in front of SearchByProjection (Frame::isInFrustum + MapPoint::PredictScale, as driven by
// Tracking::SearchLocalPoints). TEST INFRASTRUCTURE ONLY (see coeb_oracle.hpp for the rules and for the
// parity status: pinned to the reference itself through oracle/_ref).
//
// OpenCV arithmetic on this path, each model checked against cv2 4.13 in tests/test_oracle_vs_cv2.py:
//   cv::undistortPoints(src, dst, K, D, noArray(), K): double precision, 5 fixed-point iterations, no EPS test;
//   Mat::convertTo(CV_32F, factor) of a CV_16U depth map: one fp32 multiply per pixel;
//   mRcw * P + mtcw (cv::gemm on 3x3 . 3x1 CV_32F): fp32, left to right, no FMA ("projection arithmetic");
//   cv::norm(3x1 CV_32F): squares accumulated in double, sqrt in double, then rounded to float;
//   Mat::dot(3x1 CV_32F): products accumulated in double in index order.
#pragma once
#include "coeb_oracle_match.hpp"

namespace orc {

// cv::undistortPoints for one point (OpenCV 4.13 modules/calib3d/src/undistort.dispatch.cpp, cvUndistortPointsInternal
// with R = I, P = K, criteria = (COUNT, 5)); dist = {k1, k2, p1, p2, k3}.
static inline void undistort_point(float u, float v, float fx, float fy, float cx, float cy, const float* dist,
                                   float* xo, float* yo) {
    const double dfx = fx, dfy = fy, dcx = cx, dcy = cy;
    const double k0 = dist[0], k1 = dist[1], k2 = dist[2], k3 = dist[3], k4 = dist[4];
    const double ifx = 1. / dfx, ify = 1. / dfy;
    double x = ((double)u - dcx) * ifx, y = ((double)v - dcy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        // k[5..7] (the rational model), k[8..11] (thin prism) are zero: the numerator is exactly 1 and the extra delta
        // terms add +0, which cannot change a finite sum
        const double icdist = 1. / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x);
        const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    // xx = RR00 x + RR01 y + RR02 with RR = K, ww = 1 / (0 x + 0 y + 1) = 1
    *xo = (float)(dfx * x + dcx);
    *yo = (float)(dfy * y + dcy);
}

// Frame::UndistortKeyPoints (src/Frame.cc:579-609): mvKeysUn = mvKeys with pt replaced; a zero k1 means "no distortion".
static inline void undistort_keypoints(const coeb_keypoint* keys, int n, const coeb_camera& cam, const float* dist,
                                       coeb_keypoint* keys_un) {
    for (int i = 0; i < n; i++) {
        keys_un[i] = keys[i];
        if (dist && dist[0] != 0.0f) undistort_point(keys[i].x, keys[i].y, cam.fx, cam.fy, cam.cx, cam.cy, dist, &keys_un[i].x, &keys_un[i].y);
    }
}

// Depth image as Tracking::GrabImageRGBD hands it to the Frame (src/Tracking.cc:226-229): CV_32F, or a raw CV_16U map
// that convertTo scales by mDepthMapFactor (= 1 / DepthMapFactor of the settings file).
struct DepthView {
    const void* data = nullptr;
    int kind = 0;             // 0 none, 1 float32, 2 uint16
    int stride_bytes = 0;
    float factor = 1.f;       // applied to kind 2 only (a float map whose factor is not 1 is also converted by the reference:
                              // callers pass the already scaled map)
    float at(int row, int col) const {
        const uint8_t* p = (const uint8_t*)data + (size_t)row * stride_bytes;
        if (kind == 1) return ((const float*)p)[col];
        return (float)((const uint16_t*)p)[col] * factor;
    }
};

// Frame::ComputeStereoFromRGBD (src/Frame.cc:820-842). imDepth.at<float>(v, u) truncates the float coordinates.
static inline void stereo_from_rgbd(const coeb_keypoint* keys, const coeb_keypoint* keys_un, int n, const DepthView& D, float mbf,
                                    float* uright, float* depth) {
    for (int i = 0; i < n; i++) {
        uright[i] = -1.f;
        depth[i] = -1.f;
        if (!D.kind) continue;
        const float d = D.at((int)keys[i].y, (int)keys[i].x);
        if (d > 0) {
            depth[i] = d;
            uright[i] = keys_un[i].x - mbf / d;
        }
    }
}

// The local map as Tracking::SearchLocalPoints sees it, flattened.
struct LocalMapSoA {
    int n = 0;
    const float* xyz = nullptr;       // GetWorldPos(), n x 3
    const float* normal = nullptr;    // GetNormal(), n x 3
    const float* min_dist = nullptr;  // mfMinDistance (GetMinDistanceInvariance() = 0.8f * it, src/MapPoint.cc:373-377)
    const float* max_dist = nullptr;  // mfMaxDistance (GetMaxDistanceInvariance() = 1.2f * it, :379-383)
    const uint8_t* desc = nullptr;    // GetDescriptor(), n x 32
};

// glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c: 16-entry table of {1/c, log c}, degree-3 polynomial, all in double),
// restated for normal positive x so that MapPoint::PredictScale's float `log(ratio)` (src/MapPoint.cc:410; the float
// overload, <math.h> being included through opencv/cv.h) is the same float wherever it is evaluated. Pinned against
// the C library's logf by orc_logf_mismatches() (tests/test_oracle_golden.py); -ffp-contract=off keeps z*invc-1 unfused.
static inline float glibc_logf(float x) {
    static const double T[16][2] = {
        {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
        {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
        {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
        {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
        {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
        {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
        {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
        {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
    uint32_t ix;
    std::memcpy(&ix, &x, 4);
    if (ix == 0x3f800000u) return 0.f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    float zf;
    std::memcpy(&zf, &iz, 4);
    const double z = zf, r = z * T[i][0] - 1, y0 = T[i][1] + (double)k * 0x1.62e42fefa39efp-1, r2 = r * r;
    double y = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
    y = -0x1.00ea348b88334p-2 * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}

struct FrustumOut {   // the MapPoint fields isInFrustum writes (src/Frame.cc:492-498)
    uint8_t in_view;
    float proj_x, proj_y, proj_xr, view_cos;
    int level;
};

// Frame::isInFrustum (src/Frame.cc:445-501) for map point i. Tcw = [Rcw | tcw] 3x4 row-major, Ow = mOw.
static inline FrustumOut is_in_frustum(const LocalMapSoA& M, int i, const float* Tcw, const float* Ow, const coeb_camera& cam,
                                       float viewingCosLimit, float logScaleFactor, int nScaleLevels) {
    FrustumOut o{};
    o.in_view = 0;
    const float* P = M.xyz + 3 * (size_t)i;
    const float PcX = Tcw[0] * P[0] + Tcw[1] * P[1] + Tcw[2] * P[2] + Tcw[3];
    const float PcY = Tcw[4] * P[0] + Tcw[5] * P[1] + Tcw[6] * P[2] + Tcw[7];
    const float PcZ = Tcw[8] * P[0] + Tcw[9] * P[1] + Tcw[10] * P[2] + Tcw[11];
    if (PcZ < 0.0f) return o;
    const float invz = 1.0f / PcZ;
    const float u = cam.fx * PcX * invz + cam.cx;
    const float v = cam.fy * PcY * invz + cam.cy;
    if (u < cam.min_x || u > cam.max_x) return o;
    if (v < cam.min_y || v > cam.max_y) return o;
    const float maxDistance = 1.2f * M.max_dist[i], minDistance = 0.8f * M.min_dist[i];
    const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
    double s = 0;
    for (int k = 0; k < 3; k++) s += (double)PO[k] * (double)PO[k];
    const float dist = (float)std::sqrt(s);
    if (dist < minDistance || dist > maxDistance) return o;
    const float* Pn = M.normal + 3 * (size_t)i;
    double dot = 0;
    for (int k = 0; k < 3; k++) dot += (double)PO[k] * (double)Pn[k];
    const float viewCos = (float)(dot / (double)dist);
    if (viewCos < viewingCosLimit) return o;
    // MapPoint::PredictScale (src/MapPoint.cc:402-417): float log and float ceil
    const float ratio = M.max_dist[i] / dist;
    int nScale = 0;   // a zero, subnormal, infinite or NaN ratio (dist or mfMaxDistance degenerate) is undefined in the reference
    if (ratio >= FLT_MIN && ratio <= FLT_MAX) nScale = (int)std::ceil(glibc_logf(ratio) / logScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= nScaleLevels) nScale = nScaleLevels - 1;
    o.in_view = 1;
    o.proj_x = u;
    o.proj_xr = u - cam.bf * invz;
    o.proj_y = v;
    o.level = nScale;
    o.view_cos = viewCos;
    return o;
}

// Tracking::SearchLocalPoints (src/Tracking.cc:1222-1272), second loop and the matcher call: map points with skip != 0
// (mnLastFrameSeen == current frame, or isBad()) are not projected; the others go through isInFrustum(pMP, 0.5) and
// SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th). For the matcher a skipped point has mbTrackInView == false
// (reset by the first loop for the points of this frame; a bad point is skipped by the matcher itself, :57-58).
// in_view (n bytes) is what IncreaseVisible()/nToMatch count; proj (optional, n x 5: u, v, ur, viewCos, level) is for tests.
static inline int search_local_points(const FrameView& F, const LocalMapSoA& M, const uint8_t* skip, const uint8_t* has_obs,
                                      const float* Tcw, const float* Ow, float viewingCosLimit, float th, float nnratio,
                                      int* kp_match, uint8_t* in_view, float* proj) {
    std::vector<uint8_t> tiv(M.n), bad(M.n, 0);
    std::vector<float> px(M.n), py(M.n), pxr(M.n), vc(M.n);
    std::vector<int> lvl(M.n);
    const float logScale = F.nlevels > 1 ? glibc_logf(F.scale[1]) : 1.f;   // mfLogScaleFactor = log(mfScaleFactor) (src/Frame.cc:151)
    int nToMatch = 0;
    for (int i = 0; i < M.n; i++) {
        FrustumOut o{};
        if (!skip[i]) o = is_in_frustum(M, i, Tcw, Ow, F.cam, viewingCosLimit, logScale, F.nlevels);
        tiv[i] = o.in_view; px[i] = o.proj_x; py[i] = o.proj_y; pxr[i] = o.proj_xr; vc[i] = o.view_cos; lvl[i] = o.level;
        if (in_view) in_view[i] = o.in_view;
        if (proj) { proj[5 * i] = o.proj_x; proj[5 * i + 1] = o.proj_y; proj[5 * i + 2] = o.proj_xr; proj[5 * i + 3] = o.view_cos; proj[5 * i + 4] = (float)o.level; }
        nToMatch += o.in_view;
    }
    if (nToMatch == 0) return 0;
    MapPointsSoA mp;
    mp.n = M.n; mp.track_in_view = tiv.data(); mp.bad = bad.data(); mp.has_obs = has_obs;
    mp.proj_x = px.data(); mp.proj_y = py.data(); mp.proj_xr = pxr.data(); mp.level = lvl.data(); mp.view_cos = vc.data();
    mp.desc = M.desc;
    return search_by_projection_map(F, mp, th, nnratio, kp_match);
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
// (src/ORBmatcher.cc:1473-1600; Tracking::Relocalization). Per keyframe map point i: valid = pMP && !isBad() &&
// !sAlreadyFound.count(pMP); xyz = GetWorldPos(); min/max_dist = mfMinDistance / mfMaxDistance; angle = pKF->mvKeysUn[i].angle;
// desc = GetDescriptor(). Tcw = CurrentFrame.mTcw (3x4), Ow = -Rcw^T tcw as the reference evaluates it (:1479).
// kp_match encodes CurrentFrame.mvpMapPoints: -1 empty, any other negative value = holds a MapPoint (every non-null entry blocks,
// :1546-1547); on return >= 0 is the keyframe map point assigned by this call. Note that, unlike the frame-to-frame overload,
// the reference does not reject points behind the camera here (no invzc < 0 test).
static inline int search_by_projection_reloc(const FrameView& C, int n, const uint8_t* valid, const float* xyz, const float* min_dist,
                                             const float* max_dist, const float* angle, const uint8_t* desc, const float* Tcw,
                                             const float* Ow, float th, int ORBdist, bool checkOri, int* kp_match) {
    int nmatches = 0;
    std::vector<int> rotHist[COEB_HISTO_LENGTH];
    const float logScale = C.nlevels > 1 ? glibc_logf(C.scale[1]) : 1.f;
    std::vector<int> vIndices2;
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;
        const float* P = xyz + 3 * (size_t)i;
        const float xc = Tcw[0] * P[0] + Tcw[1] * P[1] + Tcw[2] * P[2] + Tcw[3];
        const float yc = Tcw[4] * P[0] + Tcw[5] * P[1] + Tcw[6] * P[2] + Tcw[7];
        const float zc = Tcw[8] * P[0] + Tcw[9] * P[1] + Tcw[10] * P[2] + Tcw[11];
        const float invzc = (float)(1.0 / (double)zc);
        const float u = C.cam.fx * xc * invzc + C.cam.cx;
        const float v = C.cam.fy * yc * invzc + C.cam.cy;
        if (u < C.cam.min_x || u > C.cam.max_x) continue;
        if (v < C.cam.min_y || v > C.cam.max_y) continue;
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)PO[k] * (double)PO[k];
        const float dist3D = (float)std::sqrt(s);
        const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float ratio = max_dist[i] / dist3D;   // MapPoint::PredictScale (src/MapPoint.cc:402-417)
        int nPredictedLevel = 0;
        if (ratio >= FLT_MIN && ratio <= FLT_MAX) nPredictedLevel = (int)std::ceil(glibc_logf(ratio) / logScale);
        if (nPredictedLevel < 0) nPredictedLevel = 0;
        else if (nPredictedLevel >= C.nlevels) nPredictedLevel = C.nlevels - 1;
        const float radius = th * C.scale[nPredictedLevel];
        C.features_in_area(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (kp_match[i2] != KP_FREE) continue;
            const int dist = hamming256(dMP, C.desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= ORBdist) {
            kp_match[bestIdx2] = i;
            nmatches++;
            if (checkOri) rotHist[rot_bin(angle[i], C.kps[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int sizes[COEB_HISTO_LENGTH], ind1, ind2, ind3;
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, COEB_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int k : rotHist[i]) { kp_match[k] = KP_FREE; nmatches--; }
        }
    }
    return nmatches;
}

// The search half of ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th) (src/ORBmatcher.cc:826-961;
// LocalMapping::SearchInNeighbors): for every map point the keypoint of the keyframe it would be fused into (bestIdx with
// bestDist <= TH_LOW), or -1. The pointer-graph half (Replace / AddObservation / AddMapPoint, :938-957) stays with the caller and is
// applied in list order from these indices: the search itself reads nothing those calls change.
// valid[i] = pMP && !pMP->isBad() && !pMP->IsInKeyFrame(pKF). F is the keyframe (mvKeysUn, mvuRight, bounds, scale factors).
// chi2_tests = false gives the search of Fuse(KeyFrame* pKF, cv::Mat Scw, vpPoints, th, vpReplacePoint) (:963-1093): same projection
// and filters with [Rcw | tcw] and Ow taken from the decomposed Sim3 (:973-977), valid = !isBad() && !spAlreadyFound.count(pMP), and no
// reprojection-error test.
static inline int fuse_search(const FrameView& F, const LocalMapSoA& M, const uint8_t* valid, const float* Tcw, const float* Ow, float th,
                              bool chi2_tests, int* best_idx) {
    int nFused = 0;
    const float logScale = F.nlevels > 1 ? glibc_logf(F.scale[1]) : 1.f;
    std::vector<int> vIndices;
    for (int i = 0; i < M.n; i++) {
        best_idx[i] = -1;
        if (!valid[i]) continue;
        const float* P = M.xyz + 3 * (size_t)i;
        const float X = Tcw[0] * P[0] + Tcw[1] * P[1] + Tcw[2] * P[2] + Tcw[3];
        const float Y = Tcw[4] * P[0] + Tcw[5] * P[1] + Tcw[6] * P[2] + Tcw[7];
        const float Z = Tcw[8] * P[0] + Tcw[9] * P[1] + Tcw[10] * P[2] + Tcw[11];
        if (Z < 0.0f) continue;
        const float invz = 1 / Z;
        const float x = X * invz, y = Y * invz;
        const float u = F.cam.fx * x + F.cam.cx, v = F.cam.fy * y + F.cam.cy;
        if (!(u >= F.cam.min_x && u < F.cam.max_x && v >= F.cam.min_y && v < F.cam.max_y)) continue;   // KeyFrame::IsInImage
        const float ur = u - F.cam.bf * invz;
        const float maxDistance = 1.2f * M.max_dist[i], minDistance = 0.8f * M.min_dist[i];
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)PO[k] * (double)PO[k];
        const float dist3D = (float)std::sqrt(s);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const float* Pn = M.normal + 3 * (size_t)i;
        double dot = 0;
        for (int k = 0; k < 3; k++) dot += (double)PO[k] * (double)Pn[k];
        if (dot < 0.5 * (double)dist3D) continue;   // viewing angle below 60 degrees (:884)
        const float ratio = M.max_dist[i] / dist3D;
        int nPredictedLevel = 0;
        if (ratio >= FLT_MIN && ratio <= FLT_MAX) nPredictedLevel = (int)std::ceil(glibc_logf(ratio) / logScale);
        if (nPredictedLevel < 0) nPredictedLevel = 0;
        else if (nPredictedLevel >= F.nlevels) nPredictedLevel = F.nlevels - 1;
        const float radius = th * F.scale[nPredictedLevel];
        F.features_in_area(u, v, radius, -1, -1, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = M.desc + (size_t)i * 32;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            const coeb_keypoint& kp = F.kps[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const float sc = F.scale[kpLevel];
            const float invSigma2 = 1.0f / (sc * sc);   // mvInvLevelSigma2 (src/ORBextractor.cc:427-428, 434-435)
            const float ex = u - kp.x, ey = v - kp.y;
            if (!chi2_tests) {
                // Fuse(pKF, Scw, ...) of loop closing (:963-1093) has no reprojection-error test
            } else if (F.uright && F.uright[idx] >= 0) {
                const float er = ur - F.uright[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if ((double)(e2 * invSigma2) > 7.8) continue;
            } else {
                const float e2 = ex * ex + ey * ey;
                if ((double)(e2 * invSigma2) > 5.99) continue;
            }
            const int dist = hamming256(dMP, F.desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= COEB_TH_LOW) { best_idx[i] = bestIdx; nFused++; }
    }
    return nFused;
}

}  // namespace orc
