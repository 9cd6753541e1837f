// coeb_oracle_match.hpp -- CPU restatement of the keypoint grid, the Hamming matchers and the stereo
// matcher of the reference. TEST INFRASTRUCTURE ONLY (see coeb_oracle.hpp for the rules and for the
// parity status: pinned to the reference itself through oracle/_ref). Pointer-graph state (MapPoint*, Frame members) is flattened into the
// structure-of-arrays inputs that the C ABI (include/coeb_frontend.h) also takes.
#pragma once
#include "coeb_oracle.hpp"

namespace orc {

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1648-1664): popcount(a ^ b) over 8 x u32.
static inline int hamming256(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4);
        std::memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// The part of ORB_SLAM2::Frame the matchers read: undistorted keypoints, descriptors, uRight,
// image bounds, scale factors and the 64x48 keypoint grid.
struct FrameView {
    int n = 0;
    const coeb_keypoint* kps = nullptr;  // mvKeysUn
    const uint8_t* desc = nullptr;       // mDescriptors, n x 32
    const float* uright = nullptr;       // mvuRight (nullptr: all -1)
    coeb_camera cam{};
    const float* scale = nullptr;        // mvScaleFactors
    int nlevels = 0;
    float gw_inv = 0, gh_inv = 0;        // mfGridElementWidthInv / HeightInv (Frame.cc:233-234)
    std::vector<int> grid[COEB_GRID_COLS][COEB_GRID_ROWS];

    // Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:396-411, 558-568)
    void build_grid() {
        gw_inv = (float)COEB_GRID_COLS / (cam.max_x - cam.min_x);
        gh_inv = (float)COEB_GRID_ROWS / (cam.max_y - cam.min_y);
        for (auto& col : grid)
            for (auto& cell : col) cell.clear();
        for (int i = 0; i < n; i++) {
            int px = (int)std::round((kps[i].x - cam.min_x) * gw_inv);
            int py = (int)std::round((kps[i].y - cam.min_y) * gh_inv);
            if (px < 0 || px >= COEB_GRID_COLS || py < 0 || py >= COEB_GRID_ROWS) continue;
            grid[px][py].push_back(i);
        }
    }

    // Frame::GetFeaturesInArea (src/Frame.cc:503-556). Order: ix outer, iy inner, insertion order.
    void features_in_area(float x, float y, float r, int minLevel, int maxLevel,
                          std::vector<int>& out) const {
        out.clear();
        const int nMinCellX = std::max(0, (int)std::floor((x - cam.min_x - r) * gw_inv));
        if (nMinCellX >= COEB_GRID_COLS) return;
        const int nMaxCellX = std::min(COEB_GRID_COLS - 1, (int)std::ceil((x - cam.min_x + r) * gw_inv));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)std::floor((y - cam.min_y - r) * gh_inv));
        if (nMinCellY >= COEB_GRID_ROWS) return;
        const int nMaxCellY = std::min(COEB_GRID_ROWS - 1, (int)std::ceil((y - cam.min_y + r) * gh_inv));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (int idx : grid[ix][iy]) {
                    const coeb_keypoint& kp = kps[idx];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float dx = kp.x - x, dy = kp.y - y;
                    if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(idx);
                }
    }
};

// Flattened MapPoint fields read by SearchByProjection(Frame&, vector<MapPoint*>&, th)
// (set by Frame::isInFrustum, src/Frame.cc:492-498).
struct MapPointsSoA {
    int n = 0;
    const uint8_t* track_in_view = nullptr;  // mbTrackInView
    const uint8_t* bad = nullptr;            // isBad()
    const uint8_t* has_obs = nullptr;        // Observations() > 0
    const float* proj_x = nullptr;           // mTrackProjX
    const float* proj_y = nullptr;           // mTrackProjY
    const float* proj_xr = nullptr;          // mTrackProjXR
    const int* level = nullptr;              // mnTrackScaleLevel
    const float* view_cos = nullptr;         // mTrackViewCos
    const uint8_t* desc = nullptr;           // GetDescriptor(), n x 32
};

// kp_match[] encodes Frame::mvpMapPoints for the call:
//   -1 empty; -2 holds a MapPoint (not from this call) with Observations()>0; -3 holds one with
//   Observations()==0; >=0 index of the map point assigned during this call.
enum { KP_FREE = -1, KP_TAKEN = -2, KP_TAKEN_NOOBS = -3 };

// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1602-1643)
static inline void three_maxima(const int* histo_size, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo_size[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

static inline int rot_bin(float a1, float a2) {  // src/ORBmatcher.cc:1434-1439
    const float factor = 1.0f / COEB_HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == COEB_HISTO_LENGTH) bin = 0;
    return bin;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-137)
static inline int search_by_projection_map(const FrameView& F, const MapPointsSoA& mp, float th,
                                           float nnratio, int* kp_match) {
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < mp.n; iMP++) {
        if (!mp.track_in_view[iMP]) continue;
        if (mp.bad[iMP]) continue;
        const int nPredictedLevel = mp.level[iMP];
        float r = ((double)mp.view_cos[iMP] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos (:131-137)
        if (bFactor) r *= th;
        F.features_in_area(mp.proj_x[iMP], mp.proj_y[iMP], r * F.scale[nPredictedLevel],
                           nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* MPdescriptor = mp.desc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            const int cur = kp_match[idx];
            if (cur == KP_TAKEN || (cur >= 0 && mp.has_obs[cur])) continue;
            if (F.uright && F.uright[idx] > 0) {
                const float er = std::fabs(mp.proj_xr[iMP] - F.uright[idx]);
                if (er > r * F.scale[nPredictedLevel]) continue;
            }
            const int dist = hamming256(MPdescriptor, F.desc + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F.kps[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F.kps[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= COEB_TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            kp_match[bestIdx] = iMP;
            nmatches++;
        }
    }
    return nmatches;
}

// Inputs of SearchByProjection(Frame& cur, const Frame& last, th, bMono): the last frame's tracked
// MapPoints flattened per last-frame keypoint index.
struct LastFrameSoA {
    int n = 0;
    const uint8_t* valid = nullptr;    // mvpMapPoints[i] != NULL && !mvbOutlier[i]
    const uint8_t* has_obs = nullptr;  // pMP->Observations() > 0
    const float* xyz = nullptr;        // pMP->GetWorldPos(), n x 3
    const int* octave = nullptr;       // LastFrame.mvKeys[i].octave
    const float* angle = nullptr;      // LastFrame.mvKeysUn[i].angle
    const uint8_t* desc = nullptr;     // pMP->GetDescriptor(), n x 32
};

// tlc.z of (src/ORBmatcher.cc:1339-1350): twc = -Rcw^T tcw; tlc = Rlw twc + tlw. fp32, fixed order.
static inline float tlc_z(const float* Tcw_cur, const float* Tcw_last) {
    // T = [r00 r01 r02 tx; r10 r11 r12 ty; r20 r21 r22 tz] row-major 3x4
    float twc[3];
    for (int c = 0; c < 3; c++)
        twc[c] = -(Tcw_cur[0 * 4 + c] * Tcw_cur[3] + Tcw_cur[1 * 4 + c] * Tcw_cur[7] + Tcw_cur[2 * 4 + c] * Tcw_cur[11]);
    return Tcw_last[8] * twc[0] + Tcw_last[9] * twc[1] + Tcw_last[10] * twc[2] + Tcw_last[11];
}

// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (src/ORBmatcher.cc:1329-1471).
// The 3x3 projection is evaluated in fp32, left to right, without FMA (the reference goes through
// cv::Mat expression templates; see DESIGN.md "projection arithmetic").
static inline int search_by_projection_last(const FrameView& C, const LastFrameSoA& L,
                                            const float* Tcw_cur, const float* Tcw_last, float th,
                                            bool bMono, bool checkOri, int* kp_match) {
    int nmatches = 0;
    std::vector<int> rotHist[COEB_HISTO_LENGTH];
    const float tz = tlc_z(Tcw_cur, Tcw_last);
    const bool bForward = tz > C.cam.b && !bMono;
    const bool bBackward = -tz > C.cam.b && !bMono;
    std::vector<int> vIndices2;
    for (int i = 0; i < L.n; i++) {
        if (!L.valid[i]) continue;
        const float X = L.xyz[3 * i], Y = L.xyz[3 * i + 1], Z = L.xyz[3 * i + 2];
        const float xc = Tcw_cur[0] * X + Tcw_cur[1] * Y + Tcw_cur[2] * Z + Tcw_cur[3];
        const float yc = Tcw_cur[4] * X + Tcw_cur[5] * Y + Tcw_cur[6] * Z + Tcw_cur[7];
        const float zc = Tcw_cur[8] * X + Tcw_cur[9] * Y + Tcw_cur[10] * Z + Tcw_cur[11];
        const float invzc = (float)(1.0 / zc);
        if (invzc < 0) continue;
        const float u = C.cam.fx * xc * invzc + C.cam.cx;
        const float v = C.cam.fy * yc * invzc + C.cam.cy;
        if (u < C.cam.min_x || u > C.cam.max_x) continue;
        if (v < C.cam.min_y || v > C.cam.max_y) continue;
        const int nLastOctave = L.octave[i];
        const float radius = th * C.scale[nLastOctave];
        if (bForward) C.features_in_area(u, v, radius, nLastOctave, -1, vIndices2);
        else if (bBackward) C.features_in_area(u, v, radius, 0, nLastOctave, vIndices2);
        else C.features_in_area(u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = L.desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int cur = kp_match[i2];
            if (cur == KP_TAKEN || (cur >= 0 && L.has_obs[cur])) continue;
            if (C.uright && C.uright[i2] > 0) {
                const float ur = u - C.cam.bf * invzc;
                const float er = std::fabs(ur - C.uright[i2]);
                if (er > radius) continue;
            }
            const int dist = hamming256(dMP, C.desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= COEB_TH_HIGH) {
            kp_match[bestIdx2] = i;
            nmatches++;
            if (checkOri) rotHist[rot_bin(L.angle[i], C.kps[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int sizes[COEB_HISTO_LENGTH], ind1, ind2, ind3;
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, COEB_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < COEB_HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { kp_match[idx] = KP_FREE; nmatches--; }
    }
    return nmatches;
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:405-520)
static inline int search_for_initialization(const FrameView& F1, const FrameView& F2,
                                            float* prev_matched /* n1 x 2, in/out */,
                                            int* matches12 /* n1, out */, int windowSize,
                                            float nnratio, bool checkOri) {
    int nmatches = 0;
    for (int i = 0; i < F1.n; i++) matches12[i] = -1;
    std::vector<int> rotHist[COEB_HISTO_LENGTH];
    std::vector<int> vMatchedDistance(F2.n, INT_MAX), vnMatches21(F2.n, -1);
    std::vector<int> vIndices2;
    for (int i1 = 0; i1 < F1.n; i1++) {
        const int level1 = F1.kps[i1].octave;
        if (level1 > 0) continue;
        F2.features_in_area(prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)windowSize, level1,
                            level1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* d1 = F1.desc + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = hamming256(d1, F2.desc + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= COEB_TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) {
                    matches12[vnMatches21[bestIdx2]] = -1;
                    nmatches--;
                }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) rotHist[rot_bin(F1.kps[i1].angle, F2.kps[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int sizes[COEB_HISTO_LENGTH], ind1, ind2, ind3;
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, COEB_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < F1.n; i1++)
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = F2.kps[matches12[i1]].x;
            prev_matched[2 * i1 + 1] = F2.kps[matches12[i1]].y;
        }
    return nmatches;
}

// Frame::ComputeStereoMatches (src/Frame.cc:644-818). pyrL/pyrR are the two extractors'
// mvImagePyramid (level ROIs, no border). Returns the number of points with depth.
static inline int compute_stereo_matches(int N, const coeb_keypoint* keysL, const uint8_t* descL, int Nr,
                                         const coeb_keypoint* keysR, const uint8_t* descR,
                                         const std::vector<Image>& pyrL, const std::vector<Image>& pyrR,
                                         const float* scaleFactors, const float* invScaleFactors,
                                         float mbf, float mb, float* uRight, float* depth) {
    for (int i = 0; i < N; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (COEB_TH_HIGH + COEB_TH_LOW) / 2;
    const int nRows = pyrL[0].h;
    std::vector<std::vector<int>> vRowIndices(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = keysR[iR].y;
        const float r = 2.0f * scaleFactors[keysR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);  // reference indexes unchecked
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const coeb_keypoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<int>& vCandidates = vRowIndices[row];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = COEB_TH_HIGH;
        int bestIdxR = 0;
        const uint8_t* dL = descL + (size_t)iL * 32;
        for (int iR : vCandidates) {
            const coeb_keypoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256(dL, descR + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = keysR[bestIdxR].x;
            const float scaleFactor = invScaleFactors[kpL.octave];
            const float scaleduL = std::round(kpL.x * scaleFactor);
            const float scaledvL = std::round(kpL.y * scaleFactor);
            const float scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5, Lw = 5;
            const Image& IL = pyrL[kpL.octave];
            const Image& IR = pyrR[kpL.octave];
            const int yl = (int)(scaledvL - w), xl = (int)(scaleduL - w);
            int bestSad = INT_MAX, bestincR = 0;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + Lw - w;
            const float endu = scaleduR0 + Lw + w + 1;
            if (iniu < 0 || endu >= IR.w) continue;
            if (yl < 0 || yl + 2 * w + 1 > IL.h || xl < 0 || xl + 2 * w + 1 > IL.w ||
                yl + 2 * w + 1 > IR.h)
                continue;  // the reference's rowRange/colRange would throw here
            const float cL = (float)IL.row(yl + w)[xl + w];
            for (int incR = -Lw; incR <= Lw; incR++) {
                const int xr = (int)(scaleduR0 + incR - w);
                const float cR = (float)IR.row(yl + w)[xr + w];
                double acc = 0;  // cv::norm(IL, IR, NORM_L1) on CV_32F accumulates in double
                for (int dy = 0; dy < 2 * w + 1; dy++)
                    for (int dx = 0; dx < 2 * w + 1; dx++) {
                        const float a = (float)IL.row(yl + dy)[xl + dx] - cL;
                        const float b = (float)IR.row(yl + dy)[xr + dx] - cR;
                        acc += std::fabs((double)(a - b));
                    }
                const float dist = (float)acc;
                if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
                vDists[Lw + incR] = dist;
            }
            if (bestincR == -Lw || bestincR == Lw) continue;
            const float dist1 = vDists[Lw + bestincR - 1], dist2 = vDists[Lw + bestincR],
                        dist3 = vDists[Lw + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = uL - bestuR;
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = (float)(uL - 0.01);
                }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back({bestSad, iL});
            }
        }
    }
    if (vDistIdx.empty()) return 0;  // the reference reads vDistIdx[0] of an empty vector here
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if ((float)vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
        kept--;
    }
    return kept;
}

// DBoW2::FeatureVector flattened: `nn` vocabulary nodes in ascending id order (std::map iteration order), node k owning
// items[start[k] .. start[k+1]) = indices of the frame's features, in the order DBoW2 pushed them.
struct FeatVecCSR {
    int nn = 0;
    const int* node = nullptr;    // [nn] ascending, unique
    const int* start = nullptr;   // [nn + 1]
    const int* items = nullptr;   // [start[nn]]
    int lower_bound(int from, int id) const {
        return (int)(std::lower_bound(node + from, node + nn, id) - node);
    }
};

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (src/ORBmatcher.cc:158-288) and
// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (:522-655) as one restatement.
//   F1 = the keyframe whose features are the queries, valid1[i] = its MapPoint exists and is not bad (:190-196, :558-562);
//   F2 = the frame / keyframe searched; valid2 = nullptr for a Frame (every feature is a candidate), else MapPoint exists
//        and is not bad (:575-581); a feature of F2 already matched by an earlier query of the call is skipped (:210, :577);
//   strict_low: the Frame overload accepts bestDist1 <= TH_LOW (:229), the KeyFrame overload bestDist1 < TH_LOW (:601).
// match12[i1] = matched feature of F2 or -1. For the Frame overload the caller writes vpMapPointMatches[match12[i1]] =
// vpMapPointsKF[i1]; its rotation histogram holds bestIdxF (:246), which removes exactly the same pairs.
static inline int search_by_bow(const FrameView& F1, const FrameView& F2, const uint8_t* valid1, const uint8_t* valid2,
                                const FeatVecCSR& V1, const FeatVecCSR& V2, float nnratio, bool checkOri, bool strict_low,
                                int* match12) {
    for (int i = 0; i < F1.n; i++) match12[i] = -1;
    std::vector<uint8_t> matched2(F2.n, 0);
    std::vector<int> rotHist[COEB_HISTO_LENGTH];
    int nmatches = 0;
    int it1 = 0, it2 = 0;
    while (it1 != V1.nn && it2 != V2.nn) {
        if (V1.node[it1] == V2.node[it2]) {
            for (int p1 = V1.start[it1]; p1 < V1.start[it1 + 1]; p1++) {
                const int idx1 = V1.items[p1];
                if (!valid1[idx1]) continue;
                const uint8_t* d1 = F1.desc + (size_t)idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int p2 = V2.start[it2]; p2 < V2.start[it2 + 1]; p2++) {
                    const int idx2 = V2.items[p2];
                    if (matched2[idx2] || (valid2 && !valid2[idx2])) continue;
                    const int dist = hamming256(d1, F2.desc + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (strict_low ? bestDist1 < COEB_TH_LOW : bestDist1 <= COEB_TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        match12[idx1] = bestIdx2;
                        matched2[bestIdx2] = 1;
                        if (checkOri) rotHist[rot_bin(F1.kps[idx1].angle, F2.kps[bestIdx2].angle)].push_back(idx1);
                        nmatches++;
                    }
                }
            }
            it1++;
            it2++;
        } else if (V1.node[it1] < V2.node[it2]) {
            it1 = V1.lower_bound(it1, V2.node[it2]);
        } else {
            it2 = V2.lower_bound(it2, V1.node[it1]);
        }
    }
    if (checkOri) {
        int sizes[COEB_HISTO_LENGTH], ind1, ind2, ind3;
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, COEB_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) { match12[idx1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ORBmatcher::CheckDistEpipolarLine (src/ORBmatcher.cc:140-156): fp32 left to right, the final comparison in double
// (3.84 is a double literal). sigma2 = pKF2->mvLevelSigma2[kp2.octave].
static inline bool check_dist_epipolar_line(const coeb_keypoint& kp1, const coeb_keypoint& kp2, const float* F12, float sigma2) {
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return (double)dsqr < 3.84 * (double)sigma2;
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:657-824). free1 / free2: the keyframe's feature has no MapPoint
// (:697-701, :722-726); F12 3x3 row-major; (ex, ey) the epipole in the second image (:663-670, computed by the caller with
// the reference's own cv::Mat expressions); the `vbMatched2[bestIdx2] = true` line is commented out in the reference
// (:765), so queries do not interact. Ties: `dist > bestDist` skips, so an equal distance met later replaces the match.
static inline int search_for_triangulation(const FrameView& F1, const FrameView& F2, const uint8_t* free1, const uint8_t* free2,
                                           const FeatVecCSR& V1, const FeatVecCSR& V2, const float* F12, float ex, float ey,
                                           bool bOnlyStereo, bool checkOri, int* match12) {
    for (int i = 0; i < F1.n; i++) match12[i] = -1;
    std::vector<int> rotHist[COEB_HISTO_LENGTH];
    int nmatches = 0;
    int it1 = 0, it2 = 0;
    while (it1 != V1.nn && it2 != V2.nn) {
        if (V1.node[it1] == V2.node[it2]) {
            for (int p1 = V1.start[it1]; p1 < V1.start[it1 + 1]; p1++) {
                const int idx1 = V1.items[p1];
                if (!free1[idx1]) continue;
                const bool bStereo1 = F1.uright && F1.uright[idx1] >= 0;
                if (bOnlyStereo && !bStereo1) continue;
                const coeb_keypoint& kp1 = F1.kps[idx1];
                const uint8_t* d1 = F1.desc + (size_t)idx1 * 32;
                int bestDist = COEB_TH_LOW, bestIdx2 = -1;
                for (int p2 = V2.start[it2]; p2 < V2.start[it2 + 1]; p2++) {
                    const int idx2 = V2.items[p2];
                    if (!free2[idx2]) continue;
                    const bool bStereo2 = F2.uright && F2.uright[idx2] >= 0;
                    if (bOnlyStereo && !bStereo2) continue;
                    const int dist = hamming256(d1, F2.desc + (size_t)idx2 * 32);
                    if (dist > COEB_TH_LOW || dist > bestDist) continue;
                    const coeb_keypoint& kp2 = F2.kps[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x, distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * F2.scale[kp2.octave]) continue;
                    }
                    const float s = F2.scale[kp2.octave];
                    if (check_dist_epipolar_line(kp1, kp2, F12, s * s)) {   // mvLevelSigma2[i] = mvScaleFactor[i]^2 (src/ORBextractor.cc:427)
                        bestIdx2 = idx2;
                        bestDist = dist;
                    }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    nmatches++;
                    if (checkOri) rotHist[rot_bin(kp1.angle, F2.kps[bestIdx2].angle)].push_back(idx1);
                }
            }
            it1++;
            it2++;
        } else if (V1.node[it1] < V2.node[it2]) {
            it1 = V1.lower_bound(it1, V2.node[it2]);
        } else {
            it2 = V2.lower_bound(it2, V1.node[it1]);
        }
    }
    if (checkOri) {
        int sizes[COEB_HISTO_LENGTH], ind1, ind2, ind3;
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, COEB_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < COEB_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) { match12[idx1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// Brute-force k=2 nearest neighbour + ratio test (BASELINE.json config 5). Not a reference
// function; semantics borrowed from the SearchByBoW inner loop (src/ORBmatcher.cc:201-231) applied
// to the whole train set: strict '<' updates, first index wins ties, accept if best <= TH_LOW and
// best < ratio * second.
static inline int knn2_ratio(int nq, const uint8_t* q, int nt, const uint8_t* t, float nnratio,
                             int* best_idx, int* best_d1, int* best_d2) {
    int accepted = 0;
    for (int i = 0; i < nq; i++) {
        int d1 = 256, d2 = 256, idx = -1;
        for (int j = 0; j < nt; j++) {
            const int dist = hamming256(q + (size_t)i * 32, t + (size_t)j * 32);
            if (dist < d1) { d2 = d1; d1 = dist; idx = j; }
            else if (dist < d2) d2 = dist;
        }
        best_d1[i] = d1;
        best_d2[i] = d2;
        const bool ok = d1 <= COEB_TH_LOW && (float)d1 < nnratio * (float)d2;
        best_idx[i] = ok ? idx : -1;
        accepted += ok;
    }
    return accepted;
}

}  // namespace orc
