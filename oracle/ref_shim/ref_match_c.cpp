// ref_match_c.cpp -- C entry points around the reference's OWN Frame / MapPoint / KeyFrame / ORBmatcher classes,
// compiled unchanged from /root/reference/src/{Frame,MapPoint,KeyFrame,Map,ORBmatcher}.cc into oracle/_ref/.
//
// TEST INFRASTRUCTURE ONLY. Every function has the signature of its orc_* twin in oracle/coeb_oracle_c.cpp, so a test
// feeds the same flattened inputs to oracle B, to the reference and to the CUDA path. Here the flattened inputs are
// turned back into the pointer graph the reference works on (Frame with its grid, MapPoint objects with their tracking
// fields, mvpMapPoints) and the reference's own member function is called:
//   ref_features_in_area     Frame::GetFeaturesInArea                       src/Frame.cc:503-556
//   ref_grid_cell            Frame::AssignFeaturesToGrid / PosInGrid        src/Frame.cc:396-411, 558-568
//   ref_match_projection     ORBmatcher::SearchByProjection(F, MPs, th)     src/ORBmatcher.cc:45-137
//   ref_match_lastframe      ORBmatcher::SearchByProjection(cur, last, ..)  src/ORBmatcher.cc:1329-1471
//   ref_match_init           ORBmatcher::SearchForInitialization            src/ORBmatcher.cc:405-520
//   ref_stereo_match         Frame::ComputeStereoMatches                    src/Frame.cc:644-818
//   ref_search_local_points  Frame::isInFrustum + MapPoint::PredictScale + SearchByProjection, in the order of
//                            Tracking::SearchLocalPoints                    src/Tracking.cc:1222-1272
//   ref_undistort_keypoints  Frame::UndistortKeyPoints                      src/Frame.cc:579-609
//   ref_stereo_from_rgbd     Frame::ComputeStereoFromRGBD                   src/Frame.cc:820-842
//   ref_match_bow            ORBmatcher::SearchByBoW (both overloads)       src/ORBmatcher.cc:158-288, 522-655
//   ref_match_triangulation  ORBmatcher::SearchForTriangulation            src/ORBmatcher.cc:657-824
//   ref_match_reloc          ORBmatcher::SearchByProjection(F, KF, set..)   src/ORBmatcher.cc:1473-1600
//   ref_hamming256           ORBmatcher::DescriptorDistance                 src/ORBmatcher.cc:1648-1664
// The reference keeps the members these calls need private / protected; this one translation unit opens them with the
// usual test-harness macro AFTER every standard header has been included. The reference's own translation units are
// compiled without it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "cvshim.hpp"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "KeyFrameDatabase.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "Converter.h"
#undef private
#undef protected

#include "../../include/coeb_types.h"
#include "ref_handles.hpp"

using namespace ORB_SLAM2;

// ---- the two reference functions the five sources need from files that are not compiled ---------------------------
namespace ORB_SLAM2 {
// Converter::toDescriptorVector (src/Converter.cc): one Mat per descriptor row. Only reached from ComputeBoW.
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> v;
    v.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
}
// KeyFrameDatabase::erase is only reached from KeyFrame::SetBadFlag (mapping threads): outside the hot path.
void KeyFrameDatabase::erase(KeyFrame*) {
    std::fprintf(stderr, "ref_shim: KeyFrameDatabase::erase is outside the hot path\n");
    std::abort();
}
}  // namespace ORB_SLAM2

namespace {

Map& the_map() { static Map m; return m; }

cv::KeyPoint to_cv(const coeb_keypoint& k) { return cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id); }

cv::Mat desc_row(const uint8_t* d) {
    cv::Mat m(1, 32, CV_8UC1);
    std::memcpy(m.data, d, 32);
    return m;
}

cv::Mat pose44(const float* T34) {
    cv::Mat T = cv::Mat::eye(4, 4, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T.at<float>(r, c) = T34[4 * r + c];
    return T;
}

// A Frame with exactly one keypoint at the origin of an identity pose: the MapPoint constructor of src/MapPoint.cc:55-80
// needs some Frame to read its camera centre and scale table from; every field it derives is overwritten afterwards.
Frame& seed_frame() {
    static Frame* F = nullptr;
    if (!F) {
        F = new Frame();
        F->mnId = 0;
        F->N = 1;
        F->mvKeys.assign(1, cv::KeyPoint(0.f, 0.f, 31.f, 0.f, 0.f, 0, -1));
        F->mvKeysUn = F->mvKeys;
        F->mnScaleLevels = 1;
        F->mvScaleFactors.assign(1, 1.f);
        F->mDescriptors = cv::Mat::zeros(1, 32, CV_8UC1);
        F->SetPose(cv::Mat::eye(4, 4, CV_32F));
    }
    return *F;
}

// A MapPoint of the reference with the fields a matcher reads set explicitly.
MapPoint* new_map_point(const float* xyz, const float* normal, float min_dist, float max_dist, const uint8_t* desc, bool bad, int nobs) {
    cv::Mat pos(3, 1, CV_32F);
    pos.at<float>(0) = 0.f; pos.at<float>(1) = 0.f; pos.at<float>(2) = 1.f;
    MapPoint* p = new MapPoint(pos, &the_map(), &seed_frame(), 0);
    if (xyz) for (int k = 0; k < 3; k++) p->mWorldPos.at<float>(k) = xyz[k];
    if (normal) for (int k = 0; k < 3; k++) p->mNormalVector.at<float>(k) = normal[k];
    p->mfMinDistance = min_dist;
    p->mfMaxDistance = max_dist;
    if (desc) p->mDescriptor = desc_row(desc);
    p->mbBad = bad;
    p->nObs = nobs;
    p->mbTrackInView = false;
    p->mnLastFrameSeen = 0;
    return p;
}

// kp_match <-> Frame::mvpMapPoints (encoding of oracle/coeb_oracle_match.hpp:94-96)
struct Claims {
    MapPoint* taken;      // -2: a MapPoint not from this call with Observations() > 0
    MapPoint* taken_free; // -3: one with Observations() == 0
    Claims() {
        taken = new_map_point(nullptr, nullptr, 1.f, 1.f, nullptr, false, 1);
        taken_free = new_map_point(nullptr, nullptr, 1.f, 1.f, nullptr, false, 0);
    }
    ~Claims() { delete taken; delete taken_free; }
    void load(Frame& F, const int* kp_match, const std::vector<MapPoint*>& mps) const {
        for (int i = 0; i < F.N; i++) {
            const int c = kp_match[i];
            F.mvpMapPoints[i] = c == -1 ? nullptr : c == -2 ? taken : c == -3 ? taken_free : (c >= 0 && c < (int)mps.size() ? mps[c] : nullptr);
        }
    }
    void store(const Frame& F, int* kp_match, const std::vector<MapPoint*>& mps) const {
        std::map<MapPoint*, int> index;
        for (size_t i = 0; i < mps.size(); i++) if (mps[i]) index[mps[i]] = (int)i;
        for (int i = 0; i < F.N; i++) {
            MapPoint* p = F.mvpMapPoints[i];
            kp_match[i] = !p ? -1 : p == taken ? -2 : p == taken_free ? -3 : index.at(p);
        }
    }
};

void free_points(std::vector<MapPoint*>& v) {
    for (MapPoint* p : v) delete p;
    v.clear();
}

}  // namespace

struct ref_frame {
    Frame* F;
};

extern "C" {

int ref_hamming256(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

// Builds the Frame members the matchers read, the way the Frame constructors do (src/Frame.cc:141-157, 229-246): keys are the
// undistorted keypoints; the camera statics (shared by every Frame of the reference) are set from `cam`.
ref_frame* ref_frame_create(const coeb_keypoint* kps, const uint8_t* desc, int n, const float* uright, const coeb_camera* cam,
                            const float* scale, int nlevels) {
    Frame* F = new Frame();
    F->mnId = Frame::nNextId++;
    F->N = n;
    F->mvKeys.resize(n);
    for (int i = 0; i < n; i++) F->mvKeys[i] = to_cv(kps[i]);
    F->mvKeysUn = F->mvKeys;
    F->mDescriptors = cv::Mat(n, 32, CV_8UC1);
    if (n) std::memcpy(F->mDescriptors.data, desc, (size_t)32 * n);
    F->mvuRight.assign(n, -1.f);
    F->mvDepth.assign(n, -1.f);
    if (uright) for (int i = 0; i < n; i++) F->mvuRight[i] = uright[i];
    F->mvpMapPoints.assign(n, static_cast<MapPoint*>(nullptr));
    F->mvbOutlier.assign(n, false);
    F->mnScaleLevels = nlevels;
    F->mvScaleFactors.assign(scale, scale + nlevels);
    F->mvInvScaleFactors.resize(nlevels);
    F->mvLevelSigma2.resize(nlevels);
    F->mvInvLevelSigma2.resize(nlevels);
    for (int l = 0; l < nlevels; l++) {   // ORBextractor::ORBextractor, src/ORBextractor.cc:422-439
        F->mvLevelSigma2[l] = scale[l] * scale[l];
        F->mvInvScaleFactors[l] = 1.0f / scale[l];
        F->mvInvLevelSigma2[l] = 1.0f / F->mvLevelSigma2[l];
    }
    F->mfScaleFactor = nlevels > 1 ? scale[1] : 1.f;
    F->mfLogScaleFactor = log(F->mfScaleFactor);   // src/Frame.cc:151 (float overload)
    F->mbf = cam->bf;
    F->mb = cam->b;
    Frame::fx = cam->fx; Frame::fy = cam->fy; Frame::cx = cam->cx; Frame::cy = cam->cy;
    Frame::invfx = 1.0f / cam->fx; Frame::invfy = 1.0f / cam->fy;
    Frame::mnMinX = cam->min_x; Frame::mnMaxX = cam->max_x; Frame::mnMinY = cam->min_y; Frame::mnMaxY = cam->max_y;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);   // src/Frame.cc:231-232
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    Frame::mbInitialComputations = false;
    F->AssignFeaturesToGrid();
    return new ref_frame{F};
}
void ref_frame_destroy(ref_frame* f) {
    if (!f) return;
    delete f->F;
    delete f;
}

int ref_features_in_area(ref_frame* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    const std::vector<size_t> v = f->F->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}
int ref_grid_cell(ref_frame* f, int ix, int iy, int* out, int cap) {
    const std::vector<size_t>& c = f->F->mGrid[ix][iy];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[i] = (int)c[i];
    return (int)c.size();
}

int ref_match_projection(ref_frame* f, int n, const uint8_t* track_in_view, const uint8_t* bad, const uint8_t* has_obs,
                         const float* proj_x, const float* proj_y, const float* proj_xr, const int* level, const float* view_cos,
                         const uint8_t* desc, float th, float nnratio, int* kp_match) {
    Frame& F = *f->F;
    std::vector<MapPoint*> mps((size_t)n);
    for (int i = 0; i < n; i++) {
        MapPoint* p = new_map_point(nullptr, nullptr, 1.f, 1.f, desc + (size_t)32 * i, bad[i] != 0, has_obs[i] ? 1 : 0);
        p->mbTrackInView = track_in_view[i] != 0;   // the fields Frame::isInFrustum leaves behind (src/Frame.cc:492-498)
        p->mTrackProjX = proj_x[i];
        p->mTrackProjY = proj_y[i];
        p->mTrackProjXR = proj_xr[i];
        p->mnTrackScaleLevel = level[i];
        p->mTrackViewCos = view_cos[i];
        mps[i] = p;
    }
    Claims claims;
    claims.load(F, kp_match, mps);
    ORBmatcher matcher(nnratio);
    const int nm = matcher.SearchByProjection(F, mps, th);
    claims.store(F, kp_match, mps);
    F.mvpMapPoints.assign(F.N, static_cast<MapPoint*>(nullptr));
    free_points(mps);
    return nm;
}

int ref_match_lastframe(ref_frame* cur, int n, const uint8_t* valid, const uint8_t* has_obs, const float* xyz, const int* octave,
                        const float* angle, const uint8_t* desc, const float* Tcw_cur, const float* Tcw_last, float th, int mono,
                        int check_ori, int* kp_match) {
    Frame& C = *cur->F;
    Frame L;
    L.N = n;
    L.mvKeys.resize(n);
    for (int i = 0; i < n; i++) L.mvKeys[i] = cv::KeyPoint(0.f, 0.f, 31.f, angle[i], 0.f, octave[i], -1);
    L.mvKeysUn = L.mvKeys;
    L.mvbOutlier.assign(n, false);
    L.mvpMapPoints.assign(n, static_cast<MapPoint*>(nullptr));
    std::vector<MapPoint*> mps((size_t)n, nullptr);
    for (int i = 0; i < n; i++) {
        // every index gets an object (kp_match may name it as a pre-existing claim); only valid ones hang in the last frame
        mps[i] = new_map_point(xyz + 3 * (size_t)i, nullptr, 1.f, 1.f, desc + (size_t)32 * i, false, has_obs[i] ? 1 : 0);
        if (valid[i]) L.mvpMapPoints[i] = mps[i];
    }
    L.SetPose(pose44(Tcw_last));
    C.SetPose(pose44(Tcw_cur));
    Claims claims;
    claims.load(C, kp_match, mps);
    ORBmatcher matcher(0.9f, check_ori != 0);   // Tracking::TrackWithMotionModel, src/Tracking.cc:914
    const int nm = matcher.SearchByProjection(C, L, th, mono != 0);
    claims.store(C, kp_match, mps);
    C.mvpMapPoints.assign(C.N, static_cast<MapPoint*>(nullptr));
    free_points(mps);
    return nm;
}

int ref_match_init(ref_frame* f1, ref_frame* f2, float* prev_matched, int* matches12, int window, float nnratio, int check_ori) {
    Frame &F1 = *f1->F, &F2 = *f2->F;
    std::vector<cv::Point2f> prev((size_t)F1.N);
    for (int i = 0; i < F1.N; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < F1.N; i++) {
        matches12[i] = m12[i];
        prev_matched[2 * i] = prev[i].x;
        prev_matched[2 * i + 1] = prev[i].y;
    }
    return nm;
}

// Frame::ComputeStereoMatches on the pyramids the two reference extractors hold after their last call.
int ref_stereo_match(ref_extractor* exL, ref_extractor* exR, int N, const coeb_keypoint* keysL, const uint8_t* descL, int Nr,
                     const coeb_keypoint* keysR, const uint8_t* descR, float mbf, float mb, float* uright, float* depth) {
    Frame F;
    F.mpORBextractorLeft = exL->ex;
    F.mpORBextractorRight = exR->ex;
    F.N = N;
    F.mvKeys.resize(N);
    for (int i = 0; i < N; i++) F.mvKeys[i] = to_cv(keysL[i]);
    F.mvKeysRight.resize(Nr);
    for (int i = 0; i < Nr; i++) F.mvKeysRight[i] = to_cv(keysR[i]);
    F.mDescriptors = cv::Mat(N, 32, CV_8UC1);
    if (N) std::memcpy(F.mDescriptors.data, descL, (size_t)32 * N);
    F.mDescriptorsRight = cv::Mat(Nr, 32, CV_8UC1);
    if (Nr) std::memcpy(F.mDescriptorsRight.data, descR, (size_t)32 * Nr);
    F.mvScaleFactors = exL->ex->GetScaleFactors();
    F.mvInvScaleFactors = exL->ex->GetInverseScaleFactors();
    F.mbf = mbf;
    F.mb = mb;
    F.ComputeStereoMatches();
    int with_depth = 0;
    for (int i = 0; i < N; i++) {
        uright[i] = F.mvuRight[i];
        depth[i] = F.mvDepth[i];
        with_depth += F.mvDepth[i] > 0;
    }
    return with_depth;
}

// Tracking::SearchLocalPoints (src/Tracking.cc:1222-1272): second loop + the matcher call, on reference objects.
// `Ow_out` (optional, 3 floats) returns the camera centre Frame::UpdatePoseMatrices computed; the `Ow` argument of the
// orc_ twin is not needed here.
int ref_search_local_points(ref_frame* f, int n, const float* xyz, const float* normal, const float* min_dist, const float* max_dist,
                            const uint8_t* desc, const uint8_t* skip, const uint8_t* has_obs, const float* Tcw, float* Ow_out,
                            float cos_limit, float th, float nnratio, int* kp_match, uint8_t* in_view, float* proj) {
    Frame& F = *f->F;
    F.SetPose(pose44(Tcw));
    if (Ow_out) for (int k = 0; k < 3; k++) Ow_out[k] = F.mOw.at<float>(k);
    std::vector<MapPoint*> mps((size_t)n);
    for (int i = 0; i < n; i++) mps[i] = new_map_point(xyz + 3 * (size_t)i, normal + 3 * (size_t)i, min_dist[i], max_dist[i], desc + (size_t)32 * i, false, has_obs[i] ? 1 : 0);
    int nToMatch = 0;
    for (int i = 0; i < n; i++) {
        MapPoint* pMP = mps[i];
        bool vis = false;
        if (!skip[i] && F.isInFrustum(pMP, cos_limit)) { vis = true; nToMatch++; }
        if (in_view) in_view[i] = vis;
        if (proj) {
            proj[5 * i] = vis ? pMP->mTrackProjX : 0.f; proj[5 * i + 1] = vis ? pMP->mTrackProjY : 0.f; proj[5 * i + 2] = vis ? pMP->mTrackProjXR : 0.f;
            proj[5 * i + 3] = vis ? pMP->mTrackViewCos : 0.f; proj[5 * i + 4] = vis ? (float)pMP->mnTrackScaleLevel : 0.f;
        }
    }
    int nm = 0;
    if (nToMatch > 0) {
        Claims claims;
        claims.load(F, kp_match, mps);
        ORBmatcher matcher(nnratio);
        nm = matcher.SearchByProjection(F, mps, th);
        claims.store(F, kp_match, mps);
        F.mvpMapPoints.assign(F.N, static_cast<MapPoint*>(nullptr));
    }
    free_points(mps);
    return nm;
}

// Frame::UndistortKeyPoints (src/Frame.cc:579-609). dist5 = {k1, k2, p1, p2, k3}.
void ref_undistort_keypoints(const coeb_keypoint* keys, int n, const coeb_camera* cam, const float* dist5, coeb_keypoint* keys_un) {
    Frame F;
    F.N = n;
    F.mvKeys.resize(n);
    for (int i = 0; i < n; i++) F.mvKeys[i] = to_cv(keys[i]);
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.mK.at<float>(0, 0) = cam->fx; F.mK.at<float>(1, 1) = cam->fy; F.mK.at<float>(0, 2) = cam->cx; F.mK.at<float>(1, 2) = cam->cy;
    F.mDistCoef = cv::Mat(5, 1, CV_32F);
    for (int k = 0; k < 5; k++) F.mDistCoef.at<float>(k) = dist5 ? dist5[k] : 0.f;
    F.UndistortKeyPoints();
    for (int i = 0; i < n; i++) {
        const cv::KeyPoint& k = F.mvKeysUn[i];
        keys_un[i] = coeb_keypoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
    }
}

// Frame::ComputeStereoFromRGBD (src/Frame.cc:820-842); a raw 16-bit map is first scaled as Tracking::GrabImageRGBD does
// (src/Tracking.cc:226-229). kind: 1 float32, 2 uint16.
void ref_stereo_from_rgbd(const coeb_keypoint* keys, const coeb_keypoint* keys_un, int n, const void* depth, int kind, int w, int h,
                          int stride_bytes, float factor, float mbf, float* uright, float* depth_out) {
    Frame F;
    F.N = n;
    F.mvKeys.resize(n);
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { F.mvKeys[i] = to_cv(keys[i]); F.mvKeysUn[i] = to_cv(keys_un[i]); }
    F.mbf = mbf;
    cv::Mat imDepth(h, w, kind == 1 ? CV_32F : CV_16U, (void*)depth, (size_t)stride_bytes);
    if (kind == 2) imDepth.convertTo(imDepth, CV_32F, factor);
    F.ComputeStereoFromRGBD(imDepth);
    for (int i = 0; i < n; i++) { uright[i] = F.mvuRight[i]; depth_out[i] = F.mvDepth[i]; }
}

}  // extern "C"

// ---- KeyFrame-based searches ----------------------------------------------------------------------------------------
namespace {

// DBoW2::FeatureVector from the CSR triple of the orc_ twin (node ids ascending, items in push order).
DBoW2::FeatureVector to_featvec(int nn, const int* node, const int* start, const int* items) {
    DBoW2::FeatureVector fv;
    for (int k = 0; k < nn; k++)
        for (int j = start[k]; j < start[k + 1]; j++) fv.addFeature((DBoW2::NodeId)node[k], (unsigned int)items[j]);
    return fv;
}

// KeyFrame::KeyFrame(Frame&, Map*, KeyFrameDatabase*) (src/KeyFrame.cc:32-56) copies keypoints, descriptors, grid, feature
// vector, scale tables and pose from the Frame.
KeyFrame* new_keyframe(Frame& F, const cv::Mat& Tcw) {
    F.SetPose(Tcw);
    F.mpORBvocabulary = nullptr;
    return new KeyFrame(F, &the_map(), nullptr);
}

}  // namespace

extern "C" {

// ORBmatcher::SearchByBoW: strict_low == 0 -> (KeyFrame* pKF, Frame& F, vpMapPointMatches) (src/ORBmatcher.cc:158-288),
// valid2 must be null; strict_low != 0 -> (KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (:522-655).
int ref_match_bow(ref_frame* f1, ref_frame* f2, const uint8_t* valid1, const uint8_t* valid2, int nn1, const int* node1, const int* start1,
                  const int* items1, int nn2, const int* node2, const int* start2, const int* items2, float nnratio, int check_ori,
                  int strict_low, int* match12) {
    Frame &F1 = *f1->F, &F2 = *f2->F;
    F1.mFeatVec = to_featvec(nn1, node1, start1, items1);
    F2.mFeatVec = to_featvec(nn2, node2, start2, items2);
    KeyFrame* kf1 = new_keyframe(F1, cv::Mat::eye(4, 4, CV_32F));
    std::vector<MapPoint*> mp1((size_t)F1.N, nullptr), mp2((size_t)F2.N, nullptr);
    for (int i = 0; i < F1.N; i++)
        if (valid1[i]) { mp1[i] = new_map_point(nullptr, nullptr, 1.f, 1.f, nullptr, false, 1); kf1->AddMapPoint(mp1[i], i); }
    for (int i = 0; i < F1.N; i++) match12[i] = -1;
    ORBmatcher matcher(nnratio, check_ori != 0);
    int nm;
    if (!strict_low) {
        std::vector<MapPoint*> vpMapPointMatches;
        nm = matcher.SearchByBoW(kf1, F2, vpMapPointMatches);
        std::map<MapPoint*, int> idx1;
        for (int i = 0; i < F1.N; i++) if (mp1[i]) idx1[mp1[i]] = i;
        for (int i2 = 0; i2 < F2.N; i2++)
            if (vpMapPointMatches[i2]) match12[idx1.at(vpMapPointMatches[i2])] = i2;
    } else {
        KeyFrame* kf2 = new_keyframe(F2, cv::Mat::eye(4, 4, CV_32F));
        for (int i = 0; i < F2.N; i++)
            if (!valid2 || valid2[i]) { mp2[i] = new_map_point(nullptr, nullptr, 1.f, 1.f, nullptr, false, 1); kf2->AddMapPoint(mp2[i], i); }
        std::vector<MapPoint*> vpMatches12;
        nm = matcher.SearchByBoW(kf1, kf2, vpMatches12);
        std::map<MapPoint*, int> idx2;
        for (int i = 0; i < F2.N; i++) if (mp2[i]) idx2[mp2[i]] = i;
        for (int i1 = 0; i1 < F1.N; i1++)
            if (vpMatches12[i1]) match12[i1] = idx2.at(vpMatches12[i1]);
        delete kf2;
    }
    delete kf1;
    free_points(mp1);
    free_points(mp2);
    return nm;
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:657-824). Keyframe 1 sits at the origin and keyframe 2 at
// [I | t2w], so the reference's own epipole expression (:663-670) evaluates fx * t2w.x / t2w.z + cx; (ex_out, ey_out)
// return what it computed from the same inputs.
int ref_match_triangulation(ref_frame* f1, ref_frame* f2, const uint8_t* free1, const uint8_t* free2, int nn1, const int* node1,
                            const int* start1, const int* items1, int nn2, const int* node2, const int* start2, const int* items2,
                            const float* F12, const float* t2w, int only_stereo, int check_ori, int* match12, float* ex_out, float* ey_out) {
    Frame &F1 = *f1->F, &F2 = *f2->F;
    F1.mFeatVec = to_featvec(nn1, node1, start1, items1);
    F2.mFeatVec = to_featvec(nn2, node2, start2, items2);
    const float mb1 = F1.mb, mb2 = F2.mb;
    F1.mb = 0.f;   // mHalfBaseline = F.mb / 2 (src/KeyFrame.cc:46): the camera centre Cw is then exactly Ow
    F2.mb = 0.f;
    cv::Mat T2 = cv::Mat::eye(4, 4, CV_32F);
    for (int k = 0; k < 3; k++) T2.at<float>(k, 3) = t2w[k];
    KeyFrame* kf1 = new_keyframe(F1, cv::Mat::eye(4, 4, CV_32F));
    KeyFrame* kf2 = new_keyframe(F2, T2);
    F1.mb = mb1;
    F2.mb = mb2;
    std::vector<MapPoint*> mp1, mp2;
    for (int i = 0; i < F1.N; i++)
        if (!free1[i]) { mp1.push_back(new_map_point(nullptr, nullptr, 1.f, 1.f, nullptr, false, 1)); kf1->AddMapPoint(mp1.back(), i); }
    for (int i = 0; i < F2.N; i++)
        if (!free2[i]) { mp2.push_back(new_map_point(nullptr, nullptr, 1.f, 1.f, nullptr, false, 1)); kf2->AddMapPoint(mp2.back(), i); }
    cv::Mat F(3, 3, CV_32F);
    for (int k = 0; k < 9; k++) F.at<float>(k / 3, k % 3) = F12[k];
    if (ex_out && ey_out) {   // the reference's expression, in this translation unit's arithmetic
        cv::Mat C2 = kf2->GetRotation() * kf1->GetCameraCenter() + kf2->GetTranslation();
        const float invz = 1.0f / C2.at<float>(2);
        *ex_out = kf2->fx * C2.at<float>(0) * invz + kf2->cx;
        *ey_out = kf2->fy * C2.at<float>(1) * invz + kf2->cy;
    }
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher matcher(0.6f, check_ori != 0);   // LocalMapping::CreateNewMapPoints: ORBmatcher matcher(0.6, false)
    const int nm = matcher.SearchForTriangulation(kf1, kf2, F, pairs, only_stereo != 0);
    for (int i = 0; i < F1.N; i++) match12[i] = -1;
    for (size_t k = 0; k < pairs.size(); k++) match12[pairs[k].first] = (int)pairs[k].second;
    delete kf1;
    delete kf2;
    free_points(mp1);
    free_points(mp2);
    return nm;
}

// Relocalisation overload SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1473-1600).
// valid[i] = the keyframe's map point i exists, is not bad and is not in sAlreadyFound.
int ref_match_reloc(ref_frame* cur, int n, const uint8_t* valid, const float* xyz, const float* min_dist, const float* max_dist,
                    const float* angle, const uint8_t* desc, const float* Tcw, float* Ow_out, float th, int orb_dist, int check_ori,
                    int* kp_match) {
    Frame& C = *cur->F;
    C.SetPose(pose44(Tcw));
    if (Ow_out) for (int k = 0; k < 3; k++) Ow_out[k] = C.mOw.at<float>(k);
    Frame K;   // the keyframe's source frame: only keypoint angles and the scale tables matter
    K.N = n;
    K.mvKeys.resize(n);
    for (int i = 0; i < n; i++) K.mvKeys[i] = cv::KeyPoint(0.f, 0.f, 31.f, angle[i], 0.f, 0, -1);
    K.mvKeysUn = K.mvKeys;
    K.mvuRight.assign(n, -1.f);
    K.mvDepth.assign(n, -1.f);
    K.mDescriptors = cv::Mat::zeros(std::max(n, 1), 32, CV_8UC1);
    K.mvpMapPoints.assign(n, static_cast<MapPoint*>(nullptr));
    K.mnScaleLevels = C.mnScaleLevels; K.mfScaleFactor = C.mfScaleFactor; K.mfLogScaleFactor = C.mfLogScaleFactor;
    K.mvScaleFactors = C.mvScaleFactors; K.mvLevelSigma2 = C.mvLevelSigma2; K.mvInvLevelSigma2 = C.mvInvLevelSigma2;
    K.mbf = C.mbf; K.mb = C.mb;
    KeyFrame* kf = new_keyframe(K, cv::Mat::eye(4, 4, CV_32F));
    std::vector<MapPoint*> mps((size_t)n, nullptr);
    for (int i = 0; i < n; i++) {
        mps[i] = new_map_point(xyz + 3 * (size_t)i, nullptr, min_dist[i], max_dist[i], desc + (size_t)32 * i, false, 1);
        if (valid[i]) kf->AddMapPoint(mps[i], i);
    }
    Claims claims;
    for (int i = 0; i < C.N; i++) C.mvpMapPoints[i] = kp_match[i] == -1 ? nullptr : claims.taken;   // every non-null entry blocks (:1546-1547)
    std::set<MapPoint*> sAlreadyFound;
    ORBmatcher matcher(0.9f, check_ori != 0);
    const int nm = matcher.SearchByProjection(C, kf, sAlreadyFound, th, orb_dist);
    std::map<MapPoint*, int> index;
    for (int i = 0; i < n; i++) index[mps[i]] = i;
    for (int i = 0; i < C.N; i++) {
        MapPoint* p = C.mvpMapPoints[i];
        if (!p) kp_match[i] = -1;
        else if (p != claims.taken) kp_match[i] = index.at(p);
    }
    C.mvpMapPoints.assign(C.N, static_cast<MapPoint*>(nullptr));
    delete kf;
    free_points(mps);
    return nm;
}

// blur_flag of the RGB-D Frame constructor (src/Frame.cc:171-202): per box, Frame::detect_laplacian of the cropped
// region (:905-913), flag = mean < 4.2.
void ref_blur_flags(const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox, int* flags, double* means) {
    cv::Mat imGray(h, w, CV_8UC1, (void*)gray, (size_t)stride);
    cv::Mat imGray_copy = imGray.clone();
    Frame F;
    for (int i = 0; i < nbox; i++) {
        const float* box = boxes + 4 * i;
        cv::Mat image = imGray_copy(cv::Rect(int(box[0]), int(box[1]), int(box[2] - box[0]), int(box[3] - box[1]))).clone();
        const double cast1 = F.detect_laplacian(image);
        if (means) means[i] = cast1;
        flags[i] = cast1 < 4.2 ? 1 : 0;
    }
}

}  // extern "C"
