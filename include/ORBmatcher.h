/*
 * ORBmatcher.h -- drop-in ORB_SLAM2::ORBmatcher for the hot-path overloads, backed by the B200 C ABI
 * (coeb_frontend.h). Same constructor, statics and call signatures as the reference's
 * include/ORBmatcher.h:41-69,87-89 for
 *     DescriptorDistance, SearchByProjection(Frame&, vector<MapPoint*>&, th),
 *     SearchByProjection(Frame& cur, const Frame& last, th, bMono), SearchForInitialization(...)
 * plus ComputeStereoMatches as a free function (the reference keeps it in Frame, src/Frame.cc:644-818), and for the
 * vocabulary-guided overloads SearchByBoW(KeyFrame*, Frame&, ...), SearchByBoW(KeyFrame*, KeyFrame*, ...) and
 * SearchForTriangulation (:158-288, :522-655, :657-824): the DBoW2 FeatureVectors stay the reference's (they come out of
 * its vocabulary) and are flattened to CSR here. Free functions at the end cover the steps either side of the path:
 * FrameTailFromExtractor (UndistortKeyPoints + ComputeStereoFromRGBD + AssignFeaturesToGrid on the device) and
 * SearchLocalPoints (isInFrustum + SearchByProjection against a device-resident local map).
 * The relocalisation overload SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (:1473-1600) and
 * Fuse(KeyFrame*, vpMapPoints, th) (:826-961) are covered too. The Sim3 functions of loop closing (SearchBySim3, Fuse with Scw,
 * SearchByProjection(KeyFrame*, Scw, ...)) stay the reference's.
 *
 * The member functions are templates over the reference's Frame / MapPoint types, so this header has no
 * dependency on them: it only touches the members the reference functions touch (cited inline). The
 * adapter flattens the pointer graph into the structure-of-arrays the C ABI takes, calls the CUDA
 * kernels and writes the results back into Frame::mvpMapPoints / vnMatches12 / vbPrevMatched exactly
 * as the reference does. coeb_adapt::* are the small type shims (cv::Mat <-> raw floats); overload them
 * for other matrix types.
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "ORBextractor.h"
#include "coeb_frontend.h"

namespace coeb_adapt {

#ifdef COEB_WITH_OPENCV
inline const unsigned char* desc_row(const cv::Mat& m, int i) { return m.ptr<unsigned char>(i); }
inline void pose34(const cv::Mat& Tcw, float out[12]) {
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) out[4 * r + c] = Tcw.at<float>(r, c);
}
inline void xyz3(const cv::Mat& p, float out[3]) { out[0] = p.at<float>(0); out[1] = p.at<float>(1); out[2] = p.at<float>(2); }
#else
inline const unsigned char* desc_row(const coeb_cv::Mat& m, int i) { return m.ptr(i); }
#endif
template <class A> inline void pose34(const A& Tcw, float out[12]) { for (int i = 0; i < 12; i++) out[i] = Tcw[i]; }
template <class A> inline void xyz3(const A& p, float out[3]) { out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; }

/* Device context shared by the matchers of one thread (stream + scratch), created on first use. */
inline coeb_matcher* matcher(int device = 0) {
    static thread_local coeb_matcher* m = nullptr;
    if (!m && coeb_matcher_create(device, &m) != COEB_OK) throw std::runtime_error(std::string("coeb: ") + coeb_last_error());
    return m;
}
inline void check(int st) {
    if (st != COEB_OK) throw std::runtime_error(std::string("coeb: ") + coeb_last_error());
}

/* Does the Frame type carry a device-resident twin (a `coeb_frame* mpDeviceFrame` member, filled by FrameTailFromExtractor)? */
template <class T, class = void> struct has_device_frame : std::false_type {};
template <class T> struct has_device_frame<T, std::void_t<decltype(std::declval<const T&>().mpDeviceFrame)>> : std::true_type {};

/* The Frame fields the matchers read (mvKeysUn, mDescriptors, mvuRight, bounds, scale factors) on the device. A Frame that
 * already owns a device frame (member mpDeviceFrame, see FrameTailFromExtractor) is used as it is: no upload, no grid build;
 * otherwise the fields are uploaded for the duration of the call (RAII). */
template <class FrameT>
struct DeviceFrame {
    coeb_frame* f = nullptr;
    bool owned = true;
    explicit DeviceFrame(const FrameT& F) {
        if constexpr (has_device_frame<FrameT>::value) {
            if (F.mpDeviceFrame) { f = F.mpDeviceFrame; owned = false; return; }
        }
        const int n = (int)F.mvKeysUn.size();
        std::vector<unsigned char> desc((size_t)n * 32);
        for (int i = 0; i < n; i++) std::memcpy(&desc[(size_t)i * 32], desc_row(F.mDescriptors, i), 32);
        coeb_camera cam;
        cam.fx = F.fx; cam.fy = F.fy; cam.cx = F.cx; cam.cy = F.cy; cam.bf = F.mbf; cam.b = F.mb;
        cam.min_x = F.mnMinX; cam.max_x = F.mnMaxX; cam.min_y = F.mnMinY; cam.max_y = F.mnMaxY;
        check(coeb_frame_create(matcher(), reinterpret_cast<const coeb_keypoint*>(F.mvKeysUn.data()), desc.data(), n,
                                F.mvuRight.empty() ? nullptr : F.mvuRight.data(), &cam, F.mvScaleFactors.data(),
                                (int)F.mvScaleFactors.size(), &f));
    }
    ~DeviceFrame() { if (owned) coeb_frame_destroy(f); }
    DeviceFrame(const DeviceFrame&) = delete;
    DeviceFrame& operator=(const DeviceFrame&) = delete;
};

/* DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>, iterated in ascending node id) -> CSR. */
struct FeatVecCSR {
    std::vector<int> node, start, items;
    template <class FeatVecT>
    explicit FeatVecCSR(const FeatVecT& fv) {
        start.push_back(0);
        for (auto it = fv.begin(); it != fv.end(); ++it) {
            node.push_back((int)it->first);
            for (auto idx : it->second) items.push_back((int)idx);
            start.push_back((int)items.size());
        }
    }
    int nn() const { return (int)node.size(); }
};

#ifdef COEB_WITH_OPENCV
inline void mat33(const cv::Mat& M, float out[9]) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) out[3 * r + c] = M.at<float>(r, c); }
#endif
template <class A> inline void mat33(const A& M, float out[9]) { for (int i = 0; i < 9; i++) out[i] = M[i]; }

}  // namespace coeb_adapt

namespace ORB_SLAM2 {

template <class MapPointT> inline coeb_local_map* MakeLocalMap(const std::vector<MapPointT*>& vpLocalMapPoints);

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    /* src/ORBmatcher.cc:1648-1664. A single pair is pure host arithmetic in the reference (also called from the mapping
     * threads, src/MapPoint.cc:281); batches go through coeb_hamming256_batch. */
    template <class MatT>
    static int DescriptorDistance(const MatT& a, const MatT& b) {
        const unsigned char *pa = coeb_adapt::desc_row(a, 0), *pb = coeb_adapt::desc_row(b, 0);
        int dist = 0;
        for (int i = 0; i < 8; i++) {
            uint32_t x, y;
            std::memcpy(&x, pa + 4 * i, 4);
            std::memcpy(&y, pb + 4 * i, 4);
            dist += __builtin_popcount(x ^ y);
        }
        return dist;
    }

    /* SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th=3) -- src/ORBmatcher.cc:45-129 */
    template <class FrameT, class MapPointT>
    int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th = 3) {
        const int n = (int)vpMapPoints.size(), K = (int)F.mvKeysUn.size();
        std::vector<uint8_t> tiv(n), bad(n), obs(n), desc((size_t)n * 32);
        std::vector<float> px(n), py(n), pxr(n), vc(n);
        std::vector<int> lvl(n);
        for (int i = 0; i < n; i++) {
            MapPointT* p = vpMapPoints[i];
            tiv[i] = p->mbTrackInView;
            bad[i] = p->isBad();
            obs[i] = p->Observations() > 0;
            px[i] = p->mTrackProjX; py[i] = p->mTrackProjY; pxr[i] = p->mTrackProjXR;
            lvl[i] = p->mnTrackScaleLevel; vc[i] = p->mTrackViewCos;
            if (tiv[i] && !bad[i]) std::memcpy(&desc[(size_t)i * 32], coeb_adapt::desc_row(p->GetDescriptor(), 0), 32);
        }
        std::vector<int> state(K);
        for (int k = 0; k < K; k++) state[k] = !F.mvpMapPoints[k] ? -1 : (F.mvpMapPoints[k]->Observations() > 0 ? -2 : -3);
        coeb_adapt::DeviceFrame<FrameT> dF(F);
        int nmatches = 0;
        coeb_adapt::check(coeb_match_projection(coeb_adapt::matcher(), dF.f, n, tiv.data(), bad.data(), obs.data(), px.data(), py.data(),
                                                pxr.data(), lvl.data(), vc.data(), desc.data(), th, mfNNratio, state.data(), &nmatches));
        for (int k = 0; k < K; k++)
            if (state[k] >= 0) F.mvpMapPoints[k] = vpMapPoints[state[k]];
        return nmatches;
    }

    /* SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono) -- :1329-1471 */
    template <class FrameT>
    int SearchByProjection(FrameT& CurrentFrame, const FrameT& LastFrame, const float th, const bool bMono) {
        const int n = LastFrame.N, K = (int)CurrentFrame.mvKeysUn.size();
        std::vector<uint8_t> valid(n), obs(n), desc((size_t)n * 32);
        std::vector<float> xyz((size_t)n * 3), ang(n);
        std::vector<int> oct(n);
        for (int i = 0; i < n; i++) {
            auto* p = LastFrame.mvpMapPoints[i];
            valid[i] = p && !LastFrame.mvbOutlier[i];
            oct[i] = LastFrame.mvKeys[i].octave;
            ang[i] = LastFrame.mvKeysUn[i].angle;
            if (valid[i]) {
                obs[i] = p->Observations() > 0;
                coeb_adapt::xyz3(p->GetWorldPos(), &xyz[(size_t)i * 3]);
                std::memcpy(&desc[(size_t)i * 32], coeb_adapt::desc_row(p->GetDescriptor(), 0), 32);
            }
        }
        float Tc[12], Tl[12];
        coeb_adapt::pose34(CurrentFrame.mTcw, Tc);
        coeb_adapt::pose34(LastFrame.mTcw, Tl);
        std::vector<int> state(K);
        for (int k = 0; k < K; k++)
            state[k] = !CurrentFrame.mvpMapPoints[k] ? -1 : (CurrentFrame.mvpMapPoints[k]->Observations() > 0 ? -2 : -3);
        coeb_adapt::DeviceFrame<FrameT> dC(CurrentFrame);
        int nmatches = 0;
        coeb_adapt::check(coeb_match_lastframe(coeb_adapt::matcher(), dC.f, n, valid.data(), obs.data(), xyz.data(), oct.data(), ang.data(),
                                               desc.data(), Tc, Tl, th, bMono ? 1 : 0, mbCheckOrientation ? 1 : 0, state.data(), &nmatches));
        for (int k = 0; k < K; k++) {
            if (state[k] >= 0) CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[state[k]];
            else if (state[k] == -1) CurrentFrame.mvpMapPoints[k] = nullptr;   /* cleared by the rotation check (:1463) */
        }
        return nmatches;
    }

    /* SearchByProjection(Frame &CurrentFrame, KeyFrame* pKF, const set<MapPoint*> &sAlreadyFound, const float th, const int ORBdist)
     * -- :1473-1600 (Tracking::Relocalization). MapPoint needs the GetMinDistance() / GetMaxDistance() getters (see MakeLocalMap). */
    template <class FrameT, class KeyFrameT, class SetT>
    int SearchByProjection(FrameT& CurrentFrame, KeyFrameT* pKF, const SetT& sAlreadyFound, const float th, const int ORBdist) {
        const auto vpMPs = pKF->GetMapPointMatches();
        const int n = (int)vpMPs.size(), K = (int)CurrentFrame.mvpMapPoints.size();
        std::vector<uint8_t> valid(n), desc((size_t)n * 32);
        std::vector<float> xyz((size_t)n * 3), dmin(n), dmax(n), ang(n);
        for (int i = 0; i < n; i++) {
            auto* p = vpMPs[i];
            valid[i] = p && !p->isBad() && !sAlreadyFound.count(p);
            ang[i] = pKF->mvKeysUn[i].angle;
            if (valid[i]) {
                coeb_adapt::xyz3(p->GetWorldPos(), &xyz[(size_t)i * 3]);
                dmin[i] = p->GetMinDistance(); dmax[i] = p->GetMaxDistance();
                std::memcpy(&desc[(size_t)i * 32], coeb_adapt::desc_row(p->GetDescriptor(), 0), 32);
            }
        }
        float Tc[12], Ow[3];
        coeb_adapt::pose34(CurrentFrame.mTcw, Tc);
        coeb_adapt::xyz3(CurrentFrame.mOw, Ow);   // -Rcw^T tcw, the expression of :1479
        std::vector<int> state(K);
        for (int k = 0; k < K; k++) state[k] = CurrentFrame.mvpMapPoints[k] ? -2 : -1;
        coeb_adapt::DeviceFrame<FrameT> dC(CurrentFrame);
        int nmatches = 0;
        coeb_adapt::check(coeb_match_reloc(coeb_adapt::matcher(), dC.f, n, valid.data(), xyz.data(), dmin.data(), dmax.data(), ang.data(), desc.data(),
                                           Tc, Ow, th, ORBdist, mbCheckOrientation ? 1 : 0, state.data(), &nmatches));
        for (int k = 0; k < K; k++)
            if (state[k] >= 0) CurrentFrame.mvpMapPoints[k] = vpMPs[state[k]];
        return nmatches;
    }

    /* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize=10) -- :405-520 */
    template <class FrameT, class PointT>
    int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<PointT>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10) {
        static_assert(sizeof(PointT) == 2 * sizeof(float), "vbPrevMatched must hold packed (x, y) floats");
        vnMatches12.assign(F1.mvKeysUn.size(), -1);
        coeb_adapt::DeviceFrame<FrameT> d1(F1), d2(F2);
        int nmatches = 0;
        coeb_adapt::check(coeb_match_init(coeb_adapt::matcher(), d1.f, d2.f, reinterpret_cast<float*>(vbPrevMatched.data()),
                                          vnMatches12.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, &nmatches));
        return nmatches;
    }

    /* SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) -- :158-288 (Tracking::TrackReferenceKeyFrame,
     * Relocalization). */
    template <class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches) {
        const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches.assign(F.N, static_cast<MapPointT*>(nullptr));
        const int n1 = (int)vpMapPointsKF.size();
        std::vector<uint8_t> valid1(n1);
        for (int i = 0; i < n1; i++) valid1[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
        const coeb_adapt::FeatVecCSR v1(pKF->mFeatVec), v2(F.mFeatVec);
        coeb_adapt::DeviceFrame<KeyFrameT> d1(*pKF);
        coeb_adapt::DeviceFrame<FrameT> d2(F);
        std::vector<int> m12(n1, -1);
        int nmatches = 0;
        coeb_adapt::check(coeb_match_bow(coeb_adapt::matcher(), d1.f, d2.f, valid1.data(), nullptr, v1.nn(), v1.node.data(), v1.start.data(),
                                         v1.items.data(), v2.nn(), v2.node.data(), v2.start.data(), v2.items.data(), mfNNratio,
                                         mbCheckOrientation ? 1 : 0, /*strict_low*/ 0, m12.data(), &nmatches));
        for (int i = 0; i < n1; i++)
            if (m12[i] >= 0) vpMapPointMatches[m12[i]] = vpMapPointsKF[i];
        return nmatches;
    }

    /* SearchByBoW(KeyFrame *pKF1, KeyFrame* pKF2, vector<MapPoint*> &vpMatches12) -- :522-655 (LoopClosing::ComputeSim3). */
    template <class KeyFrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12) {
        const std::vector<MapPointT*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
        vpMatches12.assign(vp1.size(), static_cast<MapPointT*>(nullptr));
        std::vector<uint8_t> valid1(vp1.size()), valid2(vp2.size());
        for (size_t i = 0; i < vp1.size(); i++) valid1[i] = vp1[i] && !vp1[i]->isBad();
        for (size_t i = 0; i < vp2.size(); i++) valid2[i] = vp2[i] && !vp2[i]->isBad();
        const coeb_adapt::FeatVecCSR v1(pKF1->mFeatVec), v2(pKF2->mFeatVec);
        coeb_adapt::DeviceFrame<KeyFrameT> d1(*pKF1), d2(*pKF2);
        std::vector<int> m12(vp1.size(), -1);
        int nmatches = 0;
        coeb_adapt::check(coeb_match_bow(coeb_adapt::matcher(), d1.f, d2.f, valid1.data(), valid2.data(), v1.nn(), v1.node.data(),
                                         v1.start.data(), v1.items.data(), v2.nn(), v2.node.data(), v2.start.data(), v2.items.data(),
                                         mfNNratio, mbCheckOrientation ? 1 : 0, /*strict_low*/ 1, m12.data(), &nmatches));
        for (size_t i = 0; i < vp1.size(); i++)
            if (m12[i] >= 0) vpMatches12[i] = vp2[m12[i]];
        return nmatches;
    }

    /* SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) -- :657-824 (LocalMapping::CreateNewMapPoints). The
     * epipole (:663-670) is evaluated here like the reference's cv::Mat expression R2w * Cw + t2w (fp32, left to right). */
    template <class KeyFrameT, class MatT>
    int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, const MatT& F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo) {
        float Cw[3], R2w[9], t2w[3], Fm[9], C2[3];
        coeb_adapt::xyz3(pKF1->GetCameraCenter(), Cw);
        coeb_adapt::mat33(pKF2->GetRotation(), R2w);
        coeb_adapt::xyz3(pKF2->GetTranslation(), t2w);
        coeb_adapt::mat33(F12, Fm);
        for (int r = 0; r < 3; r++) C2[r] = R2w[3 * r] * Cw[0] + R2w[3 * r + 1] * Cw[1] + R2w[3 * r + 2] * Cw[2] + t2w[r];
        const float invz = 1.0f / C2[2];
        const float epipole[2] = {pKF2->fx * C2[0] * invz + pKF2->cx, pKF2->fy * C2[1] * invz + pKF2->cy};
        const int n1 = pKF1->N, n2 = pKF2->N;
        std::vector<uint8_t> free1(n1), free2(n2);
        for (int i = 0; i < n1; i++) free1[i] = !pKF1->GetMapPoint(i);
        for (int i = 0; i < n2; i++) free2[i] = !pKF2->GetMapPoint(i);
        const coeb_adapt::FeatVecCSR v1(pKF1->mFeatVec), v2(pKF2->mFeatVec);
        coeb_adapt::DeviceFrame<KeyFrameT> d1(*pKF1), d2(*pKF2);
        std::vector<int> m12(n1, -1);
        int nmatches = 0;
        coeb_adapt::check(coeb_match_triangulation(coeb_adapt::matcher(), d1.f, d2.f, free1.data(), free2.data(), v1.nn(), v1.node.data(),
                                                   v1.start.data(), v1.items.data(), v2.nn(), v2.node.data(), v2.start.data(), v2.items.data(),
                                                   Fm, epipole, bOnlyStereo ? 1 : 0, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches));
        vMatchedPairs.clear();
        vMatchedPairs.reserve(nmatches);
        for (int i = 0; i < n1; i++)
            if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
        return nmatches;
    }

    /* Fuse(KeyFrame* pKF, const vector<MapPoint*> &vpMapPoints, const float th=3.0) -- :826-961 (LocalMapping::SearchInNeighbors).
     * The search runs on the device for all map points at once; the MapPoint side effects are then applied in list order exactly
     * as the reference's loop does (:938-957): nothing they change is read by the search. */
    template <class KeyFrameT, class MapPointT>
    int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, const float th = 3.0) {
        const int n = (int)vpMapPoints.size();
        std::vector<uint8_t> valid(n);
        std::vector<MapPointT*> pts(n);
        MapPointT* any = nullptr;
        for (int i = 0; i < n; i++) {
            MapPointT* p = vpMapPoints[i];
            valid[i] = p && !p->isBad() && !p->IsInKeyFrame(pKF);
            if (valid[i] && !any) any = p;
        }
        if (!any) return 0;
        for (int i = 0; i < n; i++) pts[i] = valid[i] ? vpMapPoints[i] : any;   // invalid slots only need well-formed fields
        coeb_local_map* lm = MakeLocalMap(pts);
        float T[12], R[9], t[3], Ow[3];
        coeb_adapt::mat33(pKF->GetRotation(), R);
        coeb_adapt::xyz3(pKF->GetTranslation(), t);
        coeb_adapt::xyz3(pKF->GetCameraCenter(), Ow);
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T[4 * r + c] = R[3 * r + c]; T[4 * r + 3] = t[r]; }
        coeb_adapt::DeviceFrame<KeyFrameT> dK(*pKF);
        std::vector<int> best(n, -1);
        int nFused = 0;
        const int st = coeb_fuse_search(coeb_adapt::matcher(), dK.f, lm, valid.data(), T, Ow, th, /*chi2_tests*/ 1, best.data(), &nFused);
        coeb_local_map_destroy(lm);
        coeb_adapt::check(st);
        for (int i = 0; i < n; i++) {
            if (best[i] < 0) continue;
            MapPointT* pMP = vpMapPoints[i];
            MapPointT* pMPinKF = pKF->GetMapPoint(best[i]);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) {
                    if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                    else pMPinKF->Replace(pMP);
                }
            } else {
                pMP->AddObservation(pKF, best[i]);
                pKF->AddMapPoint(pMP, best[i]);
            }
        }
        return nFused;
    }

    static const int TH_LOW = COEB_TH_LOW;
    static const int TH_HIGH = COEB_TH_HIGH;
    static const int HISTO_LENGTH = COEB_HISTO_LENGTH;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

/* Frame::ComputeStereoMatches (src/Frame.cc:644-818) for a Frame that holds both extractors. Writes mvuRight / mvDepth. */
template <class FrameT>
inline void ComputeStereoMatches(FrameT& F) {
    const int N = (int)F.mvKeys.size(), Nr = (int)F.mvKeysRight.size();
    F.mvuRight.assign(N, -1.0f);
    F.mvDepth.assign(N, -1.0f);
    if (N == 0 || Nr == 0) return;
    std::vector<unsigned char> dl((size_t)N * 32), dr((size_t)Nr * 32);
    for (int i = 0; i < N; i++) std::memcpy(&dl[(size_t)i * 32], coeb_adapt::desc_row(F.mDescriptors, i), 32);
    for (int i = 0; i < Nr; i++) std::memcpy(&dr[(size_t)i * 32], coeb_adapt::desc_row(F.mDescriptorsRight, i), 32);
    int nmatched = 0;
    coeb_adapt::check(coeb_stereo_match(coeb_adapt::matcher(), F.mpORBextractorLeft->handle(), F.mpORBextractorRight->handle(), N,
                                        reinterpret_cast<const coeb_keypoint*>(F.mvKeys.data()), dl.data(), Nr,
                                        reinterpret_cast<const coeb_keypoint*>(F.mvKeysRight.data()), dr.data(), F.mbf, F.mb,
                                        F.mvuRight.data(), F.mvDepth.data(), &nmatched));
}

/* Tail of the RGB-D Frame constructor (src/Frame.cc:213-240) on the device, from the extractor call that has just filled
 * F.mvKeys / F.mDescriptors (src/Frame.cc:416): writes F.mvKeysUn (UndistortKeyPoints, :579-609), F.mvuRight / F.mvDepth
 * (ComputeStereoFromRGBD, :820-842) and returns the device-resident frame (grid built, AssignFeaturesToGrid :396-411) that
 * the coeb_match_* / coeb_search_local_points entry points take; the caller owns it (coeb_frame_destroy). dist5 = mDistCoef
 * as {k1, k2, p1, p2, k3}; depth as Tracking::GrabImageRGBD holds it (raw CV_16U + mDepthMapFactor, or CV_32F). */
template <class FrameT>
inline coeb_frame* FrameTailFromExtractor(FrameT& F, const float* dist5, const coeb_depth_image* depth) {
    const int N = (int)F.mvKeys.size();
    coeb_camera cam;
    cam.fx = F.fx; cam.fy = F.fy; cam.cx = F.cx; cam.cy = F.cy; cam.bf = F.mbf; cam.b = F.mb;
    cam.min_x = F.mnMinX; cam.max_x = F.mnMaxX; cam.min_y = F.mnMinY; cam.max_y = F.mnMaxY;
    F.mvKeysUn.resize(N);
    F.mvuRight.assign(N, -1.0f);
    F.mvDepth.assign(N, -1.0f);
    coeb_frame* f = nullptr;
    int n = 0;
    coeb_adapt::check(coeb_frame_from_extractor(coeb_adapt::matcher(), F.mpORBextractorLeft->handle(), 0, N, &cam, dist5, depth,
                                                reinterpret_cast<coeb_keypoint*>(F.mvKeysUn.data()), F.mvuRight.data(), F.mvDepth.data(), &n, &f));
    return f;
}

/* Tracking::SearchLocalPoints (src/Tracking.cc:1222-1272) against a device-resident local map built from
 * mvpLocalMapPoints by MakeLocalMap (rebuild it when Tracking::UpdateLocalPoints changes the vector). dF is the current
 * frame on the device (FrameTailFromExtractor or coeb_adapt::DeviceFrame). Returns the matcher's count; the MapPoint
 * bookkeeping of the reference's loops (IncreaseVisible, mnLastFrameSeen, mbTrackInView) is applied here. */
template <class MapPointT>
inline coeb_local_map* MakeLocalMap(const std::vector<MapPointT*>& vpLocalMapPoints) {
    const size_t n = vpLocalMapPoints.size();
    std::vector<float> xyz(3 * n), nrm(3 * n), dmin(n), dmax(n);
    std::vector<unsigned char> desc(32 * n);
    for (size_t i = 0; i < n; i++) {
        MapPointT* p = vpLocalMapPoints[i];
        coeb_adapt::xyz3(p->GetWorldPos(), &xyz[3 * i]);
        coeb_adapt::xyz3(p->GetNormal(), &nrm[3 * i]);
        dmin[i] = p->GetMinDistance();   // mfMinDistance / mfMaxDistance themselves (protected in include/MapPoint.h: add the two
        dmax[i] = p->GetMaxDistance();   // one-line getters; the 0.8 / 1.2 factors of src/MapPoint.cc:373-383 are applied on the device)
        std::memcpy(&desc[32 * i], coeb_adapt::desc_row(p->GetDescriptor(), 0), 32);
    }
    coeb_local_map* lm = nullptr;
    coeb_adapt::check(coeb_local_map_create(coeb_adapt::matcher(), (int)n, xyz.data(), nrm.data(), dmin.data(), dmax.data(), desc.data(), &lm));
    return lm;
}

template <class FrameT, class MapPointT>
inline int SearchLocalPoints(FrameT& F, coeb_frame* dF, coeb_local_map* lm, const std::vector<MapPointT*>& vpLocalMapPoints, float th,
                             float nnratio = 0.8f) {
    // first loop (:1225-1241): points already matched in this frame are not searched again
    for (auto& pMP : F.mvpMapPoints)
        if (pMP) {
            if (pMP->isBad()) pMP = nullptr;
            else { pMP->IncreaseVisible(); pMP->mnLastFrameSeen = F.mnId; pMP->mbTrackInView = false; }
        }
    const int n = (int)vpLocalMapPoints.size(), K = (int)F.mvpMapPoints.size();
    std::vector<uint8_t> skip(n), obs(n), in_view(n);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpLocalMapPoints[i];
        skip[i] = p->mnLastFrameSeen == F.mnId || p->isBad();   // :1249-1252
        obs[i] = p->Observations() > 0;
    }
    std::vector<int> state(K);
    for (int k = 0; k < K; k++) state[k] = !F.mvpMapPoints[k] ? -1 : (F.mvpMapPoints[k]->Observations() > 0 ? -2 : -3);
    float Tcw[12], Ow[3];
    coeb_adapt::pose34(F.mTcw, Tcw);
    coeb_adapt::xyz3(F.mOw, Ow);
    int nmatches = 0;
    coeb_adapt::check(coeb_search_local_points(coeb_adapt::matcher(), dF, lm, skip.data(), obs.data(), Tcw, Ow, 0.5f, th, nnratio, state.data(),
                                               in_view.data(), nullptr, &nmatches));
    for (int i = 0; i < n; i++) {
        if (skip[i]) continue;
        vpLocalMapPoints[i]->mbTrackInView = in_view[i] != 0;   // Frame::isInFrustum (:447, :492)
        if (in_view[i]) vpLocalMapPoints[i]->IncreaseVisible();  // :1254-1258
    }
    for (int k = 0; k < K; k++)
        if (state[k] >= 0) F.mvpMapPoints[k] = vpLocalMapPoints[state[k]];
    return nmatches;
}

}  // namespace ORB_SLAM2

#endif  // ORBMATCHER_H
