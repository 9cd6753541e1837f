/*
 * ORBmatcher.h -- drop-in ORB_SLAM2::ORBmatcher for the hot-path overloads, backed by the B200 C ABI
 * (coeb_frontend.h). Same constructor, statics and call signatures as the reference's
 * include/ORBmatcher.h:41-69,87-89 for
 *     DescriptorDistance, SearchByProjection(Frame&, vector<MapPoint*>&, th),
 *     SearchByProjection(Frame& cur, const Frame& last, th, bMono), SearchForInitialization(...)
 * plus ComputeStereoMatches as a free function (the reference keeps it in Frame, src/Frame.cc:644-818).
 * The BoW / Fuse / Sim3 / triangulation overloads need DBoW2 and KeyFrame graph state and stay the
 * reference's.
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
#endif
inline const unsigned char* desc_row(const coeb_cv::Mat& m, int i) { return m.ptr(i); }
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

/* RAII upload of the Frame fields the matchers read (mvKeysUn, mDescriptors, mvuRight, bounds, scale factors). */
template <class FrameT>
struct DeviceFrame {
    coeb_frame* f = nullptr;
    explicit DeviceFrame(const FrameT& F) {
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
    ~DeviceFrame() { coeb_frame_destroy(f); }
    DeviceFrame(const DeviceFrame&) = delete;
    DeviceFrame& operator=(const DeviceFrame&) = delete;
};

}  // namespace coeb_adapt

namespace ORB_SLAM2 {

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

}  // namespace ORB_SLAM2

#endif  // ORBMATCHER_H
