// dropin_test.cpp -- exercises the C++ drop-in classes (include/ORBextractor.h, include/ORBmatcher.h) the way the
// reference's Frame / Tracking code calls them, with mock Frame / MapPoint types that carry the same member names,
// and compares every result with the CPU oracle (linked here as the checker). Run on a B200 by tests/test_dropin_gpu.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "../../include/ORBextractor.h"
#include "../../include/ORBmatcher.h"

// ---- oracle C entry points (oracle/coeb_oracle_c.cpp) ----
extern "C" {
struct orc_extractor;
struct orc_frame;
int orc_level_image(orc_extractor* e, int level, int which, uint8_t* dst);
orc_extractor* orc_extractor_create(const coeb_orb_params* p);
void orc_extractor_destroy(orc_extractor* e);
int orc_extract(orc_extractor* e, const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox, const float* tm, int ntm,
                const int* blur_flag, int nblur, coeb_keypoint* kps_out, uint8_t* desc_out, int cap, int* n_out);
orc_frame* orc_frame_create(const coeb_keypoint* kps, const uint8_t* desc, int n, const float* uright, const coeb_camera* cam,
                            const float* scale, int nlevels);
void orc_frame_destroy(orc_frame* f);
int orc_match_projection(orc_frame* f, int n, const uint8_t* track_in_view, const uint8_t* bad, const uint8_t* has_obs, const float* proj_x,
                         const float* proj_y, const float* proj_xr, const int* level, const float* view_cos, const uint8_t* desc, float th,
                         float nnratio, int* kp_match);
int orc_match_lastframe(orc_frame* cur, int n, const uint8_t* valid, const uint8_t* has_obs, const float* xyz, const int* octave,
                        const float* angle, const uint8_t* desc, const float* Tcw_cur, const float* Tcw_last, float th, int mono,
                        int check_ori, int* kp_match);
int orc_match_init(orc_frame* f1, orc_frame* f2, float* prev_matched, int* matches12, int window, float nnratio, int check_ori);
int orc_stereo_match(orc_extractor* exL, orc_extractor* exR, int N, const coeb_keypoint* keysL, const uint8_t* descL, int Nr,
                     const coeb_keypoint* keysR, const uint8_t* descR, float mbf, float mb, float* uright, float* depth);
int orc_match_bow(orc_frame* f1, orc_frame* f2, const uint8_t* valid1, const uint8_t* valid2, int nn1, const int* node1, const int* start1,
                  const int* items1, int nn2, const int* node2, const int* start2, const int* items2, float nnratio, int check_ori,
                  int strict_low, int* match12);
int orc_match_triangulation(orc_frame* f1, orc_frame* f2, const uint8_t* free1, const uint8_t* free2, int nn1, const int* node1,
                            const int* start1, const int* items1, int nn2, const int* node2, const int* start2, const int* items2,
                            const float* F12, float ex, float ey, int only_stereo, int check_ori, int* match12);
void orc_undistort_keypoints(const coeb_keypoint* keys, int n, const coeb_camera* cam, const float* dist5, coeb_keypoint* keys_un);
void orc_stereo_from_rgbd(const coeb_keypoint* keys, const coeb_keypoint* keys_un, int n, const void* depth, int kind, int stride_bytes,
                          float factor, float mbf, float* uright, float* depth_out);
int orc_match_reloc(orc_frame* cur, int n, const uint8_t* valid, const float* xyz, const float* min_dist, const float* max_dist,
                    const float* angle, const uint8_t* desc, const float* Tcw, const float* Ow, float th, int orb_dist, int check_ori,
                    int* kp_match);
int orc_fuse_search(orc_frame* f, int n, const float* xyz, const float* normal, const float* min_dist, const float* max_dist,
                    const uint8_t* desc, const uint8_t* valid, const float* Tcw, const float* Ow, float th, int chi2_tests, int* best_idx);
int orc_search_local_points(orc_frame* f, int n, const float* xyz, const float* normal, const float* min_dist, const float* max_dist,
                            const uint8_t* desc, const uint8_t* skip, const uint8_t* has_obs, const float* Tcw, const float* Ow,
                            float cos_limit, float th, float nnratio, int* kp_match, uint8_t* in_view, float* proj);
}

using coeb_cv::KeyPoint;
using coeb_cv::Mat;
using coeb_cv::Point2f;

static uint32_t g_rng = 12345;
static uint32_t rnd() { g_rng = g_rng * 1664525u + 1013904223u; return g_rng >> 8; }
static float rndf() { return (rnd() & 0xFFFF) / 65536.f; }

static std::vector<uint8_t> make_image(int w, int h, uint32_t seed) {
    g_rng = seed;
    std::vector<float> img((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) img[(size_t)y * w + x] = 110.f + 40.f * std::sin(x * 0.013f) * std::cos(y * 0.017f);
    for (int s = 0; s < 350; s++) {
        int sw = 6 + rnd() % 50, sh = 6 + rnd() % 50, x0 = rnd() % (w - sw), y0 = rnd() % (h - sh);
        float c = (20 + rnd() % 100) * ((rnd() & 1) ? 1.f : -1.f);
        for (int y = y0; y < y0 + sh; y++)
            for (int x = x0; x < x0 + sw; x++) img[(size_t)y * w + x] += c;
    }
    std::vector<uint8_t> out((size_t)w * h);
    for (size_t i = 0; i < out.size(); i++) {
        float v = img[i] + (float)((int)(rnd() % 5) - 2);
        out[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
    return out;
}

// ---- mock reference types: only the members the hot-path functions touch ----
struct MapPoint {
    bool mbTrackInView = true, bad = false;
    int nObs = 1, mnTrackScaleLevel = 0;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 1;
    Mat desc;
    float pos[3] = {0, 0, 1}, normal[3] = {0, 0, -1}, mfMinDistance = 0.1f, mfMaxDistance = 10.f;
    long mnLastFrameSeen = -1;
    int nVisible = 0;
    bool isBad() { return bad; }
    int Observations() { return nObs; }
    const Mat& GetDescriptor() { return desc; }
    const float* GetWorldPos() { return pos; }
    const float* GetNormal() { return normal; }
    float GetMinDistance() { return mfMinDistance; }
    float GetMaxDistance() { return mfMaxDistance; }
    void IncreaseVisible() { nVisible++; }
    // Fuse side effects (src/MapPoint.cc): recorded, not simulated
    const void* inKF = nullptr; int obsIdx = -1; MapPoint* replacedBy = nullptr;
    bool IsInKeyFrame(const void* kf) { return inKF == kf; }
    void AddObservation(const void* kf, int idx) { inKF = kf; obsIdx = idx; nObs++; }
    void Replace(MapPoint* other) { replacedBy = other; bad = true; }
};

struct Frame {
    int N = 0;
    std::vector<KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvuRight, mvDepth, mvScaleFactors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    float fx = 535.4f, fy = 539.2f, cx = 320.1f, cy = 247.6f, mbf = 40.f, mb = 40.f / 535.4f;
    float mnMinX = 0, mnMaxX = 640, mnMinY = 0, mnMaxY = 480;
    float mTcw[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, mOw[3] = {0, 0, 0};
    long mnId = 7;
    std::map<unsigned, std::vector<unsigned> > mFeatVec;   // DBoW2::FeatureVector
    ORB_SLAM2::ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    coeb_frame* mpDeviceFrame = nullptr;   // device twin kept by the Frame; the adapters use it when set
    // KeyFrame-only members used by SearchByBoW / SearchForTriangulation (a KeyFrame is built from a Frame, src/KeyFrame.cc:31-58)
    float Rcw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tcw[3] = {0, 0, 0};
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(int i) { return mvpMapPoints[i]; }
    void AddMapPoint(MapPoint* p, int i) { mvpMapPoints[i] = p; }
    const float* GetCameraCenter() { return mOw; }
    const float* GetRotation() { return Rcw; }
    const float* GetTranslation() { return tcw; }
};

// stand-in vocabulary: node = index of the nearest of 64 pseudo-random 256-bit centres
static void make_featvec(Frame& F) {
    static std::vector<uint8_t> centres;
    if (centres.empty()) { uint32_t save = g_rng; g_rng = 4242; centres.resize(64 * 32); for (auto& b : centres) b = (uint8_t)rnd(); g_rng = save; }
    F.mFeatVec.clear();
    for (int i = 0; i < F.N; i++) {
        int best = 0, bd = 999;
        for (int c = 0; c < 64; c++) {
            int d = 0;
            for (int b = 0; b < 32; b++) d += __builtin_popcount((unsigned)(F.mDescriptors.ptr(i)[b] ^ centres[c * 32 + b]));
            if (d < bd) { bd = d; best = c; }
        }
        F.mFeatVec[(unsigned)(5 * best + 1)].push_back((unsigned)i);
    }
}

static int g_fail = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { g_fail++; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

static coeb_camera cam_of(const Frame& F) {
    coeb_camera c;
    c.fx = F.fx; c.fy = F.fy; c.cx = F.cx; c.cy = F.cy; c.bf = F.mbf; c.b = F.mb; c.min_x = F.mnMinX; c.max_x = F.mnMaxX; c.min_y = F.mnMinY; c.max_y = F.mnMaxY;
    return c;
}

static void fill_frame(Frame& F, ORB_SLAM2::ORBextractor& ex, const std::vector<uint8_t>& img, int w, int h,
                       std::vector<std::vector<float> >& box, std::vector<Point2f>& tm, std::vector<int>& blur) {
    Mat gray = coeb_adapt::wrap_u8(h, w, const_cast<uint8_t*>(img.data()), (size_t)w), none, mask_result;
    ex(gray, none, none, none, F.mvKeys, F.mDescriptors, box, tm, mask_result, blur);   // src/Frame.cc:416
    F.N = (int)F.mvKeys.size();
    F.mvKeysUn = F.mvKeys;                                                                // k1 == 0 (src/Frame.cc:581-585)
    F.mvScaleFactors = ex.GetScaleFactors();
    F.mvpMapPoints.assign(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    F.mvuRight.assign(F.N, -1.f);
    F.mnMaxX = (float)w; F.mnMaxY = (float)h;
}

int main() {
    const int w = 640, h = 480;
    coeb_orb_params p = {1000, 1.2f, 8, 20, 7};
    ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    orc_extractor* oex = orc_extractor_create(&p);

    // ---- extraction with injected person boxes, COEB call form ----
    std::vector<uint8_t> img = make_image(w, h, 7);
    std::vector<std::vector<float> > box = {{100, 60, 300, 400}, {350, 100, 500, 380}};
    std::vector<Point2f> tm = {Point2f(150.5f, 100.2f), Point2f(400.7f, 200.9f), Point2f(20.f, 20.f)};
    std::vector<int> blur = {1, 0};
    Frame F;
    fill_frame(F, ex, img, w, h, box, tm, blur);
    std::vector<coeb_keypoint> okp(4096);
    std::vector<uint8_t> odesc(4096 * 32);
    int on = 0;
    const float boxes_flat[8] = {100, 60, 300, 400, 350, 100, 500, 380};
    const float tm_flat[6] = {150.5f, 100.2f, 400.7f, 200.9f, 20.f, 20.f};
    orc_extract(oex, img.data(), w, h, w, boxes_flat, 2, tm_flat, 3, blur.data(), 2, okp.data(), odesc.data(), 4096, &on);
    EXPECT(on == F.N && on > 300, "keypoint count %d vs oracle %d", F.N, on);
    EXPECT(on == F.N && std::memcmp(F.mvKeys.data(), okp.data(), sizeof(coeb_keypoint) * on) == 0, "keypoints differ");
    EXPECT(on == F.N && std::memcmp(F.mDescriptors.ptr(0), odesc.data(), (size_t)32 * on) == 0, "descriptors differ");
    // classic 4-argument form == COEB form without boxes
    {
        std::vector<KeyPoint> k4; Mat d4, gray = coeb_adapt::wrap_u8(h, w, img.data(), (size_t)w), none;
        ex(gray, none, k4, d4);
        int n4 = 0;
        orc_extract(oex, img.data(), w, h, w, nullptr, 0, nullptr, 0, nullptr, 0, okp.data(), odesc.data(), 4096, &n4);
        EXPECT(n4 == (int)k4.size() && std::memcmp(k4.data(), okp.data(), sizeof(coeb_keypoint) * n4) == 0, "4-arg operator() differs");
        EXPECT(ex.GetLevels() == 8 && ex.GetScaleFactors().size() == 8 && std::fabs(ex.GetScaleFactor() - 1.2f) < 1e-6f, "getters");
        // the public pyramid member fetches a level the first time it is read after an extraction (src/Frame.cc:651, 741 read it)
        EXPECT(ex.mvImagePyramid.size() == 8 && ex.mvImagePyramid[1].cols == 533 && ex.mvImagePyramid[7].rows == 134, "pyramid geometry");
        EXPECT(std::memcmp(ex.mvImagePyramid[0].ptr(100), img.data() + 100 * w, w) == 0, "level 0 of the lazy pyramid is not the input image");
        {
            std::vector<uint8_t> lvl(533 * 400);
            orc_level_image(oex, 1, 0, lvl.data());
            bool same = true;
            for (int y = 0; y < 400; y++) same = same && std::memcmp(ex.mvImagePyramid[1].ptr(y), lvl.data() + (size_t)y * 533, 533) == 0;
            EXPECT(same, "level 1 of the lazy pyramid differs from the oracle");
        }
        // empty image: silent return, outputs untouched (src/ORBextractor.cc:1096-1097)
        Mat empty; size_t before = k4.size();
        ex(empty, none, k4, d4);
        EXPECT(k4.size() == before, "empty image must be a no-op");
        // box outside the image: the reference's cv::Mat ROI throws; so does the drop-in
        bool threw = false;
        std::vector<std::vector<float> > badbox = {{600, 400, 700, 500}};
        std::vector<Point2f> t1 = {Point2f(610.f, 410.f)};
        std::vector<int> b1 = {1};
        Mat mr;
        try { ex(gray, none, none, none, k4, d4, badbox, t1, mr, b1); } catch (const std::exception&) { threw = true; }
        EXPECT(threw, "out-of-image box must throw");
    }
    // re-extract F (the calls above reused the extractor)
    fill_frame(F, ex, img, w, h, box, tm, blur);
    orc_extract(oex, img.data(), w, h, w, boxes_flat, 2, tm_flat, 3, blur.data(), 2, okp.data(), odesc.data(), 4096, &on);

    // ---- SearchByProjection(Frame&, vector<MapPoint*>&, th): local-map tracking (src/Tracking.cc:1261-1270) ----
    const int NM = 3000;
    std::vector<MapPoint> mps(NM);
    std::vector<MapPoint*> vp(NM);
    for (int i = 0; i < NM; i++) {
        MapPoint& m = mps[i];
        const int k = rnd() % F.N;
        coeb_adapt::create_u8(m.desc, 1, 32);
        std::memcpy(m.desc.ptr(0), F.mDescriptors.ptr(k), 32);
        for (int b = 0, nb = rnd() % 30; b < nb; b++) { int bit = rnd() % 256; m.desc.ptr(0)[bit >> 3] ^= (uint8_t)(1 << (bit & 7)); }
        m.mTrackProjX = F.mvKeysUn[k].pt.x + (rndf() - 0.5f) * 8.f;
        m.mTrackProjY = F.mvKeysUn[k].pt.y + (rndf() - 0.5f) * 8.f;
        m.mTrackProjXR = m.mTrackProjX - 10.f;
        m.mnTrackScaleLevel = std::min(F.mvKeysUn[k].octave + (int)(rnd() & 1), 7);
        m.mTrackViewCos = 0.9f + 0.1f * rndf();
        m.mbTrackInView = (rnd() % 20) != 0;
        m.bad = (rnd() % 50) == 0;
        m.nObs = (rnd() % 30) == 0 ? 0 : 2;
        vp[i] = &m;
    }
    MapPoint held_obs, held_noobs;
    held_noobs.nObs = 0;
    for (int k = 0; k < F.N; k += 7) F.mvpMapPoints[k] = (k % 14) ? &held_obs : &held_noobs;
    {
        coeb_camera cam = cam_of(F);
        orc_frame* of = orc_frame_create(okp.data(), odesc.data(), on, F.mvuRight.data(), &cam, F.mvScaleFactors.data(), 8);
        std::vector<uint8_t> tiv(NM), bad(NM), obs(NM), desc((size_t)NM * 32);
        std::vector<float> px(NM), py(NM), pxr(NM), vc(NM);
        std::vector<int> lvl(NM), state(F.N);
        for (int i = 0; i < NM; i++) {
            tiv[i] = mps[i].mbTrackInView; bad[i] = mps[i].bad; obs[i] = mps[i].nObs > 0; px[i] = mps[i].mTrackProjX; py[i] = mps[i].mTrackProjY;
            pxr[i] = mps[i].mTrackProjXR; vc[i] = mps[i].mTrackViewCos; lvl[i] = mps[i].mnTrackScaleLevel;
            std::memcpy(&desc[(size_t)i * 32], mps[i].desc.ptr(0), 32);
        }
        for (int k = 0; k < F.N; k++) state[k] = !F.mvpMapPoints[k] ? -1 : (F.mvpMapPoints[k]->nObs > 0 ? -2 : -3);
        const int nref = orc_match_projection(of, NM, tiv.data(), bad.data(), obs.data(), px.data(), py.data(), pxr.data(), lvl.data(), vc.data(),
                                              desc.data(), 3.f, 0.8f, state.data());
        ORB_SLAM2::ORBmatcher matcher(0.8f);
        const int ngot = matcher.SearchByProjection(F, vp, 3);
        EXPECT(nref == ngot && nref > 50, "SearchByProjection(map): %d vs oracle %d", ngot, nref);
        int bad_ptr = 0;
        for (int k = 0; k < F.N; k++) {
            MapPoint* want = state[k] >= 0 ? vp[state[k]] : (state[k] == -1 ? nullptr : (state[k] == -2 ? &held_obs : &held_noobs));
            bad_ptr += F.mvpMapPoints[k] != want;
        }
        EXPECT(bad_ptr == 0, "SearchByProjection(map): %d mvpMapPoints entries differ", bad_ptr);
        orc_frame_destroy(of);
    }

    // ---- SearchByProjection(cur, last, th, bMono): motion-model tracking (src/Tracking.cc:935-957) ----
    {
        Frame last = F;   // the last frame observed the same scene; its map points project near the current keypoints
        Frame cur;
        std::vector<std::vector<float> > nobox; std::vector<Point2f> notm; std::vector<int> noblur;
        fill_frame(cur, ex, img, w, h, nobox, notm, noblur);
        std::vector<MapPoint> lmp(last.N);
        const float a = 0.01f;
        const float R[9] = {std::cos(a), 0, std::sin(a), 0, 1, 0, -std::sin(a), 0, std::cos(a)}, t[3] = {0.02f, -0.01f, 0.03f};
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) cur.mTcw[4 * r + c] = R[3 * r + c]; cur.mTcw[4 * r + 3] = t[r]; }
        for (int i = 0; i < last.N; i++) {
            MapPoint& m = lmp[i];
            const int k = rnd() % cur.N;
            const float z = 0.8f + 5.f * rndf();
            const float pc[3] = {(cur.mvKeysUn[k].pt.x - cur.cx) * z / cur.fx - t[0], (cur.mvKeysUn[k].pt.y - cur.cy) * z / cur.fy - t[1], z - t[2]};
            for (int c = 0; c < 3; c++) m.pos[c] = R[c] * pc[0] + R[3 + c] * pc[1] + R[6 + c] * pc[2];   // R^T (pc - t)
            coeb_adapt::create_u8(m.desc, 1, 32);
            std::memcpy(m.desc.ptr(0), cur.mDescriptors.ptr(k), 32);
            for (int b = 0, nb = rnd() % 25; b < nb; b++) { int bit = rnd() % 256; m.desc.ptr(0)[bit >> 3] ^= (uint8_t)(1 << (bit & 7)); }
            m.nObs = (rnd() % 25) == 0 ? 0 : 3;
            last.mvpMapPoints[i] = (rnd() % 5) ? &m : nullptr;
            last.mvbOutlier[i] = (rnd() % 11) == 0;
            last.mvKeys[i].octave = last.mvKeysUn[i].octave = cur.mvKeysUn[k].octave;
        }
        coeb_camera cam = cam_of(cur);
        orc_frame* oc = orc_frame_create(reinterpret_cast<coeb_keypoint*>(cur.mvKeysUn.data()), cur.mDescriptors.ptr(0), cur.N, cur.mvuRight.data(), &cam,
                                         cur.mvScaleFactors.data(), 8);
        std::vector<uint8_t> valid(last.N), obs(last.N), desc((size_t)last.N * 32);
        std::vector<float> xyz((size_t)last.N * 3), ang(last.N);
        std::vector<int> oct(last.N), state(cur.N, -1);
        for (int i = 0; i < last.N; i++) {
            MapPoint* m = last.mvpMapPoints[i];
            valid[i] = m && !last.mvbOutlier[i];
            oct[i] = last.mvKeys[i].octave; ang[i] = last.mvKeysUn[i].angle;
            if (valid[i]) { obs[i] = m->nObs > 0; std::memcpy(&xyz[(size_t)i * 3], m->pos, 12); std::memcpy(&desc[(size_t)i * 32], m->desc.ptr(0), 32); }
        }
        const int nref = orc_match_lastframe(oc, last.N, valid.data(), obs.data(), xyz.data(), oct.data(), ang.data(), desc.data(), cur.mTcw, last.mTcw,
                                             15.f, 0, 1, state.data());
        ORB_SLAM2::ORBmatcher matcher(0.9f, true);
        const int ngot = matcher.SearchByProjection(cur, last, 15, false);
        EXPECT(nref == ngot && nref > 50, "SearchByProjection(last): %d vs oracle %d", ngot, nref);
        int bad_ptr = 0;
        for (int k = 0; k < cur.N; k++) bad_ptr += cur.mvpMapPoints[k] != (state[k] >= 0 ? last.mvpMapPoints[state[k]] : nullptr);
        EXPECT(bad_ptr == 0, "SearchByProjection(last): %d mvpMapPoints entries differ", bad_ptr);

        // ---- SearchForInitialization (src/Tracking.cc:667-668) ----
        Frame F2;
        std::vector<uint8_t> img2((size_t)w * h);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) img2[(size_t)y * w + x] = img[(size_t)std::min(std::max(y - 6, 0), h - 1) * w + std::min(std::max(x - 12, 0), w - 1)];
        fill_frame(F2, ex, img2, w, h, nobox, notm, noblur);
        std::vector<Point2f> prev(cur.N);
        for (int i = 0; i < cur.N; i++) prev[i] = cur.mvKeysUn[i].pt;
        std::vector<float> prev_ref((size_t)cur.N * 2);
        std::memcpy(prev_ref.data(), prev.data(), sizeof(float) * 2 * cur.N);
        coeb_camera cam2 = cam_of(F2);
        orc_frame* o2 = orc_frame_create(reinterpret_cast<coeb_keypoint*>(F2.mvKeysUn.data()), F2.mDescriptors.ptr(0), F2.N, nullptr, &cam2,
                                         F2.mvScaleFactors.data(), 8);
        std::vector<int> m12ref(cur.N), m12;
        const int iref = orc_match_init(oc, o2, prev_ref.data(), m12ref.data(), 100, 0.9f, 1);
        const int igot = matcher.SearchForInitialization(cur, F2, prev, m12, 100);
        EXPECT(iref == igot && iref > 10, "SearchForInitialization: %d vs oracle %d", igot, iref);
        EXPECT(m12 == m12ref, "vnMatches12 differs");
        EXPECT(std::memcmp(prev.data(), prev_ref.data(), sizeof(float) * 2 * cur.N) == 0, "vbPrevMatched differs");
        orc_frame_destroy(oc);
        orc_frame_destroy(o2);
    }

    // ---- vocabulary-guided matchers + the steps either side of the path ----
    {
        std::vector<std::vector<float> > nobox; std::vector<Point2f> notm; std::vector<int> noblur;
        Frame K1, K2;
        std::vector<uint8_t> img2((size_t)w * h);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) img2[(size_t)y * w + x] = img[(size_t)std::min(std::max(y + 3, 0), h - 1) * w + std::min(std::max(x - 7, 0), w - 1)];
        fill_frame(K1, ex, img, w, h, nobox, notm, noblur);
        fill_frame(K2, ex, img2, w, h, nobox, notm, noblur);
        make_featvec(K1); make_featvec(K2);
        std::vector<MapPoint> mp1(K1.N), mp2(K2.N);
        for (int i = 0; i < K1.N; i++) { mp1[i].bad = (rnd() % 17) == 0; K1.mvpMapPoints[i] = (rnd() % 6) ? &mp1[i] : nullptr; }
        for (int i = 0; i < K2.N; i++) { mp2[i].bad = (rnd() % 19) == 0; K2.mvpMapPoints[i] = (rnd() % 5) ? &mp2[i] : nullptr; }
        coeb_camera c1 = cam_of(K1), c2 = cam_of(K2);
        orc_frame* o1 = orc_frame_create(reinterpret_cast<coeb_keypoint*>(K1.mvKeysUn.data()), K1.mDescriptors.ptr(0), K1.N, K1.mvuRight.data(), &c1, K1.mvScaleFactors.data(), 8);
        orc_frame* o2 = orc_frame_create(reinterpret_cast<coeb_keypoint*>(K2.mvKeysUn.data()), K2.mDescriptors.ptr(0), K2.N, K2.mvuRight.data(), &c2, K2.mvScaleFactors.data(), 8);
        coeb_adapt::FeatVecCSR v1(K1.mFeatVec), v2(K2.mFeatVec);
        std::vector<uint8_t> valid1(K1.N), valid2(K2.N), free1(K1.N), free2(K2.N);
        for (int i = 0; i < K1.N; i++) { valid1[i] = K1.mvpMapPoints[i] && !K1.mvpMapPoints[i]->bad; free1[i] = !K1.mvpMapPoints[i]; }
        for (int i = 0; i < K2.N; i++) { valid2[i] = K2.mvpMapPoints[i] && !K2.mvpMapPoints[i]->bad; free2[i] = !K2.mvpMapPoints[i]; }
        std::vector<int> m12(K1.N);
        ORB_SLAM2::ORBmatcher bow(0.7f, true);
        // SearchByBoW(KeyFrame*, Frame&) (src/Tracking.cc:830-833): K2 plays the current frame
        {
            const int nref = orc_match_bow(o1, o2, valid1.data(), nullptr, v1.nn(), v1.node.data(), v1.start.data(), v1.items.data(), v2.nn(), v2.node.data(),
                                           v2.start.data(), v2.items.data(), 0.7f, 1, 0, m12.data());
            std::vector<MapPoint*> got;
            const int ngot = bow.SearchByBoW(&K1, K2, got);
            EXPECT(ngot == nref && nref > 20, "SearchByBoW(KF, F): %d vs oracle %d", ngot, nref);
            std::vector<MapPoint*> want(K2.N, nullptr);
            for (int i = 0; i < K1.N; i++) if (m12[i] >= 0) want[m12[i]] = K1.mvpMapPoints[i];
            EXPECT(got == want, "SearchByBoW(KF, F): vpMapPointMatches differs");
        }
        // SearchByBoW(KeyFrame*, KeyFrame*) (src/LoopClosing.cc:240-260)
        {
            const int nref = orc_match_bow(o1, o2, valid1.data(), valid2.data(), v1.nn(), v1.node.data(), v1.start.data(), v1.items.data(), v2.nn(),
                                           v2.node.data(), v2.start.data(), v2.items.data(), 0.7f, 1, 1, m12.data());
            std::vector<MapPoint*> got;
            const int ngot = bow.SearchByBoW(&K1, &K2, got);
            EXPECT(ngot == nref && nref > 10, "SearchByBoW(KF, KF): %d vs oracle %d", ngot, nref);
            int bad_ptr = 0;
            for (int i = 0; i < K1.N; i++) bad_ptr += got[i] != (m12[i] >= 0 ? K2.mvpMapPoints[m12[i]] : nullptr);
            EXPECT(bad_ptr == 0, "SearchByBoW(KF, KF): %d vpMatches12 entries differ", bad_ptr);
        }
        // SearchForTriangulation (src/LocalMapping.cc:252-260); K2 is K1 shifted by (7, -3) px: F12 of a pure image shift
        {
            const float F12[9] = {0, 0, -3.f, 0, 0, -7.f, 3.f, 7.f, 0};
            K1.mOw[0] = 0.3f; K1.mOw[1] = -0.1f; K1.mOw[2] = 0.05f;
            K2.tcw[0] = 0.1f; K2.tcw[1] = 0.02f; K2.tcw[2] = 0.5f;
            float C2[3];
            for (int r = 0; r < 3; r++) C2[r] = K2.Rcw[3 * r] * K1.mOw[0] + K2.Rcw[3 * r + 1] * K1.mOw[1] + K2.Rcw[3 * r + 2] * K1.mOw[2] + K2.tcw[r];
            const float invz = 1.0f / C2[2], exx = K2.fx * C2[0] * invz + K2.cx, eyy = K2.fy * C2[1] * invz + K2.cy;
            const int nref = orc_match_triangulation(o1, o2, free1.data(), free2.data(), v1.nn(), v1.node.data(), v1.start.data(), v1.items.data(), v2.nn(),
                                                     v2.node.data(), v2.start.data(), v2.items.data(), F12, exx, eyy, 0, 1, m12.data());
            std::vector<std::pair<size_t, size_t> > pairs;
            const int ngot = bow.SearchForTriangulation(&K1, &K2, F12, pairs, false);
            EXPECT(ngot == nref && nref > 3, "SearchForTriangulation: %d vs oracle %d", ngot, nref);
            size_t q = 0; int bad_pair = 0;
            for (int i = 0; i < K1.N; i++) if (m12[i] >= 0) { bad_pair += q >= pairs.size() || pairs[q].first != (size_t)i || pairs[q].second != (size_t)m12[i]; q++; }
            EXPECT(bad_pair == 0 && q == pairs.size(), "SearchForTriangulation: vMatchedPairs differs");
        }
        orc_frame_destroy(o1);
        orc_frame_destroy(o2);

        // Frame constructor tail on the device (src/Frame.cc:213-240) with the TUM1 distortion and a raw 16-bit depth map
        Frame T;
        T.fx = 517.306408f; T.fy = 516.469215f; T.cx = 318.643040f; T.cy = 255.313989f; T.mb = T.mbf / T.fx;
        T.mpORBextractorLeft = &ex;
        fill_frame(T, ex, img, w, h, nobox, notm, noblur);
        const float dist5[5] = {0.262383f, -0.953104f, -0.005358f, 0.002628f, 1.163314f};
        std::vector<uint16_t> depth((size_t)w * h);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) depth[(size_t)y * w + x] = ((x / 16 + y / 16) % 7 == 0) ? 0 : (uint16_t)(4000 + 37 * ((x / 8) % 50) + 91 * ((y / 8) % 40));
        coeb_depth_image di;
        di.data = depth.data(); di.kind = 2; di.stride_bytes = w * 2; di.factor = 1.0f / 5000.0f; di.on_device = 0; di.width = w; di.height = h;
        coeb_frame* dT = ORB_SLAM2::FrameTailFromExtractor(T, dist5, &di);
        T.mpDeviceFrame = dT;   // later matcher calls on T reuse it instead of uploading the frame again
        coeb_camera ct = cam_of(T);
        std::vector<coeb_keypoint> un(T.N);
        std::vector<float> ur(T.N), dp(T.N);
        orc_undistort_keypoints(reinterpret_cast<coeb_keypoint*>(T.mvKeys.data()), T.N, &ct, dist5, un.data());
        orc_stereo_from_rgbd(reinterpret_cast<coeb_keypoint*>(T.mvKeys.data()), un.data(), T.N, depth.data(), 2, w * 2, di.factor, T.mbf, ur.data(), dp.data());
        EXPECT(std::memcmp(un.data(), T.mvKeysUn.data(), sizeof(coeb_keypoint) * T.N) == 0, "mvKeysUn differs");
        EXPECT(std::memcmp(ur.data(), T.mvuRight.data(), 4 * T.N) == 0 && std::memcmp(dp.data(), T.mvDepth.data(), 4 * T.N) == 0, "mvuRight / mvDepth (RGB-D) differ");
        int withd = 0;
        for (int i = 0; i < T.N; i++) withd += T.mvDepth[i] > 0;
        EXPECT(withd > T.N / 2 && withd < T.N, "depth coverage %d of %d", withd, T.N);

        // Tracking::SearchLocalPoints against a resident local map (src/Tracking.cc:1222-1272)
        const int NL = 2500;
        std::vector<MapPoint> lmp(NL);
        std::vector<MapPoint*> vl(NL);
        const float a = 0.02f;
        const float R[9] = {std::cos(a), 0, std::sin(a), 0, 1, 0, -std::sin(a), 0, std::cos(a)}, t[3] = {0.05f, -0.02f, 0.04f};
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T.mTcw[4 * r + c] = R[3 * r + c]; T.mTcw[4 * r + 3] = t[r]; }
        for (int c = 0; c < 3; c++) T.mOw[c] = -(R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2]);
        for (int i = 0; i < NL; i++) {
            MapPoint& m = lmp[i];
            const int k = rnd() % T.N;
            const bool truth = (rnd() % 3) != 0;
            const float u = truth ? T.mvKeysUn[k].pt.x + (rndf() - 0.5f) * 6.f : rndf() * 800.f - 80.f;
            const float v = truth ? T.mvKeysUn[k].pt.y + (rndf() - 0.5f) * 6.f : rndf() * 600.f - 60.f;
            const float z = (0.7f + 5.f * rndf()) * ((!truth && (rnd() % 9) == 0) ? -1.f : 1.f);
            const float pc[3] = {(u - T.cx) * z / T.fx - t[0], (v - T.cy) * z / T.fy - t[1], z - t[2]};
            for (int c = 0; c < 3; c++) m.pos[c] = R[c] * pc[0] + R[3 + c] * pc[1] + R[6 + c] * pc[2];
            float po[3], dist = 0;
            for (int c = 0; c < 3; c++) { po[c] = m.pos[c] - T.mOw[c]; dist += po[c] * po[c]; }
            dist = std::sqrt(dist);
            float nn = 0;
            for (int c = 0; c < 3; c++) { m.normal[c] = po[c] / dist + (rndf() - 0.5f) * (truth ? 0.4f : 2.5f); nn += m.normal[c] * m.normal[c]; }
            for (int c = 0; c < 3; c++) m.normal[c] /= std::sqrt(nn);
            const int lvl = std::min(T.mvKeysUn[k].octave + (int)(rnd() & 1), 7);
            m.mfMaxDistance = dist * std::pow(1.2f, lvl - 0.5f) * ((!truth && (rnd() % 11) == 0) ? 30.f : 1.f);
            m.mfMinDistance = m.mfMaxDistance / T.mvScaleFactors[7];
            coeb_adapt::create_u8(m.desc, 1, 32);
            std::memcpy(m.desc.ptr(0), T.mDescriptors.ptr(k), 32);
            for (int b = 0, nb = truth ? rnd() % 30 : 120; b < nb; b++) { int bit = rnd() % 256; m.desc.ptr(0)[bit >> 3] ^= (uint8_t)(1 << (bit & 7)); }
            m.bad = (rnd() % 40) == 0;
            m.nObs = (rnd() % 30) == 0 ? 0 : 2;
            m.mnLastFrameSeen = (rnd() % 25) == 0 ? T.mnId : -1;
            vl[i] = &m;
        }
        std::vector<float> xyz(3 * NL), nrm(3 * NL), dmin(NL), dmax(NL);
        std::vector<uint8_t> ldesc(32 * NL), skip(NL), obs(NL), inview(NL);
        for (int i = 0; i < NL; i++) {
            std::memcpy(&xyz[3 * i], lmp[i].pos, 12); std::memcpy(&nrm[3 * i], lmp[i].normal, 12);
            dmin[i] = lmp[i].mfMinDistance; dmax[i] = lmp[i].mfMaxDistance;
            std::memcpy(&ldesc[32 * i], lmp[i].desc.ptr(0), 32);
            skip[i] = lmp[i].mnLastFrameSeen == T.mnId || lmp[i].bad; obs[i] = lmp[i].nObs > 0;
        }
        orc_frame* ot = orc_frame_create(un.data(), T.mDescriptors.ptr(0), T.N, ur.data(), &ct, T.mvScaleFactors.data(), 8);
        std::vector<int> state(T.N, -1);
        const int nref = orc_search_local_points(ot, NL, xyz.data(), nrm.data(), dmin.data(), dmax.data(), ldesc.data(), skip.data(), obs.data(), T.mTcw, T.mOw,
                                                 0.5f, 3.f, 0.8f, state.data(), inview.data(), nullptr);
        coeb_local_map* lm = ORB_SLAM2::MakeLocalMap(vl);
        const int ngot = ORB_SLAM2::SearchLocalPoints(T, dT, lm, vl, 3.f);
        EXPECT(ngot == nref && nref > 100, "SearchLocalPoints: %d vs oracle %d", ngot, nref);
        int bad_ptr = 0, bad_view = 0, nview = 0;
        for (int k = 0; k < T.N; k++) bad_ptr += T.mvpMapPoints[k] != (state[k] >= 0 ? vl[state[k]] : nullptr);
        for (int i = 0; i < NL; i++) { if (!skip[i]) { bad_view += (lmp[i].mbTrackInView != (inview[i] != 0)) || (lmp[i].nVisible != (int)inview[i]); nview += inview[i]; } }
        EXPECT(bad_ptr == 0, "SearchLocalPoints: %d mvpMapPoints entries differ", bad_ptr);
        EXPECT(bad_view == 0 && nview > 300 && nview < NL, "SearchLocalPoints: visibility bookkeeping (%d wrong, %d in view)", bad_view, nview);
        // ---- relocalisation SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (src/Tracking.cc:1436-1545) ----
        {
            Frame KFr;   // plays the candidate keyframe: its map points are the local-map points above, one per keypoint slot
            KFr.N = 1200;
            KFr.mvKeysUn.resize(KFr.N);
            KFr.mvpMapPoints.assign(KFr.N, nullptr);
            std::set<MapPoint*> already;
            for (int i = 0; i < KFr.N; i++) {
                KFr.mvKeysUn[i].angle = 360.f * rndf();
                if (rnd() % 9) KFr.mvpMapPoints[i] = &lmp[i];
                if ((rnd() % 13) == 0) already.insert(&lmp[i]);
            }
            for (int k = 0; k < T.N; k++) T.mvpMapPoints[k] = (k % 6 == 0) ? &lmp[NL - 1 - (k % 50)] : nullptr;   // occupied entries block
            std::vector<uint8_t> rvalid(KFr.N);
            std::vector<float> rang(KFr.N);
            for (int i = 0; i < KFr.N; i++) {
                MapPoint* p = KFr.mvpMapPoints[i];
                rvalid[i] = p && !p->bad && !already.count(p);
                rang[i] = KFr.mvKeysUn[i].angle;
            }
            std::vector<int> rstate(T.N);
            std::vector<MapPoint*> before = T.mvpMapPoints;
            for (int k = 0; k < T.N; k++) rstate[k] = T.mvpMapPoints[k] ? -2 : -1;
            const int rref = orc_match_reloc(ot, KFr.N, rvalid.data(), xyz.data(), dmin.data(), dmax.data(), rang.data(), ldesc.data(), T.mTcw, T.mOw, 10.f,
                                             100, 1, rstate.data());
            ORB_SLAM2::ORBmatcher rm(0.75f, true);
            const int rgot = rm.SearchByProjection(T, &KFr, already, 10.f, 100);
            EXPECT(rgot == rref && rref > 3, "SearchByProjection(reloc): %d vs oracle %d", rgot, rref);
            int bad_ptr = 0;
            for (int k = 0; k < T.N; k++) bad_ptr += T.mvpMapPoints[k] != (rstate[k] >= 0 ? KFr.mvpMapPoints[rstate[k]] : before[k]);
            EXPECT(bad_ptr == 0, "SearchByProjection(reloc): %d mvpMapPoints entries differ", bad_ptr);
        }
        // ---- Fuse(KeyFrame*, vpMapPoints, th) (src/LocalMapping.cc:484-515): search on the device, side effects on the host ----
        {
            for (int k = 0; k < T.N; k++) T.mvpMapPoints[k] = (k % 4 == 0) ? &lmp[NL - 1 - (k % 60)] : nullptr;
            for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T.Rcw[3 * r + c] = T.mTcw[4 * r + c]; T.tcw[r] = T.mTcw[4 * r + 3]; }
            std::vector<uint8_t> fvalid(NL);
            for (int i = 0; i < NL; i++) { lmp[i].inKF = (i % 17 == 0) ? (const void*)&T : nullptr; lmp[i].replacedBy = nullptr; fvalid[i] = !lmp[i].bad && lmp[i].inKF != (const void*)&T; }
            std::vector<int> fbest(NL);
            const int fref = orc_fuse_search(ot, NL, xyz.data(), nrm.data(), dmin.data(), dmax.data(), ldesc.data(), fvalid.data(), T.mTcw, T.mOw, 3.f, 1, fbest.data());
            std::vector<MapPoint*> beforeKF = T.mvpMapPoints;
            std::vector<MapPoint> snapshot = lmp;
            ORB_SLAM2::ORBmatcher fm;
            const int fgot = fm.Fuse(&T, vl, 3.f);
            EXPECT(fgot == fref && fref > 100, "Fuse: %d vs oracle %d", fgot, fref);
            std::vector<MapPoint> after = lmp;
            std::vector<MapPoint*> afterKF = T.mvpMapPoints;
            // the reference's loop (:938-957) replayed from the oracle's indices on the restored state
            lmp = snapshot;
            T.mvpMapPoints = beforeKF;
            for (int i = 0; i < NL; i++) {
                if (fbest[i] < 0) continue;
                MapPoint* pMP = vl[i];
                MapPoint* in = T.GetMapPoint(fbest[i]);
                if (in) {
                    if (!in->isBad()) { if (in->Observations() > pMP->Observations()) pMP->Replace(in); else in->Replace(pMP); }
                } else { pMP->AddObservation(&T, fbest[i]); T.AddMapPoint(pMP, fbest[i]); }
            }
            int wrong = 0;
            for (int i = 0; i < NL; i++)
                wrong += after[i].bad != lmp[i].bad || after[i].replacedBy != lmp[i].replacedBy || after[i].obsIdx != lmp[i].obsIdx || after[i].nObs != lmp[i].nObs;
            for (int k = 0; k < T.N; k++) wrong += afterKF[k] != T.mvpMapPoints[k];
            EXPECT(wrong == 0, "Fuse: %d map points with unexpected side effects", wrong);
        }
        coeb_local_map_destroy(lm);
        T.mpDeviceFrame = nullptr;
        coeb_frame_destroy(dT);
        orc_frame_destroy(ot);
    }

    // ---- DescriptorDistance ----
    EXPECT(ORB_SLAM2::ORBmatcher::DescriptorDistance(F.mDescriptors.row(0), F.mDescriptors.row(0)) == 0, "distance to self");
    EXPECT(ORB_SLAM2::ORBmatcher::TH_HIGH == 100 && ORB_SLAM2::ORBmatcher::TH_LOW == 50 && ORB_SLAM2::ORBmatcher::HISTO_LENGTH == 30, "constants");

    // ---- ComputeStereoMatches with two extractors (src/Frame.cc:644-818) ----
    {
        const int sw = 1241, sh = 376;
        std::vector<uint8_t> L = make_image(sw, sh, 21), Rr((size_t)sw * sh);
        for (int y = 0; y < sh; y++) {
            const int d = 4 + ((y / 24) * 13) % 60;
            for (int x = 0; x < sw; x++) Rr[(size_t)y * sw + x] = L[(size_t)y * sw + std::min(x + d, sw - 1)];
        }
        coeb_orb_params sp = {2000, 1.2f, 8, 20, 7};
        ORB_SLAM2::ORBextractor exL(2000, 1.2f, 8, 20, 7), exR(2000, 1.2f, 8, 20, 7);
        orc_extractor *oL = orc_extractor_create(&sp), *oR = orc_extractor_create(&sp);
        Frame S;
        S.mpORBextractorLeft = &exL; S.mpORBextractorRight = &exR;
        S.mbf = 386.1448f; S.mb = 386.1448f / 718.856f;
        Mat gl = coeb_adapt::wrap_u8(sh, sw, L.data(), (size_t)sw), gr = coeb_adapt::wrap_u8(sh, sw, Rr.data(), (size_t)sw), none;
        exL(gl, none, S.mvKeys, S.mDescriptors);
        exR(gr, none, S.mvKeysRight, S.mDescriptorsRight);
        std::vector<coeb_keypoint> kl(8192), kr(8192);
        std::vector<uint8_t> dl(8192 * 32), dr(8192 * 32);
        int nl = 0, nr = 0;
        orc_extract(oL, L.data(), sw, sh, sw, nullptr, 0, nullptr, 0, nullptr, 0, kl.data(), dl.data(), 8192, &nl);
        orc_extract(oR, Rr.data(), sw, sh, sw, nullptr, 0, nullptr, 0, nullptr, 0, kr.data(), dr.data(), 8192, &nr);
        EXPECT(nl == (int)S.mvKeys.size() && nr == (int)S.mvKeysRight.size(), "stereo extraction counts");
        std::vector<float> ur(nl), dp(nl);
        const int sref = orc_stereo_match(oL, oR, nl, kl.data(), dl.data(), nr, kr.data(), dr.data(), S.mbf, S.mb, ur.data(), dp.data());
        ORB_SLAM2::ComputeStereoMatches(S);
        int withd = 0;
        for (int i = 0; i < nl; i++) withd += S.mvDepth[i] > 0;
        EXPECT(withd == sref && sref > 30, "stereo: %d points with depth vs oracle %d", withd, sref);
        EXPECT(std::memcmp(S.mvuRight.data(), ur.data(), sizeof(float) * nl) == 0 && std::memcmp(S.mvDepth.data(), dp.data(), sizeof(float) * nl) == 0,
               "mvuRight / mvDepth differ");
        orc_extractor_destroy(oL);
        orc_extractor_destroy(oR);
    }
    orc_extractor_destroy(oex);
    std::printf(g_fail ? "dropin_test: %d FAILURES\n" : "dropin_test: PASS\n", g_fail);
    return g_fail ? 1 : 0;
}
