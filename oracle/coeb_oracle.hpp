// coeb_oracle.hpp -- CPU restatement ("oracle B") of the COEB-SLAM front-end hot path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked into, imported by, or called from the
// product library (coeb-slam_b200/). Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs may use it, and there only as the checker / the timed CPU arm.
//
// PARITY STATUS: pinned to the reference itself. The reference ships no tests, golden vectors or
// known-answer fixtures, but its hot-path sources (src/ORBextractor.cc,
// ORBmatcher.cc, Frame.cc, MapPoint.cc, KeyFrame.cc, Map.cc) compile unchanged against the test-only
// OpenCV stand-in of oracle/ref_shim/ into oracle/_ref (recipe: `make -C oracle _ref`), and
// tests/test_ref_parity_cpu.py requires this restatement to equal them bit for bit (extraction
// against the monotonic-heap build, because the reference's octree tie order depends on heap
// addresses). The OpenCV primitives the reference calls are restated here as
// integer models and pinned against the installed OpenCV 4.13.0 (python cv2) in
// tests/test_oracle_vs_cv2.py; the shim uses these same models.
//
// Every function cites the reference file:line it restates (paths relative to /root/reference).
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "../include/coeb_types.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// OpenCV scalar helpers (third-party, not under /root/reference; OpenCV 4.13.0 semantics)
// ---------------------------------------------------------------------------------------------
// cvRound: round-half-to-even (SSE cvtss2si / lrint under the default rounding mode).
static inline int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(float v) { int i = (int)v; return i - (i > v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(float v) { int i = (int)v; return i + (i < v); }

// cv::fastAtan2(y, x), degrees. Scalar fp32 path of OpenCV core/mathfuncs_core (atan_f32).
// The four polynomial constants are float*float products; no FMA (build with -ffp-contract=off).
static inline float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;  // tight rows, stride == w
    void alloc(int w_, int h_) { w = w_; h = h_; px.assign((size_t)w * h, 0); }
    const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
    uint8_t* row(int y) { return px.data() + (size_t)y * w; }
};

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 -- the legacy 11-bit fixed-point
// path (imgproc/resize.cpp: HResizeLinear<uchar,int,short,2048> + VResizeLinear + FixedPtCast<22>).
// Called by the reference at src/ORBextractor.cc:1356.
static inline void resize_linear_8u(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst,
                                    int dw, int dh, int dstride) {
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * (size_t)dw), beta(2 * (size_t)dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cv_round((1.f - fx) * 2048.f);
        alpha[2 * dx + 1] = (short)cv_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = (short)cv_round((1.f - fy) * 2048.f);
        beta[2 * dy + 1] = (short)cv_round(fy * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    auto hpass = [&](int sy, std::vector<int>& out) {
        sy = std::min(std::max(sy, 0), sh - 1);
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int s1 = sx + 1 < sw ? S[sx + 1] : S[sx];  // alpha[1]==0 whenever sx was clamped
            out[dx] = S[sx] * alpha[2 * dx] + s1 * alpha[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; dy++) {
        hpass(yofs[dy], r0);
        hpass(yofs[dy] + 1, r1);
        const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 -- OpenCV's
// bit-exact fixed-point smoothing: Q8 kernel {18,34,48,56,48,34,18}, horizontal pass to Q8.8,
// vertical pass to Q16.16, round-to-nearest. Called by the reference at src/ORBextractor.cc:1318.
static inline void gaussian7x7_8u(const uint8_t* src, int w, int h, int sstride, uint8_t* dst,
                                  int dstride) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    // The CPU baseline times this function, so it is written the way a tuned CPU filter is: each row is first copied
    // into a padded buffer with its 3 reflected pixels per side (no index arithmetic in the inner loops, which the
    // compiler then vectorises); the arithmetic is unchanged.
    std::vector<uint16_t> hbuf((size_t)w * h);
    std::vector<uint8_t> prow((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int k = 0; k < 3; k++) { prow[k] = S[reflect101(k - 3, w)]; prow[w + 3 + k] = S[reflect101(w + k, w)]; }
        std::memcpy(prow.data() + 3, S, w);
        const uint8_t* p = prow.data();
        uint16_t* H = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; x++)
            H[x] = (uint16_t)(18 * (p[x] + p[x + 6]) + 34 * (p[x + 1] + p[x + 5]) + 48 * (p[x + 2] + p[x + 4]) + 56 * p[x + 3]);
    }
    for (int y = 0; y < h; y++) {
        uint8_t* D = dst + (size_t)y * dstride;
        const uint16_t* r[7];
        for (int k = 0; k < 7; k++) r[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        for (int x = 0; x < w; x++) {
            const uint32_t acc = (uint32_t)K[0] * ((uint32_t)r[0][x] + r[6][x]) + (uint32_t)K[1] * ((uint32_t)r[1][x] + r[5][x]) +
                                 (uint32_t)K[2] * ((uint32_t)r[2][x] + r[4][x]) + (uint32_t)K[3] * r[3][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// cv::cvtColor(src, dst, CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) for 8-bit images (imgproc
// color_rgb: RGB2Gray<uchar>, 15-bit BT.601 weights 9798/19235/3735, round to nearest). Called by the reference at
// src/Tracking.cc:212-224 ("next" row 3 of SURVEY.md section 8f).
static inline void rgb_to_gray_8u(const uint8_t* src, int w, int h, int sstride, int channels, bool bgr, uint8_t* dst, int dstride) {
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            const int c0 = S[x * channels], c1 = S[x * channels + 1], c2 = S[x * channels + 2];
            const int r = bgr ? c2 : c0, b = bgr ? c0 : c2;
            D[x] = (uint8_t)((r * 9798 + c1 * 19235 + b * 3735 + (1 << 14)) >> 15);
        }
    }
}

// Frame::detect_laplacian on a box of the gray image (src/Frame.cc:173-202, 905-913): cv::Laplacian(roi.clone(), CV_16U)
// (3x3 cross, BORDER_REFLECT_101 on the cloned ROI, negative responses saturate to 0), cv::abs, cv::mean. The box is
// blurred (blur_flag = 1) if the mean is < 4.2. "next" row 2 of SURVEY.md section 8f. Returns the mean; -1 if the box is
// empty or leaves the image (the reference's cv::Mat ROI would throw).
static inline double laplacian_box_mean(const uint8_t* gray, int w, int h, int stride, const float* box) {
    const int x0 = (int)box[0], y0 = (int)box[1], bw = (int)(box[2] - box[0]), bh = (int)(box[3] - box[1]);
    if (x0 < 0 || y0 < 0 || bw <= 0 || bh <= 0 || x0 + bw > w || y0 + bh > h) return -1.0;
    unsigned long long sum = 0;
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < bw; x++) {
            auto px = [&](int xx, int yy) { return (int)gray[(size_t)(y0 + reflect101(yy, bh)) * stride + x0 + reflect101(xx, bw)]; };
            const int lap = px(x, y - 1) + px(x, y + 1) + px(x - 1, y) + px(x + 1, y) - 4 * px(x, y);
            sum += (unsigned)std::max(lap, 0);
        }
    return (double)sum / ((double)bw * (double)bh);
}

// ---------------------------------------------------------------------------------------------
// cv::FAST(roi, keypoints, threshold, nonmaxSuppression=true) == FAST_t<16> (features2d/fast.cpp,
// fast_score.cpp). Called per cell by the reference at src/ORBextractor.cc:831,836.
// ---------------------------------------------------------------------------------------------
static const int kFastDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kFastDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// cornerScore<16>: largest threshold for which the pixel is still a FAST-9 corner, seeded with
// the detection threshold.
static inline int fast_corner_score(const int d[25], int threshold) {
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], d[k + 2]);
        a = std::min(a, d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, d[k + 4]);
        a = std::min(a, d[k + 5]);
        a = std::min(a, d[k + 6]);
        a = std::min(a, d[k + 7]);
        a = std::min(a, d[k + 8]);
        a0 = std::max(a0, std::min(a, d[k]));
        a0 = std::max(a0, std::min(a, d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], d[k + 2]);
        b = std::max(b, d[k + 3]);
        b = std::max(b, d[k + 4]);
        b = std::max(b, d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, d[k + 6]);
        b = std::max(b, d[k + 7]);
        b = std::max(b, d[k + 8]);
        b0 = std::min(b0, std::max(b, d[k]));
        b0 = std::min(b0, std::max(b, d[k + 9]));
    }
    return -b0 - 1;
}

struct FastPt { int x, y, score; };

// Literal FAST-9/16 with 3x3 non-max suppression on a ROI view (stride may exceed w).
static inline void fast9_nms(const uint8_t* img, int w, int h, int stride, int threshold,
                             std::vector<FastPt>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<uint8_t> score((size_t)w * h, 0);
    std::vector<uint8_t> corner((size_t)w * h, 0);
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = kFastDy[k] * stride + kFastDx[k];
    uint8_t thr_tab[511];
    for (int i = -255; i <= 255; i++) thr_tab[i + 255] = (uint8_t)(i < -threshold ? 1 : (i > threshold ? 2 : 0));
    for (int y = 3; y < h - 3; y++) {
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = img + (size_t)y * stride + x;
            const int v = p[0];
            {   // FAST_t's early rejection: a 9-arc contains one pixel of every opposite pair (k, k+8), all of one sign.
                // Like OpenCV, the darker(1)/brighter(2) class comes from a 511-entry table indexed by q - v.
                const uint8_t* tab = &thr_tab[255 - v];
                auto cls = [&](int k) { return (int)tab[p[off[k]]]; };
                int dd = cls(0) | cls(8);
                if (!dd) continue;
                dd &= cls(2) | cls(10);
                dd &= cls(4) | cls(12);
                dd &= cls(6) | cls(14);
                if (!dd) continue;
                dd &= cls(1) | cls(9);
                dd &= cls(3) | cls(11);
                dd &= cls(5) | cls(13);
                dd &= cls(7) | cls(15);
                if (!dd) continue;
            }
            int d[25];
            uint32_t brighter = 0, darker = 0;  // ring pixel brighter / darker than centre by > th
            for (int k = 0; k < 16; k++) {
                const int q = p[off[k]];
                d[k] = v - q;
                if (q > v + threshold) brighter |= 1u << k;
                if (q < v - threshold) darker |= 1u << k;
            }
            auto has_arc9 = [](uint32_t m) {   // 9 contiguous set bits in a circular 16-bit mask
                m |= m << 16;
                uint32_t a = m & (m >> 1);
                a &= a >> 2;
                a &= a >> 4;
                a &= m >> 8;
                return (a & 0xFFFFu) != 0;
            };
            if (has_arc9(brighter) || has_arc9(darker)) {
                for (int k = 16; k < 25; k++) d[k] = d[k - 16];
                corner[(size_t)y * w + x] = 1;
                score[(size_t)y * w + x] = (uint8_t)fast_corner_score(d, threshold);
            }
        }
    }
    for (int y = 3; y < h - 3; y++) {
        for (int x = 3; x < w - 3; x++) {
            if (!corner[(size_t)y * w + x]) continue;
            const int s = score[(size_t)y * w + x];
            const uint8_t* c = &score[(size_t)y * w + x];
            if (s > c[1] && s > c[-1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                s > c[w - 1] && s > c[w] && s > c[w + 1])
                out.push_back({x, y, s});
        }
    }
}

// ---------------------------------------------------------------------------------------------
// ORB pattern (data) -- the reference stores it as 512 cv::Point (src/ORBextractor.cc:455-457).
// ---------------------------------------------------------------------------------------------
static const int kOrbPattern[1024] = {
#include "../include/coeb_orb_pattern.inc"
};

static const int HALF_PATCH_SIZE = 15;  // src/ORBextractor.cc:75
static const int PATCH_SIZE = 31;       // src/ORBextractor.cc:74
static const int EDGE_THRESHOLD = 19;   // src/ORBextractor.cc:76

// IC_Angle (src/ORBextractor.cc:80-107): intensity-centroid orientation on the unblurred level.
static inline float ic_angle(const Image& im, float ptx, float pty, const int* umax) {
    int m_01 = 0, m_10 = 0;
    const int cx = cv_round(ptx), cy = cv_round(pty);
    const uint8_t* center = im.row(cy) + cx;
    const int step = im.w;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
}

// computeOrbDescriptor (src/ORBextractor.cc:109-156): steered rBRIEF on the blurred level.
static inline void orb_descriptor(const Image& blurred, float ptx, float pty, float angle_deg,
                                  uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const int cx = cv_round(ptx), cy = cv_round(pty);
    const uint8_t* center = blurred.row(cy) + cx;
    const int step = blurred.w;
    const int* pat = kOrbPattern;
    auto sample = [&](int idx) -> int {
        const float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
        const int yy = cv_round(px * b + py * a);
        const int xx = cv_round(px * a - py * b);
        return center[yy * step + xx];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            int t0 = sample(2 * k), t1 = sample(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree (src/ORBextractor.cc:489-769)
// ---------------------------------------------------------------------------------------------
struct Cand {  // a FAST candidate in minBorder-relative level coordinates
    float x, y, response;
};

struct QuadNode {
    int x0, x1, y0, y1;  // UL.x, UR.x, UL.y, BR.y
    std::vector<int> keys;  // indices into the candidate array, parent order preserved
    bool no_more = false;
    long seq = 0;  // creation sequence number: replaces the reference's heap address in the
                   // (size, pointer) sort of :691 -- see octree_distribute() below
    std::list<QuadNode>::iterator self;
};

// ExtractorNode::DivideNode (src/ORBextractor.cc:489-544)
static inline void quad_divide(const QuadNode& p, const std::vector<Cand>& c, QuadNode ch[4]) {
    const int halfX = (int)std::ceil((float)(p.x1 - p.x0) / 2);
    const int halfY = (int)std::ceil((float)(p.y1 - p.y0) / 2);
    const int xm = p.x0 + halfX, ym = p.y0 + halfY;
    ch[0].x0 = p.x0; ch[0].x1 = xm; ch[0].y0 = p.y0; ch[0].y1 = ym;    // n1 top-left
    ch[1].x0 = xm; ch[1].x1 = p.x1; ch[1].y0 = p.y0; ch[1].y1 = ym;    // n2 top-right
    ch[2].x0 = p.x0; ch[2].x1 = xm; ch[2].y0 = ym; ch[2].y1 = p.y1;    // n3 bottom-left
    ch[3].x0 = xm; ch[3].x1 = p.x1; ch[3].y0 = ym; ch[3].y1 = p.y1;    // n4 bottom-right
    for (int k : p.keys) {
        const Cand& kp = c[k];
        int q;
        if (kp.x < (float)xm) q = (kp.y < (float)ym) ? 0 : 2;
        else q = (kp.y < (float)ym) ? 1 : 3;
        ch[q].keys.push_back(k);
    }
    for (int q = 0; q < 4; q++) ch[q].no_more = ch[q].keys.size() == 1;
}

// How often the heap-address tie-break of :691 can matter (read by tests/test_ref_parity_cpu.py through
// orc_octree_tie_stats): a careful-phase round whose sorted list holds equal sizes among the entries it expands
// ("order tie": the children are pushed in another order, so the output ORDER may differ) or whose >= N cut-off falls
// inside a run of equal sizes ("cut tie": another node is expanded, so the output SET may differ).
struct OctreeTieStats { long calls = 0, careful_rounds = 0, rounds_with_order_tie = 0, rounds_with_cut_tie = 0; };
inline thread_local OctreeTieStats g_tie_stats;   // one instance across the translation units that include this header

// Returns indices (into cands) of the retained key of each final node, in final list order.
//
// Tie-break rule (documented deviation from a non-deterministic reference): the reference sorts
// pair<int, ExtractorNode*> (:691), so equal-size nodes are ordered by heap address. Here the
// second key is the node's creation sequence number, i.e. "stable by size, later-created first"
// when walked from the back -- which is what ascending addresses from a fresh heap would give.
static inline std::vector<int> octree_distribute(const std::vector<Cand>& cands, int minX, int maxX,
                                                 int minY, int maxY, int N) {
    std::vector<int> result;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1 || cands.empty()) return result;
    const float hX = (float)(maxX - minX) / nIni;

    std::list<QuadNode> nodes;
    std::vector<QuadNode*> roots(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        QuadNode n;
        n.x0 = (int)(hX * (float)i);
        n.x1 = (int)(hX * (float)(i + 1));
        n.y0 = 0;
        n.y1 = maxY - minY;
        n.seq = seq++;
        nodes.push_back(n);
        roots[i] = &nodes.back();
    }
    for (size_t i = 0; i < cands.size(); i++) {
        int r = (int)(cands[i].x / hX);
        r = std::min(std::max(r, 0), nIni - 1);  // the reference indexes unchecked (:576)
        roots[r]->keys.push_back((int)i);
    }
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->no_more = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }

    typedef std::pair<int, long> SizeSeq;                      // (vKeys.size(), creation seq)
    std::vector<std::pair<SizeSeq, QuadNode*>> expandable;     // nodes with >1 key from last round
    auto push_children = [&](QuadNode ch[4]) {
        for (int q = 0; q < 4; q++) {
            if (ch[q].keys.empty()) continue;
            ch[q].seq = seq++;
            nodes.push_front(ch[q]);
            nodes.front().self = nodes.begin();
            if (nodes.front().keys.size() > 1)
                expandable.push_back({{(int)nodes.front().keys.size(), nodes.front().seq}, &nodes.front()});
        }
    };

    g_tie_stats.calls++;
    bool finish = false;
    while (!finish) {
        int prevSize = (int)nodes.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {       // full pass (:613-672)
            if (it->no_more) { ++it; continue; }
            QuadNode ch[4];
            quad_divide(*it, cands, ch);
            size_t before = expandable.size();
            push_children(ch);
            nToExpand += (int)(expandable.size() - before);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {       // careful phase (:680-744)
            while (!finish) {
                prevSize = (int)nodes.size();
                auto prev = expandable;
                expandable.clear();
                std::sort(prev.begin(), prev.end(),
                          [](const std::pair<SizeSeq, QuadNode*>& a, const std::pair<SizeSeq, QuadNode*>& b) {
                              return a.first < b.first;
                          });
                int j_stop = -1;
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    QuadNode ch[4];
                    quad_divide(*prev[j].second, cands, ch);
                    push_children(ch);
                    nodes.erase(prev[j].second->self);
                    if ((int)nodes.size() >= N) { j_stop = j; break; }
                }
                {   // tie statistics only; no effect on the result
                    g_tie_stats.careful_rounds++;
                    const int first = j_stop < 0 ? 0 : j_stop;
                    bool order_tie = false;
                    for (int j = (int)prev.size() - 1; j > first; j--)
                        if (prev[j].first.first == prev[j - 1].first.first) { order_tie = true; break; }
                    if (order_tie) g_tie_stats.rounds_with_order_tie++;
                    if (j_stop > 0 && prev[j_stop].first.first == prev[j_stop - 1].first.first) g_tie_stats.rounds_with_cut_tie++;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            }
        }
    }
    result.reserve(nodes.size());
    for (auto& n : nodes) {  // retain the best response per node, first wins ties (:748-766)
        int best = n.keys[0];
        float maxResponse = cands[best].response;
        for (size_t k = 1; k < n.keys.size(); k++)
            if (cands[n.keys[k]].response > maxResponse) {
                best = n.keys[k];
                maxResponse = cands[best].response;
            }
        result.push_back(best);
    }
    return result;
}

// ---------------------------------------------------------------------------------------------
// ORBextractor (src/ORBextractor.cc:418-477, 771-904, 1088-1450)
// ---------------------------------------------------------------------------------------------
struct IRect { int x0, y0, x1, y1; };

struct StageTimes {  // seconds, accumulated over calls (CPU baseline breakdown)
    double pyramid = 0, fast = 0, octree = 0, angle = 0, blur = 0, desc = 0, total = 0;
    long frames = 0;
};

class Extractor {
public:
    int nfeatures;
    double scaleFactor;  // the reference member is a double initialised from a float (ORBextractor.h)
    int nlevels;
    int iniThFAST, minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel;
    int umax[HALF_PATCH_SIZE + 1];

    // per-call state kept for stage-by-stage parity checks
    std::vector<Image> pyramid, blurred;
    std::vector<std::vector<Cand>> candidates;            // after the optional pre-octree cull
    std::vector<std::vector<coeb_keypoint>> level_keys;   // after octree+angle(+post cull), level coords
    coeb_dyn_info dyn;
    StageTimes times;

    // ORBextractor::ORBextractor (src/ORBextractor.cc:418-477)
    Extractor(int nf, float sf, int nl, int iniTh, int minTh)
        : nfeatures(nf), scaleFactor(sf), nlevels(nl), iniThFAST(iniTh), minThFAST(minTh) {
        mvScaleFactor.resize(nl);
        mvLevelSigma2.resize(nl);
        mvScaleFactor[0] = 1.0f;
        mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nl; i++) {
            mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nl);
        mvInvLevelSigma2.resize(nl);
        for (int i = 0; i < nl; i++) {
            mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
            mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
        }
        mnFeaturesPerLevel.resize(nl);
        float factor = (float)(1.0f / scaleFactor);
        float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int level = 0; level < nl - 1; level++) {
            mnFeaturesPerLevel[level] = cv_round(nDesired);
            sum += mnFeaturesPerLevel[level];
            nDesired *= factor;
        }
        mnFeaturesPerLevel[nl - 1] = std::max(nfeatures - sum, 0);

        int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
        int vmin = cv_ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= HALF_PATCH_SIZE; v++) umax[v] = 0;
        for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    void level_size(int w, int h, int level, int& lw, int& lh) const {  // :1348-1349
        float s = mvInvScaleFactor[level];
        lw = cv_round((float)w * s);
        lh = cv_round((float)h * s);
    }

    // box -> mask decision (src/ORBextractor.cc:1101-1195). The mask itself is kept as the list of
    // zero-filled rectangles; mask_at() answers `*mask.ptr<uchar>(y, x) == 0`.
    // Returns false if a box lies outside the image (the reference's cv::Mat ROI would throw).
    bool classify_boxes(int w, int h, const float* boxes, int nbox, const float* tm, int ntm,
                        const int* blur_flag, int nblur) {
        std::memset(&dyn, 0, sizeof(dyn));
        float area = 0;
        for (int b = 0; b < nbox; b++) {
            const float xmin = boxes[4 * b], ymin = boxes[4 * b + 1], xmax = boxes[4 * b + 2],
                        ymax = boxes[4 * b + 3];
            const int rx = (int)xmin, ry = (int)ymin, rw = (int)(xmax - xmin), rh = (int)(ymax - ymin);
            if (rx < 0 || ry < 0 || rw < 0 || rh < 0 || rx + rw > w || ry + rh > h) return false;
            if ((int)xmax > w || (int)ymax > h) return false;
            const float area_box = (xmax - xmin) * (ymax - ymin);
            size_t count = 0;
            bool mark = false;
            for (int t = 0; t < ntm; t++) {
                const int tx = (int)tm[2 * t], ty = (int)tm[2 * t + 1];  // Mat::ptr(int,int) truncation
                if (tx >= rx && tx < rx + rw && ty >= ry && ty < ry + rh) count++;
                if ((float)(count * 10000) > area_box) {                 // "layer 1" (:1145-1161)
                    mark = true;
                    break;
                }
            }
            const int bf = b < nblur ? blur_flag[b] : 0;
            if (!mark && bf == 1 && count > 0) mark = true;               // "layer 2" (:1168-1184)
            if (mark) {
                area = area + area_box;
                if (dyn.n_dynamic < COEB_MAX_BOXES) {
                    int* r = dyn.rect[dyn.n_dynamic];
                    r[0] = (int)xmin; r[1] = (int)ymin; r[2] = (int)xmax; r[3] = (int)ymax;
                }
                dyn.n_dynamic++;
            }
        }
        dyn.area = area;
        dyn.area_flag = area > 200000;                                    // :1192
        return dyn.n_dynamic <= COEB_MAX_BOXES;
    }

    bool mask_is_zero(int x, int y) const {
        for (int i = 0; i < dyn.n_dynamic; i++) {
            const int* r = dyn.rect[i];
            if (x >= r[0] && x < r[2] && y >= r[1] && y < r[3]) return true;
        }
        return false;
    }

    // CheckMovingKeyPoints / CheckMovingKeyPoints_finall predicate (src/ORBextractor.cc:1391-1397,
    // 1426-1440): true if the keypoint must be erased.
    bool is_moving(float ptx, float pty, int level, int w0, int h0) const {
        float scale = level != 0 ? mvScaleFactor[level] : 1.f;
        float sx = ptx * scale, sy = pty * scale;
        if (sx >= (float)(w0 - 1)) sx = (float)(w0 - 1);
        if (sy >= (float)(h0 - 1)) sy = (float)(h0 - 1);
        return mask_is_zero((int)sx, (int)sy);
    }

    // ComputePyramid (src/ORBextractor.cc:1344-1367). The 19-px reflect-101 border the reference
    // adds is never read on this path (FAST reads [16,dim-16), IC_Angle [4,dim-4), the blur clones
    // the ROI, descriptors stay within 19 px), so levels are stored without it.
    void compute_pyramid(const uint8_t* gray, int w, int h, int stride) {
        pyramid.resize(nlevels);
        pyramid[0].alloc(w, h);
        for (int y = 0; y < h; y++) std::memcpy(pyramid[0].row(y), gray + (size_t)y * stride, w);
        for (int l = 1; l < nlevels; l++) {
            int lw, lh;
            level_size(w, h, l, lw, lh);
            pyramid[l].alloc(lw, lh);
            resize_linear_8u(pyramid[l - 1].px.data(), pyramid[l - 1].w, pyramid[l - 1].h,
                             pyramid[l - 1].w, pyramid[l].px.data(), lw, lh, lw);
        }
    }

    // Cell loop of ComputeKeyPointsOctTree (src/ORBextractor.cc:793-850).
    void detect_level(int level, int thIni, int thMin, std::vector<Cand>& out) const {
        out.clear();
        const Image& im = pyramid[level];
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
        const float W = 30;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) return;
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        std::vector<FastPt> cell;
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, chh = (int)maxY - y0;
                const uint8_t* roi = im.row(y0) + x0;
                fast9_nms(roi, cw, chh, im.w, thIni, cell);
                if (cell.empty()) fast9_nms(roi, cw, chh, im.w, thMin, cell);
                for (const FastPt& p : cell)
                    out.push_back({(float)p.x + (float)(j * wCell), (float)p.y + (float)(i * hCell),
                                   (float)p.score});
            }
        }
    }

    // ORBextractor::operator() (src/ORBextractor.cc:1088-1342), minus the debug drawing/imshow.
    // Returns 0, or a negative coeb_status.
    int extract(const uint8_t* gray, int w, int h, int stride, const float* boxes, int nbox,
                const float* tm, int ntm, const int* blur_flag, int nblur,
                std::vector<coeb_keypoint>& kps, std::vector<uint8_t>& desc);
};

}  // namespace orc
