/*
 * ORBextractor.h -- drop-in ORB_SLAM2::ORBextractor for COEB-SLAM, backed by the B200 C ABI
 * (coeb_frontend.h). Header-only adapter: same class name, constructor, call operators and getters as
 * the reference's include/ORBextractor.h:45-128, so src/Frame.cc:413-419 and src/Tracking.cc:120-126,
 * 434-465 compile unchanged against it. All arithmetic happens in libcoeb_frontend.so (sm_100a CUDA);
 * nothing is computed here and there is no CPU fallback (construction throws without a B200).
 *
 * With OpenCV available define COEB_WITH_OPENCV before including: the adapter then speaks cv::Mat,
 * cv::KeyPoint, cv::InputArray. Without it (this repo's CI has no OpenCV C++ headers) the minimal
 * layout-compatible stand-ins of namespace coeb_cv are used; cv::KeyPoint and coeb_keypoint share the
 * 28-byte field order {pt.x, pt.y, size, angle, response, octave, class_id}.
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "coeb_frontend.h"

#ifdef COEB_WITH_OPENCV
#include <opencv2/core/core.hpp>
namespace coeb_cv = cv;
#else
namespace coeb_cv {
struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
/* 8-bit single-channel matrix: just enough of cv::Mat for the extractor boundary. */
class Mat {
public:
    int rows, cols;
    size_t step;
    unsigned char* data;
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, unsigned char* d, size_t s) : rows(r), cols(c), step(s), data(d) {}
    bool empty() const { return !data || rows <= 0 || cols <= 0; }
    void create(int r, int c) { store_.assign((size_t)r * c, 0); rows = r; cols = c; step = (size_t)c; data = store_.data(); }
    void release() { store_.clear(); rows = cols = 0; step = 0; data = nullptr; }
    unsigned char* ptr(int r) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r) const { return data + (size_t)r * step; }
    Mat row(int r) const { return Mat(1, cols, data + (size_t)r * step, step); }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), step(o.step), data(o.data), store_(o.store_) { if (!store_.empty()) data = store_.data(); }
    Mat& operator=(const Mat& o) {
        rows = o.rows; cols = o.cols; step = o.step; store_ = o.store_; data = store_.empty() ? o.data : store_.data();
        return *this;
    }
private:
    std::vector<unsigned char> store_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace coeb_cv
#endif

/* The two places where cv::Mat and the stand-in differ in spelling (an 8-bit single-channel matrix), so that every other line of
 * the adapters is the same text in both modes. */
namespace coeb_adapt {
#ifdef COEB_WITH_OPENCV
#ifndef CV_Assert
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("assertion failed: " #expr); } while (0)
#endif
inline void create_u8(cv::Mat& m, int rows, int cols) { m.create(rows, cols, CV_8UC1); }
inline cv::Mat wrap_u8(int rows, int cols, unsigned char* data, size_t step) { return cv::Mat(rows, cols, CV_8UC1, data, step); }
#else
inline void create_u8(coeb_cv::Mat& m, int rows, int cols) { m.create(rows, cols); }
inline coeb_cv::Mat wrap_u8(int rows, int cols, unsigned char* data, size_t step) { return coeb_cv::Mat(rows, cols, data, step); }
#endif
}  // namespace coeb_adapt

namespace ORB_SLAM2 {

static_assert(sizeof(coeb_cv::KeyPoint) == sizeof(coeb_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI returns");

class ORBextractor;

/* `mvImagePyramid` of the reference is a public std::vector<cv::Mat> that Frame::ComputeStereoMatches reads
 * (src/Frame.cc:651, 741, 753, 758). Here the levels live in HBM, so the member is a view that fetches a level from the
 * device the first time it is indexed after an extraction: reference code that says `mvImagePyramid[l]` compiles unchanged and
 * sees the right pixels (at the price of one copy per level touched); coeb_stereo_match reads the levels where they are. */
class LazyPyramid {
public:
    explicit LazyPyramid(ORBextractor* owner) : owner_(owner) {}
    size_t size() const { return mats_.size(); }
    bool empty() const { return mats_.empty(); }
    void resize(size_t n) { mats_.resize(n); fresh_.assign(n, false); }
    inline coeb_cv::Mat& operator[](size_t level);
    void invalidate() { fresh_.assign(mats_.size(), false); }
private:
    ORBextractor* owner_;
    std::vector<coeb_cv::Mat> mats_;
    std::vector<bool> fresh_;
};

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    /* reference: include/ORBextractor.h:50-51, src/ORBextractor.cc:418-477. `device` selects the GPU. */
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
        : mvImagePyramid(this), ex_(nullptr), nlevels_(nlevels), scaleFactor_(scaleFactor), nfeatures_(nfeatures), extracted_(false) {
        coeb_orb_params p;
        p.nfeatures = nfeatures; p.scale_factor = scaleFactor; p.nlevels = nlevels; p.ini_th_fast = iniThFAST; p.min_th_fast = minThFAST;
        check(coeb_extractor_create(&p, device, &ex_));
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        check(coeb_extractor_tables(ex_, nullptr, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                    mvInvLevelSigma2.data(), nullptr));
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() { coeb_extractor_destroy(ex_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    /* Classic ORB-SLAM2 form: operator()(image, mask, keypoints, descriptors). The mask is ignored, as in the reference. */
    void operator()(coeb_cv::InputArray image, coeb_cv::InputArray /*mask*/, std::vector<coeb_cv::KeyPoint>& keypoints,
                    coeb_cv::OutputArray descriptors) {
        run(image, nullptr, 0, nullptr, 0, nullptr, 0, keypoints, descriptors);
    }

    /* COEB form (include/ORBextractor.h:73-75, src/ORBextractor.cc:1088): `mask`, `img`, `imD` and `mask_result` are
     * accepted and, as in the reference, never used for the result. */
    void operator()(coeb_cv::InputArray image, coeb_cv::InputArray /*mask*/, const coeb_cv::Mat& /*img*/, const coeb_cv::Mat& /*imD*/,
                    std::vector<coeb_cv::KeyPoint>& keypoints, coeb_cv::OutputArray descriptors, std::vector<std::vector<float> >& box,
                    std::vector<coeb_cv::Point2f> T_M, coeb_cv::Mat& /*mask_result*/, std::vector<int> blur_flag) {
        std::vector<float> flat;
        flat.reserve(box.size() * 4);
        for (size_t b = 0; b < box.size(); b++) {
            if (box[b].size() < 4) throw std::invalid_argument("ORBextractor: a box needs xmin,ymin,xmax,ymax");
            flat.insert(flat.end(), box[b].begin(), box[b].begin() + 4);
        }
        static_assert(sizeof(coeb_cv::Point2f) == 2 * sizeof(float), "Point2f must be two packed floats");
        run(image, flat.data(), (int)box.size(), T_M.empty() ? nullptr : &T_M[0].x, (int)T_M.size(),
            blur_flag.empty() ? nullptr : blur_flag.data(), (int)blur_flag.size(), keypoints, descriptors);
    }

    int GetLevels() { return nlevels_; }
    float GetScaleFactor() { return scaleFactor_; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    /* include/ORBextractor.h:99 of the reference; see LazyPyramid above. */
    LazyPyramid mvImagePyramid;
    /* Copies one level of the last extraction into `m` (w x h, 8-bit); throws if nothing has been extracted yet. */
    void FetchLevel(int level, coeb_cv::Mat& m) {
        if (!extracted_) throw std::runtime_error("ORBextractor::mvImagePyramid read before the first extraction");
        int w = 0, h = 0;
        check(coeb_pyramid_level(ex_, 0, level, 0, nullptr, &w, &h, nullptr));
        coeb_adapt::create_u8(m, h, w);
        check(coeb_pyramid_level_copy(ex_, 0, level, 0, m.ptr(0)));
    }
    void SyncPyramidToHost() { for (int l = 0; l < nlevels_; l++) (void)mvImagePyramid[l]; }

    coeb_extractor* handle() { return ex_; }

private:
    static void check(int st) {
        if (st != COEB_OK) throw std::runtime_error(std::string("coeb: ") + coeb_last_error());
    }
    void run(coeb_cv::InputArray image_in, const float* boxes, int nbox, const float* tm, int ntm, const int* blur, int nblur,
             std::vector<coeb_cv::KeyPoint>& keypoints, coeb_cv::OutputArray descriptors) {
#ifdef COEB_WITH_OPENCV
        cv::Mat image = image_in.getMat();
        if (image.empty()) return;                          /* src/ORBextractor.cc:1096-1097 */
        CV_Assert(image.type() == CV_8UC1);
        const unsigned char* pix = image.data; const int w = image.cols, h = image.rows; const int stride = (int)image.step;
#else
        const coeb_cv::Mat& image = image_in;
        if (image.empty()) return;
        const unsigned char* pix = image.data; const int w = image.cols, h = image.rows; const int stride = (int)image.step;
#endif
        int cap = 0;                                         /* the octree may return more than nfeatures: ask the library for the bound */
        check(coeb_extractor_max_keypoints(ex_, w, h, &cap));
        kp_buf_.resize(cap);
        desc_buf_.resize((size_t)cap * 32);
        int n = 0;
        check(coeb_extract(ex_, pix, w, h, stride, boxes, nbox, tm, ntm, blur, nblur, kp_buf_.data(), desc_buf_.data(), cap, &n));
        extracted_ = true;
        mvImagePyramid.invalidate();                         /* the reference rebuilds the pyramid on every call (:1351) */
        keypoints.resize(n);
        if (n) std::memcpy(static_cast<void*>(keypoints.data()), kp_buf_.data(), sizeof(coeb_keypoint) * n);
#ifdef COEB_WITH_OPENCV
        if (n == 0) { descriptors.release(); return; }       /* :1296-1297 */
        descriptors.create(n, 32, CV_8U);
        cv::Mat d = descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), &desc_buf_[(size_t)i * 32], 32);
#else
        if (n == 0) { descriptors.release(); return; }
        descriptors.create(n, 32);
        std::memcpy(descriptors.ptr(0), desc_buf_.data(), (size_t)n * 32);
#endif
    }

    coeb_extractor* ex_;
    int nlevels_;
    float scaleFactor_;
    int nfeatures_;
    bool extracted_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<coeb_keypoint> kp_buf_;
    std::vector<unsigned char> desc_buf_;
};

inline coeb_cv::Mat& LazyPyramid::operator[](size_t level) {
    if (level >= mats_.size()) throw std::out_of_range("ORBextractor::mvImagePyramid: level out of range");
    if (!fresh_[level]) {
        owner_->FetchLevel((int)level, mats_[level]);
        fresh_[level] = true;
    }
    return mats_[level];
}

}  // namespace ORB_SLAM2

#endif  // ORBEXTRACTOR_H
