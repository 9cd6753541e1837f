// cvshim.cpp -- bodies of the OpenCV functions the reference's hot-path sources call, for the oracle/_ref build.
//
// TEST INFRASTRUCTURE ONLY. The image primitives forward to the integer models of oracle/coeb_oracle.hpp
// (pinned bit for bit to OpenCV 4.13.0 by tests/test_oracle_vs_cv2.py and the goldens in tests/golden/):
//   cv::resize INTER_LINEAR 8UC1      -> orc::resize_linear_8u     (called at src/ORBextractor.cc:1356)
//   cv::copyMakeBorder REFLECT_101    -> reflect101 indexing        (:1358, :1363)
//   cv::GaussianBlur 7x7 sigma 2      -> orc::gaussian7x7_8u        (:1318)
//   cv::FAST(roi, kps, th, true)      -> orc::fast9_nms             (:831, :836)
//   cv::fastAtan2                     -> orc::fast_atan2            (:106)
//   cv::undistortPoints               -> orc::undistort_point       (src/Frame.cc:597, 627)
//   cv::Laplacian + abs + mean        -> the 16U-saturating aperture-1 Laplacian (src/Frame.cc:909-911)
// This file is always compiled with -ffp-contract=off: it stands for library code, which is not built with the
// application's -march=native.
// Functions the hot path never reaches (ProcessMovingObject's goodFeaturesToTrack / LK / findFundamentalMat, Sobel,
// filter2D ...) abort with a message: linking them silently to something else would fake coverage.
#include "cvshim.hpp"

#include "../coeb_oracle.hpp"
#include "../coeb_oracle_frame.hpp"

namespace cv {

static void not_in_scope(const char* what) {
    std::fprintf(stderr, "cvshim: %s is not modelled (outside the hot path of SURVEY.md section 8)\n", what);
    std::abort();
}

// ---------------------------------------------------------------------------------------------
// Mat members
// ---------------------------------------------------------------------------------------------
template <typename T> static void fill_all(Mat& m, double v) {
    const int n = m.cols * m.channels();
    for (int y = 0; y < m.rows; y++) {
        T* p = m.ptr<T>(y);
        for (int x = 0; x < n; x++) p[x] = (T)v;
    }
}
Mat& Mat::setTo(const Scalar& s) {
    if (empty()) return *this;
    const double v = s[0];  // every caller on the path fills with one value
    switch (depth()) {
        case CV_8U: fill_all<uchar>(*this, v); break;
        case CV_8S: fill_all<signed char>(*this, v); break;
        case CV_16U: fill_all<ushort>(*this, v); break;
        case CV_16S: fill_all<short>(*this, v); break;
        case CV_32S: fill_all<int>(*this, v); break;
        case CV_32F: fill_all<float>(*this, v); break;
        case CV_64F: fill_all<double>(*this, v); break;
    }
    return *this;
}

template <typename S> static double load_as_double(const uchar* p) { return (double)*(const S*)p; }
static double load_elem(const Mat& m, int y, int x) {
    const uchar* p = m.data + (size_t)y * m.step + (size_t)x * m.elemSize1();
    switch (m.depth()) {
        case CV_8U: return load_as_double<uchar>(p);
        case CV_8S: return load_as_double<signed char>(p);
        case CV_16U: return load_as_double<ushort>(p);
        case CV_16S: return load_as_double<short>(p);
        case CV_32S: return load_as_double<int>(p);
        case CV_32F: return load_as_double<float>(p);
        default: return load_as_double<double>(p);
    }
}

// cv::Mat::convertTo. The conversions the path uses: 8U -> 32F (exact, src/Frame.cc:742,759), 16U -> 32F with a
// scale factor (float multiply, src/Tracking.cc:229), float -> float identity.
void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
    if (rtype < 0) rtype = type();
    const int ddepth = CV_MAT_DEPTH(rtype);
    Mat src = *this;  // keeps the source alive when dst aliases it
    Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
    const int n = cols * channels();
    for (int y = 0; y < rows; y++) {
        for (int x = 0; x < n; x++) {
            const double s = load_elem(src, y, x);
            uchar* q = out.data + (size_t)y * out.step + (size_t)x * out.elemSize1();
            if (ddepth == CV_32F) {
                // cvt32f: float(src) * float(alpha) + float(beta) in float (core/convert_scale.simd.hpp)
                float f = (float)s;
                if (alpha != 1 || beta != 0) f = f * (float)alpha + (float)beta;
                *(float*)q = f;
            } else if (ddepth == CV_64F) {
                *(double*)q = s * alpha + beta;
            } else {
                const double v = s * alpha + beta;
                const long r = lrint(v);
                switch (ddepth) {
                    case CV_8U: *q = (uchar)std::min(std::max(r, 0L), 255L); break;
                    case CV_8S: *(signed char*)q = (signed char)std::min(std::max(r, -128L), 127L); break;
                    case CV_16U: *(ushort*)q = (ushort)std::min(std::max(r, 0L), 65535L); break;
                    case CV_16S: *(short*)q = (short)std::min(std::max(r, -32768L), 32767L); break;
                    default: *(int*)q = (int)r; break;
                }
            }
        }
    }
    dst = out;
}

Mat Mat::reshape(int cn, int rows_) const {
    assert(isContinuous());
    Mat m = *this;
    if (cn == 0) cn = channels();
    const size_t total_elems = (size_t)rows * cols * channels();
    const int nrows = rows_ ? rows_ : rows;
    m.rows = nrows;
    m.cols = (int)(total_elems / ((size_t)nrows * cn));
    m.flags = CV_MAKETYPE(depth(), cn);
    m.step = (size_t)m.cols * m.elemSize();
    return m;
}

Mat Mat::t() const {
    Mat out(cols, rows, type());
    const size_t es = elemSize();
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) std::memcpy(out.data + (size_t)x * out.step + (size_t)y * es, data + (size_t)y * step + (size_t)x * es, es);
    return out;
}
Mat Mat::inv(int) const { not_in_scope("Mat::inv"); return Mat(); }
Mat Mat::mul(const Mat&, double) const { not_in_scope("Mat::mul"); return Mat(); }
double Mat::dot(const Mat& m) const {
    // cv::Mat::dot for CV_32F accumulates in double (core/matmul.simd.hpp dotProd_32f)
    double s = 0;
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) s += load_elem(*this, y, x) * load_elem(m, y, x);
    return s;
}

// ---------------------------------------------------------------------------------------------
// Matrix arithmetic. The reference only combines 3x3 / 3x1 / 4x4 CV_32F matrices (poses and points). Each operator is
// evaluated on its own in fp32 without FMA: a product element is ((a0*b0 + a1*b1) + a2*b2) ..., then `+ t`
// adds one more float. That is the "projection arithmetic" oracle/coeb_oracle_frame.hpp pins against cv::gemm on
// 20 000 random cases (tests/golden/frame_tail.npz); cv::MatExpr folds `A*B + C` into one gemm call whose small-matrix
// path evaluates exactly this sum order.
// ---------------------------------------------------------------------------------------------
template <typename F> static Mat binary_float(const Mat& a, const Mat& b, F f) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    Mat out(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) {
        if (a.depth() == CV_32F) {
            const float *p = a.ptr<float>(y), *q = b.ptr<float>(y);
            float* o = out.ptr<float>(y);
            for (int x = 0; x < n; x++) o[x] = (float)f((float)p[x], (float)q[x]);
        } else if (a.depth() == CV_64F) {
            const double *p = a.ptr<double>(y), *q = b.ptr<double>(y);
            double* o = out.ptr<double>(y);
            for (int x = 0; x < n; x++) o[x] = f(p[x], q[x]);
        } else {
            not_in_scope("integer matrix arithmetic");
        }
    }
    return out;
}
MatExpr operator+(const Mat& a, const Mat& b) {
    if (a.depth() == CV_32F) return MatExpr(binary_float(a, b, [](float x, float y) { return x + y; }));
    return MatExpr(binary_float(a, b, [](double x, double y) { return x + y; }));
}
MatExpr operator-(const Mat& a, const Mat& b) {
    if (a.depth() == CV_32F) return MatExpr(binary_float(a, b, [](float x, float y) { return x - y; }));
    return MatExpr(binary_float(a, b, [](double x, double y) { return x - y; }));
}
MatExpr operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows && a.type() == b.type() && a.channels() == 1);
    Mat out(a.rows, b.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            if (a.depth() == CV_32F) {
                float s = a.at<float>(i, 0) * b.at<float>(0, j);
                for (int k = 1; k < a.cols; k++) s = s + a.at<float>(i, k) * b.at<float>(k, j);
                out.at<float>(i, j) = s;
            } else if (a.depth() == CV_64F) {
                double s = a.at<double>(i, 0) * b.at<double>(0, j);
                for (int k = 1; k < a.cols; k++) s = s + a.at<double>(i, k) * b.at<double>(k, j);
                out.at<double>(i, j) = s;
            } else {
                not_in_scope("integer matrix product");
            }
        }
    return MatExpr(out);
}
MatExpr operator*(const Mat& a, double s) {
    Mat out(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < n; x++) {
            if (a.depth() == CV_32F) out.ptr<float>(y)[x] = (float)(a.ptr<float>(y)[x] * s);   // double scale, rounded once
            else if (a.depth() == CV_64F) out.ptr<double>(y)[x] = a.ptr<double>(y)[x] * s;
            else not_in_scope("integer matrix scaling");
        }
    return MatExpr(out);
}
MatExpr operator*(double s, const Mat& a) { return a * s; }
MatExpr operator/(const Mat& a, double s) { return a * (1.0 / s); }   // MatExpr: A / s is A scaled by 1/s
MatExpr operator-(const Mat& a) { return a * -1.0; }
MatExpr abs(const Mat& a) {
    Mat out(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < n; x++) {
            switch (a.depth()) {
                case CV_8U: out.ptr<uchar>(y)[x] = a.ptr<uchar>(y)[x]; break;
                case CV_16U: out.ptr<ushort>(y)[x] = a.ptr<ushort>(y)[x]; break;
                case CV_16S: { int v = std::abs((int)a.ptr<short>(y)[x]); out.ptr<short>(y)[x] = (short)std::min(v, 32767); break; }
                case CV_32F: out.ptr<float>(y)[x] = std::fabs(a.ptr<float>(y)[x]); break;
                case CV_64F: out.ptr<double>(y)[x] = std::fabs(a.ptr<double>(y)[x]); break;
                default: not_in_scope("abs of this depth");
            }
        }
    return MatExpr(out);
}

// ---------------------------------------------------------------------------------------------
// InputArray / OutputArray
// ---------------------------------------------------------------------------------------------
bool _InputArray::empty() const {
    switch (kind) {
        case MAT: return ((const Mat*)obj)->empty();
        case VEC_P2F: return ((const std::vector<Point2f>*)obj)->empty();
        case VEC_UCHAR: return ((const std::vector<uchar>*)obj)->empty();
        case VEC_FLOAT: return ((const std::vector<float>*)obj)->empty();
        default: return true;
    }
}
Mat _InputArray::getMat() const {
    switch (kind) {
        case MAT: return *(const Mat*)obj;
        case VEC_P2F: { auto* v = (std::vector<Point2f>*)obj; return v->empty() ? Mat() : Mat((int)v->size(), 1, CV_32FC2, v->data()); }
        case VEC_UCHAR: { auto* v = (std::vector<uchar>*)obj; return v->empty() ? Mat() : Mat((int)v->size(), 1, CV_8UC1, v->data()); }
        case VEC_FLOAT: { auto* v = (std::vector<float>*)obj; return v->empty() ? Mat() : Mat((int)v->size(), 1, CV_32FC1, v->data()); }
        default: return Mat();
    }
}
void _OutputArray::create(int rows, int cols, int type) const {
    switch (kind) {
        case MAT: ((Mat*)obj)->create(rows, cols, type); break;
        case VEC_P2F: ((std::vector<Point2f>*)obj)->resize((size_t)rows * cols); break;
        case VEC_UCHAR: ((std::vector<uchar>*)obj)->resize((size_t)rows * cols); break;
        case VEC_FLOAT: ((std::vector<float>*)obj)->resize((size_t)rows * cols); break;
        default: break;
    }
}
void _OutputArray::release() const {
    switch (kind) {
        case MAT: ((Mat*)obj)->release(); break;
        case VEC_P2F: ((std::vector<Point2f>*)obj)->clear(); break;
        case VEC_UCHAR: ((std::vector<uchar>*)obj)->clear(); break;
        case VEC_FLOAT: ((std::vector<float>*)obj)->clear(); break;
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------
// core
// ---------------------------------------------------------------------------------------------
float fastAtan2(float y, float x) { return orc::fast_atan2(y, x); }

// cv::norm. NORM_L2 of CV_32F accumulates the squares in double and returns sqrt in double (core/norm.cpp, normL2_32f);
// NORM_L1 of a CV_32F difference accumulates |a-b| in double. Both exact for the path's integer-valued SAD patches.
double norm(InputArray a_, int normType) {
    Mat a = a_.getMat();
    double s = 0;
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < n; x++) {
            const double v = load_elem(a, y, x);
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(v);
            else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
double norm(InputArray a_, InputArray b_, int normType) {
    Mat a = a_.getMat(), b = b_.getMat();
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    double s = 0;
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < n; x++) {
            double v;
            if (a.depth() == CV_32F) v = (double)(a.ptr<float>(y)[x] - b.ptr<float>(y)[x]);   // float difference, as normDiffL1_32f
            else v = load_elem(a, y, x) - load_elem(b, y, x);
            if (normType == NORM_L2) s += v * v;
            else if (normType == NORM_L1) s += std::fabs(v);
            else s = std::max(s, std::fabs(v));
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
Scalar mean(InputArray a_) {
    Mat a = a_.getMat();
    assert(a.channels() == 1);
    double s = 0;
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) s += load_elem(a, y, x);
    return Scalar(a.total() ? s / (double)a.total() : 0);
}
void multiply(InputArray, InputArray, OutputArray, double) { not_in_scope("cv::multiply"); }
void sqrt(InputArray, OutputArray) { not_in_scope("cv::sqrt"); }

// ---------------------------------------------------------------------------------------------
// imgproc / features2d
// ---------------------------------------------------------------------------------------------
void resize(InputArray src_, OutputArray dst_, Size dsize, double fx, double fy, int interpolation) {
    Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && fx == 0 && fy == 0);
    dst_.create(dsize.height, dsize.width, src.type());   // no reallocation when the header already has this shape (:1354-1356)
    Mat dst = dst_.getMat();
    orc::resize_linear_8u(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar&) {
    Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && (borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    // A view is treated as isolated in both calls: at :1358 the flag says so; at :1363 the source is the caller's
    // whole image. (Without BORDER_ISOLATED OpenCV would read a view's surroundings instead of reflecting.)
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    // the source may be the centre of dst itself (:1358): take a private copy before the border is written
    Mat s = src.clone();
    for (int y = 0; y < dst.rows; y++) {
        const uchar* srow = s.ptr(orc::reflect101(y - top, s.rows));
        uchar* drow = dst.ptr(y);
        for (int x = 0; x < dst.cols; x++) drow[x] = srow[orc::reflect101(x - left, s.cols)];
    }
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
    Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat s = src.clone();   // the reference blurs in place (:1318); the model reads a private copy (a clone is its own image: reflection at its edges)
    dst_.create(s.rows, s.cols, s.type());
    Mat dst = dst_.getMat();
    orc::gaussian7x7_8u(s.data, s.cols, s.rows, (int)s.step, dst.data, (int)dst.step);
}

void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    Mat im = image_.getMat();
    assert(im.type() == CV_8UC1 && nonmaxSuppression);
    std::vector<orc::FastPt> pts;   // no static cache: under the monotonic heap nothing may outlive the call that allocated it
    orc::fast9_nms(im.data, im.cols, im.rows, (int)im.step, threshold, pts);
    keypoints.clear();
    keypoints.reserve(pts.size());
    for (const orc::FastPt& p : pts) keypoints.push_back(KeyPoint((float)p.x, (float)p.y, 7.f, -1.f, (float)p.score));
}

// cv::KeyPointsFilter::retainBest (features2d/keypoint.cpp): keep the npoints strongest and every keypoint that ties
// with the weakest kept one. Only reached from the dead ComputeKeyPointsOld (src/ORBextractor.cc:906-1076).
void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int n_points) {
    if (n_points >= 0 && keypoints.size() > (size_t)n_points) {
        if (n_points == 0) { keypoints.clear(); return; }
        std::nth_element(keypoints.begin(), keypoints.begin() + n_points - 1, keypoints.end(),
                         [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        const float ambiguous = keypoints[n_points - 1].response;
        auto new_end = std::partition(keypoints.begin() + n_points, keypoints.end(),
                                      [ambiguous](const KeyPoint& k) { return k.response >= ambiguous; });
        keypoints.resize(new_end - keypoints.begin());
    }
}

void cvtColor(InputArray src_, OutputArray dst_, int code, int) {
    Mat src = src_.getMat();
    const bool bgr = (code == COLOR_BGR2GRAY || code == COLOR_BGRA2GRAY);
    assert(src.depth() == CV_8U && (src.channels() == 3 || src.channels() == 4));
    assert(code == COLOR_BGR2GRAY || code == COLOR_RGB2GRAY || code == COLOR_BGRA2GRAY || code == COLOR_RGBA2GRAY);
    Mat s = src.clone();
    dst_.create(s.rows, s.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    orc::rgb_to_gray_8u(s.data, s.cols, s.rows, (int)s.step, s.channels(), bgr, dst.data, (int)dst.step);
}

// cv::Laplacian(gray, lap, CV_16U) with the default aperture 1: kernel [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101,
// result saturated to 16U (negative values become 0). src/Frame.cc:909.
void Laplacian(InputArray src_, OutputArray dst_, int ddepth, int ksize, double scale, double delta, int borderType) {
    Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && ddepth == CV_16U && ksize == 1 && scale == 1 && delta == 0 && borderType == BORDER_DEFAULT);
    Mat s = src.clone();
    dst_.create(s.rows, s.cols, CV_16UC1);
    Mat dst = dst_.getMat();
    for (int y = 0; y < s.rows; y++) {
        const uchar* r0 = s.ptr(orc::reflect101(y - 1, s.rows));
        const uchar* r1 = s.ptr(y);
        const uchar* r2 = s.ptr(orc::reflect101(y + 1, s.rows));
        ushort* d = dst.ptr<ushort>(y);
        for (int x = 0; x < s.cols; x++) {
            const int v = r0[x] + r2[x] + r1[orc::reflect101(x - 1, s.cols)] + r1[orc::reflect101(x + 1, s.cols)] - 4 * r1[x];
            d[x] = (ushort)std::min(std::max(v, 0), 65535);
        }
    }
}
void Sobel(InputArray, OutputArray, int, int, int, int, double, double, int) { not_in_scope("cv::Sobel"); }
void filter2D(InputArray, OutputArray, int, InputArray, Point, double, int) { not_in_scope("cv::filter2D"); }

// cv::undistortPoints(mat, mat, K, dist, Mat(), K) on an N x 1 CV_32FC2 array (src/Frame.cc:596-598, 626-628)
void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray D_, InputArray R_, InputArray P_) {
    Mat src = src_.getMat(), K = K_.getMat(), D = D_.getMat(), P = P_.getMat();
    assert(src.type() == CV_32FC2 && K.type() == CV_32F && R_.empty());
    assert(!P.empty() && P.data == K.data);   // the reference passes mK as the new projection
    Mat s = src.clone();
    dst_.create(s.rows, s.cols, CV_32FC2);
    Mat dst = dst_.getMat();
    float dist[5] = {0, 0, 0, 0, 0};
    const int nd = (int)D.total();
    for (int i = 0; i < nd && i < 5; i++) dist[i] = D.rows == 1 ? D.at<float>(0, i) : D.at<float>(i, 0);
    const float fx = K.at<float>(0, 0), fy = K.at<float>(1, 1), cx = K.at<float>(0, 2), cy = K.at<float>(1, 2);
    const int n = s.rows * s.cols;
    for (int i = 0; i < n; i++) {
        const float* p = (const float*)(s.data + (size_t)(i / s.cols) * s.step) + 2 * (i % s.cols);
        float* q = (float*)(dst.data + (size_t)(i / dst.cols) * dst.step) + 2 * (i % dst.cols);
        orc::undistort_point(p[0], p[1], fx, fy, cx, cy, dist, &q[0], &q[1]);
    }
}

void goodFeaturesToTrack(InputArray, OutputArray, int, double, double, InputArray, int, bool, double) { not_in_scope("cv::goodFeaturesToTrack"); }
void cornerSubPix(InputArray, InputOutputArray, Size, Size, TermCriteria) { not_in_scope("cv::cornerSubPix"); }
void calcOpticalFlowPyrLK(InputArray, InputArray, InputArray, InputOutputArray, OutputArray, OutputArray, Size, int, TermCriteria, int, double) {
    not_in_scope("cv::calcOpticalFlowPyrLK");
}
Mat findFundamentalMat(InputArray, InputArray, OutputArray, int, double, double) { not_in_scope("cv::findFundamentalMat"); return Mat(); }
Mat findFundamentalMat(InputArray, InputArray, int, double, double, OutputArray) { not_in_scope("cv::findFundamentalMat"); return Mat(); }

}  // namespace cv
