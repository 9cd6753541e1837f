// cvshim.hpp -- the slice of the OpenCV C++ API that the reference's hot-path sources use, so that
// /root/reference/src/{ORBextractor,ORBmatcher,Frame}.cc compile UNCHANGED into oracle/_ref/ (no OpenCV C++
// headers or libraries exist in this image).
//
// TEST INFRASTRUCTURE ONLY (see oracle/coeb_oracle.hpp's header). This file is original: it declares the
// classes and functions the reference names (cv::Mat with ROI views, Point_, KeyPoint, InputArray ...) and
// cvshim.cpp backs the image primitives (resize, copyMakeBorder, GaussianBlur, FAST, fastAtan2, ...) with the
// integer models of oracle/coeb_oracle.hpp, which tests/test_oracle_vs_cv2.py pins to OpenCV 4.13.0 bit for bit.
// What runs through the reference's own code is therefore everything the reference wrote: control flow,
// container handling, float expressions, call order. What does not: the bodies of the OpenCV primitives.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <set>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_NUMBER CV_TERMCRIT_ITER
#define CV_TERMCRIT_EPS 2
#define CV_GRAY2RGB 8
#define CV_GRAY2BGR 8
#define CV_BGR2RGB 4
#define CV_RGB2GRAY 7
#define CV_BGR2GRAY 6
#define CV_RGBA2GRAY 11
#define CV_BGRA2GRAY 10

// OpenCV's C rounding helpers live in the global namespace (core/fast_math.hpp): cvRound is round-half-to-even
// (cvtsd2si / lrint under the default rounding mode).
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> static inline T saturate_cast(int v);
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
template <> inline ushort saturate_cast<ushort>(int v) { return (ushort)(v < 0 ? 0 : (v > 65535 ? 65535 : v)); }
template <> inline short saturate_cast<short>(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

using std::string;
typedef std::string String;

template <typename T> struct DataDepth;
template <> struct DataDepth<uchar> { enum { value = CV_8U }; };
template <> struct DataDepth<signed char> { enum { value = CV_8S }; };
template <> struct DataDepth<ushort> { enum { value = CV_16U }; };
template <> struct DataDepth<short> { enum { value = CV_16S }; };
template <> struct DataDepth<int> { enum { value = CV_32S }; };
template <> struct DataDepth<float> { enum { value = CV_32F }; };
template <> struct DataDepth<double> { enum { value = CV_64F }; };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
// Point_<float> * float is evaluated in float (core/types.hpp uses saturate_cast<float>(a.x*b)); the int and
// double overloads convert the factor first, as the header's do.
static inline Point_<float> operator*(const Point_<float>& a, float b) { return Point_<float>(a.x * b, a.y * b); }
static inline Point_<float> operator*(float a, const Point_<float>& b) { return Point_<float>(b.x * a, b.y * a); }
static inline Point_<float> operator*(const Point_<float>& a, int b) { return Point_<float>(a.x * (float)b, a.y * (float)b); }
static inline Point_<float> operator*(const Point_<float>& a, double b) { return Point_<float>((float)(a.x * b), (float)(a.y * b)); }
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;
typedef Size_<int> Size2i;
typedef Size_<float> Size2f;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1 = 0, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
};
typedef Scalar_<double> Scalar;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT_MIN, INT_MAX); }
};

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int m, double e) : type(t), maxCount(m), epsilon(e) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8 };
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_GRAY2BGR = 8, COLOR_GRAY2RGB = 8, COLOR_BGRA2GRAY = 10, COLOR_RGBA2GRAY = 11 };

class Mat;

// Mat::zeros / Mat::ones / arithmetic results. Assigning one to an existing Mat of the same shape writes INTO
// that Mat's memory (cv::MatExpr semantics: `desc = Mat::zeros(...)` on a rowRange view keeps the view,
// src/ORBextractor.cc:1081). Arithmetic is evaluated eagerly into `value`.
struct MatExpr;

class Mat {
public:
    int flags;       // the type (depth + channels)
    int dims;
    int rows, cols;
    uchar* data;
    size_t step;     // bytes per row
    int* refcount;   // start of the malloc'ed block that holds the pixels, or null for user data / empty

    Mat() : flags(0), dims(2), rows(0), cols(0), data(nullptr), step(0), refcount(nullptr) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s); }
    Mat(Size sz, int type, const Scalar& s) : Mat() { create(sz.height, sz.width, type); setTo(s); }
    Mat(int r, int c, int type, void* user, size_t step_ = 0) : Mat() {
        flags = type; rows = r; cols = c; data = (uchar*)user;
        step = step_ ? step_ : (size_t)c * elemSize();
    }
    Mat(const Mat& m) : flags(m.flags), dims(m.dims), rows(m.rows), cols(m.cols), data(m.data), step(m.step), refcount(m.refcount) {
        if (refcount) __atomic_add_fetch(refcount, 1, __ATOMIC_RELAXED);
    }
    Mat(const Mat& m, const Rect& r) : Mat(m) {
        assert(r.x >= 0 && r.y >= 0 && r.width >= 0 && r.height >= 0 && r.x + r.width <= m.cols && r.y + r.height <= m.rows);
        data += (size_t)r.y * step + (size_t)r.x * elemSize();
        rows = r.height; cols = r.width;
    }
    Mat(const MatExpr& e);
    template <typename T> explicit Mat(const std::vector<T>& v) : Mat() {
        create((int)v.size(), 1, DataDepth<T>::value);
        if (!v.empty()) std::memcpy(data, v.data(), v.size() * sizeof(T));
    }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m) {
        if (this != &m) {
            if (m.refcount) __atomic_add_fetch(m.refcount, 1, __ATOMIC_RELAXED);
            release();
            flags = m.flags; dims = m.dims; rows = m.rows; cols = m.cols; data = m.data; step = m.step; refcount = m.refcount;
        }
        return *this;
    }
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Scalar& s) { return setTo(s); }

    void release() {
        if (refcount && __atomic_sub_fetch(refcount, 1, __ATOMIC_ACQ_REL) == 0) std::free(refcount);
        refcount = nullptr; data = nullptr; rows = cols = 0; step = 0;
    }
    void create(int r, int c, int type) {
        type &= 0xFFF;
        if (data && rows == r && cols == c && this->type() == type) return;
        release();
        flags = type; rows = r; cols = c;
        step = (size_t)c * elemSize();
        const size_t bytes = step * (size_t)r;
        if (bytes == 0) return;
        // pixels live in a malloc'ed block behind a 64-byte header that carries the reference count (plain
        // malloc, not operator new: the monotonic-heap variant of the _ref build replaces operator new only
        // for the reference's own containers, see ref_c.cpp)
        uchar* block = (uchar*)std::malloc(64 + bytes + 64);
        refcount = (int*)block;
        *refcount = 1;
        data = block + 64;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }

    int type() const { return flags & 0xFFF; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return sz[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }

    Mat operator()(const Rect& r) const { return Mat(*this, r); }
    Mat operator()(Range rr, Range cr) const {
        Mat m = *this;
        if (rr.start != INT_MIN) m = m.rowRange(rr.start, rr.end);
        if (cr.start != INT_MIN) m = m.colRange(cr.start, cr.end);
        return m;
    }
    Mat rowRange(int s, int e) const { return Mat(*this, Rect(0, s, cols, e - s)); }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    Mat colRange(int s, int e) const { return Mat(*this, Rect(s, 0, e - s, rows)); }
    Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat col(int j) const { return colRange(j, j + 1); }

    Mat clone() const { Mat m; copyToMat(m); return m; }
    void copyToMat(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step == step) return;
        dst.create(rows, cols, type());   // a view of the right shape is written in place (src/KeyFrame.cc:80-81)
        const size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; y++) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, rb);
    }
    inline void copyTo(const class _OutputArray& dst) const;
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat& setTo(const Scalar& s);
    Mat reshape(int cn, int rows_ = 0) const;
    Mat t() const;
    Mat inv(int method = 0) const;
    Mat mul(const Mat& m, double scale = 1) const;
    double dot(const Mat& m) const;

    static MatExpr zeros(int r, int c, int type);
    static MatExpr zeros(Size sz, int type);
    static MatExpr ones(int r, int c, int type);
    static MatExpr ones(Size sz, int type);
    static MatExpr eye(int r, int c, int type);

    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    template <typename T> T* ptr(int i, int j) { return (T*)(data + (size_t)i * step) + j; }
    template <typename T> const T* ptr(int i, int j) const { return (const T*)(data + (size_t)i * step) + j; }
    template <typename T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step))[j]; }
    template <typename T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step))[j]; }
    template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }
    template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
};

struct MatExpr {
    enum Kind { VALUE = 0, FILL = 1 } kind;
    Mat value;           // VALUE: the evaluated result
    int r, c, type;      // FILL: shape and fill value
    double fill;
    bool eye;
    MatExpr() : kind(VALUE), r(0), c(0), type(0), fill(0), eye(false) {}
    explicit MatExpr(const Mat& m) : kind(VALUE), value(m), r(m.rows), c(m.cols), type(m.type()), fill(0), eye(false) {}
    Mat eval() const { Mat m; m = *this; return m; }
    Mat t() const { return eval().t(); }
    Mat inv(int method = 0) const { return eval().inv(method); }
    operator Mat() const { return eval(); }
    template <typename T> T& at(int i, int j) { materialise(); return value.at<T>(i, j); }
    Mat row(int i) const { return eval().row(i); }
    Mat col(int j) const { return eval().col(j); }
    Mat rowRange(int s, int e) const { return eval().rowRange(s, e); }
    Mat colRange(int s, int e) const { return eval().colRange(s, e); }
    Mat clone() const { return eval().clone(); }
    void materialise() { if (kind == FILL) { Mat m; m = *this; value = m; kind = VALUE; } }
};

inline Mat::Mat(const MatExpr& e) : Mat() { *this = e; }
inline Mat& Mat::operator=(const MatExpr& e) {
    if (e.kind == MatExpr::VALUE) {
        // cv::MatExpr assignment of an evaluated operation: same shape -> written in place, else rebound
        if (data && rows == e.value.rows && cols == e.value.cols && type() == e.value.type() && data != e.value.data) {
            e.value.copyToMat(*this);
        } else {
            *this = e.value;
        }
        return *this;
    }
    create(e.r, e.c, e.type);
    if (e.eye) {
        setTo(Scalar(0));
        Mat one(1, 1, e.type);
        one.setTo(Scalar(e.fill));
        for (int i = 0; i < std::min(rows, cols); i++) std::memcpy(data + (size_t)i * step + (size_t)i * elemSize(), one.data, elemSize());
    } else {
        setTo(Scalar::all(e.fill));
    }
    return *this;
}
inline MatExpr Mat::zeros(int r, int c, int type) { MatExpr e; e.kind = MatExpr::FILL; e.r = r; e.c = c; e.type = type; e.fill = 0; return e; }
inline MatExpr Mat::zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }
inline MatExpr Mat::ones(int r, int c, int type) { MatExpr e; e.kind = MatExpr::FILL; e.r = r; e.c = c; e.type = type; e.fill = 1; return e; }
inline MatExpr Mat::ones(Size sz, int type) { return ones(sz.height, sz.width, type); }
inline MatExpr Mat::eye(int r, int c, int type) { MatExpr e = ones(r, c, type); e.eye = true; return e; }

// cv::Mat_<T>(r, c) << a, b, c ... (src/Frame.cc:853)
template <typename T> class Mat_;
template <typename T> struct MatCommaInitializer_ {
    Mat_<T>* m;
    int idx;
    MatCommaInitializer_(Mat_<T>* m_) : m(m_), idx(0) {}
    template <typename U> MatCommaInitializer_& operator,(U v);
    operator Mat() const;
    operator Mat_<T>() const;
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, DataDepth<T>::value) {}
    Mat_(const Mat& m) : Mat() { if (m.type() == DataDepth<T>::value) Mat::operator=(m); else m.convertTo(*this, DataDepth<T>::value); }
    Mat_(const MatExpr& e) : Mat_(Mat(e)) {}
    T& operator()(int i, int j) { return at<T>(i, j); }
    const T& operator()(int i, int j) const { return at<T>(i, j); }
    T& operator()(int i) { return at<T>(i); }
};
template <typename T, typename U> static inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) {
    MatCommaInitializer_<T> ci(const_cast<Mat_<T>*>(&m));
    return (ci, v);
}
template <typename T> template <typename U> inline MatCommaInitializer_<T>& MatCommaInitializer_<T>::operator,(U v) {
    m->template at<T>(idx / m->cols, idx % m->cols) = (T)v;
    idx++;
    return *this;
}
template <typename T> inline MatCommaInitializer_<T>::operator Mat() const { return *m; }
template <typename T> inline MatCommaInitializer_<T>::operator Mat_<T>() const { return *m; }

// ----- matrix arithmetic (float / double), evaluated eagerly; see cvshim.cpp for the arithmetic each one pins -----
MatExpr operator+(const Mat& a, const Mat& b);
MatExpr operator-(const Mat& a, const Mat& b);
MatExpr operator*(const Mat& a, const Mat& b);   // matrix product
MatExpr operator*(const Mat& a, double s);
MatExpr operator*(double s, const Mat& a);
MatExpr operator/(const Mat& a, double s);
MatExpr operator-(const Mat& a);
static inline MatExpr operator+(const MatExpr& a, const Mat& b) { return a.eval() + b; }
static inline MatExpr operator+(const Mat& a, const MatExpr& b) { return a + b.eval(); }
static inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return a.eval() + b.eval(); }
static inline MatExpr operator-(const MatExpr& a, const Mat& b) { return a.eval() - b; }
static inline MatExpr operator-(const Mat& a, const MatExpr& b) { return a - b.eval(); }
static inline MatExpr operator-(const MatExpr& a, const MatExpr& b) { return a.eval() - b.eval(); }
static inline MatExpr operator*(const MatExpr& a, const Mat& b) { return a.eval() * b; }
static inline MatExpr operator*(const Mat& a, const MatExpr& b) { return a * b.eval(); }
static inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return a.eval() * b.eval(); }
static inline MatExpr operator*(const MatExpr& a, double s) { return a.eval() * s; }
static inline MatExpr operator*(double s, const MatExpr& a) { return s * a.eval(); }
static inline MatExpr operator/(const MatExpr& a, double s) { return a.eval() / s; }
static inline MatExpr operator-(const MatExpr& a) { return -a.eval(); }
MatExpr abs(const Mat& a);
static inline MatExpr abs(const MatExpr& a) { return abs(a.eval()); }

// ----- InputArray / OutputArray: thin handles on a Mat or a std::vector -----
class _InputArray {
public:
    enum Kind { NONE, MAT, VEC_P2F, VEC_UCHAR, VEC_FLOAT };
    Kind kind;
    void* obj;
    Mat tmp;  // holds a MatExpr argument
    _InputArray() : kind(NONE), obj(nullptr) {}
    _InputArray(const Mat& m) : kind(MAT), obj((void*)&m) {}
    _InputArray(const MatExpr& e) : kind(MAT), obj(nullptr), tmp(e) { obj = &tmp; }
    _InputArray(const std::vector<Point2f>& v) : kind(VEC_P2F), obj((void*)&v) {}
    _InputArray(const std::vector<uchar>& v) : kind(VEC_UCHAR), obj((void*)&v) {}
    _InputArray(const std::vector<float>& v) : kind(VEC_FLOAT), obj((void*)&v) {}
    bool empty() const;
    Mat getMat() const;
    Size size() const { Mat m = getMat(); return m.size(); }
    int type() const { return getMat().type(); }
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    _OutputArray(const Mat& m) : _InputArray(m) {}   // temporaries such as `a.copyTo(b.rowRange(0,3))`, as OpenCV allows
    _OutputArray(std::vector<Point2f>& v) : _InputArray(v) {}
    _OutputArray(std::vector<uchar>& v) : _InputArray(v) {}
    _OutputArray(std::vector<float>& v) : _InputArray(v) {}
    void create(int rows, int cols, int type) const;
    void create(Size sz, int type) const { create(sz.height, sz.width, type); }
    void release() const;
    bool needed() const { return kind != NONE; }
    Mat& getMatRef() const { return *(Mat*)obj; }
};
inline void Mat::copyTo(const _OutputArray& dst) const {
    if (dst.kind == _InputArray::MAT) { copyToMat(*(Mat*)dst.obj); return; }
    Mat m = dst.getMat();
    dst.create(rows, cols, type());
    m = dst.getMat();
    copyToMat(m);
}
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
static inline OutputArray noArray() { static _OutputArray none; return none; }

// ----- core / imgproc / features2d functions named by the reference -----
float fastAtan2(float y, float x);
double norm(InputArray a, int normType = NORM_L2);
double norm(InputArray a, InputArray b, int normType = NORM_L2);
Scalar mean(InputArray a);
void multiply(InputArray a, InputArray b, OutputArray c, double scale = 1);
void sqrt(InputArray a, OutputArray b);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void cvtColor(InputArray src, OutputArray dst, int code, int dstCn = 0);
void Laplacian(InputArray src, OutputArray dst, int ddepth, int ksize = 1, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
void Sobel(InputArray src, OutputArray dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
void filter2D(InputArray src, OutputArray dst, int ddepth, InputArray kernel, Point anchor = Point(-1, -1), double delta = 0, int borderType = BORDER_DEFAULT);
void undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs, InputArray R = noArray(), InputArray P = noArray());
void goodFeaturesToTrack(InputArray image, OutputArray corners, int maxCorners, double qualityLevel, double minDistance,
                         InputArray mask = noArray(), int blockSize = 3, bool useHarrisDetector = false, double k = 0.04);
void cornerSubPix(InputArray image, InputOutputArray corners, Size winSize, Size zeroZone, TermCriteria criteria);
void calcOpticalFlowPyrLK(InputArray prevImg, InputArray nextImg, InputArray prevPts, InputOutputArray nextPts, OutputArray status,
                          OutputArray err, Size winSize = Size(21, 21), int maxLevel = 3,
                          TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                          double minEigThreshold = 1e-4);
Mat findFundamentalMat(InputArray points1, InputArray points2, OutputArray mask, int method = FM_RANSAC, double param1 = 3., double param2 = 0.99);
Mat findFundamentalMat(InputArray points1, InputArray points2, int method = FM_RANSAC, double param1 = 3., double param2 = 0.99, OutputArray mask = noArray());

// drawing / GUI: no-ops (the reference calls them for debug display only, src/ORBextractor.cc:1256-1288)
static inline void circle(const Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) {}
static inline void line(const Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
static inline void rectangle(const Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
static inline void imshow(const std::string&, const Mat&) {}
static inline int waitKey(int = 0) { return -1; }
static inline bool imwrite(const std::string&, const Mat&) { return true; }

}  // namespace cv
