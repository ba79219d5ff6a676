// TEST INFRASTRUCTURE ONLY — a types-and-trampolines stand-in for <opencv2/opencv.hpp>.
//
// Purpose: let g++ compile the reference's OWN sources (src/core.cpp, src/objdetect.cpp, src/imgproc.cpp, read where
// they lie under /root/reference, unmodified) into oracle/_ref/librmcv_ref.so on an image that has no OpenCV C++
// headers or libraries.  Everything that is rm:: arithmetic (constructors, gates, geometry helpers, libstdc++ overload
// resolution of abs/atan2/sin/cos/pow/round/fmax) is then the reference's compiled code.  Everything that is OpenCV
// arithmetic is NOT restated here: each cv:: function below forwards to one callback (`rmcv_ref_cvcall`) that the
// Python side (oracle/ref_bridge.py) serves with the real OpenCV of this image (cv2 4.13.0).  The only cv:: code
// written out here is what OpenCV itself defines inline in its public headers (Point_/Size_/Rect_ members,
// `Rect_ & Rect_`, the Point_<float> -> Point_<int> conversion through saturate_cast = cvRound), because a real build
// would compile exactly that header code with the reference's own compiler flags.
//
// Overload environment: opencv2/core/cvdef.h pulls in <emmintrin.h> on x86-64 (cv_cpu_dispatch.h, CV_SSE2), which
// reaches libstdc++'s <stdlib.h> wrapper through mm_malloc.h and so puts `using std::abs` into the global namespace;
// that is what makes the reference's unqualified abs(float) a float abs.  This header includes
// <emmintrin.h> for the same reason and nothing else that would widen the global overload set.  Whether a real
// OpenCV include chain also reaches the <math.h> wrapper (`using std::atan2/sin/cos`, float overloads) cannot be
// checked on this image; oracle/Makefile builds a second library with -DRMCV_CVSTUB_WITH_MATH_H for that case and
// tests/test_ref_pin.py bounds the difference.
#pragma once

#include <emmintrin.h>
#ifdef RMCV_CVSTUB_WITH_MATH_H
#include <math.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

typedef unsigned char uchar;
typedef int64_t int64;

#define CV_PI 3.1415926535897932384626433832795

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
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32SC2 CV_MAKETYPE(CV_32S, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

// ---------------------------------------------------------------------------------------------- the one trampoline
extern "C" {
struct rmcv_ref_arr {   // a dense 2-D array handed across the callback: rows x cols elements of `type`, `step` bytes per row
    void* data;
    int32_t rows, cols, type;
    int64_t step;
    int64_t owner;      // out arrays: non-zero = the callee keeps `data` alive until rmcv_ref_release(owner) is called
};
// op: the cv:: function name.  in/out: arrays; params: scalar arguments.  The callee fills out[i].data with a pointer
// that stays valid until the next call (the stub copies it).  Returns 0 on success.
typedef int (*rmcv_ref_cvcall_t)(const char* op, const rmcv_ref_arr* in, int n_in, const double* params, int n_params,
                                 rmcv_ref_arr* out, int n_out);
typedef void (*rmcv_ref_release_t)(int64_t owner);
extern rmcv_ref_cvcall_t rmcv_ref_cvcall;
extern rmcv_ref_release_t rmcv_ref_release;
extern double rmcv_ref_tick_frequency;
}

namespace cv {

inline int elem_size_of(int type) {
    static const int d[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return d[CV_MAT_DEPTH(type)] * CV_MAT_CN(type);
}

// ---- opencv2/core/fast_math.hpp / saturate.hpp (inline header code)
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
template <typename T> inline T saturate_cast(double v) { return T(v); }
template <typename T> inline T saturate_cast(float v) { return T(v); }
template <typename T> inline T saturate_cast(int v) { return T(v); }
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <> inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>(cvRound(v)); }
template <> inline uchar saturate_cast<uchar>(float v) { return saturate_cast<uchar>(cvRound(v)); }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

// ---- opencv2/core/types.hpp (inline header code)
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename T2> operator Point_<T2>() const { return Point_<T2>(saturate_cast<T2>(x), saturate_cast<T2>(y)); }
};
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) {
    return Point_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y));
}
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) {
    return Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y));
}
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

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
    bool empty() const { return width <= 0 || height <= 0; }
    template <typename T2> operator Size_<T2>() const { return Size_<T2>(saturate_cast<T2>(width), saturate_cast<T2>(height)); }
};
typedef Size_<int> Size2i;
typedef Size_<float> Size2f;
typedef Size2i Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
    template <typename T2> operator Rect_<T2>() const {
        return Rect_<T2>(saturate_cast<T2>(x), saturate_cast<T2>(y), saturate_cast<T2>(width), saturate_cast<T2>(height));
    }
};
typedef Rect_<int> Rect2i;
typedef Rect_<float> Rect2f;
typedef Rect2i Rect;

// Rect_ intersection as written in OpenCV >= 4.5.3's types.hpp (the reference pins >= 4.8.0)
template <typename T> static inline Rect_<T>& operator&=(Rect_<T>& a, const Rect_<T>& b) {
    if (a.empty() || b.empty()) { a = Rect_<T>(); return a; }
    const Rect_<T>& Rx_min = (a.x < b.x) ? a : b;
    const Rect_<T>& Rx_max = (a.x < b.x) ? b : a;
    const Rect_<T>& Ry_min = (a.y < b.y) ? a : b;
    const Rect_<T>& Ry_max = (a.y < b.y) ? b : a;
    if ((Rx_min.x < 0 && Rx_min.x + Rx_min.width < Rx_max.x) || (Ry_min.y < 0 && Ry_min.y + Ry_min.height < Ry_max.y)) {
        a = Rect_<T>();
        return a;
    }
    a.width = std::min(Rx_min.width - (Rx_max.x - Rx_min.x), Rx_max.width);
    a.height = std::min(Ry_min.height - (Ry_max.y - Ry_min.y), Ry_max.height);
    a.x = Rx_max.x;
    a.y = Ry_max.y;
    if (a.empty()) a = Rect_<T>();
    return a;
}
template <typename T> static inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) {
    Rect_<T> c = a;
    return c &= b;
}

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar_<T> all(T v) { return Scalar_<T>(v, v, v, v); }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

// ---------------------------------------------------------------------------------------------- Mat (storage only)
class Mat;
class MatExpr;
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_;

struct MatStep {   // cv::MatStep: bytes per row, converts to size_t
    size_t v = 0;
    operator size_t() const { return v; }
    MatStep& operator=(size_t s) { v = s; return *this; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;   // first element (like cv::Mat::data)
    MatStep step;            // bytes per row (like cv::Mat::step)
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext) {   // a header over the caller's buffer, like cv::Mat(rows, cols, type, data)
        rows = r; cols = c; type_ = type; step = (size_t)c * elem_size_of(type); data = static_cast<uchar*>(ext);
    }
    Mat(const MatExpr& e);
    template <typename T> Mat(const MatCommaInitializer_<T>& ci);

    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elem_size_of(type);
        auto buf = std::make_shared<std::vector<uchar>>((size_t)std::max<int64_t>(1, (int64_t)r * (int64_t)step.v), (uchar)0);
        data = buf->data();
        hold_ = buf;
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < std::min(r, c); ++i) m.set_double(i, i, 1.0);
        return m;
    }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step.v == (size_t)cols * elem_size_of(type_); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step.v); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step.v); }
    template <typename T> T& at(int i) {
        return rows == 1 ? reinterpret_cast<T*>(ptr(0))[i] : (cols == 1 ? *reinterpret_cast<T*>(ptr(i)) : reinterpret_cast<T*>(ptr(i / cols))[i % cols]);
    }
    template <typename T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }
    template <typename T> T& at(int i, int j) { return reinterpret_cast<T*>(ptr(i))[j]; }
    template <typename T> const T& at(int i, int j) const { return reinterpret_cast<const T*>(ptr(i))[j]; }
    Mat operator()(const Rect& r) const {   // a view that shares storage
        Mat m = *this;
        m.data = data + (size_t)r.y * step.v + (size_t)r.x * elem_size_of(type_);
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elem_size_of(type_));
        return m;
    }
    void copyTo(Mat& dst) const {
        if (dst.rows != rows || dst.cols != cols || dst.type_ != type_ || dst.data == nullptr) dst.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(dst.ptr(r), ptr(r), (size_t)cols * elem_size_of(type_));
    }
    void copyTo(const Mat& dst_view) const {   // cv::OutputArray built from a temporary header (a ROI)
        Mat d = dst_view;
        if (d.rows != rows || d.cols != cols || d.type_ != type_) throw std::runtime_error("cvstub: copyTo size mismatch");
        for (int r = 0; r < rows; ++r) std::memcpy(d.ptr(r), ptr(r), (size_t)cols * elem_size_of(type_));
    }
    Mat reshape(int cn, int new_rows = 0) const {
        Mat c = isContinuous() ? *this : clone();
        const int64_t elems = (int64_t)rows * cols * channels();
        Mat m = c;
        if (cn == 0) cn = channels();
        if (new_rows == 0) new_rows = rows;
        m.type_ = CV_MAKETYPE(depth(), cn);
        m.rows = new_rows;
        m.cols = (int)(elems / ((int64_t)new_rows * cn));
        m.step = (size_t)m.cols * elem_size_of(m.type_);
        return m;
    }
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const;
    double get_double(int i, int j) const {
        switch (depth()) {
            case CV_8U: return at<uchar>(i, j);
            case CV_32S: return at<int>(i, j);
            case CV_32F: return at<float>(i, j);
            case CV_64F: return at<double>(i, j);
        }
        throw std::runtime_error("cvstub: depth");
    }
    void set_double(int i, int j, double v) {
        switch (depth()) {
            case CV_8U: at<uchar>(i, j) = saturate_cast<uchar>(v); return;
            case CV_32S: at<int>(i, j) = saturate_cast<int>(v); return;
            case CV_32F: at<float>(i, j) = (float)v; return;
            case CV_64F: at<double>(i, j) = v; return;
        }
        throw std::runtime_error("cvstub: depth");
    }
    rmcv_ref_arr as_arr() const {
        rmcv_ref_arr a;
        a.data = data; a.rows = rows; a.cols = cols; a.type = type_; a.step = (int64_t)step.v; a.owner = 0;
        return a;
    }
    static Mat from_arr(const rmcv_ref_arr& a) {   // a callback result: adopted without a copy when the callee owns it
        Mat m;
        if (a.rows <= 0 || a.cols <= 0 || a.data == nullptr) {
            m.type_ = a.type;
            if (a.owner && rmcv_ref_release) rmcv_ref_release(a.owner);
            return m;
        }
        if (a.owner && rmcv_ref_release) {   // like a cv::Mat the real function would have allocated: one buffer, no extra copy
            m.rows = a.rows; m.cols = a.cols; m.type_ = a.type; m.step = (size_t)a.step; m.data = static_cast<uchar*>(a.data);
            const int64_t owner = a.owner;
            m.hold_ = std::shared_ptr<void>(nullptr, [owner](void*) { if (rmcv_ref_release) rmcv_ref_release(owner); });
            return m;
        }
        m.create(a.rows, a.cols, a.type);
        for (int r = 0; r < a.rows; ++r)
            std::memcpy(m.ptr(r), static_cast<const uchar*>(a.data) + (int64_t)r * a.step, (size_t)a.cols * elem_size_of(a.type));
        return m;
    }

private:
    std::shared_ptr<void> hold_;   // keeps the storage behind `data` alive (own allocation or a callee-owned buffer)
    int type_ = 0;
};

// one call through the trampoline: arrays in, scalars in, arrays out (deep-copied)
inline std::vector<Mat> cvcall(const char* op, const std::vector<Mat>& in, const std::vector<double>& params, int n_out) {
    if (!rmcv_ref_cvcall) throw std::runtime_error("cvstub: no OpenCV callback installed");
    std::vector<rmcv_ref_arr> ain, aout((size_t)std::max(1, n_out));
    for (const Mat& m : in) ain.push_back(m.as_arr());
    std::memset(aout.data(), 0, aout.size() * sizeof(rmcv_ref_arr));
    const int rc = rmcv_ref_cvcall(op, ain.data(), (int)ain.size(), params.data(), (int)params.size(), aout.data(), n_out);
    if (rc != 0) throw std::runtime_error(std::string("cvstub: OpenCV callback failed for ") + op);
    std::vector<Mat> out;
    for (int i = 0; i < n_out; ++i) out.push_back(Mat::from_arr(aout[(size_t)i]));
    return out;
}

inline void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
    dst = cvcall("convertTo", {*this}, {(double)rtype, alpha, beta}, 1)[0];
}

// `a - b`, `a * b` on matrices: evaluated by the real cv::subtract / cv::gemm when converted to a Mat
class MatExpr {
public:
    char op; Mat a, b;
    MatExpr(char op_, const Mat& a_, const Mat& b_) : op(op_), a(a_), b(b_) {}
    Mat eval() const { return cvcall(op == '-' ? "subtract" : "matmul", {a, b}, {}, 1)[0]; }
};
inline Mat::Mat(const MatExpr& e) { *this = e.eval(); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return MatExpr('-', a, b); }
inline MatExpr operator*(const Mat& a, const Mat& b) { return MatExpr('*', a, b); }
inline MatExpr operator*(const MatExpr& a, const Mat& b) { return MatExpr('*', a.eval(), b); }

template <typename T> struct DepthOf;
template <> struct DepthOf<uchar> { enum { value = CV_8U }; };
template <> struct DepthOf<int> { enum { value = CV_32S }; };
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DepthOf<T>::value) {}
    Mat_(int r, int c, const T& v) : Mat(r, c, DepthOf<T>::value) {
        for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) this->template at<T>(i, j) = v;
    }
};
template <typename T> class MatCommaInitializer_ {
public:
    Mat_<T> m; int idx = 0;
    explicit MatCommaInitializer_(const Mat_<T>& m_) : m(m_) {}
    template <typename T2> MatCommaInitializer_<T>& operator,(T2 v) {
        if (idx < (int)m.total()) m.template at<T>(idx / m.cols, idx % m.cols) = T(v);
        ++idx;
        return *this;
    }
    operator Mat_<T>() const { return m; }
};
template <typename T, typename T2> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 v) {
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}
template <typename T> inline Mat::Mat(const MatCommaInitializer_<T>& ci) { *this = static_cast<const Mat&>(ci.m); }

// ---------------------------------------------------------------------------------------------- proxies
class _InputArray {
public:
    enum KindFlag { NONE = 0, MAT = 1 << 16, STD_VECTOR = 3 << 16, STD_VECTOR_MAT = 5 << 16 };
    Mat m; Scalar s; bool is_scalar = false;
    _InputArray() {}
    _InputArray(const Mat& m_) : m(m_) {}
    _InputArray(const MatExpr& e) : m(e.eval()) {}
    _InputArray(const double& v) : s(v), is_scalar(true) {}
    _InputArray(const Scalar& v) : s(v), is_scalar(true) {}
    template <typename T> _InputArray(const Mat_<T>& m_) : m(m_) {}
    template <typename T> _InputArray(const MatCommaInitializer_<T>& ci) : m(ci.m) {}
    template <typename T> _InputArray(const std::vector<Point_<T>>& v) : vec_(true) {
        m.create((int)v.size(), 2, DepthOf<T>::value);
        for (size_t i = 0; i < v.size(); ++i) { m.at<T>((int)i, 0) = v[i].x; m.at<T>((int)i, 1) = v[i].y; }
    }
    template <typename T> _InputArray(const std::vector<Point3_<T>>& v) : vec_(true) {
        m.create((int)v.size(), 3, DepthOf<T>::value);
        for (size_t i = 0; i < v.size(); ++i) { m.at<T>((int)i, 0) = v[i].x; m.at<T>((int)i, 1) = v[i].y; m.at<T>((int)i, 2) = v[i].z; }
    }
    int kind() const { return vec_ ? STD_VECTOR : MAT; }
    Mat getMat() const { return m; }
private:
    bool vec_ = false;
};
typedef const _InputArray& InputArray;
typedef InputArray InputArrayOfArrays;

class _OutputArray {
public:
    enum { MAT = _InputArray::MAT };
    Mat* m = nullptr; std::vector<Mat>* vm = nullptr; std::vector<std::vector<Point>>* vvp = nullptr;
    _OutputArray(Mat& m_) : m(&m_) {}
    _OutputArray(std::vector<Mat>& v) : vm(&v) {}
    _OutputArray(std::vector<std::vector<Point>>& v) : vvp(&v) {}
    void create(Size sz, int type) const { if (m) m->create(sz.height, sz.width, type == MAT ? CV_64F : type); }
    Mat getMat() const { return m ? *m : Mat(); }
    void assign(const Mat& v) const { if (m) *m = v; }
};
typedef const _OutputArray& OutputArray;
typedef OutputArray OutputArrayOfArrays;

// ---------------------------------------------------------------------------------------------- RotatedRect
class RotatedRect {
public:
    Point2f center; Size2f size; float angle = 0;
    RotatedRect() {}
    RotatedRect(const Point2f& c, const Size2f& s, float a) : center(c), size(s), angle(a) {}
    void points(Point2f pts[]) const {   // modules/core/src/types.cpp, compiled library code -> real OpenCV (cv::boxPoints)
        Mat box(1, 5, CV_32F);
        box.at<float>(0) = center.x; box.at<float>(1) = center.y; box.at<float>(2) = size.width; box.at<float>(3) = size.height;
        box.at<float>(4) = angle;
        const Mat r = cvcall("boxPoints", {box}, {}, 1)[0];
        for (int i = 0; i < 4; ++i) { pts[i].x = r.at<float>(i, 0); pts[i].y = r.at<float>(i, 1); }
    }
};
inline RotatedRect rrect_from(const Mat& r) {
    return RotatedRect(Point2f(r.at<float>(0), r.at<float>(1)), Size2f(r.at<float>(2), r.at<float>(3)), r.at<float>(4));
}

// ---------------------------------------------------------------------------------------------- constants
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
enum { RETR_EXTERNAL = 0, RETR_LIST = 1, RETR_CCOMP = 2, RETR_TREE = 3 };
enum { CHAIN_APPROX_NONE = 1, CHAIN_APPROX_SIMPLE = 2 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { COLOR_BGR2GRAY = 6 };
enum { BORDER_CONSTANT = 0 };
enum { SOLVEPNP_ITERATIVE = 0, SOLVEPNP_IPPE = 6, SOLVEPNP_IPPE_SQUARE = 7 };
enum { IMREAD_COLOR = 1 };

// ---------------------------------------------------------------------------------------------- functions -> real OpenCV
inline double getTickFrequency() { return rmcv_ref_tick_frequency; }

inline void split(InputArray src, OutputArrayOfArrays mv) {
    std::vector<Mat> out = cvcall("split", {src.getMat()}, {}, src.getMat().channels());
    if (mv.vm) *mv.vm = out;
}
inline void inRange(InputArray src, InputArray lowerb, InputArray upperb, OutputArray dst) {
    std::vector<double> p;
    for (int i = 0; i < 4; ++i) p.push_back(lowerb.s.val[i]);
    for (int i = 0; i < 4; ++i) p.push_back(upperb.s.val[i]);
    dst.assign(cvcall("inRange", {src.getMat()}, p, 1)[0]);
}
inline Mat getStructuringElement(int shape, Size ksize) {
    return cvcall("getStructuringElement", {}, {(double)shape, (double)ksize.width, (double)ksize.height}, 1)[0];
}
inline void morphologyEx(InputArray src, OutputArray dst, int op, InputArray kernel) {
    dst.assign(cvcall("morphologyEx", {src.getMat(), kernel.getMat()}, {(double)op}, 1)[0]);
}
inline void findContours(InputArray image, OutputArrayOfArrays contours, int mode, int method) {
    const std::vector<Mat> r = cvcall("findContours", {image.getMat()}, {(double)mode, (double)method}, 2);
    const Mat& pts = r[0]; const Mat& off = r[1];   // points [N x 2] int32, offsets [1 x (n+1)] int32
    std::vector<std::vector<Point>>& out = *contours.vvp;
    out.clear();
    const int n = off.cols > 0 ? off.cols - 1 : 0;
    for (int k = 0; k < n; ++k) {
        std::vector<Point> c;
        for (int i = off.at<int>(k); i < off.at<int>(k + 1); ++i) c.emplace_back(pts.at<int>(i, 0), pts.at<int>(i, 1));
        out.push_back(std::move(c));
    }
}
inline double contourArea(InputArray contour, bool oriented = false) {
    return cvcall("contourArea", {contour.getMat()}, {(double)oriented}, 1)[0].at<double>(0);
}
inline RotatedRect fitEllipseDirect(InputArray points) { return rrect_from(cvcall("fitEllipseDirect", {points.getMat()}, {}, 1)[0]); }
inline RotatedRect fitEllipse(InputArray points) { return rrect_from(cvcall("fitEllipse", {points.getMat()}, {}, 1)[0]); }
inline RotatedRect minAreaRect(InputArray points) { return rrect_from(cvcall("minAreaRect", {points.getMat()}, {}, 1)[0]); }
inline Rect boundingRect(InputArray array) {
    const Mat r = cvcall("boundingRect", {array.getMat()}, {}, 1)[0];
    return Rect(r.at<int>(0), r.at<int>(1), r.at<int>(2), r.at<int>(3));
}
inline Scalar mean(InputArray src) {
    const Mat r = cvcall("mean", {src.getMat()}, {}, 1)[0];
    return Scalar(r.at<double>(0), r.at<double>(1), r.at<double>(2), r.at<double>(3));
}
inline Mat getAffineTransform(const Point2f src[], const Point2f dst[]) {
    Mat a(3, 2, CV_32F), b(3, 2, CV_32F);
    for (int i = 0; i < 3; ++i) { a.at<float>(i, 0) = src[i].x; a.at<float>(i, 1) = src[i].y; b.at<float>(i, 0) = dst[i].x; b.at<float>(i, 1) = dst[i].y; }
    return cvcall("getAffineTransform", {a, b}, {}, 1)[0];
}
inline void warpAffine(InputArray src, OutputArray dst, InputArray M, Size dsize, int flags = INTER_LINEAR,
                       int borderMode = BORDER_CONSTANT, const Scalar& borderValue = Scalar()) {
    (void)borderValue;
    dst.assign(cvcall("warpAffine", {src.getMat(), M.getMat()}, {(double)dsize.width, (double)dsize.height, (double)flags, (double)borderMode}, 1)[0]);
}
inline void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    dst.assign(cvcall("resize", {src.getMat()}, {(double)dsize.width, (double)dsize.height, fx, fy, (double)interpolation}, 1)[0]);
}
inline void LUT(InputArray src, InputArray lut, OutputArray dst) { dst.assign(cvcall("LUT", {src.getMat(), lut.getMat()}, {}, 1)[0]); }
inline void cvtColor(InputArray src, OutputArray dst, int code) { dst.assign(cvcall("cvtColor", {src.getMat()}, {(double)code}, 1)[0]); }
inline Mat imread(const std::string&, int = IMREAD_COLOR) { throw std::runtime_error("cvstub: imread is file I/O, out of scope"); }
inline void setIdentity(Mat& m, const Scalar& s = Scalar(1)) {   // modules/core/src/matrix_operations.cpp: zero, then the diagonal
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) m.set_double(i, j, i == j ? s.val[0] : 0.0);
}
inline void Rodrigues(InputArray src, OutputArray dst) { dst.assign(cvcall("Rodrigues", {src.getMat()}, {}, 1)[0]); }
inline bool solvePnP(InputArray objectPoints, InputArray imagePoints, InputArray cameraMatrix, InputArray distCoeffs, OutputArray rvec,
                     OutputArray tvec, bool useExtrinsicGuess = false, int flags = SOLVEPNP_ITERATIVE) {
    const std::vector<Mat> r = cvcall("solvePnP", {objectPoints.getMat(), imagePoints.getMat(), cameraMatrix.getMat(), distCoeffs.getMat()},
                                      {(double)useExtrinsicGuess, (double)flags}, 2);
    rvec.assign(r[0]); tvec.assign(r[1]);
    return true;
}

// cv::KalmanFilter (modules/video/src/kalman.cpp): the state lives here, predict()/correct() run in the real OpenCV
class KalmanFilter {
public:
    Mat statePre, statePost, transitionMatrix, controlMatrix, measurementMatrix, processNoiseCov, measurementNoiseCov, errorCovPre,
        gain, errorCovPost;
    KalmanFilter() {}
    KalmanFilter(int dynamParams, int measureParams, int controlParams = 0, int type = CV_32F) { init(dynamParams, measureParams, controlParams, type); }
    void init(int DP, int MP, int CP = 0, int type = CV_32F) {
        statePre = Mat::zeros(DP, 1, type); statePost = Mat::zeros(DP, 1, type);
        transitionMatrix = Mat::eye(DP, DP, type);
        processNoiseCov = Mat::eye(DP, DP, type);
        measurementMatrix = Mat::zeros(MP, DP, type);
        measurementNoiseCov = Mat::eye(MP, MP, type);
        errorCovPre = Mat::zeros(DP, DP, type); errorCovPost = Mat::zeros(DP, DP, type);
        gain = Mat::zeros(DP, MP, type);
        if (CP > 0) controlMatrix = Mat::zeros(DP, CP, type);
    }
    const Mat& predict() {
        const std::vector<Mat> r = cvcall("kalmanPredict", state(), {}, 4);
        statePre = r[0]; statePost = r[1]; errorCovPre = r[2]; errorCovPost = r[3];
        return statePre;
    }
    const Mat& correct(const Mat& measurement) {
        std::vector<Mat> in = state();
        in.push_back(measurement);
        const std::vector<Mat> r = cvcall("kalmanCorrect", in, {}, 3);
        statePost = r[0]; errorCovPost = r[1]; gain = r[2];
        return statePost;
    }
private:
    std::vector<Mat> state() const {
        return {statePre, statePost, transitionMatrix, measurementMatrix, processNoiseCov, measurementNoiseCov, errorCovPre, errorCovPost, gain};
    }
};

}  // namespace cv
