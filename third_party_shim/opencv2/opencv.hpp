// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the OpenCV headers the reference includes.
//
// Purpose: let the UNMODIFIED reference sources compile in place (oracle/Makefile):
//   src/toed/cpu_toed.cpp                          -> oracle/_ref/libtoed_ref.so
//   src/Stereo_Matches.cpp, src/utility.cpp, src/EdgeClusterer.cpp -> oracle/_ref/libstereo_ref.so
// Only what those files execute on the no-GT, SIFT-off path is implemented for real (cv::Mat storage, at<>,
// convertTo, clone/row, mean/sum/dot/norm, the MatExpr forms used by Utility::get_patch_similarity, cv::Sobel
// 3x3); everything else they merely mention (SIFT, KeyPoint, cvtColor, buildPyramid, ...) is a declaration-level
// stub that aborts if it is ever called.  The arithmetic of the implemented primitives follows OpenCV 4.x
// (type mix documented inline); cv2 4.13 is used in tests/test_oracle_stereo.py to check the same primitives.
// Nothing here is product code.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH_MASK 7
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

[[noreturn]] inline void shim_unsupported(const char* what)
{
    std::fprintf(stderr, "ref_shim: %s is not implemented (not on the no-GT / SIFT-off stereo path)\n", what);
    std::abort();
}

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
};
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, double s) { return Point_<T>((T)(a.x * s), (T)(a.y * s)); }
template <typename T> inline Point_<T> operator*(double s, const Point_<T>& a) { return Point_<T>((T)(a.x * s), (T)(a.y * s)); }
typedef Point_<double> Point2d;
typedef Point_<float> Point2f;
struct Point3d { double x, y, z; Point3d() : x(0), y(0), z(0) {} Point3d(double a, double b, double c) : x(a), y(b), z(c) {} };
// cv::norm(Point_<T>) = sqrt((double)x*x + (double)y*y)   (core/types.hpp)
template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double operator[](int i) const { return val[i]; }
};
struct Vec3b { uchar v[3]; uchar& operator[](int i) { return v[i]; } uchar operator[](int i) const { return v[i]; } };
template <typename T> using Ptr = std::shared_ptr<T>;
enum { NORM_L2 = 4, COLOR_HSV2BGR = 54 };

class Mat;
// The lazy expression forms used by the reference: (Mat - scalar), .mul(), (expr / scalar).  value = a*alpha + beta
struct MatExpr {
    std::shared_ptr<Mat> a;
    double alpha, beta;
    inline Mat eval() const;
    inline Mat mul(const MatExpr& o) const;
    inline operator Mat() const;
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); fill(s[0]); }
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : rows(r), cols(c), data((uchar*)ext), type_(type) { step = stp ? stp : (size_t)c * elemSize(); }
    int type() const { return type_; }
    int depth() const { return type_ & CV_MAT_DEPTH_MASK; }
    int channels() const { return 1 + (type_ >> CV_CN_SHIFT); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t elemSize() const { static const int sz[] = {1, 1, 2, 2, 4, 4, 8}; return (size_t)sz[depth()] * channels(); }
    size_t total() const { return (size_t)rows * cols; }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[step * (size_t)(r > 0 ? r : 1) + 64](), std::default_delete<uchar[]>());
        data = buf_.get();
    }
    template <typename T> T& at(int i, int j) { return *reinterpret_cast<T*>(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    template <typename T> const T& at(int i, int j) const { return *reinterpret_cast<const T*>(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int i = 0; i < rows; ++i) std::memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elemSize());
        return m;
    }
    Mat row(int i) const { Mat m; m.rows = 1; m.cols = cols; m.type_ = type_; m.step = step; m.data = data + (size_t)i * step; m.buf_ = buf_; return m; }
    double get(int i, int j) const
    {
        switch (depth()) {
        case CV_8U: return at<uchar>(i, j);
        case CV_32F: return at<float>(i, j);
        case CV_64F: return at<double>(i, j);
        case CV_32S: return at<int>(i, j);
        default: shim_unsupported("Mat depth");
        }
    }
    void fill(double v) { for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) set(i, j, v); }
    void set(int i, int j, double v)
    {
        switch (depth()) {
        case CV_8U: at<uchar>(i, j) = (uchar)v; break;
        case CV_32F: at<float>(i, j) = (float)v; break;
        case CV_64F: at<double>(i, j) = v; break;
        case CV_32S: at<int>(i, j) = (int)v; break;
        default: shim_unsupported("Mat depth");
        }
    }
    // convertTo with unit scale (the only form the reference uses): saturate_cast between 8U/32F/64F
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const
    {
        if (channels() != 1) shim_unsupported("multi-channel convertTo");
        Mat out(rows, cols, CV_MAKETYPE(rtype & CV_MAT_DEPTH_MASK, 1));
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) {
                double v = get(i, j);
                if (alpha != 1.0 || beta != 0.0) v = v * alpha + beta;
                if (out.depth() == CV_8U) { v = std::nearbyint(v); v = v < 0 ? 0 : (v > 255 ? 255 : v); }
                out.set(i, j, v);
            }
        dst = out;
    }
    // Mat::dot for CV_32F / CV_64F.  OpenCV accumulates 32F products in float SIMD lanes whose width depends on the
    // build; a double accumulator is used here (differences ~1e-7, the reason NCC parity is quoted to 1e-5).
    double dot(const Mat& o) const
    {
        double r = 0;
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) r += depth() == CV_32F ? (double)(at<float>(i, j) * o.at<float>(i, j)) : get(i, j) * o.get(i, j);
        return r;
    }
    MatExpr mul_expr() const;
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < r && i < c; ++i) m.set(i, i, 1.0); return m; }
    int type_ = 0;
    std::shared_ptr<uchar> buf_;
};

// Mat - scalar: cv::subtract(Mat, Scalar).  For CV_32F sources OpenCV narrows the scalar to float and works in float.
inline MatExpr operator-(const Mat& a, double s) { return MatExpr{std::make_shared<Mat>(a), 1.0, -s}; }
inline MatExpr operator/(const MatExpr& e, double s) { return MatExpr{e.a, e.alpha * (1.0 / s), e.beta * (1.0 / s)}; }   // MatOp_AddEx::multiply(e, 1./s)
inline Mat MatExpr::eval() const
{
    Mat out(a->rows, a->cols, a->type());
    for (int i = 0; i < a->rows; ++i)
        for (int j = 0; j < a->cols; ++j) {
            if (a->depth() == CV_32F) {
                // alpha == 1: cv::add(a, Scalar) in float; otherwise convertTo(alpha, beta): cvtScale 32f->32f, float a/b
                float v = a->at<float>(i, j);
                out.at<float>(i, j) = (alpha == 1.0) ? v + (float)beta : v * (float)alpha + (float)beta;
            } else out.set(i, j, a->get(i, j) * alpha + beta);
        }
    return out;
}
inline MatExpr::operator Mat() const { return eval(); }
inline Mat MatExpr::mul(const MatExpr& o) const
{
    Mat x = eval(), y = o.eval(), out(x.rows, x.cols, x.type());
    for (int i = 0; i < x.rows; ++i)
        for (int j = 0; j < x.cols; ++j) {
            if (x.depth() == CV_32F) out.at<float>(i, j) = x.at<float>(i, j) * y.at<float>(i, j);
            else out.set(i, j, x.get(i, j) * y.get(i, j));
        }
    return out;
}

// cv::sum / cv::mean: double accumulation for every depth
inline Scalar sum(const Mat& m) { double s = 0; for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) s += m.get(i, j); return Scalar(s); }
inline Scalar mean(const Mat& m) { return Scalar(m.total() ? sum(m)[0] / (double)m.total() : 0.0); }
inline double norm(const Mat& a, const Mat& b, int)
{
    double s = 0;
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) { double d = a.get(i, j) - b.get(i, j); s += d * d; }
    return std::sqrt(s);
}

template <typename T> struct MatCommaInit_;
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64FC1 : (sizeof(T) == 4 ? CV_32FC1 : CV_8UC1)) {}
    MatCommaInit_<T> operator<<(T v);
};
template <typename T> struct MatCommaInit_ {
    Mat_<T> m; int k;
    MatCommaInit_& operator,(T v) { m.template at<T>(k / m.cols, k % m.cols) = v; ++k; return *this; }
    operator Mat() const { return m; }
};
template <typename T> inline MatCommaInit_<T> Mat_<T>::operator<<(T v) { MatCommaInit_<T> c{*this, 0}; c, v; return c; }

// cv::Sobel(src CV_32F, ddepth CV_32F, dx, dy, ksize = 3, scale, delta = 0, BORDER_REFLECT_101): separable
// [-1 0 1] x [1 2 1] with the scale folded into the smoothing kernel (exact in FP32 on 8-bit data)
inline void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1.0, double delta = 0.0, int = 4)
{
    if (src.depth() != CV_32F || ddepth != CV_32F || ksize != 3 || delta != 0.0 || dx + dy != 1) shim_unsupported("this Sobel configuration");
    const int H = src.rows, W = src.cols;
    Mat out(H, W, CV_32FC1);
    auto R = [](int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); };
    const float s1 = (float)scale, s2 = (float)(2.0 * scale);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int xm = R(x - 1, W), xp = R(x + 1, W), ym = R(y - 1, H), yp = R(y + 1, H);
            float v;
            if (dx == 1) v = (src.at<float>(ym, xp) - src.at<float>(ym, xm)) * s1 + (src.at<float>(y, xp) - src.at<float>(y, xm)) * s2 + (src.at<float>(yp, xp) - src.at<float>(yp, xm)) * s1;
            else v = (src.at<float>(yp, xm) - src.at<float>(ym, xm)) * s1 + (src.at<float>(yp, x) - src.at<float>(ym, x)) * s2 + (src.at<float>(yp, xp) - src.at<float>(ym, xp)) * s1;
            out.at<float>(y, x) = v;
        }
    dst = out;
}

template <typename T> inline std::ostream& operator<<(std::ostream& os, const Point_<T>& p) { return os << "[" << p.x << ", " << p.y << "]"; }
inline std::ostream& operator<<(std::ostream& os, const Mat& m)
{
    os << "[";
    for (int i = 0; i < m.rows; ++i) { for (int j = 0; j < m.cols; ++j) os << m.get(i, j) << (j + 1 < m.cols ? ", " : ""); os << (i + 1 < m.rows ? ";\n " : ""); }
    return os << "]";
}

// ---- declaration-level stubs (never executed on the pinned path) ----
struct KeyPoint {
    Point2f pt; float size, angle;
    KeyPoint() : size(0), angle(-1) {}
    KeyPoint(Point2f p, float s, float a = -1) : pt(p), size(s), angle(a) {}
};
struct DMatch { int queryIdx, trainIdx; float distance; };
class SIFT {
public:
    static Ptr<SIFT> create(int = 0, int = 3, double = 0.04, double = 10, double = 1.6) { return std::make_shared<SIFT>(); }
    void compute(const Mat&, std::vector<KeyPoint>&, Mat&) { shim_unsupported("cv::SIFT::compute"); }
};
inline void cvtColor(const Mat&, Mat&, int) { shim_unsupported("cv::cvtColor"); }
inline void buildPyramid(const Mat&, std::vector<Mat>&, int) { shim_unsupported("cv::buildPyramid"); }
template <typename E> inline void eigen2cv(const E&, Mat&) { shim_unsupported("cv::eigen2cv"); }
inline void undistort(const Mat&, Mat&, const Mat&, const Mat&) { shim_unsupported("cv::undistort"); }
inline bool imwrite(const std::string&, const Mat&) { shim_unsupported("cv::imwrite"); }

}  // namespace cv
