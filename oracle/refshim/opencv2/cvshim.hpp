// cvshim.hpp -- TEST INFRASTRUCTURE, not product code.
//
// The smallest `cv` surface that lets the reference's icp.cpp, pointcloud.cpp
// and map.cpp compile UNMODIFIED, by path, from /root/reference (SURVEY.md 8c
// "Route B"); OpenCV's C++ headers are not installed in this image.  Written
// from the reference's call sites, not from OpenCV sources.  The arithmetic
// that OpenCV owns (3x3 products, inverse, determinant, SVD) is written the
// way tests/test_oracle_cv2.py pins it against cv2 4.13; viz is a no-op.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC3 16
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_32FC3 21
#define CV_64F 6
#define CV_64FC1 6

typedef unsigned char uchar; // OpenCV puts it in the global namespace (icp.cpp:101 uses it unqualified)

namespace cv {
using ::uchar;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> operator Point_<U>() const;
};
template <typename T> static inline int shim_round(T v) { return (int)std::lrint((double)v); }
template <typename T> template <typename U> Point_<T>::operator Point_<U>() const
{
    // OpenCV converts with saturate_cast: float -> int rounds to nearest (pointcloud.cpp:65)
    if (std::is_integral<U>::value && !std::is_integral<T>::value) return Point_<U>((U)shim_round(x), (U)shim_round(y));
    return Point_<U>((U)x, (U)y);
}
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b) { static_assert(N == 2, ""); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(N == 3, ""); val[0] = a; val[1] = b; val[2] = c; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    bool operator==(const Vec &o) const { for (int i = 0; i < N; ++i) if (val[i] != o.val[i]) return false; return true; }
    bool operator!=(const Vec &o) const { return !(*this == o); }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;
typedef Vec<int, 2> Vec2i;
template <typename T, int N> std::ostream &operator<<(std::ostream &os, const Vec<T, N> &v)
{
    os << "[";
    for (int i = 0; i < N; ++i) os << (i ? ", " : "") << +v.val[i];
    return os << "]";
}

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    operator Vec<T, 3>() const { return Vec<T, 3>(x, y, z); }
    Point3_ &operator+=(const Point3_ &o) { x += o.x; y += o.y; z += o.z; return *this; }
    Point3_ &operator-=(const Point3_ &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    bool operator==(const Point3_ &o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const Point3_ &o) const { return !(*this == o); }
};
template <typename T> Point3_<T> operator+(const Point3_<T> &a, const Point3_<T> &b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T> &a, const Point3_<T> &b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T> &a) { return Point3_<T>(-a.x, -a.y, -a.z); }
template <typename T> std::ostream &operator<<(std::ostream &os, const Point3_<T> &p) { return os << "[" << p.x << ", " << p.y << ", " << p.z << "]"; }
typedef Point3_<float> Point3f;
typedef Point3_<int> Point3i;

struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} int area() const { return width * height; } };
typedef Size Size2i;
struct Rect { int x, y, width, height; Rect(int x_ = 0, int y_ = 0, int w = 0, int h = 0) : x(x_), y(y_), width(w), height(h) {} };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {} };

static inline int shim_elem_size(int type)
{
    switch (type) { case CV_8UC3: return 3; case CV_16UC1: return 2; case CV_32FC1: return 4; case CV_32FC3: return 12; case CV_64F: return 8; }
    return 4;
}

// Dense 2-D matrix: shared storage + (data pointer, step) view, enough for ROI / col views.
struct Mat {
    int rows, cols, type_;
    size_t step;            // bytes per row
    unsigned char *data;
    std::shared_ptr<std::vector<unsigned char>> owner;
    Mat() : rows(0), cols(0), type_(CV_32FC1), step(0), data(nullptr) {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void *ext) : rows(r), cols(c), type_(t), step((size_t)c * shim_elem_size(t)), data((unsigned char *)ext) {}
    explicit Mat(const Point3f &p) { create(3, 1, CV_32FC1); at<float>(0, 0) = p.x; at<float>(1, 0) = p.y; at<float>(2, 0) = p.z; }
    void create(int r, int c, int t)
    {
        rows = r; cols = c; type_ = t; step = (size_t)c * shim_elem_size(t);
        owner = std::make_shared<std::vector<unsigned char>>((size_t)r * step + 16);
        data = owner->data();
    }
    int type() const { return type_; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    template <typename T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T &at(Point2i p) { return at<T>(p.y, p.x); }
    Mat operator()(const Rect &r) const
    {
        Mat v = *this;
        v.rows = r.height; v.cols = r.width;
        v.data = data + (size_t)r.y * step + (size_t)r.x * shim_elem_size(type_);
        return v;
    }
    Mat col(int c) const { return (*this)(Rect(c, 0, 1, rows)); }
    Mat clone() const { Mat m(rows, cols, type_); copyTo(m); return m; }
    void copyTo(Mat dst) const
    {
        if (dst.empty() || dst.rows != rows || dst.cols != cols) return; // only the in-place ROI form is used (icp.cpp:228,232)
        const size_t rb = (size_t)cols * shim_elem_size(type_);
        for (int r = 0; r < rows; ++r) std::memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, rb);
    }
    Mat t() const
    {
        Mat m(cols, rows, type_);
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m.at<float>(c, r) = at<float>(r, c);
        return m;
    }
    Mat &operator*=(double s) { for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) at<float>(r, c) = (float)(at<float>(r, c) * s); return *this; }
    Mat &operator*=(const Mat &o);
    Mat inv() const;
};

// CV_32F product.  Inner dimension <= 4 (every 3x3 * 3x3 / 3x3 * 3xN product in the reference):
// float, no FMA, left to right -- pinned against cv2.gemm.  Larger inner dimension (the
// N-pair cross-covariance, icp.cpp:212): double accumulation in index order, rounded once.
static inline Mat operator*(const Mat &a, const Mat &b)
{
    Mat d(a.rows, b.cols, CV_32FC1);
    const int len = a.cols;
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            if (len <= 4) {
                float s = a.at<float>(i, 0) * b.at<float>(0, j);
                for (int k = 1; k < len; ++k) s = s + a.at<float>(i, k) * b.at<float>(k, j);
                d.at<float>(i, j) = s;
            } else {
                double s = 0.0;
                for (int k = 0; k < len; ++k) s += (double)a.at<float>(i, k) * (double)b.at<float>(k, j);
                d.at<float>(i, j) = (float)s;
            }
        }
    return d;
}
inline Mat &Mat::operator*=(const Mat &o) { Mat r = (*this) * o; r.copyTo(*this); return *this; }

static inline double determinant(const Mat &m)
{
    double m00 = m.at<float>(0, 0), m01 = m.at<float>(0, 1), m02 = m.at<float>(0, 2);
    double m10 = m.at<float>(1, 0), m11 = m.at<float>(1, 1), m12 = m.at<float>(1, 2);
    double m20 = m.at<float>(2, 0), m21 = m.at<float>(2, 1), m22 = m.at<float>(2, 2);
    return m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
}
inline Mat Mat::inv() const
{
    Mat D(3, 3, CV_32FC1);
    double d = determinant(*this);
    if (d == 0.0) { for (int i = 0; i < 9; ++i) D.at<float>(i / 3, i % 3) = 0.f; return D; }
    d = 1.0 / d;
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = at<float>(i / 3, i % 3);
    const double T[9] = {(S[4] * S[8] - S[5] * S[7]) * d, (S[2] * S[7] - S[1] * S[8]) * d, (S[1] * S[5] - S[2] * S[4]) * d,
                         (S[5] * S[6] - S[3] * S[8]) * d, (S[0] * S[8] - S[2] * S[6]) * d, (S[2] * S[3] - S[0] * S[5]) * d,
                         (S[3] * S[7] - S[4] * S[6]) * d, (S[1] * S[6] - S[0] * S[7]) * d, (S[0] * S[4] - S[1] * S[3]) * d};
    for (int i = 0; i < 9; ++i) D.at<float>(i / 3, i % 3) = (float)T[i];
    return D;
}

// cv::SVD of a 3x3 CV_32F: one-sided Jacobi in double, outputs rounded to float, w descending.
struct SVD {
    Mat u, w, vt;
    explicit SVD(const Mat &src)
    {
        double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, sv[3];
        for (int i = 0; i < 9; ++i) A[i] = src.at<float>(i / 3, i % 3);
        const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
        for (int sweep = 0; sweep < 30; ++sweep) {
            bool changed = false;
            for (int k = 0; k < 3; ++k) {
                const int p = P[k], q = Q[k];
                double al = 0, be = 0, ga = 0;
                for (int r = 0; r < 3; ++r) { al += A[3 * r + p] * A[3 * r + p]; be += A[3 * r + q] * A[3 * r + q]; ga += A[3 * r + p] * A[3 * r + q]; }
                if (std::fabs(ga) <= 2.220446049250313e-16 * std::sqrt(al * be)) continue;
                changed = true;
                double zeta = (be - al) / (2.0 * ga);
                double t = 1.0 / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                if (zeta < 0) t = -t;
                double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < 3; ++r) {
                    double a = A[3 * r + p], b = A[3 * r + q];
                    A[3 * r + p] = c * a - s * b; A[3 * r + q] = s * a + c * b;
                    double va = V[3 * r + p], vb = V[3 * r + q];
                    V[3 * r + p] = c * va - s * vb; V[3 * r + q] = s * va + c * vb;
                }
            }
            if (!changed) break;
        }
        for (int k = 0; k < 3; ++k) sv[k] = std::sqrt(A[k] * A[k] + A[3 + k] * A[3 + k] + A[6 + k] * A[6 + k]);
        int order[3] = {0, 1, 2};
        for (int i = 0; i < 2; ++i) for (int j = i + 1; j < 3; ++j) if (sv[order[j]] > sv[order[i]]) std::swap(order[i], order[j]);
        u.create(3, 3, CV_32FC1); w.create(3, 1, CV_32FC1); vt.create(3, 3, CV_32FC1);
        for (int k = 0; k < 3; ++k) {
            const int o = order[k];
            w.at<float>(k, 0) = (float)sv[o];
            for (int r = 0; r < 3; ++r) {
                u.at<float>(r, k) = (float)(sv[o] > 0 ? A[3 * r + o] / sv[o] : 0.0);
                vt.at<float>(k, r) = (float)V[3 * r + o];
            }
        }
    }
};

namespace viz {
enum { POINT_SIZE = 0 };
struct Color {
    Color() {}
    Color(const Scalar &) {}
    static Color red() { return Color(); }
    static Color green() { return Color(); }
    static Color yellow() { return Color(); }
};
struct Widget { void setRenderingProperty(int, double) {} };
struct WCloud : Widget { WCloud(const Mat &, const Mat &) {} WCloud(const Mat &, const Color &) {} };
struct WLine : Widget { WLine(const Point3f &, const Point3f &, const Color &) {} };
struct WCube : Widget { WCube(const Vec3d &, const Vec3d &, bool, const Color &) {} };
struct WSphere : Widget { WSphere(const Vec3d &, double, int, const Color &) {} };
struct Viz3d {
    Viz3d() {}
    explicit Viz3d(const std::string &) {}
    void removeAllWidgets() {}
    void removeWidget(const std::string &) {}
    void showWidget(const std::string &, const Widget &) {}
    void spinOnce(int = 1, bool = false) {}
};
} // namespace viz
} // namespace cv
