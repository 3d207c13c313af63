// cv_min.hpp -- the few cv:: value types the drop-in headers traffic in, for builds where OpenCV's C++
// headers are not installed (this image).  With real OpenCV available, define ICPB200_USE_OPENCV and the
// genuine <opencv2/...> types are used instead; the wrappers only rely on the members declared here.
#pragma once
#ifdef ICPB200_USE_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/viz.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#ifndef CV_8UC3
#define CV_8UC3 16
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_32FC3 21
#endif

typedef unsigned char uchar;

namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point3f {
    float x = 0, y = 0, z = 0;
    Point3f() {}
    Point3f(float a, float b, float c) : x(a), y(b), z(c) {}
    Point3f &operator+=(const Point3f &o) { x += o.x; y += o.y; z += o.z; return *this; }
    Point3f &operator-=(const Point3f &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    bool operator==(const Point3f &o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const Point3f &o) const { return !(*this == o); }
};
inline Point3f operator+(Point3f a, const Point3f &b) { return a += b; }
inline Point3f operator-(Point3f a, const Point3f &b) { return a -= b; }
inline Point3f operator-(const Point3f &a) { return Point3f(-a.x, -a.y, -a.z); }
inline std::ostream &operator<<(std::ostream &os, const Point3f &p) { return os << "[" << p.x << ", " << p.y << ", " << p.z << "]"; }
struct Point3i { int x = 0, y = 0, z = 0; Point3i() {} Point3i(int a, int b, int c) : x(a), y(b), z(c) {} };
struct Vec3b {
    uchar val[3] = {0, 0, 0};
    Vec3b() {}
    Vec3b(uchar a, uchar b, uchar c) { val[0] = a; val[1] = b; val[2] = c; }
    uchar &operator[](int i) { return val[i]; }
    const uchar &operator[](int i) const { return val[i]; }
    bool operator==(const Vec3b &o) const { return val[0] == o.val[0] && val[1] == o.val[1] && val[2] == o.val[2]; }
    bool operator!=(const Vec3b &o) const { return !(*this == o); }
};
inline std::ostream &operator<<(std::ostream &os, const Vec3b &v) { return os << "[" << +v[0] << ", " << +v[1] << ", " << +v[2] << "]"; }
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };

// Row-major dense matrix, continuous storage (what the wrappers need from cv::Mat).
struct Mat {
    int rows = 0, cols = 0;
    int type_ = CV_32FC1;
    std::shared_ptr<std::vector<uchar>> store;
    uchar *data = nullptr;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void *ext) : rows(r), cols(c), type_(t), data((uchar *)ext) {}
    static int elem(int t) { return t == CV_8UC3 ? 3 : t == CV_16UC1 ? 2 : t == CV_32FC3 ? 12 : 4; }
    void create(int r, int c, int t)
    {
        rows = r; cols = c; type_ = t;
        store = std::make_shared<std::vector<uchar>>((size_t)r * c * elem(t), 0);
        data = store->data();
    }
    int type() const { return type_; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    template <typename T> T &at(int r, int c) { return reinterpret_cast<T *>(data)[(size_t)r * cols + c]; }
    template <typename T> const T &at(int r, int c) const { return reinterpret_cast<const T *>(data)[(size_t)r * cols + c]; }
    Mat clone() const { Mat m(rows, cols, type_); std::memcpy(m.data, data, (size_t)rows * cols * elem(type_)); return m; }
};

namespace viz {
struct Color { static Color red() { return Color(); } static Color green() { return Color(); } };
struct Viz3d {
    Viz3d() {}
    explicit Viz3d(const std::string &) {}
    void removeAllWidgets() {}
    void spinOnce(int = 1, bool = false) {}
};
} // namespace viz
} // namespace cv
#endif
