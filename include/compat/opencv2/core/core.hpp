/*
 * Minimal, dependency-free stand-in for the handful of OpenCV value types that the sobfu / kfusion public
 * headers use (cv::Vec3i/Vec3f, cv::Matx33f, cv::Ptr, a small dense cv::Mat, tick counters).
 * OpenCV is not a dependency of sobfu_b200; callers that do have OpenCV simply put it first on the
 * include path.  Only what the drop-in boundary needs is provided.
 */
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <vector>

#define CV_32FC1 5
#define CV_32FC2 13
#define CV_32FC3 21
#define CV_32FC4 29
#define CV_32SC1 4

namespace cv {

template <typename T, int n>
struct Vec {
    T val[n];
    Vec() { for (int i = 0; i < n; ++i) val[i] = T(0); }
    Vec(T a, T b) { static_assert(n >= 2, ""); val[0] = a; val[1] = b; for (int i = 2; i < n; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(n >= 3, ""); val[0] = a; val[1] = b; val[2] = c; for (int i = 3; i < n; ++i) val[i] = T(0); }
    Vec(T a, T b, T c, T d) { static_assert(n >= 4, ""); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    template <typename U>
    Vec(const Vec<U, n> &o) { for (int i = 0; i < n; ++i) val[i] = static_cast<T>(o.val[i]); }
    static Vec all(T v) { Vec r; for (int i = 0; i < n; ++i) r.val[i] = v; return r; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    Vec operator+(const Vec &o) const { Vec r; for (int i = 0; i < n; ++i) r.val[i] = val[i] + o.val[i]; return r; }
    Vec operator-(const Vec &o) const { Vec r; for (int i = 0; i < n; ++i) r.val[i] = val[i] - o.val[i]; return r; }
    Vec operator*(T s) const { Vec r; for (int i = 0; i < n; ++i) r.val[i] = val[i] * s; return r; }
};
typedef Vec<int, 3> Vec3i;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<double, 3> Vec3d;

template <typename T, int m, int n>
struct Matx {
    T val[m * n];
    Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
    static Matx eye() { Matx r; for (int i = 0; i < (m < n ? m : n); ++i) r.val[i * n + i] = T(1); return r; }
    T &operator()(int i, int j) { return val[i * n + j]; }
    const T &operator()(int i, int j) const { return val[i * n + j]; }
    Matx<T, n, m> t() const { Matx<T, n, m> r; for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) r(j, i) = (*this)(i, j); return r; }
    template <int k>
    Matx<T, m, k> operator*(const Matx<T, n, k> &o) const {
        Matx<T, m, k> r;
        for (int i = 0; i < m; ++i) for (int j = 0; j < k; ++j) { T s = 0; for (int q = 0; q < n; ++q) s += (*this)(i, q) * o(q, j); r(i, j) = s; }
        return r;
    }
    Vec<T, m> operator*(const Vec<T, n> &v) const {
        Vec<T, m> r;
        for (int i = 0; i < m; ++i) { T s = 0; for (int q = 0; q < n; ++q) s += (*this)(i, q) * v[q]; r[i] = s; }
        return r;
    }
};
typedef Matx<float, 3, 3> Matx33f;

/* cv::Ptr: shared ownership, implicitly constructible from a raw pointer as in OpenCV */
template <typename T>
struct Ptr {
    std::shared_ptr<T> p;
    Ptr() {}
    Ptr(T *raw) : p(raw) {}
    Ptr(const std::shared_ptr<T> &s) : p(s) {}
    T *operator->() const { return p.get(); }
    T &operator*() const { return *p; }
    T *get() const { return p.get(); }
    bool empty() const { return !p; }
    operator bool() const { return (bool)p; }
    void release() { p.reset(); }
};
template <typename T, typename... A>
Ptr<T> makePtr(A &&... a) { return Ptr<T>(new T(std::forward<A>(a)...)); }

enum { DECOMP_LU = 0, DECOMP_SVD = 1 };

/* element types as OpenCV encodes them: depth + ((channels - 1) << 3) */
#ifndef CV_8U
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#endif

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size &o) const { return width == o.width && height == o.height; }
};
struct MatStep {
    size_t v;
    MatStep(size_t s = 0) : v(s) {}
    operator size_t() const { return v; }
};

/* dense matrix with OpenCV's public members (rows, cols, data, step): n-d float fields for the download / print helpers
 * (at<T>(k,j,i), ptr<T>()) and 8/16-bit images for the application (imread, copyTo with a mask) */
class Mat {
public:
    int rows, cols, dims;
    unsigned char *data;
    MatStep step;                                   /* bytes per row (2-d) */
    Mat() : rows(0), cols(0), dims(0), data(nullptr), type_(CV_32FC1) {}
    Mat(int ndims, const int *sizes, int type) : type_(type) { sz_.assign(sizes, sizes + ndims); alloc(); }
    Mat(int r, int c, int type) : type_(type) { sz_ = {r, c}; alloc(); }
    Mat(Size s, int type) : type_(type) { sz_ = {s.height, s.width}; alloc(); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size s, int type) { return Mat(s, type); }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < (r < c ? r : c); ++i) m.at<float>(i, i) = 1.f; return m; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { static const size_t b[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return b[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || total() == 0; }
    template <typename T> T *ptr(int y = 0) { return reinterpret_cast<T *>(data + (size_t)y * step.v); }
    template <typename T> const T *ptr(int y = 0) const { return reinterpret_cast<const T *>(data + (size_t)y * step.v); }
    template <typename T> T &at(int i0) { return reinterpret_cast<T *>(data)[i0]; }
    template <typename T> T &at(int i0, int i1) { return reinterpret_cast<T *>(data)[(size_t)i0 * sz_[1] + i1]; }
    template <typename T> T &at(int i0, int i1, int i2) { return reinterpret_cast<T *>(data)[((size_t)i0 * sz_[1] + i1) * sz_[2] + i2]; }
    size_t total() const { size_t t = 1; for (int s : sz_) t *= s; return sz_.empty() ? 0 : t; }
    /* dst(x) = src(x) where mask(x) != 0 (8-bit single-channel mask of the same size); dst is (re)allocated when its size or
     * type differs, as cv::Mat::copyTo does -- then the unmasked elements are zero */
    void copyTo(Mat &dst, const Mat &mask) const {
        if (dst.rows != rows || dst.cols != cols || dst.type_ != type_ || !dst.data) dst = Mat(rows, cols, type_);
        const size_t es = elemSize();
        for (int y = 0; y < rows; ++y) {
            const unsigned char *m = mask.data + (size_t)y * mask.step.v;
            for (int x = 0; x < cols; ++x)
                if (m[(size_t)x * mask.elemSize()]) std::memcpy(dst.data + (size_t)y * dst.step.v + x * es, data + (size_t)y * step.v + x * es, es);
        }
    }
    void copyTo(Mat &dst) const {
        dst = Mat(dims, sz_.data(), type_);
        if (data) std::memcpy(dst.data, data, total() * elemSize());
    }
    Mat clone() const { Mat m; copyTo(m); return m; }
    std::vector<int> sz_;
private:
    void alloc() {
        buf_ = std::make_shared<std::vector<unsigned char>>(total() * elemSize(), 0);
        data = buf_->empty() ? nullptr : buf_->data();
        dims = (int)sz_.size();
        rows = dims == 2 ? sz_[0] : (dims == 1 ? sz_[0] : -1);
        cols = dims == 2 ? sz_[1] : (dims == 1 ? 1 : -1);
        step = MatStep(dims >= 1 ? (size_t)sz_.back() * elemSize() : 0);
    }
    int type_;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
inline Mat operator*(float s, const Mat &m) { Mat o(m.rows, m.cols, CV_32FC1); for (size_t i = 0; i < m.total(); ++i) o.ptr<float>()[i] = s * m.ptr<float>()[i]; return o; }
inline Mat operator-(const Mat &a, const Mat &b) { Mat o(a.rows, a.cols, CV_32FC1); for (size_t i = 0; i < a.total(); ++i) o.ptr<float>()[i] = a.ptr<float>()[i] - b.ptr<float>()[i]; return o; }
inline std::ostream &operator<<(std::ostream &os, const Mat &m) { for (size_t i = 0; i < m.total(); ++i) os << m.ptr<float>()[i] << (i + 1 < m.total() ? ", " : ""); return os; }
/* dense solve is only referenced from dead code in the reference (solver.cpp:107-158) */
inline bool solve(const Mat &, const Mat &, Mat &, int) { std::fprintf(stderr, "cv::solve: not provided by the compat shim\n"); return false; }

inline int64_t getTickCount() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline double getTickFrequency() { return 1e9; }

}  // namespace cv
