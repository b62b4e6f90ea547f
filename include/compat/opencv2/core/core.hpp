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

/* dense n-d float matrix, enough for field download / print helpers (at<T>(k,j,i), ptr<T>()) */
class Mat {
public:
    Mat() : type_(CV_32FC1) {}
    Mat(int ndims, const int *sizes, int type) : type_(type) { sz_.assign(sizes, sizes + ndims); alloc(); }
    Mat(int rows, int cols, int type) : type_(type) { sz_ = {rows, cols}; alloc(); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < (r < c ? r : c); ++i) m.at<float>(i, i) = 1.f; return m; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return 4u * channels(); }
    template <typename T> T *ptr() { return reinterpret_cast<T *>(buf_->data()); }
    template <typename T> const T *ptr() const { return reinterpret_cast<const T *>(buf_->data()); }
    template <typename T> T &at(int i0) { return ptr<T>()[i0]; }
    template <typename T> T &at(int i0, int i1) { return ptr<T>()[(size_t)i0 * sz_[1] + i1]; }
    template <typename T> T &at(int i0, int i1, int i2) { return ptr<T>()[((size_t)i0 * sz_[1] + i1) * sz_[2] + i2]; }
    int rows() const { return sz_.empty() ? 0 : sz_[0]; }
    int cols() const { return sz_.size() < 2 ? 0 : sz_[1]; }
    size_t total() const { size_t t = 1; for (int s : sz_) t *= s; return sz_.empty() ? 0 : t; }
    std::vector<int> sz_;
private:
    void alloc() { buf_ = std::make_shared<std::vector<unsigned char>>(total() * elemSize(), 0); }
    int type_;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
inline Mat operator*(float s, const Mat &m) { Mat r = m; Mat o(m.rows(), m.cols(), CV_32FC1); for (size_t i = 0; i < m.total(); ++i) o.ptr<float>()[i] = s * m.ptr<float>()[i]; return o; }
inline Mat operator-(const Mat &a, const Mat &b) { Mat o(a.rows(), a.cols(), CV_32FC1); for (size_t i = 0; i < a.total(); ++i) o.ptr<float>()[i] = a.ptr<float>()[i] - b.ptr<float>()[i]; return o; }
inline std::ostream &operator<<(std::ostream &os, const Mat &m) { for (size_t i = 0; i < m.total(); ++i) os << m.ptr<float>()[i] << (i + 1 < m.total() ? ", " : ""); return os; }
/* dense solve is only referenced from dead code in the reference (solver.cpp:107-158) */
inline bool solve(const Mat &, const Mat &, Mat &, int) { std::fprintf(stderr, "cv::solve: not provided by the compat shim\n"); return false; }

inline int64_t getTickCount() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline double getTickFrequency() { return 1e9; }

}  // namespace cv
