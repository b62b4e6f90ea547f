/* Minimal cv::Affine3f (rotation + translation) for the sobfu / kfusion public headers. */
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
template <typename T>
struct Affine3 {
    typedef Matx<T, 3, 3> Mat3;
    typedef Vec<T, 3> Vec3;
    Mat3 R;
    Vec3 t;
    Affine3() : R(Mat3::eye()) {}
    Affine3(const Mat3 &R_, const Vec3 &t_ = Vec3()) : R(R_), t(t_) {}
    static Affine3 Identity() { return Affine3(); }
    Mat3 rotation() const { return R; }
    Vec3 translation() const { return t; }
    void rotation(const Mat3 &r) { R = r; }
    void translation(const Vec3 &v) { t = v; }
    Affine3 translate(const Vec3 &v) const { Affine3 r = *this; r.t = r.t + v; return r; }
    Affine3 inv() const { Affine3 r; r.R = R.t(); r.t = (r.R * t) * T(-1); return r; }
    Affine3 operator*(const Affine3 &o) const { Affine3 r; r.R = R * o.R; r.t = R * o.t + t; return r; }
    Vec3 operator*(const Vec3 &v) const { return R * v + t; }
};
typedef Affine3<float> Affine3f;
}  // namespace cv
