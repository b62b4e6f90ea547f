/* POD stand-ins for the two PCL point types that appear in kfusion signatures (same sizes as PCL's). */
#pragma once
#include <boost/shared_ptr.hpp>
namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, pad_; PointXYZ() : x(0), y(0), z(0), pad_(1.f) {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad_(1.f) {} };
struct alignas(16) Normal { float normal_x, normal_y, normal_z, pad_; float curvature, pad2_[3]; };
struct alignas(16) PointNormal { float x, y, z, pad_; float normal_x, normal_y, normal_z, pad2_; float curvature, pad3_[3]; };
}  // namespace pcl
