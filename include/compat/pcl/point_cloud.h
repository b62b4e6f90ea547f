#pragma once
#include <pcl/point_types.h>
#include <vector>
namespace pcl {
template <typename P>
struct PointCloud {
    std::vector<P> points;
    unsigned width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    void push_back(const P &p) { points.push_back(p); }
    void resize(size_t n) { points.resize(n); }
    P &operator[](size_t i) { return points[i]; }
    const P &operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl
