#pragma once
#include <pcl/PCLPointCloud2.h>
#include <pcl/point_cloud.h>
#include <cstring>
namespace pcl {
template <typename P>
void toPCLPointCloud2(const PointCloud<P> &c, PCLPointCloud2 &m) {
    m.width = (unsigned)c.size(); m.height = 1; m.point_step = sizeof(P); m.row_step = m.point_step * m.width;
    m.data.resize((size_t)m.row_step);
    if (!c.points.empty()) std::memcpy(m.data.data(), c.points.data(), m.data.size());
}
template <typename P>
void fromPCLPointCloud2(const PCLPointCloud2 &m, PointCloud<P> &c) {
    const size_t n = m.point_step ? m.data.size() / m.point_step : 0;
    c.points.resize(n);
    c.width = m.width; c.height = m.height;
    for (size_t i = 0; i < n; ++i) std::memcpy(&c.points[i], m.data.data() + i * m.point_step, sizeof(P) < m.point_step ? sizeof(P) : m.point_step);
}
}  // namespace pcl
