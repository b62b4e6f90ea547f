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
}  // namespace pcl
