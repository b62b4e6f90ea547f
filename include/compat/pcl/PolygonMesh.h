#pragma once
#include <pcl/PCLPointCloud2.h>
#include <memory>
#include <vector>
namespace pcl {
struct Vertices { std::vector<uint32_t> vertices; };
struct PolygonMesh {
    typedef std::shared_ptr<PolygonMesh> Ptr;
    PCLPointCloud2 cloud;
    std::vector<Vertices> polygons;
};
}  // namespace pcl
