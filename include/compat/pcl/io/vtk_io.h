/* pcl::io::saveVTKFile for polygon meshes (used by src/apps/demo.cpp:237-246): legacy ASCII VTK polydata, the layout
 * pcl/io/vtk_io.cpp writes -- POINTS, VERTICES (one per point), POLYGONS.  PCL is not a dependency of sobfu_b200. */
#pragma once
#include <pcl/PolygonMesh.h>
#include <sobfu_b200_io.hpp>

#include <string>

namespace pcl {
namespace io {
inline int saveVTKFile(const std::string &file_name, const pcl::PolygonMesh &mesh, unsigned precision = 5) {
    const size_t step = mesh.cloud.point_step ? mesh.cloud.point_step : 16, n = mesh.cloud.data.size() / step;
    if (n == 0) { std::fprintf(stderr, "[pcl::io::saveVTKFile] Input point cloud has no data!\n"); return -1; }
    std::vector<uint32_t> poly;
    int k = mesh.polygons.empty() ? 3 : (int)mesh.polygons[0].vertices.size();
    poly.reserve(mesh.polygons.size() * k);
    for (const auto &p : mesh.polygons) {
        if ((int)p.vertices.size() != k) { std::fprintf(stderr, "[pcl::io::saveVTKFile] polygons of mixed size are not supported\n"); return -1; }
        poly.insert(poly.end(), p.vertices.begin(), p.vertices.end());
    }
    try {
        sobfu_b200::io::write_vtk_polydata(file_name, reinterpret_cast<const float *>(mesh.cloud.data.data()), n, step / sizeof(float),
                                           poly.data(), mesh.polygons.size(), k, (int)precision);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "[pcl::io::saveVTKFile] %s\n", e.what());
        return -1;
    }
    return 0;
}
}  // namespace io
}  // namespace pcl
