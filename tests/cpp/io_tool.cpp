// CPU-only test helper for the host-side file formats (include/sobfu_b200_io.hpp and the OpenCV / Boost / PCL / VTK stand-ins
// built on it).  Driven by tests/test_io_cpu.py.
//   io_tool imread <png> <flags> <out.raw>       cv::imread -> "rows cols type\n" + pixel bytes
//   io_tool imwrite16 <out.png> <cols> <rows>     cv::imwrite of a 16-bit ramp
//   io_tool ini <file>                            the application's option set (demo.cpp:84-160) -> NAME=value lines
//   io_tool vtk <out.vtk>                         pcl::io::saveVTKFile of a two-triangle mesh
//   io_tool vti <out.vti>                         vtkXMLImageDataWriter of a 3x2x2 field with 4 components
//   io_tool glob <dir>                            cv::glob + sort
#include <boost/program_options.hpp>
#include <opencv2/highgui/highgui.hpp>
#include <pcl/conversions.h>
#include <pcl/io/vtk_io.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <vtkImageData.h>
#include <vtkSmartPointer.h>
#include <vtkXMLImageDataWriter.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

namespace po = boost::program_options;

static int cmd_ini(const char *path) {
    int dims[3] = {0, 0, 0}, ksz = 0, start = 0, max_iter = 0, s = 0;
    float size[3] = {0, 0, 0}, maxw = 0, gdf = 0, fx = 0, fy = 0, cx = 0, cy = 0, trunc_depth = 0, sd = 0, ss = 0, mun = 0, lambda = 0, alpha = 0, wreg = 0;
    po::options_description desc("parameters");
    desc.add_options()("VOL_DIMS_X", po::value<int>(&dims[0]), "")("VOL_DIMS_Y", po::value<int>(&dims[1]), "")("VOL_DIMS_Z", po::value<int>(&dims[2]), "");
    desc.add_options()("VOL_SIZE_X", po::value<float>(&size[0]), "")("VOL_SIZE_Y", po::value<float>(&size[1]), "")("VOL_SIZE_Z", po::value<float>(&size[2]), "");
    desc.add_options()("TSDF_TRUNC_DIST", po::value<float>(), "truncation distance (voxels)");
    desc.add_options()("ETA", po::value<float>(), "expected object thickness (voxels)");
    desc.add_options()("TSDF_MAX_WEIGHT", po::value<float>(&maxw), "")("GRADIENT_DELTA_FACTOR", po::value<float>(&gdf), "");
    desc.add_options()("INTR_FX", po::value<float>(&fx), "")("INTR_FY", po::value<float>(&fy), "")("INTR_CX", po::value<float>(&cx), "")("INTR_CY", po::value<float>(&cy), "");
    desc.add_options()("TRUNC_DEPTH", po::value<float>(&trunc_depth), "")("VOL_POSE_T_Z", po::value<float>(), "");
    desc.add_options()("BILATERAL_SIGMA_DEPTH", po::value<float>(&sd), "")("BILATERAL_SIGMA_SPATIAL", po::value<float>(&ss), "")("BILATERAL_KERNEL_SIZE", po::value<int>(&ksz), "");
    desc.add_options()("START_FRAME", po::value<int>(&start), "")("MAX_ITER", po::value<int>(&max_iter), "")("MAX_UPDATE_NORM", po::value<float>(&mun), "");
    desc.add_options()("S", po::value<int>(&s), "")("LAMBDA", po::value<float>(&lambda), "")("ALPHA", po::value<float>(&alpha), "")("W_REG", po::value<float>(&wreg), "");
    po::variables_map vm;
    std::ifstream f(path);
    try {
        po::store(po::parse_config_file(f, desc), vm);
        po::notify(vm);
    } catch (const po::error &e) {
        std::printf("ERROR %s\n", e.what());
        return 3;
    }
    std::printf("VOL_DIMS=%d %d %d\nVOL_SIZE=%.9g %.9g %.9g\n", dims[0], dims[1], dims[2], size[0], size[1], size[2]);
    if (vm.count("TSDF_TRUNC_DIST")) std::printf("TSDF_TRUNC_DIST=%.9g\n", vm["TSDF_TRUNC_DIST"].as<float>());
    if (vm.count("ETA")) std::printf("ETA=%.9g\n", vm["ETA"].as<float>());
    if (vm.count("VOL_POSE_T_Z")) std::printf("VOL_POSE_T_Z=%.9g\n", vm["VOL_POSE_T_Z"].as<float>());
    std::printf("TSDF_MAX_WEIGHT=%.9g\nGRADIENT_DELTA_FACTOR=%.9g\nINTR=%.9g %.9g %.9g %.9g\nTRUNC_DEPTH=%.9g\n", maxw, gdf, fx, fy, cx, cy, trunc_depth);
    std::printf("BILATERAL=%.9g %.9g %d\nSTART_FRAME=%d\nMAX_ITER=%d\nMAX_UPDATE_NORM=%.9g\nS=%d\nLAMBDA=%.9g\nALPHA=%.9g\nW_REG=%.9g\n", sd, ss, ksz, start, max_iter,
                mun, s, lambda, alpha, wreg);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const std::string cmd = argv[1];
    if (cmd == "imread" && argc == 5) {
        cv::Mat m = cv::imread(argv[2], std::atoi(argv[3]));
        if (!m.data) { std::printf("EMPTY\n"); return 0; }
        FILE *f = std::fopen(argv[4], "wb");
        std::fprintf(f, "%d %d %d\n", m.rows, m.cols, m.type());
        for (int y = 0; y < m.rows; ++y) std::fwrite(m.data + (size_t)y * m.step, 1, (size_t)m.cols * m.elemSize(), f);
        std::fclose(f);
        std::printf("OK %d %d %d\n", m.rows, m.cols, m.type());
        return 0;
    }
    if (cmd == "imwrite16" && argc == 5) {
        const int cols = std::atoi(argv[3]), rows = std::atoi(argv[4]);
        cv::Mat m(rows, cols, CV_16UC1);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) m.ptr<unsigned short>(y)[x] = (unsigned short)((y * 257 + x * 3) & 0xffff);
        cv::Mat mask = cv::Mat::zeros(m.size(), CV_8UC1), masked = cv::Mat::zeros(m.size(), m.type());
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) mask.ptr<unsigned char>(y)[x] = (unsigned char)((x + y) % 3 == 0 ? 255 : 0);
        m.copyTo(masked, mask);                      // as the application does with the object masks (demo.cpp:304-308)
        return cv::imwrite(argv[2], masked) ? 0 : 1;
    }
    if (cmd == "ini") return cmd_ini(argv[2]);
    if (cmd == "vtk") {
        pcl::PointCloud<pcl::PointXYZ> cloud;
        const float P[6][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0.5f, 0.25f, -1.5f}, {1e-3f, 123456.789f, 2}, {3, 2, 1}};
        for (auto &p : P) cloud.push_back(pcl::PointXYZ(p[0], p[1], p[2]));
        cloud.width = 6; cloud.height = 1;
        pcl::PolygonMesh mesh;
        pcl::toPCLPointCloud2(cloud, mesh.cloud);
        mesh.polygons.resize(2);
        mesh.polygons[0].vertices = {0, 1, 2};
        mesh.polygons[1].vertices = {3, 4, 5};
        pcl::PointCloud<pcl::PointXYZ> back;
        pcl::fromPCLPointCloud2(mesh.cloud, back);
        if (back.size() != 6 || back[4].y != 123456.789f) return 1;
        return pcl::io::saveVTKFile(argv[2], mesh) == 0 ? 0 : 1;
    }
    if (cmd == "vti") {
        vtkSmartPointer<vtkImageData> image = vtkSmartPointer<vtkImageData>::New();
        image->SetDimensions(3, 2, 2);
        image->AllocateScalars(VTK_FLOAT, 4);
        float *p = static_cast<float *>(image->GetScalarPointer());
        for (int i = 0; i < 3 * 2 * 2 * 4; ++i) p[i] = 0.5f * i;
        vtkSmartPointer<vtkXMLImageDataWriter> writer = vtkSmartPointer<vtkXMLImageDataWriter>::New();
        writer->SetFileName(argv[2]);
        writer->SetInputData(image);
        return writer->Write() ? 0 : 1;
    }
    if (cmd == "glob") {
        std::vector<cv::String> files;
        cv::glob(argv[2], files);
        std::sort(files.begin(), files.end());
        for (auto &f : files) std::printf("%s\n", f.c_str());
        return 0;
    }
    return 2;
}
