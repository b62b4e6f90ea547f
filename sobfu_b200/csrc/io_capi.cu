// io_capi.cu -- host-only C entry points for the application layer's file formats (include/sobfu_b200_io.hpp): the z-slab capable
// Python driver (sobfu_b200/app.py) reads depth maps / masks and writes meshes through the same code as the C++ applications.
// Nothing here touches the GPU.
#include <opencv2/highgui/highgui.hpp>      // include/compat: cv::imread with OpenCV's flag semantics on top of sobfu_b200_io.hpp
#include <sobfu_b200.h>

#include <cstring>
#include <string>

namespace {
thread_local std::string g_io_error;
int io_fail(const std::string &what) {
    g_io_error = what;
    return SOBFU_B200_EINVAL;
}
}  // namespace

extern "C" const char *sobfu_b200_io_last_error(void) { return g_io_error.c_str(); }

// cv::imread(path, CV_LOAD_IMAGE_ANYDEPTH) as src/apps/demo.cpp:301 uses it: 16-bit grey depth map in millimetres.  out == NULL:
// only the size is returned.
extern "C" int sobfu_b200_read_depth_png(const char *path, unsigned short *out, int capacity_pixels, int *cols, int *rows) {
    if (!path || !cols || !rows) return io_fail("read_depth_png: null argument");
    const cv::Mat m = cv::imread(path, CV_LOAD_IMAGE_ANYDEPTH);
    if (!m.data) return io_fail(std::string("image could not be read: ") + path);
    if (m.type() != CV_16UC1) return io_fail(std::string("not a 16-bit depth map: ") + path);
    *cols = m.cols;
    *rows = m.rows;
    if (!out) return 0;
    if ((long long)m.cols * m.rows > capacity_pixels) return io_fail("read_depth_png: buffer too small");
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr<unsigned short>(y), (size_t)m.cols * 2);
    return 0;
}
// cv::imread(path, CV_8U) (8-bit grey), the object masks of demo.cpp:303
extern "C" int sobfu_b200_read_mask_png(const char *path, unsigned char *out, int capacity_pixels, int *cols, int *rows) {
    if (!path || !cols || !rows) return io_fail("read_mask_png: null argument");
    const cv::Mat m = cv::imread(path, CV_8U);
    if (!m.data) return io_fail(std::string("image could not be read: ") + path);
    *cols = m.cols;
    *rows = m.rows;
    if (!out) return 0;
    if ((long long)m.cols * m.rows > capacity_pixels) return io_fail("read_mask_png: buffer too small");
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr<unsigned char>(y), (size_t)m.cols);
    return 0;
}
extern "C" int sobfu_b200_write_depth_png(const char *path, const unsigned short *depth, int cols, int rows) {
    if (!path || !depth || cols <= 0 || rows <= 0) return io_fail("write_depth_png: bad argument");
    try {
        sobfu_b200::io::write_png(path, depth, cols, rows, 1, 16);
    } catch (const std::exception &e) {
        return io_fail(e.what());
    }
    return 0;
}
// pcl::io::saveVTKFile of a triangle soup (demo.cpp:237-246): n vertices of `stride_floats` floats (x y z first), consecutive triples
// are triangles
extern "C" int sobfu_b200_write_vtk_mesh(const char *path, const float *vertices, long long n_vertices, int stride_floats) {
    if (!path || (!vertices && n_vertices > 0) || n_vertices < 0 || stride_floats < 3) return io_fail("write_vtk_mesh: bad argument");
    try {
        sobfu_b200::io::write_vtk_polydata(path, vertices, (size_t)n_vertices, (size_t)stride_floats, nullptr, (size_t)(n_vertices / 3), 3, 5);
    } catch (const std::exception &e) {
        return io_fail(e.what());
    }
    return 0;
}
