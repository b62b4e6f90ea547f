// sobfu_b200_shim.hpp -- header-only C++ drop-in for the part of dgrzech/sobfu's host API that sits on the solver hot
// path.  Same namespaces, class names, method names and signatures as the reference headers
//     include/sobfu/{params,solver,vector_fields,reductor,sob_fusion}.hpp
//     include/kfusion/{types,internal,precomp}.hpp, include/kfusion/cuda/{device_memory,device_array,tsdf_volume,
//     marching_cubes,imgproc}.hpp
// so that src/sobfu/sob_fusion.cpp-style callers and the reference's gtest sources (test/*.cpp) compile unchanged; every
// body forwards to the C ABI of libsobfu_b200.so (include/sobfu_b200.h).  The reference-named headers in include/sobfu/
// and include/kfusion/ simply include this file.  Third-party value types (cv::Vec3i, cv::Ptr, cv::Affine3f, pcl points)
// come from the caller's OpenCV/PCL or from the dependency-free stand-ins in include/compat/.
//
// Error behaviour follows the reference: any failure prints the message and exits (kfusion::cuda::error,
// src/kfusion/device_memory.cpp:7-10), there are no exceptions and no return codes at this level.
#pragma once

#include <cuda_runtime.h>
#include <sobfu_b200.h>

#include <opencv2/core/affine.hpp>
#include <opencv2/core/core.hpp>
#include <pcl/PolygonMesh.h>
#include <pcl/conversions.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

// headers the reference's own headers pull in (directly or through OpenCV / PCL / Boost) and its sources rely on
#include <boost/filesystem.hpp>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#ifndef KF_EXPORTS
#define KF_EXPORTS
#endif

struct Mat4f {
    float4 data[4];
};

namespace kfusion {
namespace cuda {
inline void error(const char *msg, const char *file, const int line, const char *func = "") {
    std::cout << "KinFu2 error: " << msg << "\t" << file << ":" << line << " " << func << std::endl;
    exit(0);
}
inline void ___shimCheck(int rc, const char *file, int line) {
    if (rc != 0) error(sobfu_b200_last_error(), file, line);
}
inline void ___cudaSafeCall(cudaError_t err, const char *file, const int line, const char *func = "") {
    if (cudaSuccess != err) error(cudaGetErrorString(err), file, line, func);
}
static inline int divUp(int total, int grain) { return (total + grain - 1) / grain; }
}  // namespace cuda
}  // namespace kfusion
#define SOBFU_SHIM_CALL(expr) kfusion::cuda::___shimCheck((expr), __FILE__, __LINE__)
#ifndef cudaSafeCall
#define cudaSafeCall(expr) kfusion::cuda::___cudaSafeCall(expr, __FILE__, __LINE__, __func__)
#endif

namespace kfusion {
typedef cv::Matx33f Mat3f;
typedef cv::Vec3f Vec3f;
typedef cv::Vec3i Vec3i;
typedef cv::Affine3f Affine3f;

struct Intr {
    float fx, fy, cx, cy;
    Intr() {}
    Intr(float fx_, float fy_, float cx_, float cy_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_) {}
    Intr operator()(int level) const { int div = 1 << level; return Intr(fx / div, fy / div, cx / div, cy / div); }
};

namespace cuda {
// reference-counted device blob (include/kfusion/cuda/device_memory.hpp:20-102)
class DeviceMemory {
public:
    DeviceMemory() {}
    DeviceMemory(size_t bytes) { create(bytes); }
    DeviceMemory(void *ptr, size_t bytes) : user_(ptr), bytes_(bytes) {}
    void create(size_t bytes) {
        if (bytes == bytes_ && (blob_ || user_)) return;
        release();
        if (bytes > 0) {
            void *p = nullptr;
            cudaSafeCall(cudaMalloc(&p, bytes));
            blob_ = std::shared_ptr<void>(p, [](void *q) { cudaFree(q); });
            bytes_ = bytes;
        }
    }
    void release() { blob_.reset(); user_ = nullptr; bytes_ = 0; }
    void copyTo(DeviceMemory &other) const {
        if (empty()) { other.release(); return; }
        other.create(bytes_);
        cudaSafeCall(cudaMemcpy(other.raw(), raw(), bytes_, cudaMemcpyDeviceToDevice));
        cudaSafeCall(cudaDeviceSynchronize());
    }
    void upload(const void *host, size_t bytes) {
        create(bytes);
        cudaSafeCall(cudaMemcpy(raw(), host, bytes, cudaMemcpyHostToDevice));
        cudaSafeCall(cudaDeviceSynchronize());
    }
    void download(void *host) const {
        cudaSafeCall(cudaMemcpy(host, raw(), bytes_, cudaMemcpyDeviceToHost));
        cudaSafeCall(cudaDeviceSynchronize());
    }
    void swap(DeviceMemory &o) { std::swap(blob_, o.blob_); std::swap(user_, o.user_); std::swap(bytes_, o.bytes_); }
    template <class T> T *ptr() { return static_cast<T *>(raw()); }
    template <class T> const T *ptr() const { return static_cast<const T *>(raw()); }
    bool empty() const { return raw() == nullptr; }
    size_t sizeBytes() const { return bytes_; }

private:
    void *raw() const { return user_ ? user_ : blob_.get(); }
    std::shared_ptr<void> blob_;
    void *user_ = nullptr;
    size_t bytes_ = 0;
};

template <class T>
class DeviceArray : public DeviceMemory {
public:
    DeviceArray() {}
    DeviceArray(size_t n) : DeviceMemory(n * sizeof(T)) {}
    DeviceArray(T *ptr, size_t n) : DeviceMemory(ptr, n * sizeof(T)) {}
    void create(size_t n) { DeviceMemory::create(n * sizeof(T)); }
    void upload(const T *host, size_t n) { DeviceMemory::upload(host, n * sizeof(T)); }
    void download(T *host) const { DeviceMemory::download(host); }
    T *ptr() { return DeviceMemory::ptr<T>(); }
    const T *ptr() const { return DeviceMemory::ptr<T>(); }
    operator T *() { return ptr(); }
    operator const T *() const { return ptr(); }
    size_t size() const { return sizeBytes() / sizeof(T); }
};

// dense (unpitched) 2-D device image: rows x cols of T
template <class T>
class DeviceArray2D : public DeviceMemory {
public:
    DeviceArray2D() {}
    DeviceArray2D(int rows, int cols) { create(rows, cols); }
    void create(int rows, int cols) { rows_ = rows; cols_ = cols; DeviceMemory::create((size_t)rows * cols * sizeof(T)); }
    void upload(const void *host, size_t host_step, int rows, int cols) {
        create(rows, cols);
        cudaSafeCall(cudaMemcpy2D(DeviceMemory::ptr<T>(), step(), host, host_step, cols * sizeof(T), rows, cudaMemcpyHostToDevice));
    }
    void download(void *host, size_t host_step) const {
        cudaSafeCall(cudaMemcpy2D(host, host_step, DeviceMemory::ptr<T>(), step(), cols_ * sizeof(T), rows_, cudaMemcpyDeviceToHost));
    }
    T *ptr(int y = 0) { return DeviceMemory::ptr<T>() + (size_t)y * cols_; }
    const T *ptr(int y = 0) const { return DeviceMemory::ptr<T>() + (size_t)y * cols_; }
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    size_t step() const { return (size_t)cols_ * sizeof(T); }

private:
    int rows_ = 0, cols_ = 0;
};

typedef DeviceMemory CudaData;
typedef DeviceArray2D<unsigned short> Depth;
typedef DeviceArray2D<float> Dists;
typedef DeviceArray<pcl::PointXYZ> Vertices;
typedef DeviceArray<pcl::Normal> Norms;
struct Surface {
    Vertices vertices;
    Norms normals;
};
inline void waitAllDefaultStream() { cudaSafeCall(cudaDeviceSynchronize()); }

// device management of include/kfusion/kinfu.hpp:25-30 (src/kfusion/core.cpp:8-212), as the application's main() uses it
inline int getCudaEnabledDeviceCount() {
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e == cudaErrorInsufficientDriver) return -1;
    if (e == cudaErrorNoDevice) return 0;
    cudaSafeCall(e);
    return count;
}
inline void setDevice(int device) { cudaSafeCall(cudaSetDevice(device)); }
inline std::string getDeviceName(int device) {
    cudaDeviceProp prop;
    cudaSafeCall(cudaGetDeviceProperties(&prop, device));
    return prop.name;
}
inline bool checkIfPreFermiGPU(int device) {
    if (device < 0) cudaSafeCall(cudaGetDevice(&device));
    cudaDeviceProp prop;
    cudaSafeCall(cudaGetDeviceProperties(&prop, device));
    return prop.major < 2;
}
// one line per device, the reference's format (core.cpp:189-212); FP32 lanes per SM: 128 on every architecture this library runs on
inline void printShortCudaDeviceInfo(int device) {
    const int count = getCudaEnabledDeviceCount();
    const bool valid = device >= 0 && device < count;
    int driver = 0, runtime = 0;
    cudaSafeCall(cudaDriverGetVersion(&driver));
    cudaSafeCall(cudaRuntimeGetVersion(&runtime));
    for (int dev = valid ? device : 0; dev < (valid ? device + 1 : count); ++dev) {
        cudaDeviceProp prop;
        cudaSafeCall(cudaGetDeviceProperties(&prop, dev));
        std::printf("Device %d:  \"%s\"  %.0fMb, sm_%d%d%s, %d cores, Driver/Runtime ver.%d.%d/%d.%d\n", dev, prop.name,
                    (float)prop.totalGlobalMem / 1048576.0f, prop.major, prop.minor, prop.major < 2 ? " (pre-Fermi)" : "",
                    128 * prop.multiProcessorCount, driver / 1000, driver % 100, runtime / 1000, runtime % 100);
    }
    std::fflush(stdout);
}
inline void printCudaDeviceInfo(int device) { printShortCudaDeviceInfo(device); }
}  // namespace cuda

// timers of include/kfusion/types.hpp:100-122 (src/kfusion/core.cpp:214-236)
struct ScopeTime {
    const char *name;
    double start;
    ScopeTime(const char *name_) : name(name_), start((double)cv::getTickCount()) {}
    ~ScopeTime() { std::cout << "Time(" << name << ") = " << ((double)cv::getTickCount() - start) * 1000.0 / cv::getTickFrequency() << "ms" << std::endl; }
};
struct SampledScopeTime {
    enum { EACH = 34 };
    SampledScopeTime(double &time_ms) : time_ms_(time_ms), start((double)cv::getTickCount()) {}
    ~SampledScopeTime() {
        static int i_ = 0;
        time_ms_ += ((double)cv::getTickCount() - start) * 1000.0 / cv::getTickFrequency();
        if (i_ % EACH == 0 && i_) {
            std::cout << "avg. frame time = " << time_ms_ / EACH << "ms (" << 1000.f * EACH / time_ms_ << "fps)" << std::endl;
            time_ms_ = 0.0;
        }
        ++i_;
    }

private:
    SampledScopeTime(const SampledScopeTime &);
    SampledScopeTime &operator=(const SampledScopeTime &);
    double &time_ms_;
    double start;
};

namespace device {
typedef int3 Vec3i;
typedef float3 Vec3f;
using kfusion::cuda::DeviceArray;
using kfusion::cuda::DeviceArray2D;
using kfusion::cuda::divUp;

// POD view of a TSDF volume (include/kfusion/internal.hpp:59-78)
struct TsdfVolume {
    float2 *const data;
    const int3 dims;
    const float3 voxel_size;
    const float trunc_dist, eta, max_weight;
    TsdfVolume(float2 *const data_, int3 dims_, float3 voxel_size_, float trunc_dist_, float eta_, float max_weight_)
        : data(data_), dims(dims_), voxel_size(voxel_size_), trunc_dist(trunc_dist_), eta(eta_), max_weight(max_weight_) {}
};
inline void clear_volume(TsdfVolume &v) { SOBFU_SHIM_CALL(sobfu_b200_tsdf_clear(v.data, v.dims.x, v.dims.y, v.dims.z)); }
}  // namespace device

template <typename D, typename S>
inline D device_cast(const S &source) {
    return *reinterpret_cast<const D *>(source.val);
}
}  // namespace kfusion

// ---- sobfu parameters (include/sobfu/params.hpp:7-38) ------------------------------------------------------------
struct Params {
    int cols = 640, rows = 480;
    cv::Vec3i volume_dims;
    cv::Vec3f volume_size;
    cv::Affine3f volume_pose;
    kfusion::Intr intr;
    float icp_truncate_depth_dist;
    float bilateral_sigma_depth, bilateral_sigma_spatial;
    int bilateral_kernel_size;
    float tsdf_trunc_dist, eta;
    float tsdf_max_weight;
    float gradient_delta_factor;
    int start_frame = 0;
    int verbosity = 0;
    int s, max_iter;
    float max_update_norm, lambda, alpha, w_reg;
    cv::Vec3f voxel_sizes() {
        return cv::Vec3f(volume_size[0] / volume_dims[0], volume_size[1] / volume_dims[1], volume_size[2] / volume_dims[2]);
    }
};

namespace kfusion {
namespace cuda {
// include/kfusion/cuda/imgproc.hpp:11-23
inline void depthBilateralFilter(const Depth &in, Depth &out, int ksz, float sigma_spatial, float sigma_depth) {
    out.create(in.rows(), in.cols());
    SOBFU_SHIM_CALL(sobfu_b200_depth_bilateral(in.ptr(), in.step(), out.ptr(), out.step(), in.cols(), in.rows(), ksz, sigma_spatial, sigma_depth));
}
inline void depthTruncation(Depth &depth, float threshold) {
    SOBFU_SHIM_CALL(sobfu_b200_depth_truncate(depth.ptr(), depth.step(), depth.cols(), depth.rows(), threshold));
}
inline void computeDists(const Depth &depth, Dists &dists, const Intr &intr) {
    dists.create(depth.rows(), depth.cols());
    SOBFU_SHIM_CALL(sobfu_b200_compute_dists(depth.ptr(), depth.step(), dists.ptr(), dists.step(), depth.cols(), depth.rows(), intr.fx, intr.fy,
                                             intr.cx, intr.cy));
}

// include/kfusion/cuda/tsdf_volume.hpp:17-92
class TsdfVolume {
public:
    TsdfVolume(const Params &params)
        : trunc_dist_(params.tsdf_trunc_dist), eta_(params.eta), max_weight_(params.tsdf_max_weight), dims_(params.volume_dims),
          size_(params.volume_size), pose_(params.volume_pose), gradient_delta_factor_(params.gradient_delta_factor) {
        create(dims_);
    }
    virtual ~TsdfVolume() {}
    void create(const Vec3i &dims) {
        dims_ = dims;
        data_.create((size_t)dims_[0] * dims_[1] * dims_[2] * 2 * sizeof(float));
        clear();
    }
    Vec3i getDims() const { return dims_; }
    Vec3f getVoxelSize() const { return Vec3f(size_[0] / dims_[0], size_[1] / dims_[1], size_[2] / dims_[2]); }
    const CudaData data() const { return data_; }
    CudaData data() { return data_; }
    Vec3f getSize() const { return size_; }
    void setSize(const Vec3f &size) { size_ = size; }
    float getTruncDist() const { return trunc_dist_; }
    void setTruncDist(float &distance) { trunc_dist_ = distance; }
    float getEta() const { return eta_; }
    void setEta(float &eta) { eta_ = eta; }
    float getMaxWeight() const { return max_weight_; }
    void setMaxWeight(float &weight) { max_weight_ = weight; }
    Affine3f getPose() const { return pose_; }
    void setPose(const Affine3f &pose) { pose_ = pose; }
    float getGradientDeltaFactor() const { return gradient_delta_factor_; }
    void setGradientDeltaFactor(float &factor) { gradient_delta_factor_ = factor; }
    virtual void clear() { SOBFU_SHIM_CALL(sobfu_b200_tsdf_clear(data_.ptr<float2>(), dims_[0], dims_[1], dims_[2])); }
    void swap(CudaData &data) { data_.swap(data); }
    virtual void applyAffine(const Affine3f &affine) { pose_ = affine * pose_; }
    virtual void integrate(const TsdfVolume &phi_n_psi) {
        SOBFU_SHIM_CALL(sobfu_b200_tsdf_fuse(data_.ptr<float2>(), phi_n_psi.data().ptr<float2>(), dims_[0], dims_[1], dims_[2], max_weight_));
    }
    virtual void integrate(const Dists &dists, const Affine3f &camera_pose, const Intr &intr) {
        Affine3f vol2cam = camera_pose.inv() * pose_;
        const Vec3f vs = getVoxelSize();
        const Mat3f R = vol2cam.rotation();
        const Vec3f t = vol2cam.translation();
        SOBFU_SHIM_CALL(sobfu_b200_tsdf_integrate(dists.ptr(), dists.step(), dists.cols(), dists.rows(), data_.ptr<float2>(), dims_[0], dims_[1],
                                                  dims_[2], vs.val, trunc_dist_, eta_, R.val, t.val, intr.fx, intr.fy, intr.cx, intr.cy));
    }
    virtual void initSphere(const float3 &centre, const float &radius) {
        const Vec3f vs = getVoxelSize();
        const float c[3] = {centre.x, centre.y, centre.z};
        SOBFU_SHIM_CALL(sobfu_b200_tsdf_init_sphere(data_.ptr<float2>(), dims_[0], dims_[1], dims_[2], vs.val, trunc_dist_, eta_, c, radius));
    }
    virtual void initBox(const float3 &b) {
        const Vec3f vs = getVoxelSize();
        const float v[3] = {b.x, b.y, b.z};
        SOBFU_SHIM_CALL(sobfu_b200_tsdf_init_box(data_.ptr<float2>(), dims_[0], dims_[1], dims_[2], vs.val, trunc_dist_, v));
    }
    virtual void initEllipsoid(const float3 &r) {
        const Vec3f vs = getVoxelSize();
        const float v[3] = {r.x, r.y, r.z};
        SOBFU_SHIM_CALL(sobfu_b200_tsdf_init_ellipsoid(data_.ptr<float2>(), dims_[0], dims_[1], dims_[2], vs.val, trunc_dist_, v));
    }
    virtual void initPlane(const float &z) {
        const Vec3f vs = getVoxelSize();
        SOBFU_SHIM_CALL(sobfu_b200_tsdf_init_plane(data_.ptr<float2>(), dims_[0], dims_[1], dims_[2], vs.val, trunc_dist_, z));
    }
    virtual void initTorus(const float2 &t) {
        const Vec3f vs = getVoxelSize();
        const float v[2] = {t.x, t.y};
        SOBFU_SHIM_CALL(sobfu_b200_tsdf_init_torus(data_.ptr<float2>(), dims_[0], dims_[1], dims_[2], vs.val, trunc_dist_, v));
    }
    void print_sdf_values() {
        std::vector<float2> h((size_t)dims_[0] * dims_[1] * dims_[2]);
        data_.download(h.data());
        for (int i = 0; i < dims_[0]; i++)
            for (int j = 0; j < dims_[1]; j++)
                for (int k = 0; k < dims_[2]; k++) {
                    const float2 v = h[(size_t)k * dims_[1] * dims_[0] + (size_t)j * dims_[0] + i];
                    if (v.x != 0.f) std::cout << v.x << std::endl;
                }
    }

private:
    CudaData data_;
    float trunc_dist_, eta_, max_weight_;
    Vec3i dims_;
    Vec3f size_;
    Affine3f pose_;
    float gradient_delta_factor_;
};

// include/kfusion/cuda/marching_cubes.hpp:19-56
class MarchingCubes {
public:
    enum { POINTS_PER_TRIANGLE = 3, DEFAULT_TRIANGLES_BUFFER_SIZE = 2 * 1000 * 1000 * POINTS_PER_TRIANGLE };
    typedef std::shared_ptr<MarchingCubes> Ptr;
    MarchingCubes() {}
    ~MarchingCubes() {}
    void setPose(const cv::Affine3f &pose_) { pose = pose_; }
    Surface run(const TsdfVolume &volume, DeviceArray<pcl::PointXYZ> &vertices_buffer, DeviceArray<pcl::Normal> &normals_buffer) {
        if (vertices_buffer.empty()) vertices_buffer.create(DEFAULT_TRIANGLES_BUFFER_SIZE);
        if (normals_buffer.empty()) normals_buffer.create(DEFAULT_TRIANGLES_BUFFER_SIZE);
        const Vec3i d = volume.getDims();
        const Vec3f size = volume.getSize();
        const Mat3f R = pose.rotation();
        const Vec3f t = pose.translation();
        int n_vertices = 0, n_voxels = 0;
        const int cap = (int)vertices_buffer.size();
        SOBFU_SHIM_CALL(sobfu_b200_marching_cubes(volume.data().ptr<float2>(), d[0], d[1], d[2], size.val, R.val, t.val, vertices_buffer.ptr(),
                                                  normals_buffer.ptr(), cap, &n_vertices, nullptr, nullptr, nullptr, cap / 3, &n_voxels));
        std::cout << "no. of active voxels: " << n_voxels << std::endl;
        Surface s;
        if (!n_voxels) return s;
        if (n_vertices > cap) n_vertices = cap;
        s.vertices = DeviceArray<pcl::PointXYZ>(vertices_buffer.ptr(), n_vertices);
        s.normals = DeviceArray<pcl::Normal>(normals_buffer.ptr(), n_vertices);
        return s;
    }

private:
    cv::Affine3f pose;
};
}  // namespace cuda
}  // namespace kfusion

// ---- fields, differentiators, reductor, solver (include/sobfu/{vector_fields,reductor,solver}.hpp) ---------------
namespace sobfu {
namespace device {
struct VectorField {
    VectorField(float4 *const data_, const int3 dims_) : data(data_), dims(dims_) {}
    float4 *const data;
    const int3 dims;
};
typedef VectorField DeformationField;
typedef VectorField TsdfGradient;
typedef VectorField Laplacian;
typedef VectorField PotentialGradient;
struct Jacobian {
    Jacobian(Mat4f *const data_, int3 dims_) : data(data_), dims(dims_) {}
    Mat4f *const data;
    const int3 dims;
};
inline void clear(VectorField &f) { SOBFU_SHIM_CALL(sobfu_b200_clear_field(f.data, f.dims.x, f.dims.y, f.dims.z)); }
inline void init_identity(DeformationField &f) { SOBFU_SHIM_CALL(sobfu_b200_init_identity(f.data, f.dims.x, f.dims.y, f.dims.z)); }
inline void apply(const kfusion::device::TsdfVolume &phi, kfusion::device::TsdfVolume &phi_warped, const DeformationField &psi) {
    SOBFU_SHIM_CALL(sobfu_b200_apply(phi.data, phi_warped.data, psi.data, psi.dims.x, psi.dims.y, psi.dims.z));
}
inline void estimate_inverse(DeformationField &psi, DeformationField &psi_inv) {
    SOBFU_SHIM_CALL(sobfu_b200_estimate_inverse(psi.data, psi_inv.data, psi.dims.x, psi.dims.y, psi.dims.z, 48));
}
struct TsdfDifferentiator {
    TsdfDifferentiator(kfusion::device::TsdfVolume &vol_) : vol(vol_) {}
    void calculate(TsdfGradient &grad) { SOBFU_SHIM_CALL(sobfu_b200_tsdf_gradient(vol.data, grad.data, vol.dims.x, vol.dims.y, vol.dims.z)); }
    kfusion::device::TsdfVolume vol;
};
struct SecondOrderDifferentiator {
    SecondOrderDifferentiator(DeformationField &psi_) : psi(psi_) {}
    void calculate(Laplacian &L) { SOBFU_SHIM_CALL(sobfu_b200_laplacian(psi.data, L.data, psi.dims.x, psi.dims.y, psi.dims.z)); }
    DeformationField psi;
};
struct Differentiator {
    Differentiator(DeformationField &psi_) : psi(psi_) {}
    void calculate(Jacobian &J) { SOBFU_SHIM_CALL(sobfu_b200_jacobian(psi.data, J.data, psi.dims.x, psi.dims.y, psi.dims.z, 0)); }
    void calculate_deformation_jacobian(Jacobian &J) { SOBFU_SHIM_CALL(sobfu_b200_jacobian(psi.data, J.data, psi.dims.x, psi.dims.y, psi.dims.z, 1)); }
    DeformationField psi;
};
// device-level free functions and helper structs of include/sobfu/solver.hpp:16-47,100-121 and vector_fields.hpp:258-277: the
// plumbing between the reference's Solver and its kernels.  The stages are available one by one (the whole loop is
// sobfu::cuda::Solver::estimate_psi); the three per-axis convolution launchers are provided as one call that does what the
// reference's rows + columns + depth sequence does (solver.cu:152-159).
inline void calculate_potential_gradient(kfusion::device::TsdfVolume &phi_n_psi, kfusion::device::TsdfVolume &phi_global, TsdfGradient &nabla_phi_n_o_psi,
                                         Laplacian &L, PotentialGradient &nabla_U, float w_reg) {
    SOBFU_SHIM_CALL(sobfu_b200_potential_gradient(phi_n_psi.data, phi_global.data, nabla_phi_n_o_psi.data, L.data, nabla_U.data, w_reg,
                                                  nabla_U.dims.x, nabla_U.dims.y, nabla_U.dims.z));
}
inline void update_psi(DeformationField &psi, PotentialGradient &nabla_U_S, float4 *updates, float alpha) {
    SOBFU_SHIM_CALL(sobfu_b200_update_psi(psi.data, nabla_U_S.data, updates, alpha, psi.dims.x, psi.dims.y, psi.dims.z));
}
// nabla_U_S = S *x nabla_U + S *y nabla_U + S *z nabla_U with the seven HOST taps h_S_i (convolution_rows / _columns / _depth)
inline void convolve_sobolev(PotentialGradient &nabla_U_S, PotentialGradient &nabla_U, const float *h_S_i) {
    SOBFU_SHIM_CALL(sobfu_b200_sobolev_filter(nabla_U_S.data, nabla_U.data, h_S_i, nabla_U.dims.x, nabla_U.dims.y, nabla_U.dims.z));
}
struct SpatialGradients {
    SpatialGradients(TsdfGradient *nabla_phi_n_, TsdfGradient *nabla_phi_n_o_psi_, Jacobian *J_, Jacobian *J_inv_, Laplacian *L_, Laplacian *L_o_psi_inv_,
                     PotentialGradient *nabla_U_, PotentialGradient *nabla_U_S_)
        : nabla_phi_n(nabla_phi_n_), nabla_phi_n_o_psi(nabla_phi_n_o_psi_), J(J_), J_inv(J_inv_), L(L_), L_o_psi_inv(L_o_psi_inv_), nabla_U(nabla_U_),
          nabla_U_S(nabla_U_S_) {}
    TsdfGradient *nabla_phi_n, *nabla_phi_n_o_psi;
    Jacobian *J, *J_inv;
    Laplacian *L, *L_o_psi_inv;
    PotentialGradient *nabla_U, *nabla_U_S;
};
// include/sobfu/reductor.hpp:24-50
struct Reductor {
    Reductor(int3 dims_, float vsz_, float trunc_dist_) : dims(dims_), vsz(vsz_), trunc_dist(trunc_dist_), no_voxels(dims_.x * dims_.y * dims_.z) {
        cudaSafeCall(cudaMalloc((void **)&updates, (size_t)no_voxels * sizeof(float4)));
    }
    ~Reductor() { cudaFree(updates); }
    float data_energy(float2 *phi_global_data, float2 *phi_n_data) {
        float e = 0.f;
        SOBFU_SHIM_CALL(sobfu_b200_data_energy(phi_global_data, phi_n_data, no_voxels, &e));
        return e;
    }
    float reg_energy_sobolev(Mat4f *J_data) {
        float e = 0.f;
        SOBFU_SHIM_CALL(sobfu_b200_reg_energy(J_data, no_voxels, &e));
        return e;
    }
    float2 max_update_norm() {
        float v = 0.f, i = 0.f;
        SOBFU_SHIM_CALL(sobfu_b200_max_update_norm(updates, no_voxels, &v, &i, nullptr));
        return make_float2(v, i);
    }
    int3 dims;
    float vsz, trunc_dist;
    int no_voxels;
    float4 *updates;
};
}  // namespace device
}  // namespace sobfu

// global-scope helper structs of include/sobfu/solver.hpp:16-47
struct SolverParams {
    int verbosity, max_iter, s;
    float max_update_norm, lambda, alpha, w_reg;
};
struct SDFs {
    SDFs(kfusion::device::TsdfVolume &phi_global_, kfusion::device::TsdfVolume &phi_global_psi_inv_, kfusion::device::TsdfVolume &phi_n_,
         kfusion::device::TsdfVolume &phi_n_psi_)
        : phi_global(phi_global_), phi_global_psi_inv(phi_global_psi_inv_), phi_n(phi_n_), phi_n_psi(phi_n_psi_) {}
    kfusion::device::TsdfVolume phi_global, phi_global_psi_inv, phi_n, phi_n_psi;
};
struct Differentiators {
    Differentiators(sobfu::device::TsdfDifferentiator &tsdf_diff_, sobfu::device::Differentiator &diff_, sobfu::device::Differentiator &diff_inv_,
                    sobfu::device::SecondOrderDifferentiator &second_order_diff_)
        : tsdf_diff(tsdf_diff_), diff(diff_), diff_inv(diff_inv_), second_order_diff(second_order_diff_) {}
    sobfu::device::TsdfDifferentiator tsdf_diff;
    sobfu::device::Differentiator diff, diff_inv;
    sobfu::device::SecondOrderDifferentiator second_order_diff;
};

namespace sobfu {
namespace cuda {
class VectorField {
public:
    VectorField(cv::Vec3i dims_) : dims(dims_) {
        data.create((size_t)dims[0] * dims[1] * dims[2] * sizeof(float4));
        clear();
    }
    ~VectorField() {}
    cv::Vec3i get_dims() const { return dims; }
    kfusion::cuda::CudaData get_data() { return data; }
    const kfusion::cuda::CudaData get_data() const { return data; }
    void set_data(kfusion::cuda::CudaData &d) { data = d; }
    void clear() { SOBFU_SHIM_CALL(sobfu_b200_clear_field(data.ptr<float4>(), dims[0], dims[1], dims[2])); }
    void print() {
        std::vector<float4> h((size_t)dims[0] * dims[1] * dims[2]);
        data.download(h.data());
        std::cout << "--- FIELD ---" << std::endl;
        for (int i = 0; i < dims[0]; i++)
            for (int j = 0; j < dims[1]; j++)
                for (int k = 0; k < dims[2]; k++) {
                    const float4 v = h[((size_t)k * dims[1] + j) * dims[0] + i];
                    if (fabs(v.x) > 1e-5f || fabs(v.y) > 1e-5f || fabs(v.z) > 1e-5f)
                        std::cout << "(x,y,z)=(" << i << ", " << j << ", " << k << "), (u,v,w)=(" << v.x << ", " << v.y << "," << v.z << ")" << std::endl;
                }
    }
    int get_no_nans() {
        std::vector<float4> h((size_t)dims[0] * dims[1] * dims[2]);
        data.download(h.data());
        int n = 0;
        for (const float4 &v : h) n += (std::isnan(v.x) || std::isnan(v.y) || std::isnan(v.z)) ? 1 : 0;
        return n;
    }

protected:
    kfusion::cuda::CudaData data;
    cv::Vec3i dims;
};
class DeformationField : public VectorField {
public:
    DeformationField(cv::Vec3i dims_) : VectorField(dims_) { clear(); }
    ~DeformationField() {}
    void clear() { SOBFU_SHIM_CALL(sobfu_b200_init_identity(data.ptr<float4>(), dims[0], dims[1], dims[2])); }
    void apply(const cv::Ptr<kfusion::cuda::TsdfVolume> phi, cv::Ptr<kfusion::cuda::TsdfVolume> phi_psi) {
        SOBFU_SHIM_CALL(sobfu_b200_apply(phi->data().ptr<float2>(), phi_psi->data().ptr<float2>(), data.ptr<float4>(), dims[0], dims[1], dims[2]));
    }
    void get_inverse(sobfu::cuda::DeformationField &psi_inv) {
        SOBFU_SHIM_CALL(sobfu_b200_estimate_inverse(data.ptr<float4>(), psi_inv.get_data().ptr<float4>(), dims[0], dims[1], dims[2], 48));
    }
};
typedef VectorField TsdfGradient;
typedef VectorField Laplacian;
typedef VectorField PotentialGradient;

// include/sobfu/solver.hpp:56-67
class Solver {
public:
    Solver(Params &params) {
        sobfu_b200_params p;
        for (int i = 0; i < 3; ++i) { p.dims[i] = params.volume_dims[i]; p.voxel_size[i] = params.voxel_sizes()[i]; }
        p.trunc_dist = params.tsdf_trunc_dist; p.eta = params.eta; p.max_weight = params.tsdf_max_weight;
        p.verbosity = params.verbosity; p.max_iter = params.max_iter; p.s = params.s;
        p.max_update_norm = params.max_update_norm; p.lambda = params.lambda; p.alpha = params.alpha; p.w_reg = params.w_reg;
        SOBFU_SHIM_CALL(sobfu_b200_solver_create(&h_, &p));
    }
    ~Solver() { sobfu_b200_solver_destroy(h_); }
    Solver(const Solver &) = delete;
    Solver &operator=(const Solver &) = delete;
    void estimate_psi(const cv::Ptr<kfusion::cuda::TsdfVolume> phi_global, cv::Ptr<kfusion::cuda::TsdfVolume> phi_global_psi_inv,
                      const cv::Ptr<kfusion::cuda::TsdfVolume> phi_n, cv::Ptr<kfusion::cuda::TsdfVolume> phi_n_psi,
                      std::shared_ptr<sobfu::cuda::DeformationField> psi, std::shared_ptr<sobfu::cuda::DeformationField> psi_inv) {
        SOBFU_SHIM_CALL(sobfu_b200_solver_estimate_psi(h_, phi_global->data().ptr<float2>(), phi_global_psi_inv->data().ptr<float2>(),
                                                       phi_n->data().ptr<float2>(), phi_n_psi->data().ptr<float2>(), psi->get_data().ptr<float4>(),
                                                       psi_inv->get_data().ptr<float4>(), &info));
    }
    sobfu_b200_solve_info info;   // extension: outcome of the last call

private:
    sobfu_b200_solver *h_ = nullptr;
};
}  // namespace cuda
}  // namespace sobfu

// ---- per-frame pipeline (include/sobfu/sob_fusion.hpp:17-74, src/sobfu/sob_fusion.cpp:71-183) --------------------
class SobFusion {
public:
    SobFusion(const Params &params_) : frame_counter_(0), params(params_) {
        poses_.push_back(cv::Affine3f::Identity());
        mc = cv::Ptr<kfusion::cuda::MarchingCubes>(new kfusion::cuda::MarchingCubes());
        mc->setPose(params.volume_pose);
    }
    ~SobFusion() {}
    Params &getParams() { return params; }
    std::shared_ptr<sobfu::cuda::DeformationField> getDeformationField() { return psi; }
    pcl::PolygonMesh::Ptr get_phi_global_mesh() { return get_mesh(phi_global); }
    pcl::PolygonMesh::Ptr get_phi_global_psi_inv_mesh() { return get_mesh(phi_global_psi_inv); }
    pcl::PolygonMesh::Ptr get_phi_n_mesh() { return get_mesh(phi_n); }
    pcl::PolygonMesh::Ptr get_phi_n_psi_mesh() { return get_mesh(phi_n_psi); }

    bool operator()(const kfusion::cuda::Depth &depth) {
        std::cout << "--- FRAME NO. " << frame_counter_ << " ---" << std::endl;
        kfusion::cuda::depthBilateralFilter(depth, depth_f_, params.bilateral_kernel_size, params.bilateral_sigma_spatial, params.bilateral_sigma_depth);
        kfusion::cuda::depthTruncation(depth_f_, params.icp_truncate_depth_dist);
        kfusion::cuda::computeDists(depth_f_, dists_, params.intr);
        typedef kfusion::cuda::TsdfVolume Vol;
        if (frame_counter_ == 0) {
            phi_global = cv::Ptr<Vol>(new Vol(params));
            phi_global->integrate(dists_, poses_.back(), params.intr);
            phi_global_psi_inv = cv::Ptr<Vol>(new Vol(params));
            phi_n = cv::Ptr<Vol>(new Vol(params));
            phi_n_psi = cv::Ptr<Vol>(new Vol(params));
            psi = std::make_shared<sobfu::cuda::DeformationField>(params.volume_dims);
            psi_inv = std::make_shared<sobfu::cuda::DeformationField>(params.volume_dims);
            solver = std::make_shared<sobfu::cuda::Solver>(params);
            return ++frame_counter_, true;
        }
        phi_n->clear();
        phi_n->integrate(dists_, poses_.back(), params.intr);
        if (frame_counter_ < params.start_frame) {
            phi_global->integrate(*phi_n);
            return ++frame_counter_, true;
        }
        solver->estimate_psi(phi_global, phi_global_psi_inv, phi_n, phi_n_psi, psi, psi_inv);
        phi_global->integrate(*phi_n_psi);
        return ++frame_counter_, true;
    }

    cv::Ptr<kfusion::cuda::TsdfVolume> phi_global, phi_global_psi_inv, phi_n, phi_n_psi;
    std::shared_ptr<sobfu::cuda::DeformationField> psi, psi_inv;
    std::shared_ptr<sobfu::cuda::Solver> solver;

private:
    pcl::PolygonMesh::Ptr get_mesh(cv::Ptr<kfusion::cuda::TsdfVolume> vol) {
        kfusion::cuda::DeviceArray<pcl::PointXYZ> vb;
        kfusion::cuda::DeviceArray<pcl::Normal> nb;
        kfusion::cuda::Surface s = mc->run(*vol, vb, nb);
        pcl::PolygonMesh::Ptr mesh(new pcl::PolygonMesh());
        const size_t n = s.vertices.size();
        if (n == 0) return mesh;
        pcl::PointCloud<pcl::PointXYZ> cloud;
        cloud.width = (unsigned)n; cloud.height = 1;
        cloud.points.resize(n);
        s.vertices.download(cloud.points.data());
        pcl::toPCLPointCloud2(cloud, mesh->cloud);
        mesh->polygons.resize(n / 3);
        for (size_t i = 0; i < n / 3; ++i) {
            pcl::Vertices v;
            v.vertices = {(uint32_t)(3 * i), (uint32_t)(3 * i + 1), (uint32_t)(3 * i + 2)};
            mesh->polygons[i] = v;
        }
        return mesh;
    }
    int frame_counter_;
    Params params;
    std::vector<cv::Affine3f> poses_;
    kfusion::cuda::Depth depth_f_;
    kfusion::cuda::Dists dists_;
    cv::Ptr<kfusion::cuda::MarchingCubes> mc;
};
