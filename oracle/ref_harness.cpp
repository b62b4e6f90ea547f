/*
 * ref_harness.cpp -- C entry points over the UNMODIFIED reference (dgrzech/sobfu) built from
 * /root/reference by oracle/build_ref.sh into oracle/_ref/libsobfu_ref.so.
 *
 * TEST INFRASTRUCTURE ONLY: used to (a) dump golden vectors from the reference's own CUDA
 * (oracle/make_golden.py), (b) check sobfu_b200 against the reference on the GPU box, and (c) run the
 * reference arm of bench.py (--impl reference).  It calls the reference strictly through its public host
 * API: sobfu::cuda::Solver, sobfu::cuda::DeformationField, kfusion::cuda::TsdfVolume,
 * kfusion::cuda::MarchingCubes and the kfusion::cuda image-processing free functions.
 */
#include <kfusion/cuda/imgproc.hpp>
#include <kfusion/cuda/marching_cubes.hpp>
#include <kfusion/cuda/tsdf_volume.hpp>
#include <kfusion/internal.hpp>
#include <kfusion/precomp.hpp>
#include <sobfu/params.hpp>
#include <sobfu/reductor.hpp>
#include <sobfu/solver.hpp>
#include <sobfu/vector_fields.hpp>

#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <vector>

namespace {
struct Ref {
    Params params;
    cv::Ptr<kfusion::cuda::TsdfVolume> vol[4]; /* 0 phi_global, 1 phi_global_psi_inv, 2 phi_n, 3 phi_n_psi */
    std::shared_ptr<sobfu::cuda::DeformationField> psi, psi_inv;
    std::shared_ptr<sobfu::cuda::Solver> solver;
    cv::Ptr<kfusion::cuda::MarchingCubes> mc;
    kfusion::cuda::Depth depth_in, depth_f;
    kfusion::cuda::Dists dists;
    size_t N;
};
}  // namespace

extern "C" {

void *ref_create(int X, int Y, int Z, float sx, float sy, float sz, float trunc_dist, float eta, float max_weight,
                 int verbosity, int max_iter, int s, float max_update_norm, float lambda, float alpha, float w_reg,
                 float pose_tx, float pose_ty, float pose_tz, float fx, float fy, float cx, float cy) {
    Ref *r = new Ref();
    Params &p = r->params;
    p.volume_dims = cv::Vec3i(X, Y, Z);
    p.volume_size = cv::Vec3f(sx, sy, sz);
    p.volume_pose = cv::Affine3f().translate(cv::Vec3f(pose_tx, pose_ty, pose_tz));
    p.intr = kfusion::Intr(fx, fy, cx, cy);
    p.icp_truncate_depth_dist = 0.f;
    p.bilateral_sigma_depth = 0.f;
    p.bilateral_sigma_spatial = 0.f;
    p.bilateral_kernel_size = 0;
    p.tsdf_trunc_dist = trunc_dist;
    p.eta = eta;
    p.tsdf_max_weight = max_weight;
    p.gradient_delta_factor = 0.5f;
    p.verbosity = verbosity;
    p.max_iter = max_iter;
    p.s = s;
    p.max_update_norm = max_update_norm;
    p.lambda = lambda;
    p.alpha = alpha;
    p.w_reg = w_reg;
    r->N = (size_t)X * Y * Z;
    for (int i = 0; i < 4; ++i) r->vol[i] = cv::Ptr<kfusion::cuda::TsdfVolume>(new kfusion::cuda::TsdfVolume(p));
    r->psi = std::make_shared<sobfu::cuda::DeformationField>(p.volume_dims);
    r->psi_inv = std::make_shared<sobfu::cuda::DeformationField>(p.volume_dims);
    r->solver = std::make_shared<sobfu::cuda::Solver>(p);
    return r;
}

void ref_destroy(void *h) { delete (Ref *)h; }

void ref_upload_tsdf(void *h, int which, const void *host) {
    Ref *r = (Ref *)h;
    cudaMemcpy(r->vol[which]->data().ptr<float2>(), host, r->N * sizeof(float2), cudaMemcpyHostToDevice);
}
void ref_download_tsdf(void *h, int which, void *host) {
    Ref *r = (Ref *)h;
    cudaMemcpy(host, r->vol[which]->data().ptr<float2>(), r->N * sizeof(float2), cudaMemcpyDeviceToHost);
}
void ref_upload_psi(void *h, int which, const void *host) {
    Ref *r = (Ref *)h;
    auto &f = which ? r->psi_inv : r->psi;
    cudaMemcpy(f->get_data().ptr<float4>(), host, r->N * sizeof(float4), cudaMemcpyHostToDevice);
}
void ref_download_psi(void *h, int which, void *host) {
    Ref *r = (Ref *)h;
    auto &f = which ? r->psi_inv : r->psi;
    cudaMemcpy(host, f->get_data().ptr<float4>(), r->N * sizeof(float4), cudaMemcpyDeviceToHost);
}
void *ref_tsdf_devptr(void *h, int which) { return ((Ref *)h)->vol[which]->data().ptr<float2>(); }
void *ref_psi_devptr(void *h, int which) {
    Ref *r = (Ref *)h;
    return (which ? r->psi_inv : r->psi)->get_data().ptr<float4>();
}

void ref_psi_clear(void *h, int which) {
    Ref *r = (Ref *)h;
    (which ? r->psi_inv : r->psi)->clear();
    cudaDeviceSynchronize();
}
void ref_tsdf_clear(void *h, int which) { ((Ref *)h)->vol[which]->clear(); cudaDeviceSynchronize(); }
void ref_init_sphere(void *h, int which, float cx, float cy, float cz, float radius) {
    ((Ref *)h)->vol[which]->initSphere(make_float3(cx, cy, cz), radius);
}

/* TsdfVolume::initBox / initEllipsoid / initPlane / initTorus (tsdf_volume.cpp:108-146) */
void ref_init_shape(void *h, int which, int shape, float a, float b, float c) {
    kfusion::cuda::TsdfVolume &v = *((Ref *)h)->vol[which];
    if (shape == 0) v.initBox(make_float3(a, b, c));
    else if (shape == 1) v.initEllipsoid(make_float3(a, b, c));
    else if (shape == 2) v.initPlane(a);
    else v.initTorus(make_float2(a, b));
}

/* sobfu::cuda::Solver::estimate_psi (solver.cpp:69); returns device milliseconds of the call */
float ref_estimate_psi(void *h) {
    Ref *r = (Ref *)h;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    r->solver->estimate_psi(r->vol[0], r->vol[1], r->vol[2], r->vol[3], r->psi, r->psi_inv);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ms;
}

/* DeformationField::apply (vector_fields.cpp:106): out = in o psi */
void ref_apply(void *h, int which_psi, int in, int out) {
    Ref *r = (Ref *)h;
    (which_psi ? r->psi_inv : r->psi)->apply(r->vol[in], r->vol[out]);
}
/* DeformationField::get_inverse (vector_fields.cpp:95); psi_inv must hold the starting guess */
void ref_get_inverse(void *h) {
    Ref *r = (Ref *)h;
    r->psi->get_inverse(*r->psi_inv);
    cudaDeviceSynchronize();
}
/* TsdfVolume::integrate(const TsdfVolume&) (tsdf_volume.cpp:84) */
void ref_fuse(void *h, int dst, int src) { ((Ref *)h)->vol[dst]->integrate(*((Ref *)h)->vol[src]); }

/* depth preprocessing exactly as SobFusion::operator() does it (sob_fusion.cpp:78-91) then
 * TsdfVolume::integrate(dists, pose = identity, intr) (sob_fusion.cpp:130) */
void ref_depth_to_dists(void *h, const unsigned short *depth_host, int cols, int rows, int ksz, float sigma_spatial,
                        float sigma_depth, float trunc_depth, unsigned short *filtered_out, float *dists_out) {
    Ref *r = (Ref *)h;
    r->depth_in.create(rows, cols);
    r->depth_in.upload(depth_host, cols * sizeof(unsigned short), rows, cols);
    kfusion::cuda::depthBilateralFilter(r->depth_in, r->depth_f, ksz, sigma_spatial, sigma_depth);
    kfusion::cuda::depthTruncation(r->depth_f, trunc_depth);
    kfusion::cuda::computeDists(r->depth_f, r->dists, r->params.intr);
    cudaDeviceSynchronize();
    if (filtered_out) r->depth_f.download(filtered_out, cols * sizeof(unsigned short));
    if (dists_out) r->dists.download(dists_out, cols * sizeof(float));
}
void ref_integrate_dists(void *h, int which) {
    Ref *r = (Ref *)h;
    r->vol[which]->integrate(r->dists, cv::Affine3f::Identity(), r->params.intr);
}

/* test-API entry points used by the reference's gtest harness */
void ref_tsdf_gradient(void *h, int which, void *grad_host) {
    Ref *r = (Ref *)h;
    cv::Vec3i d = r->params.volume_dims;
    int3 dims = make_int3(d[0], d[1], d[2]);
    cv::Vec3f v = r->params.voxel_sizes();
    float4 *g;
    cudaMalloc(&g, r->N * sizeof(float4));
    kfusion::device::TsdfVolume vol(r->vol[which]->data().ptr<float2>(), dims, make_float3(v[0], v[1], v[2]),
                                    r->params.tsdf_trunc_dist, r->params.eta, r->params.tsdf_max_weight);
    sobfu::device::TsdfGradient grad(g, dims);
    sobfu::device::TsdfDifferentiator diff(vol);
    diff.calculate(grad);
    cudaMemcpy(grad_host, g, r->N * sizeof(float4), cudaMemcpyDeviceToHost);
    cudaFree(g);
}
void ref_laplacian(void *h, void *L_host) {
    Ref *r = (Ref *)h;
    cv::Vec3i d = r->params.volume_dims;
    int3 dims = make_int3(d[0], d[1], d[2]);
    float4 *g;
    cudaMalloc(&g, r->N * sizeof(float4));
    sobfu::device::DeformationField psi(r->psi->get_data().ptr<float4>(), dims);
    sobfu::device::Laplacian L(g, dims);
    sobfu::device::SecondOrderDifferentiator diff(psi);
    diff.calculate(L);
    cudaMemcpy(L_host, g, r->N * sizeof(float4), cudaMemcpyDeviceToHost);
    cudaFree(g);
}
void ref_jacobian(void *h, int mode, void *J_host) {
    Ref *r = (Ref *)h;
    cv::Vec3i d = r->params.volume_dims;
    int3 dims = make_int3(d[0], d[1], d[2]);
    Mat4f *g;
    cudaMalloc(&g, r->N * sizeof(Mat4f));
    cudaMemset(g, 0, r->N * sizeof(Mat4f));
    sobfu::device::DeformationField psi(r->psi->get_data().ptr<float4>(), dims);
    sobfu::device::Jacobian J(g, dims);
    sobfu::device::Differentiator diff(psi);
    if (mode == 0) diff.calculate(J); else diff.calculate_deformation_jacobian(J);
    cudaMemcpy(J_host, g, r->N * sizeof(Mat4f), cudaMemcpyDeviceToHost);
    cudaFree(g);
}
float ref_data_energy(void *h, int a, int b) {
    Ref *r = (Ref *)h;
    cv::Vec3i d = r->params.volume_dims;
    sobfu::device::Reductor red(make_int3(d[0], d[1], d[2]), r->params.voxel_sizes()[0], r->params.tsdf_trunc_dist);
    return red.data_energy(r->vol[a]->data().ptr<float2>(), r->vol[b]->data().ptr<float2>());
}

/* MarchingCubes::run (marching_cubes.cpp:24); returns #vertices, copies up to cap float4 vertices/normals */
int ref_marching_cubes(void *h, int which, void *verts_host, void *normals_host, int cap) {
    Ref *r = (Ref *)h;
    if (!r->mc) {
        r->mc = cv::Ptr<kfusion::cuda::MarchingCubes>(new kfusion::cuda::MarchingCubes());
        r->mc->setPose(r->params.volume_pose);
    }
    kfusion::cuda::DeviceArray<pcl::PointXYZ> vb;
    kfusion::cuda::DeviceArray<pcl::Normal> nb;
    kfusion::cuda::Surface s = r->mc->run(*r->vol[which], vb, nb);
    cudaDeviceSynchronize();
    int n = (int)s.vertices.size();
    int m = n < cap ? n : cap;
    if (m > 0 && verts_host) cudaMemcpy(verts_host, s.vertices.ptr(), (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost);
    if (m > 0 && normals_host) cudaMemcpy(normals_host, (const float4 *)s.normals.ptr(), (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost);
    return n;
}

}  // extern "C"
