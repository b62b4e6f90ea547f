// dropin_test.cpp -- exercises the C++ drop-in headers (include/sobfu/*.hpp, include/kfusion/**) the way the reference's
// callers do: the per-frame pipeline of src/sobfu/sob_fusion.cpp through SobFusion::operator(), the field/differentiator
// entry points of test/deformation_field_test.cpp, and mesh extraction.  Written for this repo (the reference's own
// test/*.cpp are additionally compiled against the same headers by oracle/build_dropin_tests.sh where the reference
// checkout is available).
#include <gtest/gtest.h>

#include <kfusion/cuda/imgproc.hpp>
#include <kfusion/cuda/marching_cubes.hpp>
#include <kfusion/internal.hpp>
#include <kfusion/precomp.hpp>
#include <sobfu/params.hpp>
#include <sobfu/sob_fusion.hpp>
#include <sobfu/solver.hpp>

#include <cmath>
#include <vector>

namespace {
std::vector<unsigned short> sphere_depth(int cols, int rows, float fx, float fy, float cx, float cy, float sx, float r) {
    std::vector<unsigned short> d((size_t)cols * rows, 0);
    for (int v = 0; v < rows; ++v)
        for (int u = 0; u < cols; ++u) {
            const double dx = (u - cx) / fx, dy = (v - cy) / fy, cz = 0.5;
            const double a = dx * dx + dy * dy + 1.0, b = -2.0 * (dx * sx + cz), c = sx * sx + cz * cz - (double)r * r;
            const double disc = b * b - 4 * a * c;
            if (disc > 0) d[(size_t)v * cols + u] = (unsigned short)std::lround((-b - std::sqrt(disc)) / (2 * a) * 1000.0);
        }
    return d;
}
}  // namespace

// compile-time check of the device-level helpers of include/sobfu/solver.hpp (never called: the stages are exercised through the
// C ABI by the Python parity tests)
static void device_level_api_compiles(kfusion::device::TsdfVolume &a, kfusion::device::TsdfVolume &b, sobfu::device::DeformationField &psi,
                                      sobfu::device::TsdfGradient &g, sobfu::device::Laplacian &L, sobfu::device::PotentialGradient &u,
                                      sobfu::device::PotentialGradient &us, sobfu::device::Jacobian &J, float4 *updates, const float *taps) {
    SolverParams sp{0, 1, 7, 1e-3f, 0.1f, 0.1f, 0.2f};
    SDFs sdfs(a, b, a, b);
    sobfu::device::TsdfDifferentiator td(a);
    sobfu::device::Differentiator d(psi), di(psi);
    sobfu::device::SecondOrderDifferentiator sd(psi);
    Differentiators diffs(td, d, di, sd);
    sobfu::device::SpatialGradients sg(&g, &g, &J, &J, &L, &L, &u, &us);
    sobfu::device::calculate_potential_gradient(sdfs.phi_n_psi, sdfs.phi_global, *sg.nabla_phi_n_o_psi, *sg.L, *sg.nabla_U, sp.w_reg);
    sobfu::device::convolve_sobolev(*sg.nabla_U_S, *sg.nabla_U, taps);
    sobfu::device::update_psi(psi, *sg.nabla_U_S, updates, sp.alpha);
    (void)diffs;
}

class DropInTest : public ::testing::Test {
protected:
    void SetUp() override {
        params.cols = 160; params.rows = 120;
        params.volume_dims = cv::Vec3i::all(48);
        params.volume_size = cv::Vec3f::all(0.75f);
        params.volume_pose = cv::Affine3f().translate(cv::Vec3f(-0.375f, -0.375f, 0.1f));
        params.intr = kfusion::Intr(142.6f, 142.6f, 80.f, 60.f);
        params.icp_truncate_depth_dist = 1.f;
        params.bilateral_sigma_depth = 0.005f; params.bilateral_sigma_spatial = 4.5f; params.bilateral_kernel_size = 7;
        params.tsdf_trunc_dist = 6.f * params.voxel_sizes()[0];
        params.eta = 3.f * params.voxel_sizes()[0];
        params.tsdf_max_weight = 128.f;
        params.gradient_delta_factor = 0.5f;
        params.start_frame = 1; params.verbosity = 1;
        params.s = 7; params.max_iter = 20; params.max_update_norm = 1e-10f; params.lambda = 0.1f; params.alpha = 0.05f; params.w_reg = 0.6f;
    }
    Params params;
};

TEST_F(DropInTest, FramePipelineReducesTheDataEnergy) {
    SobFusion fusion(params);
    kfusion::cuda::Depth depth;
    for (int f = 0; f < 3; ++f) {
        const std::vector<unsigned short> h = sphere_depth(params.cols, params.rows, params.intr.fx, params.intr.fy, params.intr.cx, params.intr.cy, 0.004f * f, 0.15f);
        depth.upload(h.data(), params.cols * sizeof(unsigned short), params.rows, params.cols);
        ASSERT_TRUE(fusion(depth));
    }
    ASSERT_EQ(fusion.solver->info.iters, 20);
    ASSERT_EQ(fusion.getDeformationField()->get_no_nans(), 0);
    // the warped live frame is closer to the model than the live frame itself
    const int3 dims = kfusion::device_cast<int3>(params.volume_dims);
    sobfu::device::Reductor red(dims, params.voxel_sizes()[0], params.tsdf_trunc_dist);
    const float before = red.data_energy(fusion.phi_global->data().ptr<float2>(), fusion.phi_n->data().ptr<float2>());
    const float after = red.data_energy(fusion.phi_global->data().ptr<float2>(), fusion.phi_n_psi->data().ptr<float2>());
    ASSERT_LT(after, before);
    pcl::PolygonMesh::Ptr mesh = fusion.get_phi_global_mesh();
    ASSERT_GT(mesh->polygons.size(), (size_t)500);
    ASSERT_EQ(mesh->cloud.width, (unsigned)(3 * mesh->polygons.size()));
}

TEST_F(DropInTest, IdentityFieldAndDifferentiators) {
    const int n = 48 * 48 * 48;
    const int3 dims = make_int3(48, 48, 48);
    sobfu::cuda::DeformationField psi(params.volume_dims);
    std::vector<float4> h(n);
    psi.get_data().download(h.data());
    ASSERT_NEAR(h[5 + 48 * (7 + 48 * 9)].x, 5.f, 1e-5f);
    ASSERT_NEAR(h[5 + 48 * (7 + 48 * 9)].y, 7.f, 1e-5f);
    ASSERT_NEAR(h[5 + 48 * (7 + 48 * 9)].z, 9.f, 1e-5f);
    kfusion::cuda::CudaData J_data;
    J_data.create((size_t)n * sizeof(Mat4f));
    sobfu::device::DeformationField psi_device(psi.get_data().ptr<float4>(), dims);
    sobfu::device::Jacobian J(J_data.ptr<Mat4f>(), dims);
    sobfu::device::Differentiator diff(psi_device);
    diff.calculate(J);
    std::vector<Mat4f> hj(n);
    J_data.download(hj.data());
    const Mat4f m = hj[20 + 48 * (21 + 48 * 22)];
    ASSERT_NEAR(m.data[0].x, 1.f, 1e-5f); ASSERT_NEAR(m.data[1].y, 1.f, 1e-5f); ASSERT_NEAR(m.data[2].z, 1.f, 1e-5f);
    ASSERT_NEAR(m.data[0].y, 0.f, 1e-5f); ASSERT_NEAR(m.data[1].z, 0.f, 1e-5f); ASSERT_NEAR(m.data[2].x, 0.f, 1e-5f);
}

int main(int argc, char **argv) {
    ::testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
