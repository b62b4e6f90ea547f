"""Per-frame kernels around the solver: TSDF clear / analytic sphere / projective integration /
fusion and the depth-image preparation.  Against the oracle with a stated tolerance (these kernels use the GPU's
approximate units exactly like the reference: __fdividef, __expf, sqrtf under --prec-sqrt=false), and -- where
oracle/_ref is built -- against the reference's own CUDA bit for bit, including a whole frame sequence through
SobFusion::operator()."""
import os

import numpy as np
import pytest

from tests.common import assert_bits, f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(built):
    import torch
    import sobfu_b200 as sf
    from oracle import pyoracle as orc
    return sf, orc, torch


def synth_depth(cols, rows, fx, fy, cx, cy, centre, radius, noise_seed=None):
    u, v = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    dx, dy = (u - cx) / fx, (v - cy) / fy
    c = np.asarray(centre, dtype=np.float64)
    a = dx * dx + dy * dy + 1.0
    b = -2.0 * (dx * c[0] + dy * c[1] + c[2])
    disc = b * b - 4 * a * (float(c @ c) - radius * radius)
    t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0.0)
    d = np.where(disc > 0, np.round(t * 1000.0), 0)
    if noise_seed is not None:
        d = np.where(d > 0, d + np.random.RandomState(noise_seed).randint(-2, 3, d.shape), 0)
    return np.ascontiguousarray(d.astype(np.uint16))


CAM = dict(cols=160, rows=120, fx=142.6, fy=142.6, cx=80.0, cy=60.0)


def make_params(sf, dims=(48, 48, 48), iters=6):
    size = 0.75
    p = sf.Params(cols=CAM["cols"], rows=CAM["rows"], volume_dims=dims, volume_size=(size,) * 3,
                  intr=sf.Intr(CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]), icp_truncate_depth_dist=1.0, bilateral_sigma_depth=0.005,
                  bilateral_sigma_spatial=4.5, bilateral_kernel_size=7, tsdf_max_weight=128.0, start_frame=2, s=7, max_iter=iters,
                  max_update_norm=1e-10, lambda_=0.1, alpha=0.05, w_reg=0.6)
    vs = p.voxel_sizes()
    p.tsdf_trunc_dist, p.eta = float(f32(6) * vs[0]), float(f32(3) * vs[0])
    p.volume_pose = sf.Affine3f().translate((-size / 2, -size / 2, 0.1))
    return p


def test_depth_ops_and_tsdf_kernels_against_the_oracle(env):
    sf, orc, torch = env
    p = make_params(sf)
    depth = synth_depth(CAM["cols"], CAM["rows"], CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"], (0.0, 0.0, 0.5), 0.15, noise_seed=3)
    d_dev = torch.from_numpy(depth.view(np.int16)).cuda().view(torch.uint16)
    filt = sf.depthBilateralFilter(d_dev, 7, 4.5, 0.005)
    o_filt = orc.bilateral(depth, 7, 4.5, 0.005)
    got = filt.cpu().view(torch.int16).numpy().view(np.uint16)
    assert np.abs(got.astype(np.int32) - o_filt.astype(np.int32)).max() <= 1           # __expf / fast division, rounded to mm
    sf.depthTruncation(filt, 0.52)
    got = filt.cpu().view(torch.int16).numpy().view(np.uint16).copy()
    assert got.max() <= 520 and np.array_equal(got, orc.truncate_depth(got.copy(), 0.52))
    dists = sf.computeDists(filt, p.intr)
    o_d = orc.compute_dists(got, CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"])
    assert np.abs(dists.cpu().numpy() - o_d).max() < 1e-6

    vol = sf.TsdfVolume(p)
    vol.integrate(dists, sf.Affine3f(), p.intr)
    vs = p.voxel_sizes()
    o_vol = np.zeros((48, 48, 48, 2), f32)
    orc.tsdf_integrate(dists.cpu().numpy(), o_vol, vs, p.tsdf_trunc_dist, p.eta, np.eye(3), p.volume_pose.t, CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"])
    g = vol.data().cpu().numpy()
    # the weight is a threshold on psdf: allow flips only where psdf sits within float noise of -eta
    # and a handful of voxels whose projection lands within float noise of a pixel border sample the neighbouring pixel
    flips = g[..., 1] != o_vol[..., 1]
    assert flips.mean() < 1e-4
    dv = np.abs(g[..., 0] - o_vol[..., 0])[~flips]
    assert (dv > 2e-5).mean() < 1e-3 and np.quantile(dv, 0.999) < 2e-5
    assert (g[..., 1] > 0).sum() > 1000

    sph = sf.TsdfVolume(p)
    sph.initSphere((0.375, 0.36, 0.4), 0.15)
    o_sph = orc.tsdf_init_sphere((48, 48, 48), vs, p.tsdf_trunc_dist, p.eta, (0.375, 0.36, 0.4), 0.15)
    gs = sph.data().cpu().numpy()
    assert (gs[..., 1] != o_sph[..., 1]).mean() < 1e-4 and np.abs(gs[..., 0] - o_sph[..., 0]).max() < 2e-5

    vol.integrate(sph)       # running-average fusion
    o_f = orc.tsdf_fuse(g.copy(), gs, 128.0)         # same inputs as the GPU (the integration above differs on a few voxels)
    gf = vol.data().cpu().numpy()
    same_w = gf[..., 1] == o_f[..., 1]
    assert same_w.mean() > 0.9999 and np.abs(gf[..., 0] - o_f[..., 0])[same_w].max() < 2e-5
    vol.clear()
    assert float(vol.data().abs().max()) == 0.0


def test_frame_kernels_bit_exact_against_the_reference_cuda(env):
    sf, orc, torch = env
    if not os.path.exists(orc.REF):
        pytest.skip("oracle/_ref not built")
    p = make_params(sf)
    dims = p.volume_dims
    ref = orc.Reference(dims, p.volume_size, p.tsdf_trunc_dist, p.eta, p.tsdf_max_weight, 0, p.max_iter, 7, p.max_update_norm, p.lambda_, p.alpha,
                        p.w_reg, pose_t=tuple(p.volume_pose.t), intr=(CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]))
    depth = synth_depth(CAM["cols"], CAM["rows"], CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"], (0.01, -0.02, 0.5), 0.15, noise_seed=5)
    r_filt, r_dists = ref.depth_to_dists(depth, 7, 4.5, 0.005, 1.0)
    d_dev = torch.from_numpy(depth.view(np.int16)).cuda().view(torch.uint16)
    filt = sf.depthBilateralFilter(d_dev, 7, 4.5, 0.005)
    sf.depthTruncation(filt, 1.0)
    dists = sf.computeDists(filt, p.intr)
    assert np.array_equal(filt.cpu().view(torch.int16).numpy().view(np.uint16), r_filt)
    assert_bits(dists.cpu().numpy(), r_dists, "dists")
    ref.integrate_dists(ref.N)
    vol = sf.TsdfVolume(p)
    vol.integrate(dists, sf.Affine3f(), p.intr)
    assert_bits(vol.data().cpu().numpy(), ref.download_tsdf(ref.N), "projective integration")
    ref.init_sphere(ref.GLOBAL, (0.375, 0.36, 0.4), 0.15)
    sph = sf.TsdfVolume(p)
    sph.initSphere((0.375, 0.36, 0.4), 0.15)
    assert_bits(sph.data().cpu().numpy(), ref.download_tsdf(ref.GLOBAL), "initSphere")
    ref.fuse(ref.GLOBAL, ref.N)
    sph.integrate(vol)
    assert_bits(sph.data().cpu().numpy(), ref.download_tsdf(ref.GLOBAL), "fusion")
    ref.close()


SHAPE_CASES = [("box", (0.21, 0.15, 0.1)), ("ellipsoid", (0.25, 0.18, 0.12)), ("plane", (0.31,)), ("torus", (0.2, 0.07))]


@pytest.mark.parametrize("shape,prm", SHAPE_CASES)
def test_primitive_initialisers(env, shape, prm):
    """TsdfVolume::initBox / initEllipsoid / initPlane / initTorus (tsdf_volume.cu:181-247, 277-334): within float tolerance of
    the oracle (approximate sqrt / division on the GPU, as in the reference) and bit-exact against the reference's CUDA"""
    sf, orc, torch = env
    p = make_params(sf, dims=(48, 40, 36))
    vol = sf.TsdfVolume(p)
    call = {"box": vol.initBox, "ellipsoid": vol.initEllipsoid, "plane": vol.initPlane, "torus": vol.initTorus}[shape]
    call(prm if shape != "plane" else prm[0])
    got = vol.data().cpu().numpy()
    want = orc.tsdf_init_shape(p.volume_dims, p.voxel_sizes(), f32(p.tsdf_trunc_dist), shape, prm)
    assert np.array_equal(got[..., 1], want[..., 1]) and (got[..., 1] == 1).all()
    assert np.abs(got[..., 0] - want[..., 0]).max() < 2e-5
    assert got[..., 0].min() < 0 < got[..., 0].max()                      # the surface is inside the volume
    if os.path.exists(orc.REF):
        ref = orc.Reference(p.volume_dims, p.volume_size, p.tsdf_trunc_dist, p.eta, p.tsdf_max_weight, 0, p.max_iter, 7, p.max_update_norm, p.lambda_,
                            p.alpha, p.w_reg, pose_t=tuple(p.volume_pose.t), intr=(CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]))
        ref.init_shape(ref.GLOBAL, shape, prm)
        assert_bits(got, ref.download_tsdf(ref.GLOBAL), "init " + shape)
        ref.close()


def test_frame_sequence_through_sobfusion_matches_the_reference(env):
    """SobFusion::operator() (sob_fusion.cpp:71-145) on 5 synthetic frames: init, rigid fusion, then 3 solver frames with a
    warm-started psi -- every volume and field equal to the reference's own CUDA, bit for bit"""
    sf, orc, torch = env
    if not os.path.exists(orc.REF):
        pytest.skip("oracle/_ref not built")
    p = make_params(sf, iters=8)
    ref = orc.Reference(p.volume_dims, p.volume_size, p.tsdf_trunc_dist, p.eta, p.tsdf_max_weight, 0, p.max_iter, 7, p.max_update_norm, p.lambda_,
                        p.alpha, p.w_reg, pose_t=tuple(p.volume_pose.t), intr=(CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]))
    fusion = sf.SobFusion(p)
    for f in range(5):
        depth = synth_depth(CAM["cols"], CAM["rows"], CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"], (0.004 * f, 0.0, 0.5), 0.15 + 0.002 * f)
        fusion(torch.from_numpy(depth.view(np.int16)).cuda().view(torch.uint16))
        ref.depth_to_dists(depth, 7, 4.5, 0.005, 1.0)
        if f == 0:
            ref.integrate_dists(ref.GLOBAL)
        else:
            ref.tsdf_clear(ref.N)
            ref.integrate_dists(ref.N)
            if f < p.start_frame:
                ref.fuse(ref.GLOBAL, ref.N)
            else:
                ref.estimate_psi()
                ref.fuse(ref.GLOBAL, ref.N_PSI)
        assert_bits(fusion.phi_global.data().cpu().numpy(), ref.download_tsdf(ref.GLOBAL), "frame %d phi_global" % f)
        if f >= p.start_frame:
            assert_bits(fusion.psi.get_data().cpu().numpy(), ref.download_psi(0), "frame %d psi" % f)
            assert_bits(fusion.psi_inv.get_data().cpu().numpy(), ref.download_psi(1), "frame %d psi_inv" % f)
            assert_bits(fusion.phi_n_psi.data().cpu().numpy(), ref.download_tsdf(ref.N_PSI), "frame %d phi_n_psi" % f)
            assert_bits(fusion.phi_global_psi_inv.data().cpu().numpy(), ref.download_tsdf(ref.GLOBAL_PSI_INV), "frame %d phi_global_psi_inv" % f)
    ref.close()
