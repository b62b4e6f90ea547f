"""Pins of the CPU oracle that need no GPU:
  (a) the reference's own asserting gtest cases (test/deformation_field_test.cpp, test/reductions_test.cpp), restated
      against the oracle at the fixtures' native 64^3 / 0.25 m;
  (b) the golden vectors dumped from the reference's own CUDA (tests/golden/*.npz, made by oracle/make_golden.py):
      the oracle must reproduce them bit for bit;
  (c) domain properties of the restated operators."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.common import assert_bits, f32, random_field, sphere_pair, wavy_psi

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
D64 = (64, 64, 64)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- (a) the reference's gtest assertions ----------------------------------------------------------------------
def test_clear_is_identity_with_x_fastest_layout():
    """DeformationFieldTest.ClearTest (deformation_field_test.cpp:92-108): at(k, j, i) == (i, j, k)"""
    psi = orc.init_identity(*D64)
    k, j, i = np.meshgrid(np.arange(64), np.arange(64), np.arange(64), indexing="ij")
    assert np.abs(psi[..., 0] - i).max() < 1e-5 and np.abs(psi[..., 1] - j).max() < 1e-5 and np.abs(psi[..., 2] - k).max() < 1e-5
    assert psi.reshape(-1, 4)[5 + 64 * (7 + 64 * 9)].tolist() == [5.0, 7.0, 9.0, 0.0]


def test_tsdf_gradient_magnitude_of_a_sphere():
    """DeformationFieldTest.TsdfGradientTest (:111-149): |grad| ~ voxel / trunc = 0.1 on non-truncated interior voxels"""
    vs = f32(0.25) / f32(64)
    phi = orc.tsdf_init_sphere(D64, (vs,) * 3, 10 * vs, 2 * vs, (0.125, 0.125, 0.125), 0.05)
    g = orc.tsdf_gradient(phi)
    t = phi[..., 0]
    ok = np.ones(t.shape, bool)
    for ax in range(3):          # non-truncated voxel whose six neighbours are not truncated either
        for sh in (1, -1):
            ok &= np.abs(np.roll(t, sh, ax)) < 1.0
    ok &= np.abs(t) < 1.0
    ok[[0, -1], :, :] = ok[:, [0, -1], :] = ok[:, :, [0, -1]] = False
    n = np.linalg.norm(g[..., :3], axis=-1)[ok]
    assert n.size > 1000 and np.abs(n - 0.1).max() < 0.15 and abs(float(n.mean()) - 0.1) < 0.01


def test_jacobian_of_uniform_and_identity_fields():
    """UniformFieldJacobianTest (:152-196) and JacobianTestSimple (:199-249)"""
    psi = np.zeros((64, 64, 64, 4), f32)
    psi[..., :3] = (1.5, -2.0, 0.25)
    assert np.abs(orc.jacobian(psi, 0)[..., :3, :3]).max() < 1e-5
    J = orc.jacobian(orc.init_identity(*D64), 0)
    inner = J[1:-1, 1:-1, 1:-1, :3, :3]
    assert np.abs(inner - np.eye(3, dtype=f32)).max() < 1e-5
    assert np.abs(orc.jacobian(orc.init_identity(*D64), 1)[..., :3, :3]).max() < 1e-5   # displacement of identity is 0


def test_jacobian_and_negative_laplacian_of_an_analytic_field():
    """JacobianLaplacianTestComplicated (:252-336): psi = (x(1-y), exp(-z)+y, z) in voxel units scaled to [0,1]"""
    n = 64
    k, j, i = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    h = 1.0 / n
    x, y, z = i * h, j * h, k * h
    psi = np.zeros((n, n, n, 4), f32)
    psi[..., 0], psi[..., 1], psi[..., 2] = x * (1 - y), np.exp(-z) + y, z
    J = orc.jacobian(psi, 0)[2:-2, 2:-2, 2:-2] / h          # central differences are per voxel
    xs, ys, zs = x[2:-2, 2:-2, 2:-2], y[2:-2, 2:-2, 2:-2], z[2:-2, 2:-2, 2:-2]
    assert np.abs(J[..., 0, 0] - (1 - ys)).max() < 0.1 and np.abs(J[..., 0, 1] + xs).max() < 0.1
    assert np.abs(J[..., 1, 1] - 1).max() < 0.1 and np.abs(J[..., 1, 2] + np.exp(-zs)).max() < 0.1
    assert np.abs(J[..., 2, 2] - 1).max() < 0.1
    L = orc.laplacian(psi)[2:-2, 2:-2, 2:-2] / (h * h)      # L is the NEGATIVE laplacian (vector_fields.cu:335)
    assert np.abs(L[..., 0]).max() < 0.1 and np.abs(L[..., 1] + np.exp(-zs)).max() < 0.1 and np.abs(L[..., 2]).max() < 0.1


def test_data_energy_of_constant_volumes():
    """ReductionsTest.DataTermTest (reductions_test.cpp:86-100): phi_global = 1, phi_n = 0 -> 0.5 * 64^3 = 131072"""
    a = np.zeros((64, 64, 64, 2), f32)
    a[..., 0] = 1
    b = np.zeros_like(a)
    assert abs(orc.data_energy(a, b) - 131072.0) < 0.1


# ---- (b) golden vectors from the reference CUDA -------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_the_reference_cuda(path):
    g = np.load(path)
    dims = tuple(int(v) for v in g["dims"])
    psi0 = wavy_psi(dims, amp=float(g["psi0_wavy"])) if float(g["psi0_wavy"]) else orc.init_identity(*dims)
    o = orc.estimate_psi(g["phi_global"], g["phi_n"], psi0, int(g["iters"]), -1.0, 7, float(g["lam"]), float(g["alpha"]), float(g["w_reg"]))
    for k in ("psi", "psi_inv", "phi_n_psi", "phi_global_psi_inv"):
        assert sha(o[k]) == str(g["sha_" + k]), k
        if bool(g["full"]):
            assert_bits(o[k], g[k], k)
    assert sha(orc.tsdf_gradient(g["phi_n"])) == str(g["sha_grad_n"])
    assert sha(orc.laplacian(o["psi"])) == str(g["sha_lap"])
    assert sha(orc.jacobian(o["psi"], 0)[..., :3, :]) == str(g["sha_jac0"])
    assert sha(orc.jacobian(o["psi"], 1)[..., :3, :]) == str(g["sha_jac1"])
    assert orc.data_energy(g["phi_global"], o["phi_n_psi"]) == float(g["e_data"])     # same reduction tree, same bits


def test_golden_files_exist():
    assert len(GOLDEN) >= 4


# ---- (c) properties ---------------------------------------------------------------------------------------------
def test_filter_preserves_constants_and_is_linear_in_exact_cases():
    dims = (12, 10, 9)
    taps = orc.sobolev_taps(7, 0.1)
    c = np.zeros((9, 10, 12, 4), f32)
    c[..., :3] = (0.5, -2.0, 8.0)
    out = orc.sobolev_filter(c, taps)      # sum of three unit-sum 1-D filters -> 3x the constant (to rounding)
    assert np.allclose(out[..., :3], 3 * c[..., :3], rtol=1e-6)
    f = random_field(dims, 1)
    assert_bits(orc.sobolev_filter(2 * f, taps), 2 * orc.sobolev_filter(f, taps), "scaling by 2 commutes exactly")
    assert np.all(out[..., 3] == 0)


def test_identity_warp_and_inverse_are_fixed_points():
    dims = (14, 11, 9)
    pg, pn, vs, trunc, eta = sphere_pair(dims)
    ident = orc.init_identity(*dims)
    w = orc.apply(pn, ident)
    assert_bits(w, pn, "warping by the identity returns the volume")
    assert_bits(orc.estimate_inverse(ident, ident.copy(), 48), ident, "inverse of the identity")
    psi = wavy_psi(dims, amp=0.3)
    inv = orc.estimate_inverse(psi, ident.copy(), 48)
    # psi(psi_inv(x)) ~ x away from the border
    Z, Y, X = psi.shape[:3]
    comp = np.stack([orc.apply(np.stack([psi[..., c], np.ones_like(psi[..., c])], -1), inv)[..., 0] for c in range(3)], -1)
    assert np.abs(comp - ident[..., :3])[2:-2, 2:-2, 2:-2].max() < 2e-3


def test_solver_reduces_the_data_energy():
    dims = (24, 24, 24)
    pg, pn, vs, trunc, eta = sphere_pair(dims)
    r = orc.estimate_psi(pg, pn, orc.init_identity(*dims), 40, -1.0, 7, 0.1, 0.05, 0.2, log_energies=2)
    e = r["log"][:, 2]
    assert e[-1] < 0.8 * e[0] and np.all(np.diff(r["log"][:, 0]) <= 1e-6)   # energy falls, update norms decay
    assert r["iters"] == 40 and r["converged"] == 0


def test_max_update_norm_is_round_down_sqrt():
    for x in (2.0, 3.0, 1e-12, 0.3, 12345.678):
        r = orc.lib().orc_sqrt_rd(f32(x))
        assert np.float64(r) ** 2 <= np.float64(f32(x)) < np.float64(np.nextafter(f32(r), f32(np.inf))) ** 2


def test_primitive_signed_distance_fields():
    """oracle restatement of tsdf_volume.cu:181-247, 277-334 against the closed forms they implement (double precision)"""
    dims, size, trunc = (24, 20, 28), (0.48, 0.40, 0.56), 0.05
    vs = tuple(f32(size[i]) / f32(dims[i]) for i in range(3))
    z, y, x = np.meshgrid(*[(np.arange(dims[k], dtype=np.float64) + 0.5) * float(vs[k]) for k in (2, 1, 0)], indexing="ij")
    cx, cy, cz = [dims[k] / 2.0 * float(vs[k]) for k in range(3)]
    X, Y, Z = x - cx, y - cy, z - cz
    clamp = lambda sdf: np.clip(sdf / trunc, -1.0, 1.0)  # noqa: E731
    b = (0.1, 0.08, 0.12)
    d = np.stack([np.abs(X) - b[0], np.abs(Y) - b[1], np.abs(Z) - b[2]])
    box = np.minimum(d.max(0), 0) + np.linalg.norm(np.maximum(d, 0), axis=0)
    r = (0.15, 0.1, 0.2)
    k0 = np.sqrt((X / r[0]) ** 2 + (Y / r[1]) ** 2 + (Z / r[2]) ** 2)
    k1 = np.sqrt((X / r[0] ** 2) ** 2 + (Y / r[1] ** 2) ** 2 + (Z / r[2] ** 2) ** 2)
    ell = k0 * (k0 - 1) / k1
    t = (0.12, 0.04)
    tor = np.sqrt((np.sqrt(X * X + Z * Z) - t[0]) ** 2 + Y * Y) - t[1]
    for shape, prm, sdf in (("box", b, box), ("ellipsoid", r, ell), ("plane", (0.2,), z - 0.2), ("torus", t, tor)):
        vol = orc.tsdf_init_shape(dims, vs, f32(trunc), shape, prm)
        assert (vol[..., 1] == 1).all(), shape
        assert np.abs(vol[..., 0] - clamp(sdf)).max() < 2e-5, shape


@pytest.mark.parametrize("s", [3, 7, 9, 11])
def test_filter_of_any_odd_length_against_a_numpy_restatement(s):
    """orc_sobolev_filter_r (2 * R + 1 taps; the reference compiles R = 3 only, solver.cu:211) against an independent numpy float32
    restatement of the same three sweeps: taps S[R - j] for j = -R..R accumulated from 0 with un-fused multiply and add, clamp to edge,
    (x + y) + z.  For s = 7 this is the filter the golden vectors pin against the reference CUDA."""
    dims = (13, 11, 9)
    src = random_field(dims, seed=40 + s)
    src[np.abs(src) < 1e-30] = 0                      # nothing denormal: numpy does not flush to zero
    taps = orc.sobolev_taps(s, 0.1)
    R = s // 2
    X, Y, Z = dims
    want = np.zeros_like(src)
    sums = []
    for axis, n in ((2, X), (1, Y), (0, Z)):          # array axes of [Z, Y, X, C]: x is axis 2
        acc = np.zeros(src.shape[:3] + (3,), dtype=f32)
        idx = np.arange(n)
        for j in range(-R, R + 1):
            sel = np.clip(idx + j, 0, n - 1)
            moved = np.take(src[..., :3], sel, axis=axis)
            acc = (acc + (moved * f32(taps[R - j])).astype(f32)).astype(f32)
        sums.append(acc)
    want[..., :3] = ((sums[0] + sums[1]).astype(f32) + sums[2]).astype(f32)
    got = orc.sobolev_filter(src, taps)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
