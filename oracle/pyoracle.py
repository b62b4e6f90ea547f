"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) and of the reference CUDA build
(oracle/_ref/libsobfu_ref.so).  TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
REF = os.path.join(HERE, "_ref", "libsobfu_ref.so")
REF_MCSYNC = os.path.join(HERE, "_ref", "libsobfu_ref_mcsync.so")   # + the __syncwarp the reference's MC compaction lacks

_P, _I, _F = C.c_void_p, C.c_int, C.c_float
_FP = C.POINTER(C.c_float)


class SolveResult(C.Structure):
    _fields_ = [("iters", C.c_int), ("max_norm", C.c_float), ("max_idx", C.c_float), ("converged", C.c_int)]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "sobfu_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "_build/liboracle.so"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_sqrt_rd.restype = C.c_float
        L.orc_sqrt_rd.argtypes = [C.c_float]
        L.orc_data_energy.restype = C.c_float
        L.orc_reg_energy.restype = C.c_float
        _lib = L
    return _lib


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def dims_of(a):
    Z, Y, X = a.shape[:3]
    return X, Y, Z


def sobolev_taps(s, lam):
    t = np.zeros(16, dtype=np.float32)
    rc = lib().orc_sobolev_taps(int(s), C.c_float(lam), _p(t))
    if rc != 0:
        raise ValueError("no taps for s=%d lambda=%g" % (s, lam))
    return t[:s].copy()


def sobolev_taps_computed(s, lam):
    """Restatement of what the reference's unused get_3d_sobolev_filter builds (solver.cpp:107-158): (Id - lambda * L) S = delta on an
    s^3 grid with the 7-point Laplacian truncated at the faces, solved densely; then the separation the tables of
    decompose_sobolev_filter (solver.cpp:160-251) come from -- first left singular vector of the s x s^2 unfolding -- and the
    unit-sum normalisation in fp32, left to right (solver.cpp:253-261).  numpy.linalg in double: independent of the product's
    conjugate-gradient / power-iteration implementation."""
    n = s ** 3
    A = np.zeros((n, n))
    for i in range(n):
        z = i // (s * s)
        y = (i - z * s * s) // s
        x = i - s * (y + s * z)
        A[i, i] = 1.0 + 6.0 * float(np.float32(lam))
        for dx, dy, dz in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
            xx, yy, zz = x + dx, y + dy, z + dz
            if 0 <= xx < s and 0 <= yy < s and 0 <= zz < s:
                A[i, xx + s * (yy + s * zz)] = -float(np.float32(lam))
    v = np.zeros(n)
    v[n // 2] = 1.0
    S = np.linalg.solve(A, v).reshape(s, s * s)
    u = np.linalg.svd(S)[0][:, 0]
    u = u * np.sign(u[s // 2])
    raw = u.astype(np.float32)
    total = np.float32(0)
    for t in raw:
        total = np.float32(total + t)
    return raw, (raw / total).astype(np.float32)


def init_identity(X, Y, Z):
    psi = np.empty((Z, Y, X, 4), dtype=np.float32)
    lib().orc_init_identity(_p(psi), X, Y, Z)
    return psi


def apply(phi, psi):
    X, Y, Z = dims_of(phi)
    out = np.empty_like(phi)
    lib().orc_apply(_p(phi), _p(out), _p(psi), X, Y, Z)
    return out


def tsdf_gradient(phi):
    X, Y, Z = dims_of(phi)
    g = np.empty((Z, Y, X, 4), dtype=np.float32)
    lib().orc_tsdf_gradient(_p(phi), _p(g), X, Y, Z)
    return g


def laplacian(psi):
    X, Y, Z = dims_of(psi)
    L = np.empty_like(psi)
    lib().orc_laplacian(_p(psi), _p(L), X, Y, Z)
    return L


def jacobian(psi, mode):
    X, Y, Z = dims_of(psi)
    J = np.zeros((Z, Y, X, 4, 4), dtype=np.float32)
    lib().orc_jacobian(_p(psi), _p(J), X, Y, Z, int(mode))
    return J


def potential_gradient(phi_n_psi, phi_global, grad, L, w_reg):
    out = np.empty_like(grad)
    lib().orc_potential_gradient(_p(phi_n_psi), _p(phi_global), _p(grad), _p(L), _p(out), C.c_float(w_reg), int(grad.size // 4))
    return out


def sobolev_filter(src, taps):
    X, Y, Z = dims_of(src)
    dst = np.empty_like(src)
    t = np.ascontiguousarray(taps, dtype=np.float32)
    assert t.ndim == 1 and t.size % 2 == 1
    lib().orc_sobolev_filter_r(_p(dst), _p(src), _p(t), int(t.size // 2), X, Y, Z)
    return dst


def update_psi(psi, g, alpha):
    upd = np.empty_like(psi)
    lib().orc_update_psi(_p(psi), _p(g), _p(upd), C.c_float(alpha), int(psi.size // 4))
    return upd


def max_update_norm(upd):
    v, i = C.c_float(), C.c_float()
    lib().orc_max_update_norm(_p(upd), int(upd.size // 4), C.byref(v), C.byref(i))
    return v.value, i.value


def data_energy(a, b):
    return float(lib().orc_data_energy(_p(a), _p(b), int(a.size // 2)))


def reg_energy(J):
    return float(lib().orc_reg_energy(_p(J), int(J.size // 16)))


def estimate_inverse(psi, psi_inv, iters=48):
    X, Y, Z = dims_of(psi)
    lib().orc_estimate_inverse(_p(psi), _p(psi_inv), X, Y, Z, int(iters))
    return psi_inv


def estimate_psi(phi_global, phi_n, psi, max_iter, max_update_norm, s, lam, alpha, w_reg, log_energies=0, taps=None):
    """returns dict(phi_n_psi, phi_global_psi_inv, psi (updated copy), psi_inv, result, log); taps: explicit filter taps, an odd
    number of them (then s / lam are ignored)"""
    X, Y, Z = dims_of(phi_global)
    psi = psi.copy()
    phi_n_psi = np.zeros_like(phi_n)
    pgpi = np.zeros_like(phi_global)
    psi_inv = np.zeros_like(psi)
    res = SolveResult()
    log = np.zeros((max(max_iter, 1), 4), dtype=np.float32)
    if taps is not None:
        t7 = np.ascontiguousarray(taps, dtype=np.float32)
        assert t7.ndim == 1 and t7.size % 2 == 1
        rc = lib().orc_estimate_psi_taps_r(_p(phi_global), _p(pgpi), _p(phi_n), _p(phi_n_psi), _p(psi), _p(psi_inv), X, Y, Z, int(max_iter),
                                           C.c_float(max_update_norm), _p(t7), int(t7.size // 2), C.c_float(alpha), C.c_float(w_reg),
                                           int(log_energies), C.byref(res), _p(log))
    else:
        rc = lib().orc_estimate_psi(_p(phi_global), _p(pgpi), _p(phi_n), _p(phi_n_psi), _p(psi), _p(psi_inv), X, Y, Z, int(max_iter),
                                    C.c_float(max_update_norm), int(s), C.c_float(lam), C.c_float(alpha), C.c_float(w_reg),
                                    int(log_energies), C.byref(res), _p(log))
    if rc != 0:
        raise ValueError("orc_estimate_psi failed: %d" % rc)
    return dict(phi_n_psi=phi_n_psi, phi_global_psi_inv=pgpi, psi=psi, psi_inv=psi_inv, iters=res.iters, max_norm=res.max_norm,
                max_idx=res.max_idx, converged=res.converged, log=log[:res.iters])


def solver_iteration(phi_global, phi_n, phi_n_psi, psi, scratch, taps, alpha, w_reg):
    X, Y, Z = dims_of(phi_global)
    v, i = C.c_float(), C.c_float()
    lib().orc_solver_iteration(_p(phi_global), _p(phi_n), _p(phi_n_psi), _p(psi), _p(scratch), _p(taps), C.c_float(alpha),
                               C.c_float(w_reg), X, Y, Z, C.byref(v), C.byref(i))
    return v.value, i.value


def tsdf_init_sphere(dims, voxel, trunc, eta, centre, radius):
    X, Y, Z = dims
    vol = np.zeros((Z, Y, X, 2), dtype=np.float32)
    lib().orc_tsdf_init_sphere(_p(vol), X, Y, Z, C.c_float(voxel[0]), C.c_float(voxel[1]), C.c_float(voxel[2]), C.c_float(trunc),
                               C.c_float(eta), C.c_float(centre[0]), C.c_float(centre[1]), C.c_float(centre[2]), C.c_float(radius))
    return vol


SHAPES = {"box": 0, "ellipsoid": 1, "plane": 2, "torus": 3}


def tsdf_init_shape(dims, voxel, trunc, shape, prm):
    """TsdfVolume::initBox / initEllipsoid / initPlane / initTorus (tsdf_volume.cu:181-247, 277-334)"""
    X, Y, Z = dims
    vol = np.zeros((Z, Y, X, 2), dtype=np.float32)
    q = list(prm) + [0.0, 0.0, 0.0]
    lib().orc_tsdf_init_shape(_p(vol), X, Y, Z, C.c_float(voxel[0]), C.c_float(voxel[1]), C.c_float(voxel[2]), C.c_float(trunc),
                              SHAPES[shape], C.c_float(q[0]), C.c_float(q[1]), C.c_float(q[2]))
    return vol


def tsdf_fuse(pg, pn, max_weight):
    lib().orc_tsdf_fuse(_p(pg), _p(pn), int(pg.size // 2), C.c_float(max_weight))
    return pg


def tsdf_integrate(dists, vol, voxel, trunc, eta, R, t, fx, fy, cx, cy):
    X, Y, Z = dims_of(vol)
    rows, cols = dists.shape
    R = np.ascontiguousarray(R, dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(t, dtype=np.float32)
    lib().orc_tsdf_integrate(_p(dists), cols, rows, _p(vol), X, Y, Z, C.c_float(voxel[0]), C.c_float(voxel[1]), C.c_float(voxel[2]),
                             C.c_float(trunc), C.c_float(eta), _p(R), _p(t), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy))
    return vol


def bilateral(depth, ksz, sigma_spatial, sigma_depth):
    rows, cols = depth.shape
    out = np.empty_like(depth)
    lib().orc_bilateral(_p(depth), _p(out), cols, rows, int(ksz), C.c_float(sigma_spatial), C.c_float(sigma_depth))
    return out


def truncate_depth(depth, max_dist):
    rows, cols = depth.shape
    lib().orc_truncate_depth(_p(depth), cols, rows, C.c_float(max_dist))
    return depth


def compute_dists(depth, fx, fy, cx, cy):
    rows, cols = depth.shape
    d = np.empty((rows, cols), dtype=np.float32)
    lib().orc_compute_dists(_p(depth), _p(d), cols, rows, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy))
    return d


# ---- the reference's own CUDA (GPU box only) -------------------------------------------------------------------
class Reference:
    """Handle on oracle/_ref/libsobfu_ref.so: the unmodified reference driven through its public host API.
    lib=REF_MCSYNC: the build whose marching-cubes compaction has the __syncwarp the reference lacks (oracle/patch_textures.py)."""

    def __init__(self, dims, size, trunc, eta, max_weight, verbosity, max_iter, s, max_update_norm, lam, alpha, w_reg,
                 pose_t=(0, 0, 0), intr=(1, 1, 0, 0), lib=None):
        lib = lib or REF
        if not os.path.exists(lib):
            raise FileNotFoundError(lib)
        L = C.CDLL(lib)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [_I] * 3 + [_F] * 6 + [_I] * 3 + [_F] * 4 + [_F] * 3 + [_F] * 4
        L.ref_estimate_psi.restype = C.c_float
        L.ref_data_energy.restype = C.c_float
        for n in ("ref_destroy", "ref_estimate_psi", "ref_get_inverse"):
            getattr(L, n).argtypes = [_P]
        for n in ("ref_upload_tsdf", "ref_download_tsdf", "ref_upload_psi", "ref_download_psi"):
            getattr(L, n).argtypes = [_P, _I, _P]
        for n in ("ref_psi_clear", "ref_tsdf_clear", "ref_integrate_dists"):
            getattr(L, n).argtypes = [_P, _I]
        L.ref_init_sphere.argtypes = [_P, _I, _F, _F, _F, _F]
        L.ref_apply.argtypes = [_P, _I, _I, _I]
        L.ref_fuse.argtypes = [_P, _I, _I]
        L.ref_depth_to_dists.argtypes = [_P, _P, _I, _I, _I, _F, _F, _F, _P, _P]
        L.ref_tsdf_gradient.argtypes = [_P, _I, _P]
        L.ref_laplacian.argtypes = [_P, _P]
        L.ref_jacobian.argtypes = [_P, _I, _P]
        L.ref_data_energy.argtypes = [_P, _I, _I]
        L.ref_marching_cubes.argtypes = [_P, _I, _P, _P, _I]
        self.L = L
        self.dims = tuple(dims)
        self.h = L.ref_create(dims[0], dims[1], dims[2], size[0], size[1], size[2], trunc, eta, max_weight, verbosity, max_iter, s,
                              max_update_norm, lam, alpha, w_reg, pose_t[0], pose_t[1], pose_t[2], intr[0], intr[1], intr[2], intr[3])

    GLOBAL, GLOBAL_PSI_INV, N, N_PSI = 0, 1, 2, 3

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def _vol(self, c):
        X, Y, Z = self.dims
        return np.empty((Z, Y, X, c), dtype=np.float32)

    def upload_tsdf(self, which, a): self.L.ref_upload_tsdf(self.h, which, _p(a))
    def upload_psi(self, which, a): self.L.ref_upload_psi(self.h, which, _p(a))

    def download_tsdf(self, which):
        a = self._vol(2)
        self.L.ref_download_tsdf(self.h, which, _p(a))
        return a

    def download_psi(self, which):
        a = self._vol(4)
        self.L.ref_download_psi(self.h, which, _p(a))
        return a

    def psi_clear(self, which=0): self.L.ref_psi_clear(self.h, which)
    def tsdf_clear(self, which): self.L.ref_tsdf_clear(self.h, which)
    def init_sphere(self, which, c, r): self.L.ref_init_sphere(self.h, which, c[0], c[1], c[2], r)

    def init_shape(self, which, shape, prm):
        q = list(prm) + [0.0, 0.0, 0.0]
        self.L.ref_init_shape.argtypes = [_P, _I, _I, _F, _F, _F]
        self.L.ref_init_shape(self.h, which, SHAPES[shape], q[0], q[1], q[2])
    def estimate_psi(self): return float(self.L.ref_estimate_psi(self.h))
    def apply(self, which_psi, src, dst): self.L.ref_apply(self.h, which_psi, src, dst)
    def get_inverse(self): self.L.ref_get_inverse(self.h)
    def fuse(self, dst, src): self.L.ref_fuse(self.h, dst, src)

    def depth_to_dists(self, depth, ksz, sigma_spatial, sigma_depth, trunc_depth):
        rows, cols = depth.shape
        f = np.empty_like(depth)
        d = np.empty((rows, cols), dtype=np.float32)
        self.L.ref_depth_to_dists(self.h, _p(depth), cols, rows, ksz, sigma_spatial, sigma_depth, trunc_depth, _p(f), _p(d))
        return f, d

    def integrate_dists(self, which): self.L.ref_integrate_dists(self.h, which)

    def tsdf_gradient(self, which):
        g = self._vol(4)
        self.L.ref_tsdf_gradient(self.h, which, _p(g))
        return g

    def laplacian(self):
        g = self._vol(4)
        self.L.ref_laplacian(self.h, _p(g))
        return g

    def jacobian(self, mode):
        X, Y, Z = self.dims
        J = np.empty((Z, Y, X, 4, 4), dtype=np.float32)
        self.L.ref_jacobian(self.h, mode, _p(J))
        return J

    def data_energy(self, a, b): return float(self.L.ref_data_energy(self.h, a, b))

    def marching_cubes(self, which, cap=6000000):
        v = np.empty((cap, 4), dtype=np.float32)
        n = np.empty((cap, 4), dtype=np.float32)
        k = self.L.ref_marching_cubes(self.h, which, _p(v), _p(n), cap)
        k = min(k, cap)
        return v[:k].copy(), n[:k].copy()


def _ref_mc_count(self, which):
    """marching cubes through the reference's own class, result left on the device (vertex count only)"""
    return int(self.L.ref_marching_cubes(self.h, which, None, None, 0))


Reference.marching_cubes_count = _ref_mc_count


def mc_tables():
    lib().orc_mc_num_verts_table.restype = C.POINTER(C.c_int)
    lib().orc_mc_tri_table.restype = C.POINTER(C.c_int)
    nv = np.ctypeslib.as_array(lib().orc_mc_num_verts_table(), shape=(256,)).copy()
    tri = np.ctypeslib.as_array(lib().orc_mc_tri_table(), shape=(256, 16)).copy()
    return nv, tri


def mc_occupied(vol, cap=None):
    X, Y, Z = dims_of(vol)
    cap = cap or X * Y * Z
    v, c, n = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    k = lib().orc_mc_occupied(_p(vol), X, Y, Z, _p(v), _p(c), _p(n), cap)
    k = min(k, cap)
    return v[:k], c[:k], n[:k]


def mc_triangles(vol, size, R, t, voxel_idx, nverts_total):
    X, Y, Z = dims_of(vol)
    verts = np.zeros((max(nverts_total, 1), 4), dtype=np.float32)
    normals = np.zeros_like(verts)
    R = np.ascontiguousarray(R, dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(t, dtype=np.float32)
    vi = np.ascontiguousarray(voxel_idx, dtype=np.int32)
    k = lib().orc_mc_triangles(_p(vol), X, Y, Z, C.c_float(size[0]), C.c_float(size[1]), C.c_float(size[2]), _p(R), _p(t), _p(vi), len(vi),
                               _p(verts), _p(normals), len(verts))
    return verts[:k], normals[:k]
