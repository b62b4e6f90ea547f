"""Host-side mirror of the reference's operator interface for the solver hot path.

Same class names, method names, argument meaning and in/out contract as dgrzech/sobfu's host classes, so that the
parity tests read like the reference's own (test/solver_test.cpp, test/deformation_field_test.cpp):

    Params                       include/sobfu/params.hpp:7-38
    TsdfVolume                   include/kfusion/cuda/tsdf_volume.hpp:17-92
    DeformationField             include/sobfu/vector_fields.hpp:52-66
    Solver                       include/sobfu/solver.hpp:56-67
    MarchingCubes                include/kfusion/cuda/marching_cubes.hpp:19-56
    depthBilateralFilter / depthTruncation / computeDists   include/kfusion/cuda/imgproc.hpp:11-23
    SobFusion                    include/sobfu/sob_fusion.hpp:17-74

PyTorch is used for device memory only (torch.Tensor as the CudaData blob); every computation is a call through the
C ABI of libsobfu_b200.so.  Volumes are tensors of shape [Z, Y, X, C] (x fastest) == the reference's linear layout.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _capi
from ._capi import check, fvec

f32 = np.float32


class _StreamBoundLib:
    """The library issues the free functions on -- and orders solver calls against -- the stream set with
    sobfu_b200_set_stream (include/sobfu_b200.h).  Every function fetched through this proxy first binds the caller's CURRENT
    torch stream, so work queued under `with torch.cuda.stream(s):` is ordered with the library's kernels."""

    def __getattr__(self, name):
        L = _capi.lib()
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            L.sobfu_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return getattr(L, name)


_BOUND = _StreamBoundLib()


def lib():
    return _BOUND


@dataclass
class Intr:
    """kfusion::Intr, include/kfusion/types.hpp:28-35"""
    fx: float = 0.0
    fy: float = 0.0
    cx: float = 0.0
    cy: float = 0.0


@dataclass
class Affine3f:
    """cv::Affine3f: rotation (3x3 row-major) + translation"""
    R: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=f32))
    t: np.ndarray = field(default_factory=lambda: np.zeros(3, dtype=f32))

    def translate(self, v):
        return Affine3f(self.R.copy(), (self.t + np.asarray(v, dtype=f32)).astype(f32))

    def inv(self):
        Rt = self.R.T.astype(f32)
        return Affine3f(Rt, (-(Rt @ self.t)).astype(f32))

    def __mul__(self, o):
        return Affine3f((self.R @ o.R).astype(f32), (self.R @ o.t + self.t).astype(f32))


@dataclass
class Params:
    """sobfu Params, include/sobfu/params.hpp:7-38 (same field names; `lambda` is spelled lambda_)"""
    cols: int = 640
    rows: int = 480
    volume_dims: tuple = (64, 64, 64)
    volume_size: tuple = (0.25, 0.25, 0.25)
    volume_pose: Affine3f = field(default_factory=Affine3f)
    intr: Intr = field(default_factory=Intr)
    icp_truncate_depth_dist: float = 0.0
    bilateral_sigma_depth: float = 0.0
    bilateral_sigma_spatial: float = 0.0
    bilateral_kernel_size: int = 0
    tsdf_trunc_dist: float = 0.0
    eta: float = 0.0
    tsdf_max_weight: float = 0.0
    gradient_delta_factor: float = 0.0
    start_frame: int = 0
    verbosity: int = 0
    s: int = 7
    max_iter: int = 0
    max_update_norm: float = 0.0
    lambda_: float = 0.1
    alpha: float = 0.0
    w_reg: float = 0.0
    compute_filter: bool = False      # extension: compute the Sobolev filter for a lambda the reference does not tabulate

    def voxel_sizes(self):
        """params.hpp:34-37 -- fp32 division, as in the reference"""
        return tuple(f32(self.volume_size[i]) / f32(self.volume_dims[i]) for i in range(3))


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _dev():
    if not torch.cuda.is_available():
        raise _capi.Sobfu200Error("sobfu_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class TsdfVolume:
    """kfusion::cuda::TsdfVolume.  data(): float32 tensor [Z, Y, X, 2] = {tsdf, weight}."""

    def __init__(self, params):
        self.dims_ = tuple(int(v) for v in params.volume_dims)
        self.size_ = tuple(f32(v) for v in params.volume_size)
        self.pose_ = params.volume_pose
        self.trunc_dist_ = f32(params.tsdf_trunc_dist)
        self.eta_ = f32(params.eta)
        self.max_weight_ = f32(params.tsdf_max_weight)
        X, Y, Z = self.dims_
        self.data_ = torch.empty((Z, Y, X, 2), dtype=torch.float32, device=_dev())
        self.clear()

    def getDims(self): return self.dims_
    def getSize(self): return self.size_
    def getVoxelSize(self): return tuple(f32(self.size_[i]) / f32(self.dims_[i]) for i in range(3))
    def getTruncDist(self): return self.trunc_dist_
    def getEta(self): return self.eta_
    def getMaxWeight(self): return self.max_weight_
    def getPose(self): return self.pose_
    def setPose(self, pose): self.pose_ = pose
    def data(self): return self.data_

    def clear(self):
        X, Y, Z = self.dims_
        check(lib().sobfu_b200_tsdf_clear(_ptr(self.data_), X, Y, Z))

    def initSphere(self, centre, radius):
        X, Y, Z = self.dims_
        check(lib().sobfu_b200_tsdf_init_sphere(_ptr(self.data_), X, Y, Z, fvec(self.getVoxelSize()), self.trunc_dist_,
                                                self.eta_, fvec(centre), f32(radius)))

    def _init_shape(self, fn, prm):
        X, Y, Z = self.dims_
        check(fn(_ptr(self.data_), X, Y, Z, fvec(self.getVoxelSize()), self.trunc_dist_, prm))

    def initBox(self, b): self._init_shape(lib().sobfu_b200_tsdf_init_box, fvec(b))                    # tsdf_volume.cpp:108-114
    def initEllipsoid(self, r): self._init_shape(lib().sobfu_b200_tsdf_init_ellipsoid, fvec(r))        # :116-122
    def initPlane(self, z): self._init_shape(lib().sobfu_b200_tsdf_init_plane, f32(z))                 # :124-130
    def initTorus(self, t): self._init_shape(lib().sobfu_b200_tsdf_init_torus, fvec(t))                # :140-146

    def integrate(self, other, camera_pose=None, intr=None):
        """integrate(TsdfVolume) = running-average fusion (tsdf_volume.cpp:84-93);
        integrate(dists, camera_pose, intr) = projective integration of a ray-length image (tsdf_volume.cpp:95-108)."""
        X, Y, Z = self.dims_
        if isinstance(other, TsdfVolume):
            check(lib().sobfu_b200_tsdf_fuse(_ptr(self.data_), _ptr(other.data_), X, Y, Z, self.max_weight_))
            return
        dists = other
        assert dists.dtype == torch.float32 and dists.dim() == 2 and dists.is_contiguous()
        vol2cam = camera_pose.inv() * self.pose_
        rows, cols = dists.shape
        check(lib().sobfu_b200_tsdf_integrate(_ptr(dists), cols * 4, cols, rows, _ptr(self.data_), X, Y, Z,
                                              fvec(self.getVoxelSize()), self.trunc_dist_, self.eta_,
                                              fvec(vol2cam.R.reshape(-1)), fvec(vol2cam.t), f32(intr.fx), f32(intr.fy),
                                              f32(intr.cx), f32(intr.cy)))


class VectorField:
    """sobfu::cuda::VectorField.  get_data(): float32 tensor [Z, Y, X, 4]."""

    def __init__(self, dims):
        self.dims = tuple(int(v) for v in dims)
        X, Y, Z = self.dims
        self.data = torch.empty((Z, Y, X, 4), dtype=torch.float32, device=_dev())
        VectorField.clear(self)

    def get_dims(self): return self.dims
    def get_data(self): return self.data

    def clear(self):
        X, Y, Z = self.dims
        check(lib().sobfu_b200_clear_field(_ptr(self.data), X, Y, Z))

    def get_no_nans(self):
        return int(torch.isnan(self.data[..., :3]).any(dim=-1).sum().item())


class DeformationField(VectorField):
    """sobfu::cuda::DeformationField: psi stores absolute voxel coordinates; clear() = identity."""

    def __init__(self, dims):
        super().__init__(dims)
        self.clear()

    def clear(self):
        X, Y, Z = self.dims
        check(lib().sobfu_b200_init_identity(_ptr(self.data), X, Y, Z))

    def apply(self, phi, phi_psi):
        X, Y, Z = self.dims
        check(lib().sobfu_b200_apply(_ptr(phi.data()), _ptr(phi_psi.data()), _ptr(self.data), X, Y, Z))

    def get_inverse(self, psi_inv):
        """48 fixed-point steps starting from whatever psi_inv holds (vector_fields.cpp:95-104)."""
        X, Y, Z = self.dims
        check(lib().sobfu_b200_estimate_inverse(_ptr(self.data), _ptr(psi_inv.data), X, Y, Z, 48))


class Solver:
    """sobfu::cuda::Solver."""

    def __init__(self, params):
        self.params = params
        p = _capi.Params()
        p.dims[:] = [int(v) for v in params.volume_dims]
        p.voxel_size[:] = [float(v) for v in params.voxel_sizes()]
        p.trunc_dist, p.eta, p.max_weight = float(params.tsdf_trunc_dist), float(params.eta), float(params.tsdf_max_weight)
        p.verbosity, p.max_iter, p.s = int(params.verbosity), int(params.max_iter), int(params.s)
        p.max_update_norm, p.lambda_ = float(params.max_update_norm), float(params.lambda_)
        p.alpha, p.w_reg = float(params.alpha), float(params.w_reg)
        self._h = C.c_void_p()
        _dev()
        # compute_filter (not a field of the reference's Params): a lambda outside the reference's tables gets the computed filter
        flags = 1 if getattr(params, "compute_filter", False) else 0
        check(lib().sobfu_b200_solver_create_ex(C.byref(self._h), C.byref(p), flags))
        self.info = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                lib().sobfu_b200_solver_destroy(h)
            except Exception:
                pass
            self._h = None

    def estimate_psi(self, phi_global, phi_global_psi_inv, phi_n, phi_n_psi, psi, psi_inv):
        """reads phi_global, phi_n; overwrites phi_n_psi, phi_global_psi_inv, psi (in place, warm start), psi_inv."""
        info = _capi.SolveInfo()
        check(lib().sobfu_b200_solver_estimate_psi(self._h, _ptr(phi_global.data()), _ptr(phi_global_psi_inv.data()),
                                                   _ptr(phi_n.data()), _ptr(phi_n_psi.data()), _ptr(psi.get_data()),
                                                   _ptr(psi_inv.get_data()), C.byref(info)))
        self.info = info
        return info

    def estimate_psi_host(self, phi_global, phi_global_psi_inv, phi_n, phi_n_psi, psi, psi_inv):
        """Same call with HOST numpy arrays (C-contiguous float32); copies happen inside the library."""
        info = _capi.SolveInfo()

        def hp(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)
        check(lib().sobfu_b200_solver_estimate_psi_host(self._h, hp(phi_global), hp(phi_global_psi_inv), hp(phi_n),
                                                        hp(phi_n_psi), hp(psi), hp(psi_inv), C.byref(info)))
        self.info = info
        return info

    def get_log(self):
        n = self.info.iters if self.info is not None else 0
        buf = (_capi.IterLog * max(n, 1))()
        check(lib().sobfu_b200_solver_get_log(self._h, buf, n))
        return [(buf[i].max_norm, buf[i].max_idx_f, buf[i].e_data, buf[i].e_reg) for i in range(n)]

    def get_taps(self):
        t = (C.c_float * 11)()
        check(lib().sobfu_b200_solver_get_taps(self._h, t))
        return np.array(list(t)[:int(self.params.s)], dtype=f32)

    def set_variant(self, v):
        check(lib().sobfu_b200_solver_set_variant(self._h, int(v)))

    def time_loop(self, iters):
        a, b, l = C.c_float(), C.c_float(), C.c_float()
        check(lib().sobfu_b200_solver_time_loop(self._h, int(iters), C.byref(a), C.byref(b), C.byref(l)))
        return a.value, b.value, l.value

    def time_phases(self, iters=50):
        """slab mode over NCCL: (A_mid, wait + A_edge, wait + B_edge, B_mid, iteration) mean ms -- see include/sobfu_b200.h"""
        out = (C.c_float * 5)()
        check(lib().sobfu_b200_solver_time_phases(self._h, int(iters), out))
        return tuple(out)

    def tail_fallbacks(self):
        """slab mode: solves whose psi^-1 / final warp left the neighbour window and were repeated on all-gathered volumes"""
        return int(lib().sobfu_b200_solver_tail_fallbacks(self._h))

    def workspace_bytes(self):
        return int(lib().sobfu_b200_solver_workspace_bytes(self._h))


# ---- depth pre-processing (include/kfusion/cuda/imgproc.hpp) ---------------------------------------------------
def depthBilateralFilter(depth_in, ksz, sigma_spatial, sigma_depth):
    """uint16 (mm) image [rows, cols] -> filtered copy"""
    assert depth_in.dtype == torch.uint16 or depth_in.dtype == torch.int16
    out = torch.empty_like(depth_in)
    rows, cols = depth_in.shape
    check(lib().sobfu_b200_depth_bilateral(_ptr(depth_in), cols * 2, _ptr(out), cols * 2, cols, rows, int(ksz),
                                           f32(sigma_spatial), f32(sigma_depth)))
    return out


def depthTruncation(depth, threshold):
    rows, cols = depth.shape
    check(lib().sobfu_b200_depth_truncate(_ptr(depth), cols * 2, cols, rows, f32(threshold)))


def computeDists(depth, intr):
    rows, cols = depth.shape
    dists = torch.empty((rows, cols), dtype=torch.float32, device=depth.device)
    check(lib().sobfu_b200_compute_dists(_ptr(depth), cols * 2, _ptr(dists), cols * 4, cols, rows, f32(intr.fx), f32(intr.fy),
                                         f32(intr.cx), f32(intr.cy)))
    return dists


class MarchingCubes:
    """kfusion::cuda::MarchingCubes: run(volume) -> (vertices [n,4], normals [n,4]) float32 tensors."""
    DEFAULT_TRIANGLES_BUFFER_SIZE = 2 * 1000 * 1000 * 3

    def __init__(self):
        self.pose = Affine3f()

    def setPose(self, pose):
        self.pose = pose

    def run(self, volume, vertex_cap=None, return_occupied=False):
        cap = int(vertex_cap or self.DEFAULT_TRIANGLES_BUFFER_SIZE)
        dev = volume.data().device
        verts = torch.empty((cap, 4), dtype=torch.float32, device=dev)
        normals = torch.empty((cap, 4), dtype=torch.float32, device=dev)
        vcap = cap // 3
        occ = torch.empty((3, vcap), dtype=torch.int32, device=dev)
        nv, nvox = C.c_int(0), C.c_int(0)
        X, Y, Z = volume.getDims()
        check(lib().sobfu_b200_marching_cubes(_ptr(volume.data()), X, Y, Z, fvec(volume.getSize()), fvec(self.pose.R.reshape(-1)),
                                              fvec(self.pose.t), _ptr(verts), _ptr(normals), cap, C.byref(nv),
                                              C.c_void_p(occ[0].data_ptr()), C.c_void_p(occ[1].data_ptr()),
                                              C.c_void_p(occ[2].data_ptr()), vcap, C.byref(nvox)))
        n = min(nv.value, cap)
        if return_occupied:
            return verts[:n], normals[:n], occ[:, :nvox.value]
        return verts[:n], normals[:n]

    def run_slab(self, slab, dims, size, z0, nz, vertex_cap=None, return_occupied=False):
        """z-slab form (sobfu_b200_marching_cubes_slab): `slab` is a float2 tensor [nz_avail, Y, X, 2] holding planes
        [z0, z0 + nz_avail) of a volume of `dims` / `size`, nz_avail = nz (+ 1: first plane of the upper neighbour).  The ranks'
        outputs concatenated in rank order equal run() on the whole volume."""
        cap = int(vertex_cap or self.DEFAULT_TRIANGLES_BUFFER_SIZE)
        dev = slab.device
        verts = torch.empty((cap, 4), dtype=torch.float32, device=dev)
        normals = torch.empty((cap, 4), dtype=torch.float32, device=dev)
        vcap = cap // 3
        occ = torch.empty((3, vcap), dtype=torch.int32, device=dev)
        nv, nvox = C.c_int(0), C.c_int(0)
        X, Y, Z = dims
        assert slab.is_contiguous() and slab.shape[1:] == (Y, X, 2), slab.shape
        check(lib().sobfu_b200_marching_cubes_slab(_ptr(slab), X, Y, Z, int(z0), int(nz), int(slab.shape[0]), fvec(size),
                                                   fvec(self.pose.R.reshape(-1)), fvec(self.pose.t), _ptr(verts), _ptr(normals), cap,
                                                   C.byref(nv), C.c_void_p(occ[0].data_ptr()), C.c_void_p(occ[1].data_ptr()),
                                                   C.c_void_p(occ[2].data_ptr()), vcap, C.byref(nvox)))
        n = min(nv.value, cap)
        if return_occupied:
            return verts[:n], normals[:n], occ[:, :nvox.value]
        return verts[:n], normals[:n]


class SobFusion:
    """Per-frame pipeline, SobFusion::operator() (src/sobfu/sob_fusion.cpp:71-145)."""

    def __init__(self, params):
        self.params = params
        self.frame_counter_ = 0
        self.poses_ = [Affine3f()]
        self.mc = MarchingCubes()
        self.mc.setPose(params.volume_pose)
        self.phi_global = self.phi_global_psi_inv = self.phi_n = self.phi_n_psi = None
        self.psi = self.psi_inv = self.solver = None

    def __call__(self, depth):
        p = self.params
        d = depthBilateralFilter(depth, p.bilateral_kernel_size, p.bilateral_sigma_spatial, p.bilateral_sigma_depth)
        depthTruncation(d, p.icp_truncate_depth_dist)
        dists = computeDists(d, p.intr)
        if self.frame_counter_ == 0:
            self.phi_global = TsdfVolume(p)
            self.phi_global.integrate(dists, self.poses_[-1], p.intr)
            self.phi_global_psi_inv = TsdfVolume(p)
            self.phi_n = TsdfVolume(p)
            self.phi_n_psi = TsdfVolume(p)
            self.psi = DeformationField(p.volume_dims)
            self.psi_inv = DeformationField(p.volume_dims)
            self.solver = Solver(p)
            self.frame_counter_ += 1
            return True
        self.phi_n.clear()
        self.phi_n.integrate(dists, self.poses_[-1], p.intr)
        if self.frame_counter_ < p.start_frame:
            self.phi_global.integrate(self.phi_n)
            self.frame_counter_ += 1
            return True
        self.solver.estimate_psi(self.phi_global, self.phi_global_psi_inv, self.phi_n, self.phi_n_psi, self.psi, self.psi_inv)
        self.phi_global.integrate(self.phi_n_psi)
        self.frame_counter_ += 1
        return True

    def get_phi_global_mesh(self):
        return self.mc.run(self.phi_global)
