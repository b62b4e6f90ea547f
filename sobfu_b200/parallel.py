"""z-slab multi-GPU mode of the solver: one process per GPU, torch.distributed for the plumbing only.

The volume is partitioned along z; rank r owns planes [r*Z/n, (r+1)*Z/n) of psi, psi^-1, phi_global, phi_global o psi^-1
and phi_n o psi, and keeps the whole phi_n (every rank integrates the 640x480 depth frame itself).  The halo exchanges
(psi: 4 planes per iteration and direction; nabla_U on the halo planes is recomputed) and the maximum of the convergence test run
inside libsobfu_b200.so -- by default inside its kernels over NVLink peer memory (CUDA IPC), else over its own NCCL communicator;
torch.distributed only carries the NCCL unique id and the IPC handles at start-up.
"""
import ctypes as C

from . import _capi
from ._capi import check
from .api import Solver, lib


def slab_range(Z, rank, nranks):
    """planes [z0, z0 + nz) owned by `rank` (pure host logic; same function the library uses)"""
    z0, nz = C.c_int(), C.c_int()
    check(lib().sobfu_b200_slab_range(int(Z), int(rank), int(nranks), C.byref(z0), C.byref(nz)))
    return z0.value, nz.value


def broadcast_unique_id(dist, src=0):
    """rank `src` creates the NCCL unique id, everybody receives it (works with the nccl and the gloo backend)"""
    buf = [None]
    if dist.get_rank() == src:
        raw = (C.c_ubyte * 128)()
        check(lib().sobfu_b200_comm_unique_id(raw))
        buf = [bytes(raw)]
    dist.broadcast_object_list(buf, src=src)
    return buf[0]


def slab_offsets(counts):
    """exclusive prefix sum over the ranks' vertex counts: where each rank's triangles start in the concatenated mesh
    ('MC needs a +1 z-plane halo and a cross-GPU exclusive offset, done on the host over G counts')"""
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += int(c)
    return out, acc


def upper_halo_plane(dist, slab, rank, nranks):
    """slab [nz, Y, X, 2] -> [nz + 1, Y, X, 2] with the first plane of the upper neighbour appended (last rank: unchanged).
    One plane travels down per rank; works with the nccl backend (device tensors) and gloo (host tensors)."""
    import torch
    if nranks == 1:
        return slab
    reqs = []
    out = slab
    if rank < nranks - 1:
        out = torch.empty((slab.shape[0] + 1,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
        out[:-1].copy_(slab)
        reqs.append(dist.irecv(out[-1], src=rank + 1))
    if rank > 0:
        first = slab[0].contiguous()
        reqs.append(dist.isend(first, dst=rank - 1))
    for r in reqs:
        r.wait()
    return out


class SlabSolver(Solver):
    """sobfu::cuda::Solver on a z-slab.  estimate_psi takes slab-local volumes/fields except phi_n (whole volume)."""

    def __init__(self, params, dist):
        super().__init__(params)
        self.rank, self.nranks = dist.get_rank(), dist.get_world_size()
        self.z0, self.nz = slab_range(params.volume_dims[2], self.rank, self.nranks)
        if self.nranks > 1:
            uid = broadcast_unique_id(dist)
            raw = (C.c_ubyte * 128).from_buffer_copy(uid)
            check(lib().sobfu_b200_solver_attach_comm(self._h, raw, self.rank, self.nranks))
            self.peer = self._attach_peers(dist)

    def _attach_peers(self, dist):
        """NVLink peer mode (include/sobfu_b200.h): every rank exports CUDA IPC handles of its psi planes and control block,
        the blocks are all-gathered, every rank maps its neighbours.  Returns False (NCCL exchange stays) when the ranks are
        not all on one node / cannot map each other; the decision is collective so that all ranks run the same protocol."""
        import os
        # Default since round 2 (profiles/r2_tuning_log.md): 256^3 on 8 GPUs 12709 vs 8416 it/s over NCCL, on 2 GPUs 4878 vs 4778;
        # SOBFU_B200_NO_PEER=1 keeps the NCCL exchange (also the automatic fallback when the IPC mappings cannot be made).
        if os.environ.get("SOBFU_B200_NO_PEER"):
            return False
        blk = (C.c_ubyte * 128)()
        ok = lib().sobfu_b200_solver_peer_export(self._h, blk) == 0
        blocks = [None] * self.nranks
        dist.all_gather_object(blocks, bytes(blk) if ok else None)
        if any(b is None for b in blocks):
            return False
        raw = (C.c_ubyte * (128 * self.nranks)).from_buffer_copy(b"".join(blocks))
        ok = lib().sobfu_b200_solver_peer_attach(self._h, raw) == 0
        oks = [None] * self.nranks
        dist.all_gather_object(oks, bool(ok))
        if not all(oks):
            lib().sobfu_b200_solver_peer_attach(self._h, None)      # every rank falls back together
            return False
        return True

    def slab_dims(self):
        X, Y, _ = self.params.volume_dims
        return (X, Y, self.nz)


class SlabFusion:
    """SobFusion::operator() (src/sobfu/sob_fusion.cpp:71-145) with the volume partitioned along z over the ranks.

    Every rank pre-processes the depth frame and integrates the whole live TSDF phi_n itself (0.3 Mpixel of input against
    a gather that may reach anywhere in the volume); phi_global, psi, psi^-1 and the warped volumes exist only as slabs.
    """

    def __init__(self, params, dist):
        import copy

        from .api import Affine3f, DeformationField, TsdfVolume
        self.params, self.dist = params, dist
        self.rank, self.nranks = dist.get_rank(), dist.get_world_size()
        self.z0, self.nz = slab_range(params.volume_dims[2], self.rank, self.nranks)
        X, Y, Z = params.volume_dims
        self.slab_params = copy.copy(params)
        self.slab_params.volume_dims = (X, Y, self.nz)
        self.slab_params.volume_size = (params.volume_size[0], params.volume_size[1], float(params.voxel_sizes()[2]) * self.nz)
        self.frame_counter_ = 0
        self.poses_ = [Affine3f()]
        from .api import MarchingCubes
        self.mc = MarchingCubes()
        self.mc.setPose(params.volume_pose)
        self._T, self._D = TsdfVolume, DeformationField
        self.phi_global = self.phi_global_psi_inv = self.phi_n = self.phi_n_psi = None
        self.psi = self.psi_inv = self.solver = None

    def _slab_of(self, vol):
        return vol.data()[self.z0:self.z0 + self.nz]

    def __call__(self, depth):
        from .api import computeDists, depthBilateralFilter, depthTruncation
        p = self.params
        d = depthBilateralFilter(depth, p.bilateral_kernel_size, p.bilateral_sigma_spatial, p.bilateral_sigma_depth)
        depthTruncation(d, p.icp_truncate_depth_dist)
        dists = computeDists(d, p.intr)
        if self.frame_counter_ == 0:
            self.phi_n = self._T(p)                                   # whole volume
            self.phi_n.integrate(dists, self.poses_[-1], p.intr)
            self.phi_global = self._T(self.slab_params)               # slabs
            self.phi_global.data().copy_(self._slab_of(self.phi_n))
            self.phi_global_psi_inv = self._T(self.slab_params)
            self.phi_n_psi = self._T(self.slab_params)
            self.psi = self._D(self.slab_params.volume_dims)
            self.psi.get_data()[..., 2] += float(self.z0)             # identity in ABSOLUTE voxel coordinates
            self.psi_inv = self._D(self.slab_params.volume_dims)
            self.solver = SlabSolver(p, self.dist)
            self.frame_counter_ += 1
            return True
        self.phi_n.clear()
        self.phi_n.integrate(dists, self.poses_[-1], p.intr)
        if self.frame_counter_ < p.start_frame:
            self.phi_n_psi.data().copy_(self._slab_of(self.phi_n))
            self.phi_global.integrate(self.phi_n_psi)
            self.frame_counter_ += 1
            return True
        self.solver.estimate_psi(self.phi_global, self.phi_global_psi_inv, self.phi_n, self.phi_n_psi, self.psi, self.psi_inv)
        self.phi_global.integrate(self.phi_n_psi)
        self.frame_counter_ += 1
        return True

    # ---- meshes (SobFusion::get_phi_*_mesh, sob_fusion.cpp:147-183): marching cubes per slab ----
    def slab_mesh(self, vol, vertex_cap=None):
        """this rank's part of the zero level set of a slab volume: (vertices [n,4], normals [n,4], offset, total) where
        `offset` is the position of the rank's first vertex in the mesh of the whole volume and `total` its vertex count.
        The parts concatenated in rank order are bit-identical to marching cubes on the whole volume on one GPU."""
        import torch
        p = self.params
        slab = upper_halo_plane(self.dist, vol.data(), self.rank, self.nranks)
        verts, normals = self.mc.run_slab(slab, p.volume_dims, p.volume_size, self.z0, self.nz, vertex_cap)
        offs, total = slab_offsets(self._all_counts(int(verts.shape[0]), verts.device))
        return verts, normals, offs[self.rank], total

    def _all_counts(self, n, device):
        """every rank's vertex count (a tensor all-gather on the ranks' device: no pickling on the per-frame path)"""
        import torch
        mine = torch.tensor([n], dtype=torch.int64, device=device)
        allc = torch.empty((self.nranks,), dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(allc, mine)
        return [int(c) for c in allc.tolist()]

    def gather_mesh(self, vol, dst=0, vertex_cap=None):
        """the whole mesh on rank `dst` (vertices, normals), None elsewhere -- for export"""
        import torch
        verts, normals, off, total = self.slab_mesh(vol, vertex_cap)
        counts = self._all_counts(int(verts.shape[0]), verts.device)
        if self.rank == dst:
            V = torch.empty((total, 4), dtype=torch.float32, device=verts.device)
            Nn = torch.empty((total, 4), dtype=torch.float32, device=verts.device)
            offs, _ = slab_offsets(counts)
            reqs = []
            for r in range(self.nranks):
                if r == dst:
                    V[offs[r]:offs[r] + counts[r]].copy_(verts)
                    Nn[offs[r]:offs[r] + counts[r]].copy_(normals)
                elif counts[r] > 0:
                    reqs.append(self.dist.irecv(V[offs[r]:offs[r] + counts[r]], src=r))
                    reqs.append(self.dist.irecv(Nn[offs[r]:offs[r] + counts[r]], src=r))
            for q in reqs:
                q.wait()
            return V, Nn
        if verts.shape[0] > 0:
            self.dist.send(verts.contiguous(), dst=dst)
            self.dist.send(normals.contiguous(), dst=dst)
        return None

    def get_phi_global_mesh(self):
        return self.slab_mesh(self.phi_global)

    def get_phi_global_psi_inv_mesh(self):
        return self.slab_mesh(self.phi_global_psi_inv)

    def get_phi_n_psi_mesh(self):
        return self.slab_mesh(self.phi_n_psi)
