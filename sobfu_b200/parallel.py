"""z-slab multi-GPU mode of the solver: one process per GPU, torch.distributed for the plumbing only.

The volume is partitioned along z; rank r owns planes [r*Z/n, (r+1)*Z/n) of psi, psi^-1, phi_global, phi_global o psi^-1
and phi_n o psi, and keeps the whole phi_n (every rank integrates the 640x480 depth frame itself).  The halo exchanges
(nabla_U: 3 planes, psi: 1 plane, per iteration and direction) and the scalar MAX all-reduce of the convergence test run
inside libsobfu_b200.so over its own NCCL communicator; torch.distributed only carries the NCCL unique id at start-up.
"""
import ctypes as C

from . import _capi
from ._capi import check, lib
from .api import Solver


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

    def slab_dims(self):
        X, Y, _ = self.params.volume_dims
        return (X, Y, self.nz)
