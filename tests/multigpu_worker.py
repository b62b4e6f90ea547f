"""torchrun worker: z-slab solve on N GPUs vs the same solve on one GPU (run by tests/test_multigpu_gpu.py).
Every rank also solves the whole volume on its own GPU and compares its slab bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SOBFU_B200_QUIET", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sobfu_b200 as sf  # noqa: E402
from tests.common import assert_bits, sphere_pair, wavy_psi  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    # the fifth case displaces psi by up to 24 planes along z: psi^-1 leaves the 16-plane neighbour window and the all-gather fallback runs
    cases = [((64, 64, 64), 12, 0.0, 0, -1.0), ((96, 40, 32), 7, 0.4, 2, -1.0), ((20, 18, 16), 5, 0.3, 1, -1.0), ((64, 32, 32), 60, 0.0, 0, None),
             ((64, 32, 64), 4, 24.0, 0, -1.0)]
    for dims, iters, wavy, verbosity, thr in cases:
        X, Y, Z = dims
        pg, pn, vs, trunc, eta = sphere_pair(dims, r=0.07)
        psi0 = wavy_psi(dims, amp=wavy) if wavy else sf_identity(dims)
        p = sf.Params(volume_dims=dims, volume_size=tuple(float(vs[i]) * dims[i] for i in range(3)), max_iter=iters, max_update_norm=-1.0,
                      s=7, lambda_=0.1, alpha=0.05, w_reg=0.3, verbosity=verbosity, tsdf_max_weight=64.0, tsdf_trunc_dist=float(trunc), eta=float(eta))
        # single-GPU solve of the whole volume (on this rank's GPU)
        full = solve(p, dims, pg, pn, psi0, None)
        if thr is None:    # early stop in the middle of a chunk: threshold = update norm of iteration 37
            p.max_update_norm = float(full["log"][36][0])
            full = solve(p, dims, pg, pn, psi0, None)
            assert full["info"].converged == 1 and full["info"].iters <= 37
        # slab solve
        slab = solve(p, dims, pg, pn, psi0, dist)
        z0, nz = slab["z0"], slab["nz"]
        assert slab["info"].iters == full["info"].iters and slab["info"].converged == full["info"].converged
        assert slab["info"].max_norm == full["info"].max_norm and slab["info"].max_idx == full["info"].max_idx
        for k in ("psi", "psi_inv", "phi_n_psi", "phi_global_psi_inv"):
            assert_bits(slab[k], full[k][z0:z0 + nz], "%s rank %d %s" % (dims, rank, k))
        assert (slab["tail_fallbacks"] > 0) == (wavy > 16.0), (dims, wavy, slab["tail_fallbacks"])
        for a, b in zip(slab["log"], full["log"]):
            assert a[0] == b[0] and a[1] == b[1]
            assert abs(a[2] - b[2]) <= 2e-5 * abs(b[2]) + 1e-6 and abs(a[3] - b[3]) <= 2e-5 * abs(b[3]) + 1e-6
        if rank == 0:
            print("slab == single GPU, bit for bit:", dims, "iters", full["info"].iters, "ranks", world, "peer mode:", slab["peer"], flush=True)
    frames_and_meshes(rank, world)
    dist.barrier()
    dist.destroy_process_group()


def frames_and_meshes(rank, world):
    """SlabFusion (depth frame -> slab TSDF -> slab solve -> slab fusion -> marching cubes per slab) against SobFusion on one GPU"""
    from sobfu_b200.parallel import SlabFusion
    import bench
    p = bench.make_params(sf, 64, 6)
    p.max_update_norm = -1.0
    one, slab = sf.SobFusion(p), SlabFusion(p, dist)
    for f in range(3):
        d = torch.from_numpy(bench.synth_depth(3 * f).view(np.int16)).cuda().view(torch.uint16)
        one(d)
        slab(d)
    z0, nz = slab.z0, slab.nz
    for name in ("phi_global", "phi_n_psi", "phi_global_psi_inv"):
        assert_bits(getattr(slab, name).data().cpu().numpy(), getattr(one, name).data().cpu().numpy()[z0:z0 + nz], "SlabFusion %s rank %d" % (name, rank))
    V, Nn = one.mc.run(one.phi_global)
    v, n, off, total = slab.get_phi_global_mesh()
    assert total == V.shape[0] and total > 1000, (total, V.shape)
    assert torch.equal(v, V[off:off + v.shape[0]]) and torch.equal(n, Nn[off:off + n.shape[0]])
    whole = slab.gather_mesh(slab.phi_global, dst=0)
    if rank == 0:
        assert torch.equal(whole[0], V) and torch.equal(whole[1], Nn)
        print("slab meshes == single GPU, bit for bit:", total, "vertices over", world, "ranks", flush=True)


def sf_identity(dims):
    X, Y, Z = dims
    z, y, x = np.meshgrid(np.arange(Z, dtype=np.float32), np.arange(Y, dtype=np.float32), np.arange(X, dtype=np.float32), indexing="ij")
    return np.ascontiguousarray(np.stack([x, y, z, np.zeros_like(x)], -1))


def solve(p, dims, pg, pn, psi0, dist_or_none):
    X, Y, Z = dims
    if dist_or_none is None:
        solver, z0, nz = sf.Solver(p), 0, Z
    else:
        solver = sf.SlabSolver(p, dist_or_none)
        z0, nz = solver.z0, solver.nz
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    d_pg, d_pn, d_psi = dev(pg[z0:z0 + nz]), dev(pn), dev(psi0[z0:z0 + nz])
    d_pgpi, d_pnp, d_inv = torch.empty_like(d_pg), torch.empty_like(d_pg), torch.empty_like(d_psi)
    import ctypes as C
    from sobfu_b200 import _capi
    info = _capi.SolveInfo()
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _capi.check(_capi.lib().sobfu_b200_solver_estimate_psi(solver._h, ptr(d_pg), ptr(d_pgpi), ptr(d_pn), ptr(d_pnp), ptr(d_psi), ptr(d_inv), C.byref(info)))
    solver.info = info
    return dict(info=info, log=solver.get_log(), psi=d_psi.cpu().numpy(), psi_inv=d_inv.cpu().numpy(), phi_n_psi=d_pnp.cpu().numpy(),
                phi_global_psi_inv=d_pgpi.cpu().numpy(), z0=z0, nz=nz, peer=bool(getattr(solver, "peer", False)),
                tail_fallbacks=solver.tail_fallbacks())


if __name__ == "__main__":
    main()
