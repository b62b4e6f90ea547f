"""torchrun worker (N >= 2 GPUs): (1) slab solve in peer mode == single-GPU solve, bit for bit, on a small volume (exercises
ranks with TWO neighbours when N >= 3); (2) it/s of a 256^3 estimate_psi (200 iterations) in peer mode and over NCCL in the
same process.  Usage: python -m torch.distributed.run --nproc-per-node N ... tests/peer_check_worker.py [dim] [iters] [Z]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SOBFU_B200_QUIET", "1")
os.environ.pop("SOBFU_B200_NO_PEER", None)   # peer mode is the default

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sobfu_b200 as sf  # noqa: E402
from tests.common import assert_bits, sphere_pair, wavy_psi  # noqa: E402
from tests.multigpu_worker import solve  # noqa: E402


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    # (1) correctness, peer mode
    dims = (64, 48, 16 * world)
    pg, pn, vs, trunc, eta = sphere_pair(dims, r=0.07)
    psi0 = wavy_psi(dims, amp=0.4)
    p = sf.Params(volume_dims=dims, volume_size=tuple(float(vs[i]) * dims[i] for i in range(3)), max_iter=9, max_update_norm=-1.0, s=7,
                  lambda_=0.1, alpha=0.05, w_reg=0.3, verbosity=0, tsdf_max_weight=64.0, tsdf_trunc_dist=float(trunc), eta=float(eta))
    full = solve(p, dims, pg, pn, psi0, None)
    slab = solve(p, dims, pg, pn, psi0, dist)
    z0, nz = slab["z0"], slab["nz"]
    assert slab["peer"], "peer mode was not attached"
    assert slab["info"].iters == full["info"].iters and slab["info"].max_norm == full["info"].max_norm
    for k in ("psi", "psi_inv", "phi_n_psi", "phi_global_psi_inv"):
        assert_bits(slab[k], full[k][z0:z0 + nz], "%s rank %d %s" % (dims, rank, k))
    if rank == 0:
        print("peer slab == single GPU, bit for bit:", dims, "ranks", world, flush=True)
    # (2) timing: peer, then NCCL
    dims = (dim, dim, int(sys.argv[3]) if len(sys.argv) > 3 else dim)     # optional third argument: Z (thin slabs on few GPUs)
    pg, pn, vs, trunc, eta = sphere_pair(dims, r=0.07)
    p = sf.Params(volume_dims=dims, volume_size=tuple(float(vs[i]) * dims[i] for i in range(3)), max_iter=iters, max_update_norm=1e-10, s=7,
                  lambda_=0.1, alpha=0.001, w_reg=0.6, verbosity=0, tsdf_max_weight=64.0, tsdf_trunc_dist=float(trunc), eta=float(eta))
    out = {}
    for mode in os.environ.get("PEER_CHECK_MODES", "peer,nccl").split(","):
        if mode == "nccl":
            os.environ["SOBFU_B200_NO_PEER"] = "1"
        solver = sf.SlabSolver(p, dist)
        z0, nz = solver.z0, solver.nz
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
        X, Y, Z = dims
        ident = np.zeros((nz, Y, X, 4), dtype=np.float32)
        zz, yy, xx = np.meshgrid(np.arange(z0, z0 + nz, dtype=np.float32), np.arange(Y, dtype=np.float32), np.arange(X, dtype=np.float32), indexing="ij")
        ident[..., 0], ident[..., 1], ident[..., 2] = xx, yy, zz
        d_pg, d_pn, d_psi = dev(pg[z0:z0 + nz]), dev(pn), dev(ident)
        d_pgpi, d_pnp, d_inv = torch.empty_like(d_pg), torch.empty_like(d_pg), torch.empty_like(d_psi)
        import ctypes as C
        from sobfu_b200 import _capi
        info = _capi.SolveInfo()
        ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        best = 1e30
        for rep in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            _capi.check(_capi.lib().sobfu_b200_solver_estimate_psi(solver._h, ptr(d_pg), ptr(d_pgpi), ptr(d_pn), ptr(d_pnp), ptr(d_psi), ptr(d_inv), C.byref(info)))
            t = torch.tensor([info.loop_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep > 0:
                best = min(best, float(t.item()))
        assert info.iters == iters
        out[mode] = {"peer_attached": bool(getattr(solver, "peer", False)), "loop_ms_per_iter": best / iters, "iters_per_s": iters / (best * 1e-3), "launches": info.launches}
        if mode == "peer" and os.environ.get("SOBFU_B200_TRACE"):
            out[mode]["trace_per_rank"] = trace_summary(solver, iters, dist, world)
        if mode == "nccl":      # where the time of an iteration goes on this rank: A_mid | wait + A_edge | wait + B_edge | B_mid | iteration
            ph = torch.tensor(solver.time_phases(50), device="cuda")
            dist.all_reduce(ph, op=dist.ReduceOp.MAX)
            out[mode]["phases_ms_max_over_ranks"] = [round(float(x), 5) for x in ph.tolist()]
            ta, tb, tl = solver.time_loop(50)
            out[mode]["whole_slab_kernels_ms"] = {"pass_a": ta, "pass_b": tb}
        del solver
    if rank == 0:
        print(json.dumps({"n_gpus": world, "dim": dim, "iters": iters, **out}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def trace_summary(solver, iters, dist, world):
    """device-side timeline of the last peer-mode solve (sobfu_b200_solver_get_trace): mean microseconds over the middle iterations"""
    import ctypes as C
    from sobfu_b200 import _capi
    cap = 2 * iters
    buf, n = (C.c_ulonglong * (8 * cap))(), C.c_int()
    _capi.check(_capi.lib().sobfu_b200_solver_get_trace(solver._h, buf, cap, C.byref(n)))
    t = np.array(buf[:8 * n.value], dtype=np.float64).reshape(-1, 8)
    res = None
    if len(t) >= 40:
        a, b = t[0::2], t[1::2]
        k = slice(len(a) // 4, 3 * len(a) // 4)
        us = lambda x: round(float(np.mean(x)) * 1e-3, 2)  # noqa: E731
        res = {"A_us": us(a[k, 1] - a[k, 0]), "B_us": us(b[k, 1] - b[k, 0]), "gap_A_to_B_us": us(b[k, 0] - a[k, 1]),
               "gap_B_to_nextA_us": us(a[1:][k, 0] - b[:-1][k, 1]), "iter_us": us(a[1:][k, 0] - a[:-1][k, 0]),
               "A_face_wait_max_us": us(a[k, 5]), "A_face_wait_sum_per_cta_us": us(a[k, 4] / np.maximum(a[k, 6], 1)),
               "B_table_wait_max_us": us(b[k, 3]), "B_table_wait_mean_us": us(b[k, 2] / np.maximum(b[k, 6], 1)),
               "B_ack_wait_max_us": us(b[k, 5]), "ctas": [int(a[k, 6].mean()), int(b[k, 6].mean())]}
    allr = [None] * world
    dist.all_gather_object(allr, res)
    return allr


if __name__ == "__main__":
    main()
