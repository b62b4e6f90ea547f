#!/usr/bin/env python3
"""bench.py -- headline benchmark of the SobolevFusion solver hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[2]): 256^3 volume, params/params_boxing.ini of the reference (dims -> 256, MAX_ITER -> 200,
MAX_UPDATE_NORM 1e-10 so that exactly 200 iterations run), synthetic 640x480 depth frames of an analytically ray-cast
sphere translating 2 mm per frame.  One "step" = one frame = one Solver::estimate_psi (200 gradient-descent iterations +
psi^-1 + the two warps).

  value  : Gvoxel-iterations/s of estimate_psi with all volumes resident in HBM (device time, CUDA events)
  e2e    : the same metric through the frame call a user of the reference makes -- SobFusion::operator()(depth) with the
           depth frame in pinned HOST memory: H2D of the frame, bilateral/truncate/dists, TSDF integration, the solver,
           TSDF fusion, and a D2H read of the solve result, all inside the timed region
  roofline: the longer of the two kernels of an iteration, with the algorithmic bytes of SURVEY.md 8d
           (pass A incl. the warp 64 B/voxel, pass B = Sobolev filter + psi update + max partials 48 B/voxel);
           both fractions are also reported as pass_a_roofline_frac / pass_b_roofline_frac
  --impl reference : the UNMODIFIED reference CUDA (oracle/_ref, built from /root/reference) on the same workload; the
           reference has no CPU solver path, so its own CUDA on one B200 is the baseline arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("SOBFU_B200_QUIET", "1")

# params/params_boxing.ini (reference), with the overrides BASELINE.json names for config 3
BOXING = dict(vol_size=0.75, trunc_vox=48.0, eta_vox=3.0, max_weight=128.0, fx=570.342, fy=570.342, cx=320.0, cy=240.0,
              trunc_depth=1.0, pose_tz=0.1, sigma_depth=0.005, sigma_spatial=4.5, ksz=7, start_frame=1,
              max_update_norm=1e-10, s=7, lam=0.1, alpha=0.001, w_reg=0.6)
# params/params_umbrella.ini (reference): BASELINE.json configs[3] names it for the 512^3 volume (MAX_ITER -> 200 as everywhere here)
UMBRELLA = dict(vol_size=1.0, trunc_vox=8.0, eta_vox=3.0, max_weight=128.0, fx=570.342, fy=570.342, cx=320.0, cy=240.0,
                trunc_depth=1.5, pose_tz=0.3, sigma_depth=0.04, sigma_spatial=4.5, ksz=7, start_frame=1,
                max_update_norm=1e-10, s=7, lam=0.1, alpha=0.001, w_reg=0.2)
COLS, ROWS = 640, 480


def params_for(dim):
    """(ini values, ini name) of the solver workload at `dim`^3: params_umbrella.ini at 512^3, params_boxing.ini otherwise"""
    return (UMBRELLA, "params_umbrella.ini") if dim >= 512 else (BOXING, "params_boxing.ini")
# algorithmic bytes per voxel-iteration in the reference's layouts; the warp of the live TSDF is fused
# into pass A here (SURVEY fuses it into pass B), so its 16 B move with it: 48 + 16 | 48 = 112
ALGO_BYTES_PASS_A = 64      # R psi 16 + R phi_n_psi 8 + R phi_global 8 + W nabla_U 16  +  warp: R phi_n 8 + W phi_n_psi 8
ALGO_BYTES_PASS_B = 48      # R nabla_U 16 + R psi 16 + W psi 16   (Sobolev filter + psi update + max-norm partials)
ALGO_BYTES_ITER = 112
# compulsory bytes of the layouts this library actually uses inside the loop (float planes; phi_n o psi never leaves the SM):
OWN_BYTES_PASS_A = 32       # R psi 12 + R phi_global 4 + R phi_n (gathers) ~4 + W nabla_U 12
OWN_BYTES_PASS_B = 36       # R nabla_U 12 + R psi 12 + W psi 12


def synth_depth(frame, radius=0.15, z0=0.5, intr=None):
    """analytically ray-cast sphere: centre (0.002*frame, 0, z0) m, ushort millimetres, background 0"""
    u, v = np.meshgrid(np.arange(COLS, dtype=np.float64), np.arange(ROWS, dtype=np.float64))
    k = intr or BOXING
    dx, dy = (u - k["cx"]) / k["fx"], (v - k["cy"]) / k["fy"]
    c = np.array([0.002 * frame, 0.0, z0])
    a = dx * dx + dy * dy + 1.0
    b = -2.0 * (dx * c[0] + dy * c[1] + c[2])
    cc = float(c @ c) - radius * radius
    disc = b * b - 4 * a * cc
    t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0.0)
    return np.ascontiguousarray(np.where(disc > 0, np.round(t * 1000.0), 0).astype(np.uint16))


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def make_params(sf, dim, iters):
    b, _ = params_for(dim)
    p = sf.Params(cols=COLS, rows=ROWS, volume_dims=(dim, dim, dim), volume_size=(b["vol_size"],) * 3,
                  intr=sf.Intr(b["fx"], b["fy"], b["cx"], b["cy"]), icp_truncate_depth_dist=b["trunc_depth"],
                  bilateral_sigma_depth=b["sigma_depth"], bilateral_sigma_spatial=b["sigma_spatial"], bilateral_kernel_size=b["ksz"],
                  tsdf_max_weight=b["max_weight"], gradient_delta_factor=0.5, start_frame=b["start_frame"], verbosity=0, s=b["s"],
                  max_iter=iters, max_update_norm=b["max_update_norm"], lambda_=b["lam"], alpha=b["alpha"], w_reg=b["w_reg"])
    vs = p.voxel_sizes()
    p.tsdf_trunc_dist = float(np.float32(b["trunc_vox"]) * vs[0])      # demo.cpp:71-72: given in voxels
    p.eta = float(np.float32(b["eta_vox"]) * vs[0])
    p.volume_pose = sf.Affine3f().translate((-b["vol_size"] / 2, -b["vol_size"] / 2, b["pose_tz"]))   # demo.cpp:73-74
    return p


def cpu_baseline(dim=160, min_seconds=4.0):
    """the oracle port (CPU restatement of the same iteration) on all host cores, bounded sample (a few seconds of wall
    clock = minutes of core time on the GPU box's 128 cores)"""
    from oracle import pyoracle as orc
    from tests.common import sphere_pair
    dims = (dim, dim, dim)
    pg, pn, vs, trunc, eta = sphere_pair(dims)
    psi = orc.init_identity(*dims)
    pnp = orc.apply(pn, psi)
    scratch = np.zeros((5,) + psi.shape, dtype=np.float32)
    taps = orc.sobolev_taps(7, 0.1)
    orc.solver_iteration(pg, pn, pnp, psi, scratch, taps, 0.001, 0.6)      # warm-up (page faults, omp pool)
    t0 = time.perf_counter()
    iters = 0
    while time.perf_counter() - t0 < min_seconds:
        orc.solver_iteration(pg, pn, pnp, psi, scratch, taps, 0.001, 0.6)
        iters += 1
    dt = time.perf_counter() - t0
    return {"value": dim ** 3 * iters / dt / 1e9, "unit": "Gvoxel-iter/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d solver iterations at %d^3 (oracle/sobfu_oracle.c, OpenMP on all host cores), %.2f s" % (iters, dim, dt)}


def measure_solver(args, rank, world, torch, dist, dim, iters, steps, warmup):
    """one solver-workload measurement (value, e2e, per-kernel roofline) at `dim`^3; returns (JSON line as a dict, fusion)"""
    import sobfu_b200 as sf
    from sobfu_b200.parallel import SlabFusion
    p = make_params(sf, dim, iters)
    fusion = sf.SobFusion(p) if world == 1 else SlabFusion(p, dist)   # z-slab over the ranks
    frames = [torch.from_numpy(synth_depth(f).view(np.int16)).pin_memory() for f in range(1 + 2 * (warmup + steps))]
    dev_depth = torch.empty((ROWS, COLS), dtype=torch.int16, device="cuda")

    def frame_step(f):
        dev_depth.copy_(frames[f], non_blocking=True)              # H2D of this step's input from pinned memory
        fusion(dev_depth.view(torch.uint16))
        return fusion.solver.info.max_norm if fusion.solver is not None and fusion.solver.info is not None else 0.0   # D2H'd by the call

    frame_step(0)                                                  # frame 0 only initialises phi_global
    f = 1
    for _ in range(warmup):
        frame_step(f); f += 1
    solver = fusion.solver
    if args.variant:
        solver.set_variant(args.variant)
    N = dim ** 3

    # ---- value: solver only, volumes resident ------------------------------------------------------------------
    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sync_all()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, loop_ms = 0, 0.0
    e0.record()
    for _ in range(steps):
        info = solver.estimate_psi(fusion.phi_global, fusion.phi_global_psi_inv, fusion.phi_n, fusion.phi_n_psi, fusion.psi, fusion.psi_inv)
        launches += info.launches
        loop_ms += info.loop_ms
        assert info.iters == iters, info.iters
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    # ---- e2e: frames from pinned host memory through SobFusion::operator() ---------------------------------------
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(steps):
        frame_step(f); f += 1
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    clk = clocks.stop()
    # ---- roofline of the dominant kernel: per-kernel device time measured on the solver's own stream ----------------
    ms_a, ms_b, ms_it = solver.time_loop(max(20, min(iters, 100)))
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_a, ms_b, ms_it], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_a, ms_b, ms_it = [float(x) for x in t.tolist()]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    Nl = N // world                                   # voxels per GPU (z-slab)
    dom, dom_ms, dom_bytes = ("pass_a", ms_a, ALGO_BYTES_PASS_A) if ms_a >= ms_b else ("pass_b", ms_b, ALGO_BYTES_PASS_B)
    achieved = dom_bytes * Nl / (dom_ms * 1e-3) / 1e9
    desc = {"pass_a": "pass_a (warp of the live TSDF + SDF-difference data-term gradient + Laplacian -> nabla_U)",
            "pass_b": "pass_b (Sobolev filter + psi update + max-norm partials)"}
    out = {
        "metric": "solver_gvoxel_iters_per_s", "value": N * iters * steps / (ms * 1e-3) / 1e9, "unit": "Gvoxel-iter/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d^3 volume, %s (dims->%d, MAX_ITER->%d), 640x480 synthetic sphere depth, 1 step = 1 frame = "
                               "estimate_psi with %d iterations" % (dim, params_for(dim)[1], dim, iters, iters), "l2": ("inputs exceed L2 (%.0f MB of solver state)" if 36 * N > 126e6 else
                          "NOT flushed: the solver state (%.0f MB) fits in the 126 MB L2 at this size; only volumes of >= 192^3 are HBM-bound") % (36 * N / 1e6),
                   "parallelism": "1 GPU" if world == 1 else "z-slab x%d" % world},
        "solver_iters_per_s": iters * steps / (ms * 1e-3), "loop_ms_per_iter": loop_ms / (steps * iters),
        "kernel_ms": {"pass_a": ms_a, "pass_b": ms_b, "iteration": ms_it if world == 1 else loop_ms / (steps * iters)},
        # whole iteration incl. halo exchanges, from the loop of the timed estimate_psi calls (device events of the library)
        "iteration_roofline_frac": ALGO_BYTES_ITER * Nl / (loop_ms / (steps * iters) * 1e-3) / 1e9 / peak,
        "pass_b_roofline_frac": ALGO_BYTES_PASS_B * Nl / (ms_b * 1e-3) / 1e9 / peak,
        "pass_a_roofline_frac": ALGO_BYTES_PASS_A * Nl / (ms_a * 1e-3) / 1e9 / peak,
        # the same two kernels against the bytes of OUR layouts (what a perfectly HBM-bound version of these kernels would move)
        "own_layout_frac": {"pass_a": OWN_BYTES_PASS_A * Nl / (ms_a * 1e-3) / 1e9 / peak, "pass_b": OWN_BYTES_PASS_B * Nl / (ms_b * 1e-3) / 1e9 / peak,
                            "bytes_per_voxel": {"pass_a": OWN_BYTES_PASS_A, "pass_b": OWN_BYTES_PASS_B}},
        "clocks": clk,
        "e2e": {"value": N * iters * steps / (ms_e2e * 1e-3) / 1e9, "unit": "Gvoxel-iter/s", "frames_per_s": steps / (ms_e2e * 1e-3),
                "h2d_bytes_per_step": COLS * ROWS * 2, "d2h_bytes_per_step": 4 + iters * 24 + 16},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": desc[dom], "achieved": achieved,
                     "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "unit": "GB/s", "frac": achieved / peak, "algorithmic_bytes_per_voxel": dom_bytes,
                     "traffic": None, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"},
    }
    if world > 1:
        out["roofline"]["traffic_source"] = "not measured at N > 1 (ncu profiles one process); the N = 1 line of the same build carries it"
        peer = bool(getattr(fusion.solver, "peer", False))
        out["config"]["multi_gpu"] = ("z-slab of %d planes per GPU; nabla_U on the 3 halo planes is recomputed locally, so an iteration needs "
                                      "one psi halo exchange (4 planes) + one scalar MAX; " % (dim // world)) + (
            "peer mode: pass B on the slab faces stores the planes straight into the neighbours' halo planes over NVLink (CUDA IPC) and "
            "signals through counters per work item (face chunks first in both passes), maxima are published to every rank's table by the "
            "last CTA; two kernels per iteration on one stream, no NCCL in the loop" if peer else
            "NCCL: grouped ncclSend/Recv with both neighbours behind the mid-slab kernels + MAX all-reduce on a second communicator")
    return out, fusion


def _fingerprint(a):
    import hashlib
    return hashlib.blake2b(np.ascontiguousarray(a).view(np.uint8), digest_size=8).hexdigest()


def _cmp(name, ours, theirs):
    """bit comparison of two float arrays (host numpy or device tensors of the same kind)"""
    if isinstance(ours, np.ndarray):
        a, b = ours.view(np.uint32), theirs.view(np.uint32)
        nbad = int(np.count_nonzero(a != b))
        maxd = float(np.nanmax(np.abs(ours - theirs))) if nbad else 0.0
        return {"bit_exact": nbad == 0, "words_differing": nbad, "max_abs_diff": maxd, "fingerprint": _fingerprint(ours)}
    import torch
    a, b = ours.contiguous().view(torch.int32), theirs.contiguous().view(torch.int32)
    nbad = int((a != b).sum().item())
    maxd = float((ours - theirs).abs().nan_to_num(0.0).max().item()) if nbad else 0.0
    return {"bit_exact": nbad == 0, "words_differing": nbad, "max_abs_diff": maxd}


def parity_block(args, rank, world, torch, dist, fusion, dim, iters):
    """AFTER the timed regions (the checker is never inside them): one estimate_psi from the identity on this run's phi_global /
    phi_n, compared bit for bit
      N = 1 : with the reference's own CUDA (oracle/_ref, unmodified sources) on the same inputs
      N > 1 : with a single-GPU solve of the whole volume on rank 0 (slabs gathered there)
    for psi, psi^-1, phi_n o psi and phi_global o psi^-1."""
    import sobfu_b200 as sf
    names = ("psi", "psi_inv", "phi_n_psi", "phi_global_psi_inv")
    solver = fusion.solver
    X = Y = Z = dim
    if world == 1:
        from oracle import pyoracle as orc
        if not os.path.exists(orc.REF):
            return {"checked": False, "why": "oracle/_ref/libsobfu_ref.so is not built on this box"}
        b, _ = params_for(dim)
        psi, psi_inv = sf.DeformationField((X, Y, Z)), sf.DeformationField((X, Y, Z))
        info = solver.estimate_psi(fusion.phi_global, fusion.phi_global_psi_inv, fusion.phi_n, fusion.phi_n_psi, psi, psi_inv)
        ours = {"psi": psi.get_data().cpu().numpy(), "psi_inv": psi_inv.get_data().cpu().numpy(),
                "phi_n_psi": fusion.phi_n_psi.data().cpu().numpy(), "phi_global_psi_inv": fusion.phi_global_psi_inv.data().cpu().numpy()}
        vs = np.float32(b["vol_size"]) / np.float32(dim)
        ref = orc.Reference((dim,) * 3, (b["vol_size"],) * 3, float(np.float32(b["trunc_vox"]) * vs), float(np.float32(b["eta_vox"]) * vs),
                            b["max_weight"], 0, iters, b["s"], b["max_update_norm"], b["lam"], b["alpha"], b["w_reg"])
        ref.upload_tsdf(ref.GLOBAL, fusion.phi_global.data().cpu().numpy())
        ref.upload_tsdf(ref.N, fusion.phi_n.data().cpu().numpy())
        ref.psi_clear(0)
        devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
        sys.stdout.flush()
        os.dup2(devnull, 1)          # the reference prints from inside its loop
        try:
            ref.estimate_psi()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
        theirs = {"psi": ref.download_psi(0), "psi_inv": ref.download_psi(1), "phi_n_psi": ref.download_tsdf(ref.N_PSI),
                  "phi_global_psi_inv": ref.download_tsdf(ref.GLOBAL_PSI_INV)}
        ref.close()
        res = {k: _cmp(k, ours[k], theirs[k]) for k in names}
        return {"checked": True, "against": "oracle/_ref: the reference's own CUDA (unmodified sources, sm_100a), same inputs, identity start",
                "volume": "%d^3" % dim, "iterations": int(info.iters), "bit_exact": all(r["bit_exact"] for r in res.values()), **res}
    # ---- N > 1: the slab solve against a single-GPU solve of the whole volume on rank 0 ----
    nz, z0 = fusion.nz, fusion.z0
    psi, psi_inv = sf.DeformationField((X, Y, nz)), sf.DeformationField((X, Y, nz))
    psi.get_data()[..., 2] += float(z0)
    info = solver.estimate_psi(fusion.phi_global, fusion.phi_global_psi_inv, fusion.phi_n, fusion.phi_n_psi, psi, psi_inv)
    mine = {"psi": psi.get_data(), "psi_inv": psi_inv.get_data(), "phi_n_psi": fusion.phi_n_psi.data(), "phi_global_psi_inv": fusion.phi_global_psi_inv.data()}

    def gather0(t):
        parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t.contiguous(), parts, dst=0)
        return torch.cat(parts, 0) if rank == 0 else None

    pg_full = gather0(fusion.phi_global.data())
    got = {k: gather0(mine[k]) for k in names}
    res = None
    if rank == 0:
        p1 = make_params(sf, dim, iters)
        vols = [sf.TsdfVolume(p1) for _ in range(3)]
        vols[0].data().copy_(pg_full)
        f_psi, f_inv = sf.DeformationField((X, Y, Z)), sf.DeformationField((X, Y, Z))
        one = sf.Solver(p1)
        info1 = one.estimate_psi(vols[0], vols[1], fusion.phi_n, vols[2], f_psi, f_inv)
        want = {"psi": f_psi.get_data(), "psi_inv": f_inv.get_data(), "phi_n_psi": vols[2].data(), "phi_global_psi_inv": vols[1].data()}
        cmp = {k: _cmp(k, got[k], want[k]) for k in names}
        res = {"checked": True, "against": "a single-GPU solve of the whole volume on rank 0 (the library's one-GPU path, itself compared with the "
                                          "reference CUDA in the N=1 line), slabs gathered from all ranks, identity start",
               "volume": "%d^3" % dim, "iterations": int(info.iters), "iterations_single_gpu": int(info1.iters),
               "max_norm_equal": bool(info.max_norm == info1.max_norm),
               "bit_exact": all(r["bit_exact"] for r in cmp.values()) and info.iters == info1.iters and info.max_norm == info1.max_norm, **cmp}
        del one, vols, f_psi, f_inv, want
    del got, pg_full
    dist.barrier()
    return res


def traffic_child(args):
    """hidden mode (run under ncu by measure_traffic): a few solver iterations at --dim, nothing printed"""
    import torch
    import sobfu_b200 as sf
    torch.cuda.set_device(0)
    p = make_params(sf, args.dim, 6)
    fusion = sf.SobFusion(p)
    dev = torch.from_numpy(synth_depth(0).view(np.int16)).cuda()
    fusion(dev.view(torch.uint16))
    dev = torch.from_numpy(synth_depth(1).view(np.int16)).cuda()
    fusion(dev.view(torch.uint16))
    torch.cuda.synchronize()


def measure_traffic(dim):
    """DRAM bytes per launch of the two loop kernels, measured in THIS run: a child process of this script (6 solver iterations)
    under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`; the timed numbers never come from a profiled process."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:pass_[ab]_",
           "-s", "4", "-c", "6", sys.executable, os.path.abspath(__file__), "--traffic-child", "--dim", str(dim)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, SOBFU_B200_QUIET="1"))
    except Exception as e:      # noqa: BLE001
        return None, "ncu failed: %s" % e
    lines = r.stdout.splitlines()
    start = next((i for i, ln in enumerate(lines) if ln.startswith('"ID"')), None)
    if start is None:
        return None, "ncu produced no table (rc %d): %s" % (r.returncode, (r.stdout + r.stderr)[-300:].replace("\n", " | "))
    acc = {}
    for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        k = "pass_a" if "pass_a" in row.get("Kernel Name", "") else ("pass_b" if "pass_b" in row.get("Kernel Name", "") else None)
        if k is None or not row.get("Metric Value"):
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "byte").lower()
        v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        acc.setdefault(k, {}).setdefault(row["ID"], 0.0)
        acc[k][row["ID"]] += v
    out = {k: sum(v.values()) / len(v) for k, v in acc.items() if v}
    return (out or None), "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum in a child process of this run (%s)" % ", ".join(
        "%s: %d launches" % (k, len(v)) for k, v in acc.items())


def run_ours(args, rank, world, torch, dist):
    out, fusion = measure_solver(args, rank, world, torch, dist, args.dim, args.iters, args.steps, args.warmup)
    if not args.no_parity:
        par = parity_block(args, rank, world, torch, dist, fusion, args.dim, args.iters)
        if rank == 0:
            out["parity"] = par
    extra_dim = args.extra_dim if args.extra_dim is not None else (512 if world == 8 else 0)
    if extra_dim and extra_dim != args.dim:
        # BASELINE.json configs[3]: the 512^3 volume, params_umbrella.ini, z-slabbed over the GPUs of the box; fewer steps, the same
        # timing rules
        del fusion
        torch.cuda.empty_cache()
        ex, fusion = measure_solver(args, rank, world, torch, dist, extra_dim, args.iters, min(args.steps, 3), 3)
        if not args.no_parity:
            par = parity_block(args, rank, world, torch, dist, fusion, extra_dim, args.iters)
            if rank == 0:
                ex["parity"] = par
        if rank == 0:
            out["extra_%d" % extra_dim] = {k: ex[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "config", "solver_iters_per_s", "loop_ms_per_iter",
                                                               "kernel_ms", "iteration_roofline_frac", "pass_a_roofline_frac", "pass_b_roofline_frac", "e2e",
                                                               "gpu_launches", "parity") if k in ex}
    del fusion
    if rank == 0:
        if world == 1 and not args.no_traffic:
            tr, how = measure_traffic(args.dim)
            r = out["roofline"]
            dom = "pass_a" if r["kernel"].startswith("pass_a") else "pass_b"
            r["traffic_source"] = how
            if tr and dom in tr:
                r["traffic"] = tr[dom]
                r["dram_frac"] = tr[dom] / (out["kernel_ms"][dom] * 1e-3) / 1e9 / r["peak"]
                out["dram_bytes_per_launch"] = tr
                out["dram_frac"] = {k: tr[k] / (out["kernel_ms"][k] * 1e-3) / 1e9 / r["peak"] for k in tr if k in out["kernel_ms"]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out), flush=True)


# params/params_snoopy.ini of the reference (BASELINE.json configs[4]: 256^3, 50-frame sequence, marching cubes every frame);
# the truncation band is given in voxels and kept at snoopy's 10 / 5
SNOOPY = dict(vol_size=0.9, trunc_vox=10.0, eta_vox=5.0, max_weight=128.0, fx=517.0, fy=517.0, cx=320.0, cy=240.0, trunc_depth=3.0,
              pose_tz=0.05, sigma_depth=0.01, sigma_spatial=4.5, ksz=7, start_frame=4, max_update_norm=1e-3, s=7, lam=0.1, alpha=0.1, w_reg=0.2)


def run_pipeline(args, rank, world, torch, dist):
    """BASELINE.json configs[4]: the whole per-frame pipeline (depth preparation, TSDF integration, solver, fusion, marching cubes on
    phi_global every frame) over a synthetic sequence of a radially pulsating, translating sphere.  One step
    = one frame from pinned host memory; value = frames/s over the frames after START_FRAME (the ones that run the solver)."""
    import sobfu_b200 as sf
    from sobfu_b200.parallel import SlabFusion
    b, dim = SNOOPY, args.dim
    p = sf.Params(cols=COLS, rows=ROWS, volume_dims=(dim, dim, dim), volume_size=(b["vol_size"],) * 3, intr=sf.Intr(b["fx"], b["fy"], b["cx"], b["cy"]),
                  icp_truncate_depth_dist=b["trunc_depth"], bilateral_sigma_depth=b["sigma_depth"], bilateral_sigma_spatial=b["sigma_spatial"],
                  bilateral_kernel_size=b["ksz"], tsdf_max_weight=b["max_weight"], gradient_delta_factor=0.5, start_frame=b["start_frame"], verbosity=0,
                  s=b["s"], max_iter=args.iters, max_update_norm=b["max_update_norm"], lambda_=b["lam"], alpha=b["alpha"], w_reg=b["w_reg"])
    vs = p.voxel_sizes()
    p.tsdf_trunc_dist, p.eta = float(np.float32(b["trunc_vox"]) * vs[0]), float(np.float32(b["eta_vox"]) * vs[0])
    p.volume_pose = sf.Affine3f().translate((-b["vol_size"] / 2, -b["vol_size"] / 2, b["pose_tz"]))
    fusion = sf.SobFusion(p) if world == 1 else SlabFusion(p, dist)
    nframes = args.frames
    frames = [torch.from_numpy(synth_depth(f, radius=0.15 + 0.01 * np.sin(2 * np.pi * f / 25.0), intr=b).view(np.int16)).pin_memory() for f in range(nframes)]
    dev_depth = torch.empty((ROWS, COLS), dtype=torch.int16, device="cuda")
    nverts, iters_run = [], []

    def frame_step(f):
        dev_depth.copy_(frames[f], non_blocking=True)
        fusion(dev_depth.view(torch.uint16))
        mesh = fusion.get_phi_global_mesh()                     # marching cubes on the canonical model, every frame
        nverts.append(int(mesh[3]) if world > 1 else int(mesh[0].shape[0]))
        if fusion.solver is not None and fusion.solver.info is not None and f >= b["start_frame"]:
            iters_run.append(int(fusion.solver.info.iters))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = b["start_frame"] + max(1, args.warmup - 2)          # rigid frames + the first solver frames (allocations, tensor maps)
    for f in range(warm):
        frame_step(f)
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sync_all()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(warm, nframes):
        frame_step(f)
    e1.record()
    sync_all()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    timed = nframes - warm
    if rank == 0:
        it = iters_run[-timed:]
        print(json.dumps({
            "metric": "pipeline_frames_per_s", "value": timed / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": timed, "warmup": warm,
            "ms_per_step": ms / timed, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d^3 volume, params_snoopy.ini (dims->%d, MAX_ITER->%d), %d-frame synthetic sequence (pulsating, translating sphere), "
                                   "marching cubes on phi_global every frame; 1 step = 1 frame from pinned host memory" % (dim, dim, args.iters, nframes),
                       "parallelism": "1 GPU" if world == 1 else "z-slab x%d (solver, fusion and marching cubes per slab)" % world},
            "solver_iterations_per_frame": {"mean": float(np.mean(it)) if it else 0.0, "min": int(min(it)) if it else 0, "max": int(max(it)) if it else 0},
            "mesh_vertices_last_frame": nverts[-1], "clocks": clk,
            "e2e": {"value": timed / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": COLS * ROWS * 2, "d2h_bytes_per_step": 4 + 16 + 8},
        }), flush=True)


def _silenced(fn):
    """the reference prints from inside the solver loop (solver.cu:115-117): keep stdout for the JSON line"""
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        return fn()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def measure_reference(dim, iters, steps, warmup):
    """solver workload through the unmodified reference CUDA and its own host classes: (line as a dict)"""
    from oracle import pyoracle as orc
    import torch
    b, ini = params_for(dim)
    vs = np.float32(b["vol_size"]) / np.float32(dim)
    ref = orc.Reference((dim,) * 3, (b["vol_size"],) * 3, float(np.float32(b["trunc_vox"]) * vs), float(np.float32(b["eta_vox"]) * vs),
                        b["max_weight"], 0, iters, b["s"], b["max_update_norm"], b["lam"], b["alpha"], b["w_reg"],
                        pose_t=(-b["vol_size"] / 2, -b["vol_size"] / 2, b["pose_tz"]), intr=(b["fx"], b["fy"], b["cx"], b["cy"]))
    frames = [synth_depth(f) for f in range(1 + 2 * (warmup + steps))]

    def frame_step(f):     # SobFusion::operator(), sob_fusion.cpp:71-145, through the reference's own classes
        ref.depth_to_dists(frames[f], b["ksz"], b["sigma_spatial"], b["sigma_depth"], b["trunc_depth"])
        if f == 0:
            ref.integrate_dists(ref.GLOBAL)
            return
        ref.tsdf_clear(ref.N)
        ref.integrate_dists(ref.N)
        ref.estimate_psi()
        ref.fuse(ref.GLOBAL, ref.N_PSI)

    def body():
        frame_step(0)
        f = 1
        for _ in range(warmup):
            frame_step(f); f += 1
        clocks = ClockSampler(0)
        torch.cuda.synchronize()
        clocks.start()
        ms = sum(ref.estimate_psi() for _ in range(steps))      # cudaEvent pair around Solver::estimate_psi
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            frame_step(f); f += 1
        torch.cuda.synchronize()
        return ms, (time.perf_counter() - t0) * 1e3, clocks.stop()

    ms, ms_e2e, clk = _silenced(body)
    ref.close()
    N = dim ** 3
    v = N * iters * steps / (ms * 1e-3) / 1e9
    return {
        "impl": "reference", "metric": "solver_gvoxel_iters_per_s", "value": v, "unit": "Gvoxel-iter/s", "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d^3 volume, %s (dims->%d, MAX_ITER->%d), 640x480 synthetic sphere depth, 1 step = 1 frame = "
                               "estimate_psi with %d iterations" % (dim, ini, dim, iters, iters), "parallelism": "1 GPU (the reference is single-GPU)"},
        "solver_iters_per_s": iters * steps / (ms * 1e-3), "clocks": clk,
        "cpu_baseline": {"value": v, "unit": "Gvoxel-iter/s", "cores": os.cpu_count(), "kind": "reference",
                         "sample": "the reference has no CPU solver path: this is its own CUDA (sm_100a build of the unmodified sources) on one B200"},
        "e2e": {"value": N * iters * steps / (ms_e2e * 1e-3) / 1e9, "unit": "Gvoxel-iter/s", "frames_per_s": steps / (ms_e2e * 1e-3),
                "h2d_bytes_per_step": COLS * ROWS * 2, "d2h_bytes_per_step": 0},
    }


def reference_pipeline(args):
    """BASELINE.json configs[4] through the reference's own classes: SobFusion::operator() (sob_fusion.cpp:71-145) + marching cubes on
    phi_global every frame (kfusion::cuda::MarchingCubes::run, the triangles stay on the device as in our arm)"""
    from oracle import pyoracle as orc
    import torch
    b, dim = SNOOPY, args.dim
    vs = np.float32(b["vol_size"]) / np.float32(dim)
    ref = orc.Reference((dim,) * 3, (b["vol_size"],) * 3, float(np.float32(b["trunc_vox"]) * vs), float(np.float32(b["eta_vox"]) * vs),
                        b["max_weight"], 0, args.iters, b["s"], b["max_update_norm"], b["lam"], b["alpha"], b["w_reg"],
                        pose_t=(-b["vol_size"] / 2, -b["vol_size"] / 2, b["pose_tz"]), intr=(b["fx"], b["fy"], b["cx"], b["cy"]))
    nframes = args.frames
    frames = [synth_depth(f, radius=0.15 + 0.01 * np.sin(2 * np.pi * f / 25.0), intr=b) for f in range(nframes)]
    nverts = []

    def frame_step(f):
        ref.depth_to_dists(frames[f], b["ksz"], b["sigma_spatial"], b["sigma_depth"], b["trunc_depth"])
        if f == 0:
            ref.integrate_dists(ref.GLOBAL)
        else:
            ref.tsdf_clear(ref.N)
            ref.integrate_dists(ref.N)
            if f < b["start_frame"]:
                ref.fuse(ref.GLOBAL, ref.N)
            else:
                ref.estimate_psi()
                ref.fuse(ref.GLOBAL, ref.N_PSI)
        nverts.append(ref.marching_cubes_count(ref.GLOBAL))

    warm = b["start_frame"] + max(1, args.warmup - 2)

    def body():
        for f in range(warm):
            frame_step(f)
        clocks = ClockSampler(0)
        torch.cuda.synchronize()
        clocks.start()
        t0 = time.perf_counter()
        for f in range(warm, nframes):
            frame_step(f)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, clocks.stop()

    ms, clk = _silenced(body)
    ref.close()
    timed = nframes - warm
    return {
        "impl": "reference", "metric": "pipeline_frames_per_s", "value": timed / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": timed, "warmup": warm,
        "ms_per_step": ms / timed, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d^3 volume, params_snoopy.ini (dims->%d, MAX_ITER->%d), %d-frame synthetic sequence (pulsating, translating sphere), "
                               "marching cubes on phi_global every frame; 1 step = 1 frame from pinned host memory" % (dim, dim, args.iters, nframes),
                   "parallelism": "1 GPU (the reference is single-GPU)"},
        "mesh_vertices_last_frame": nverts[-1], "clocks": clk,
        "cpu_baseline": {"value": timed / (ms * 1e-3), "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
                         "sample": "the reference has no CPU path: its own CUDA (sm_100a build of the unmodified sources) on one B200, wall clock over %d frames" % timed},
        "e2e": {"value": timed / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": COLS * ROWS * 2, "d2h_bytes_per_step": 4},
    }


def run_reference(args, rank, world):
    """the unmodified reference CUDA through its own host API, same frames / same parameters (rank 0 only)"""
    if rank != 0:
        return
    from oracle import pyoracle as orc
    if not os.path.exists(orc.REF):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsobfu_ref.so was not built (needs /root/reference at build time)"}))
        return
    if args.workload == "pipeline":
        print(json.dumps(reference_pipeline(args)), flush=True)
        return
    out = measure_reference(args.dim, args.iters, args.steps, args.warmup)
    extra_dim = args.extra_dim if args.extra_dim is not None else (512 if args.gpus == 8 else 0)
    if extra_dim and extra_dim != args.dim:      # BASELINE.json configs[3]: the reference fits one B200 at 512^3 (304 B/voxel = 40.8 GB)
        ex = measure_reference(extra_dim, args.iters, min(args.steps, 2), 1)
        out["extra_%d" % extra_dim] = {k: ex[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "config", "solver_iters_per_s", "e2e")}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="solver", choices=["solver", "pipeline"],
                    help="solver: BASELINE.json configs[2] (default, the headline metric); pipeline: configs[4], the per-frame pipeline incl. marching cubes")
    ap.add_argument("--frames", type=int, default=50, help="pipeline workload: length of the synthetic sequence")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant of the solver (0 default; 1 generic; 2 tiled; 4 tiled with the round-1 pass A)")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing parity check (reference CUDA at N=1, single-GPU solve at N>1)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child process that measures DRAM bytes per launch (N=1)")
    ap.add_argument("--extra-dim", type=int, default=None, help="also measure this volume size and report it under extra_<dim> (default: 512 when --gpus 8)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.traffic_child:
        return traffic_child(args)
    args.warmup = max(args.warmup, 3)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            import torch
            torch.cuda.set_device(0)
        return run_reference(args, rank, world)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    if args.workload == "pipeline":
        run_pipeline(args, rank, world, torch, dist)
    else:
        run_ours(args, rank, world, torch, dist)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
