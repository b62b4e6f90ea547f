#!/usr/bin/env python3
"""Generates the golden vectors of tests/golden/ by running the UNMODIFIED reference CUDA (oracle/_ref/libsobfu_ref.so,
built from /root/reference by oracle/build_ref.sh) on a GPU.  Test infrastructure.

    gpurun -- python oracle/make_golden.py gpurun_out/golden        # then copy gpurun_out/golden/*.npz to tests/golden/

The reference ships no golden files for this path, so its own CUDA output on fixed synthetic inputs
is the pin: tests/test_golden.py checks the CPU oracle (no GPU needed) and the sm_100a kernels (GPU) against these files.
Inputs are produced by the reference's own TsdfVolume::initSphere (approximate GPU units -> stored, not regenerated).
Small cases store full output volumes; larger ones store SHA-256 digests of the output bytes.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402
from tests.common import wavy_psi  # noqa: E402

CASES = [
    # name, dims, size, trunc_vox, eta_vox, iters, lambda, alpha, w_reg, centre_g, centre_n, radius, wavy, full
    ("c16_id", (16, 16, 16), 0.25, 5.0, 2.0, 6, 0.1, 0.01, 0.4, (0.125, 0.125, 0.125), (0.117, 0.125, 0.125), 0.06, 0.0, True),
    ("c20x17x13_wavy", (20, 17, 13), 0.25, 4.0, 2.0, 5, 0.2, 0.02, 0.2, (0.13, 0.12, 0.11), (0.12, 0.125, 0.115), 0.05, 0.4, True),
    ("c32_cfg1", (32, 32, 32), 0.25, 5.0, 2.0, 5, 0.1, 0.01, 0.4, (0.125, 0.125, 0.125), (0.117, 0.125, 0.125), 0.06, 0.0, False),
    ("c64_fixture", (64, 64, 64), 0.25, 10.0, 2.0, 20, 0.1, 0.01, 0.4, (0.13, 0.13, 0.13), (0.125, 0.13, 0.13), 0.012, 0.0, False),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for (name, dims, size, tv, ev, iters, lam, alpha, w_reg, cg, cn, rad, wavy, full) in CASES:
        vs = np.float32(size) / np.float32(dims[0])
        trunc, eta = np.float32(tv) * vs, np.float32(ev) * vs
        ref = orc.Reference(dims, (size, size, size), float(trunc), float(eta), 64.0, 2, iters, 7, -1.0, lam, alpha, w_reg)
        ref.init_sphere(ref.GLOBAL, cg, rad)
        ref.init_sphere(ref.N, cn, rad)
        pg, pn = ref.download_tsdf(ref.GLOBAL), ref.download_tsdf(ref.N)
        psi0 = wavy_psi(dims, amp=wavy) if wavy else orc.init_identity(*dims)
        ref.upload_psi(0, psi0)
        grad_n = ref.tsdf_gradient(ref.N)
        ref.estimate_psi()
        out = dict(psi=ref.download_psi(0), psi_inv=ref.download_psi(1), phi_n_psi=ref.download_tsdf(ref.N_PSI),
                   phi_global_psi_inv=ref.download_tsdf(ref.GLOBAL_PSI_INV))
        lap = ref.laplacian()
        jac0, jac1 = ref.jacobian(0), ref.jacobian(1)
        e_data = ref.data_energy(ref.GLOBAL, ref.N_PSI)
        ref.close()
        rec = dict(dims=np.array(dims), size=np.float32(size), trunc=trunc, eta=eta, iters=iters, lam=np.float32(lam),
                   alpha=np.float32(alpha), w_reg=np.float32(w_reg), phi_global=pg, phi_n=pn, psi0_wavy=np.float32(wavy),
                   e_data=np.float32(e_data), full=full)
        for k, v in out.items():
            rec["sha_" + k] = sha(v)
            if full:
                rec[k] = v
        rec["sha_grad_n"], rec["sha_lap"], rec["sha_jac0"], rec["sha_jac1"] = sha(grad_n), sha(lap), sha(jac0[..., :3, :]), sha(jac1[..., :3, :])
        if full:
            rec["grad_n"], rec["lap"] = grad_n, lap
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **rec)
        # immediate cross-check of the CPU oracle against the reference (printed, not asserted)
        o = orc.estimate_psi(pg, pn, psi0, iters, -1.0, 7, lam, alpha, w_reg)
        msg = ["%s: %s" % (k, "bit-exact" if sha(o[k]) == rec["sha_" + k] else "DIFF max|d|=%g" % np.abs(o[k] - out[k]).max()) for k in out]
        msg.append("grad: %s" % ("bit-exact" if sha(orc.tsdf_gradient(pn)) == rec["sha_grad_n"] else "DIFF"))
        msg.append("lap: %s" % ("bit-exact" if sha(orc.laplacian(out["psi"])) == rec["sha_lap"] else "DIFF"))
        msg.append("jac1: %s" % ("bit-exact" if sha(orc.jacobian(out["psi"], 1)[..., :3, :]) == rec["sha_jac1"] else "DIFF"))
        msg.append("e_data ref %.9g oracle %.9g" % (e_data, orc.data_energy(pg, out["phi_n_psi"])))
        print("[golden] %-16s oracle vs reference CUDA -> %s" % (name, "; ".join(msg)), flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
