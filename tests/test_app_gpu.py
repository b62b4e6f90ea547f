"""The application layer on the GPU: a synthetic VolumeDeform-style sequence (depth/ and color/ PNGs,
an .ini in the reference's format) is run through
  * apps/sobfu_headless.cpp (first party, C++ host over the C ABI), and
  * the reference's own src/apps/demo.cpp compiled UNCHANGED against include/ (oracle/_ref/sobfu_app_dropin, when built),
and the meshes they log are compared with the Python mirror of the same pipeline on the same frames."""
import json
import os
import subprocess

import numpy as np
import pytest

import bench
from tests.test_io_cpu import write_png

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "sobfu_b200", "_lib", "sobfu_headless")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "sobfu_app_dropin")

INI = """# synthetic sequence, 64^3
VOL_DIMS_X=64
VOL_DIMS_Y=64
VOL_DIMS_Z=64
VOL_SIZE_X=0.75
VOL_SIZE_Y=0.75
VOL_SIZE_Z=0.75
TSDF_TRUNC_DIST=6
ETA=3
TSDF_MAX_WEIGHT=128
GRADIENT_DELTA_FACTOR=0.5
INTR_FX=570.342
INTR_FY=570.342
INTR_CX=320.0
INTR_CY=240.0
TRUNC_DEPTH=1.0
VOL_POSE_T_Z=0.1
BILATERAL_SIGMA_DEPTH=0.005
BILATERAL_SIGMA_SPATIAL=4.5
BILATERAL_KERNEL_SIZE=7
START_FRAME=1
MAX_ITER=12
MAX_UPDATE_NORM=1e-10
S=7
LAMBDA=0.1
ALPHA=0.05
W_REG=0.3
"""
NFRAMES = 3


def test_applications_are_built(built):
    assert os.path.exists(APP)
    if os.path.isdir("/root/reference/src/apps"):
        assert os.path.exists(DROPIN)
        syms = subprocess.run(["nm", "-D", "--undefined-only", DROPIN], capture_output=True, text=True).stdout
        assert "sobfu_b200_solver_estimate_psi" in syms and "sobfu_b200_marching_cubes" in syms    # reference app -> our C ABI


def make_sequence(root):
    os.makedirs(os.path.join(root, "depth"))
    os.makedirs(os.path.join(root, "color"))
    frames = []
    for f in range(NFRAMES):
        d = bench.synth_depth(2 * f)
        frames.append(d)
        write_png(os.path.join(root, "depth", "depth_%06d.png" % f), d[..., None], 16, 0, level=6)
        write_png(os.path.join(root, "color", "color_%06d.png" % f), np.full((8, 8, 3), 40 * f, np.uint8), 8, 2)
    ini = os.path.join(root, "params.ini")
    open(ini, "w").write(INI)
    return frames, ini


def read_vtk(path):
    lines = open(path).read().split("\n")
    assert lines[3] == "DATASET POLYDATA"
    n = int(lines[4].split()[1])
    pts = np.array([[float(v) for v in l.split()] for l in lines[5:5 + n]], dtype=np.float64)
    j = [i for i, l in enumerate(lines) if l.startswith("POLYGONS")][0]
    npoly = int(lines[j].split()[1])
    tri = np.array([[int(v) for v in l.split()] for l in lines[j + 1:j + 1 + npoly]])
    return pts, tri


def python_pipeline(frames):
    import torch
    import sobfu_b200 as sf
    p = sf.Params(cols=640, rows=480, volume_dims=(64, 64, 64), volume_size=(0.75, 0.75, 0.75), intr=sf.Intr(570.342, 570.342, 320.0, 240.0),
                  icp_truncate_depth_dist=1.0, bilateral_sigma_depth=0.005, bilateral_sigma_spatial=4.5, bilateral_kernel_size=7, tsdf_max_weight=128.0,
                  gradient_delta_factor=0.5, start_frame=1, verbosity=0, s=7, max_iter=12, max_update_norm=1e-10, lambda_=0.1, alpha=0.05, w_reg=0.3)
    vs = p.voxel_sizes()
    p.tsdf_trunc_dist, p.eta = float(np.float32(6) * vs[0]), float(np.float32(3) * vs[0])
    p.volume_pose = sf.Affine3f().translate((-0.75 / 2, -0.75 / 2, 0.1))
    fusion = sf.SobFusion(p)
    meshes = []
    for d in frames:
        fusion(torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16))
        meshes.append(fusion.get_phi_global_mesh()[0].cpu().numpy())
    return meshes


def check_logged_meshes(out_dir, meshes):
    for f, want in enumerate(meshes):
        pts, tri = read_vtk(os.path.join(out_dir, "canonical_mesh_%06d.vtk" % f))
        assert len(pts) == len(want) and len(want) > 1000, (f, len(pts), len(want))
        assert np.abs(pts - want[:, :3]).max() < 2e-5 * max(1.0, np.abs(want[:, :3]).max())       # 5 significant digits in the file
        assert np.array_equal(tri[:, 0], np.full(len(tri), 3)) and np.array_equal(tri[:, 1:].ravel(), np.arange(len(pts)))
        if f >= 1:
            assert os.path.getsize(os.path.join(out_dir, "canonical_warped_to_live_mesh_%06d.vtk" % f)) > 1000


@pytest.mark.gpu
def test_headless_app_matches_the_python_pipeline(built, tmp_path):
    root = str(tmp_path / "seq")
    frames, ini = make_sequence(root)
    meshes = python_pipeline(frames)
    r = subprocess.run([APP, "--enable-log", "--save-field", "--json", root, ini], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, SOBFU_B200_QUIET="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["frames"] == NFRAMES and info["vertices"] == len(meshes[-1]) and info["volume"] == [64, 64, 64]
    assert r.stdout.count("--- FRAME NO.") == NFRAMES and "saved canonical_mesh_000002.vtk" in r.stdout
    check_logged_meshes(os.path.join(root, "meshes"), meshes)
    blob = open(os.path.join(root, "meshes", "field_000002.vti"), "rb").read()
    assert b'WholeExtent="0 63 0 63 0 63"' in blob and len(blob) > 64 ** 3 * 16
    # synthetic source: no input directory
    r = subprocess.run([APP, "--synthetic", "2", "--json", ini], capture_output=True, text=True, timeout=600, env=dict(os.environ, SOBFU_B200_QUIET="1"))
    assert r.returncode == 0 and json.loads(r.stdout.strip().splitlines()[-1])["frames"] == 2, r.stdout[-2000:] + r.stderr[-2000:]
    # errors follow the reference's wording and exit codes are non-zero
    r = subprocess.run([APP, str(tmp_path / "nowhere"), ini], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "does not exist" in r.stderr
    bad = str(tmp_path / "bad.ini")
    open(bad, "w").write(INI + "RHO_0=1\n")
    r = subprocess.run([APP, root, bad], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "RHO_0" in r.stderr


@pytest.mark.gpu
def test_reference_app_source_runs_on_our_library(built, tmp_path):
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/sobfu_app_dropin not built (needs /root/reference at build time)")
    root = str(tmp_path / "seq")
    frames, ini = make_sequence(root)
    meshes = python_pipeline(frames)
    r = subprocess.run([DROPIN, "--enable-log", root, ini], capture_output=True, text=True, timeout=600, env=dict(os.environ, SOBFU_B200_QUIET="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "created output directory for meshes" in r.stdout and "saved canonical_warped_to_live_mesh_000002.vtk" in r.stdout
    assert "no. of point-normal pairs in the canonical model: %d" % len(meshes[0]) in r.stdout
    check_logged_meshes(os.path.join(root, "meshes"), meshes)
