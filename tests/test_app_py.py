"""sobfu_b200/app.py, the z-slab capable driver: its host-side pieces on CPU (the .ini reader against the C++ reader's output on
the same file, depth / mask PNGs through the library's I/O entries against PNGs written here with zlib and every filter, the VTK
writer against the C++ writer) and -- opt-in until it has run on hardware -- the whole driver against the C++ application."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.test_io_cpu import INI, write_png

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ini_reader_matches_the_cpp_reader(built, tmp_path):
    from sobfu_b200 import app
    p = tmp_path / "p.ini"
    p.write_text(INI)
    kv = app.read_ini(str(p))
    assert kv["VOL_DIMS_X"] == 96 and kv["VOL_DIMS_Z"] == 64 and kv["VOL_SIZE_Y"] == 0.75 and kv["ALPHA"] == 0.1 and kv["MAX_ITER"] == 2048
    prm = app.params_from_ini(str(p), verbosity=1)
    vs = np.float32(0.9) / np.float32(96)
    assert prm.volume_dims == (96, 80, 64) and prm.verbosity == 1 and prm.start_frame == 4 and prm.s == 7
    assert prm.tsdf_trunc_dist == float(np.float32(10) * vs) and prm.eta == float(np.float32(5) * vs)
    assert np.array_equal(prm.volume_pose.t, np.array([-np.float32(0.9) / 2, -np.float32(0.75) / 2, np.float32(0.05)], dtype=np.float32))
    for bad, what in ((INI + "RHO_0=1.0\n", "RHO_0"), (INI.replace("MAX_ITER=2048", "MAX_ITER=20.5"), "MAX_ITER"), ("VOL_DIMS_X 96\n", "invalid line")):
        p.write_text(bad)
        with pytest.raises(app.AppError, match=what):
            app.read_ini(str(p))
    p.write_text("VOL_DIMS_X=8\n")
    with pytest.raises(app.AppError, match="missing option"):
        app.params_from_ini(str(p))


def test_depth_and_mask_files(built, tmp_path):
    from sobfu_b200 import app
    rng = np.random.RandomState(1)
    depth = rng.randint(0, 65536, size=(48, 64)).astype(np.uint16)
    depth[:, :30] = 700
    p = str(tmp_path / "d.png")
    write_png(p, depth[..., None], 16, 0, level=9)                      # zlib level 9, filters 0..4 cycling
    assert np.array_equal(app.read_depth(p), depth)
    mask = (rng.rand(48, 64) > 0.5).astype(np.uint8) * 255
    write_png(p, mask[..., None], 8, 0)
    assert np.array_equal(app.read_mask(p), mask)
    with pytest.raises(app.AppError, match="16-bit"):
        app.read_depth(p)                                               # an 8-bit file is not a depth map
    with pytest.raises(app.AppError, match="could not be read"):
        app.read_depth(str(tmp_path / "missing.png"))
    app.write_depth(p, depth)
    assert np.array_equal(app.read_depth(p), depth)


def test_vtk_writer_is_the_cpp_writer(built, tmp_path):
    from sobfu_b200 import app
    exe = str(tmp_path / "io_tool")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "include", "compat"),
                           os.path.join(ROOT, "tests", "cpp", "io_tool.cpp"), "-o", exe])
    subprocess.check_call([exe, "vtk", str(tmp_path / "cpp.vtk")])
    v = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1], [0.5, 0.25, -1.5, 1], [1e-3, 123456.789, 2, 1], [3, 2, 1, 1]], dtype=np.float32)
    app.write_vtk(str(tmp_path / "py.vtk"), v)
    assert open(str(tmp_path / "py.vtk")).read() == open(str(tmp_path / "cpp.vtk")).read()


def test_command_line(built):
    from sobfu_b200 import app
    o = app.parse_args(["--enable-log", "--vverbose", "/data/seq", "p.ini", "--frames", "7"])
    assert o["file_path"] == "/data/seq" and o["params_path"] == "p.ini" and o["logger"] and o["verbosity"] == 2 and o["frames"] == 7
    o = app.parse_args(["--synthetic", "5", "p.ini", "--json"])
    assert o["synthetic"] == 5 and o["params_path"] == "p.ini" and o["file_path"] is None and o["json"]
    with pytest.raises(app.AppError):
        app.parse_args(["only_one"])


@pytest.mark.gpu
def test_python_driver_matches_the_cpp_application(built, tmp_path):
    from tests.test_app_gpu import APP, make_sequence
    root = str(tmp_path / "seq")
    frames, ini = make_sequence(root)
    env = dict(os.environ, SOBFU_B200_QUIET="1")
    r = subprocess.run([APP, "--enable-log", "--out", str(tmp_path / "cpp"), root, ini], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-m", "sobfu_b200.app", "--enable-log", "--json", "--out", str(tmp_path / "py"), root, ini], capture_output=True,
                       text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["frames"] == len(frames) and info["n_gpus"] == 1
    for name in sorted(os.listdir(str(tmp_path / "cpp"))):
        assert open(os.path.join(str(tmp_path / "cpp"), name)).read() == open(os.path.join(str(tmp_path / "py"), name)).read(), name
