"""sobfu application, z-slab capable: `python -m sobfu_b200.app [OPTIONS] <file path> <ini path>` on one GPU, or under
`python -m torch.distributed.run --nproc-per-node N -m sobfu_b200.app ...` with the volume partitioned over N GPUs.

Same command line, directory layout (<file path>/depth/*.png, optional <file path>/omask/*.png), .ini option set and console
messages as the reference's src/apps/demo.cpp:30-618 and as apps/sobfu_headless.cpp (the single-GPU C++ host); the extra
flags are --synthetic N, --frames K, --out DIR, --json.  With --enable-log rank 0 writes canonical_mesh_XXXXXX.vtk /
canonical_warped_to_live_mesh_XXXXXX.vtk (legacy VTK polydata, the layout of pcl::io::saveVTKFile) from the meshes the ranks
extract per slab.

Host plumbing only: files go through the library's host-only I/O entries (the code the C++ applications use), all voxel work
through the C ABI (api.py / parallel.py).
"""
import json
import os
import sys
import time

import numpy as np

# (name, type) of every option the application declares (demo.cpp:84-160); anything else in the file is an error, as with
# boost::program_options::parse_config_file
INI_OPTIONS = {
    "VOL_DIMS_X": int, "VOL_DIMS_Y": int, "VOL_DIMS_Z": int, "VOL_SIZE_X": float, "VOL_SIZE_Y": float, "VOL_SIZE_Z": float,
    "TSDF_TRUNC_DIST": float, "ETA": float, "TSDF_MAX_WEIGHT": float, "GRADIENT_DELTA_FACTOR": float,
    "INTR_FX": float, "INTR_FY": float, "INTR_CX": float, "INTR_CY": float, "TRUNC_DEPTH": float, "VOL_POSE_T_Z": float,
    "BILATERAL_SIGMA_DEPTH": float, "BILATERAL_SIGMA_SPATIAL": float, "BILATERAL_KERNEL_SIZE": int,
    "START_FRAME": int, "MAX_ITER": int, "MAX_UPDATE_NORM": float, "S": int, "LAMBDA": float, "ALPHA": float, "W_REG": float,
}
REQUIRED = ("VOL_DIMS_X", "VOL_DIMS_Y", "VOL_DIMS_Z", "VOL_SIZE_X", "VOL_SIZE_Y", "VOL_SIZE_Z", "TSDF_TRUNC_DIST", "ETA", "VOL_POSE_T_Z",
            "MAX_ITER", "S", "LAMBDA", "ALPHA", "W_REG")


class AppError(RuntimeError):
    pass


def read_ini(path):
    """NAME=VALUE per line, '#' comments, first occurrence wins, unknown names and unparsable values are errors"""
    out = {}
    try:
        lines = open(path).read().splitlines()
    except OSError:
        raise AppError("cannot open '%s'" % path)
    for line in lines:
        line = line.split("#", 1)[0].strip()
        if not line:
            continue
        if "=" not in line:
            raise AppError("%s: the options configuration file contains an invalid line '%s'" % (path, line))
        name, val = [t.strip() for t in line.split("=", 1)]
        if name not in INI_OPTIONS:
            raise AppError("%s: unrecognised option '%s'" % (path, name))
        if name in out:
            continue
        try:
            out[name] = INI_OPTIONS[name](val)
        except ValueError:
            raise AppError("%s: the argument ('%s') for option '%s' is invalid" % (path, val, name))
    return out


def params_from_ini(path, verbosity=0):
    """sobfu Params from a reference-format .ini (demo.cpp:41-74): TSDF_TRUNC_DIST and ETA are given in voxels"""
    import sobfu_b200 as sf
    kv = read_ini(path)
    for k in REQUIRED:
        if k not in kv:
            raise AppError("%s: missing option '%s'" % (path, k))
    f32 = np.float32
    p = sf.Params(volume_dims=(kv["VOL_DIMS_X"], kv["VOL_DIMS_Y"], kv["VOL_DIMS_Z"]),
                  volume_size=tuple(float(f32(kv[k])) for k in ("VOL_SIZE_X", "VOL_SIZE_Y", "VOL_SIZE_Z")),
                  intr=sf.Intr(kv.get("INTR_FX", 0.0), kv.get("INTR_FY", 0.0), kv.get("INTR_CX", 0.0), kv.get("INTR_CY", 0.0)),
                  icp_truncate_depth_dist=kv.get("TRUNC_DEPTH", 0.0), bilateral_sigma_depth=kv.get("BILATERAL_SIGMA_DEPTH", 0.0),
                  bilateral_sigma_spatial=kv.get("BILATERAL_SIGMA_SPATIAL", 0.0), bilateral_kernel_size=kv.get("BILATERAL_KERNEL_SIZE", 0),
                  tsdf_max_weight=kv.get("TSDF_MAX_WEIGHT", 0.0), gradient_delta_factor=kv.get("GRADIENT_DELTA_FACTOR", 0.0),
                  start_frame=kv.get("START_FRAME", 0), verbosity=verbosity, s=kv["S"], max_iter=kv["MAX_ITER"],
                  max_update_norm=kv.get("MAX_UPDATE_NORM", 0.0), lambda_=kv["LAMBDA"], alpha=kv["ALPHA"], w_reg=kv["W_REG"])
    vs = p.voxel_sizes()
    p.tsdf_trunc_dist = float(f32(kv["TSDF_TRUNC_DIST"]) * vs[0])
    p.eta = float(f32(kv["ETA"]) * vs[0])
    p.volume_pose = sf.Affine3f().translate((-f32(p.volume_size[0]) / f32(2), -f32(p.volume_size[1]) / f32(2), f32(kv["VOL_POSE_T_Z"])))
    return p


# ---- files: through the library's host-only entries (include/sobfu_b200_io.hpp), the code the C++ applications use -----------
def _io_check(rc):
    if rc != 0:
        from ._capi import lib
        raise AppError(lib().sobfu_b200_io_last_error().decode())


def _read_png(fn, dtype, path):
    import ctypes as C
    cols, rows = C.c_int(), C.c_int()
    _io_check(fn(path.encode(), None, 0, C.byref(cols), C.byref(rows)))
    out = np.empty((rows.value, cols.value), dtype=dtype)
    _io_check(fn(path.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(cols), C.byref(rows)))
    return out


def read_depth(path):
    """16-bit depth map in millimetres: cv::imread(path, CV_LOAD_IMAGE_ANYDEPTH) of demo.cpp:301"""
    from ._capi import lib
    return _read_png(lib().sobfu_b200_read_depth_png, np.uint16, path)


def read_mask(path):
    """8-bit object mask: cv::imread(path, CV_8U) of demo.cpp:303"""
    from ._capi import lib
    return _read_png(lib().sobfu_b200_read_mask_png, np.uint8, path)


def write_depth(path, depth):
    import ctypes as C
    from ._capi import lib
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    _io_check(lib().sobfu_b200_write_depth_png(path.encode(), d.ctypes.data_as(C.c_void_p), d.shape[1], d.shape[0]))


def write_vtk(path, vertices):
    """legacy ASCII polydata as pcl::io::saveVTKFile lays it out: vertices [n, >= 3] float32, consecutive triples are triangles"""
    import ctypes as C
    from ._capi import lib
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    _io_check(lib().sobfu_b200_write_vtk_mesh(path.encode(), v.ctypes.data_as(C.c_void_p), v.shape[0], v.shape[1]))


def synthetic_depth(frame, params):
    """analytically ray-cast sphere of radius 0.15 m centred at (0.002 * frame, 0, 0.5) m through the .ini's intrinsics"""
    k = params.intr
    u, v = np.meshgrid(np.arange(params.cols, dtype=np.float64), np.arange(params.rows, dtype=np.float64))
    dx, dy = (u - k.cx) / k.fx, (v - k.cy) / k.fy
    c = np.array([0.002 * frame, 0.0, 0.5])
    a = dx * dx + dy * dy + 1.0
    b = -2.0 * (dx * c[0] + dy * c[1] + c[2])
    disc = b * b - 4 * a * (float(c @ c) - 0.15 ** 2)
    t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0.0)
    return np.ascontiguousarray(np.where(disc > 0, np.floor(t * 1000.0 + 0.5), 0).astype(np.uint16))


def list_files(directory):
    return sorted(os.path.join(directory, n) for n in os.listdir(directory) if os.path.isfile(os.path.join(directory, n)))


def parse_args(argv):
    o = dict(file_path=None, params_path=None, out=None, logger=False, viz=False, verbosity=0, synthetic=0, frames=-1, json=False)
    pos, i = [], 0
    while i < len(argv):
        a = argv[i]
        if a in ("-h", "--help"):
            print(__doc__)
            raise SystemExit(0)
        elif a == "--enable-log": o["logger"] = True
        elif a in ("--enable-viz", "--enable-viz-detailed"): o["viz"] = True
        elif a == "--verbose": o["verbosity"] = 1
        elif a == "--vverbose": o["verbosity"] = 2
        elif a == "--json": o["json"] = True
        elif a in ("--synthetic", "--frames") and i + 1 < len(argv):
            o[a[2:]] = int(argv[i + 1]); i += 1
        elif a == "--out" and i + 1 < len(argv):
            o["out"] = argv[i + 1]; i += 1
        else:
            pos.append(a)
        i += 1
    if o["synthetic"] > 0 and len(pos) == 1:
        o["params_path"] = pos[0]
    elif len(pos) >= 2:
        o["file_path"], o["params_path"] = pos[0], pos[1]
    else:
        raise AppError("incorrect number of arguments; please supply path to source data and .ini file")
    return o


def main(argv=None):
    o = parse_args(sys.argv[1:] if argv is None else argv)
    import torch
    import sobfu_b200 as sf
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    params = params_from_ini(o["params_path"], o["verbosity"])
    depths, masks = [], []
    if o["synthetic"] <= 0:
        if not os.path.isdir(o["file_path"]):
            raise AppError("directory '%s' does not exist" % o["file_path"])
        if not os.path.isdir(os.path.join(o["file_path"], "depth")):
            raise AppError("source directory should contain a 'depth' folder")
        depths = list_files(os.path.join(o["file_path"], "depth"))
        if os.path.isdir(os.path.join(o["file_path"], "omask")):
            masks = list_files(os.path.join(o["file_path"], "omask"))
    n_frames = o["synthetic"] if o["synthetic"] > 0 else len(depths)
    if o["frames"] >= 0:
        n_frames = min(n_frames, o["frames"])
    out_dir = o["out"] or (os.path.join(o["file_path"], "meshes") if o["file_path"] else "meshes")
    if o["logger"] and rank == 0 and not os.path.isdir(out_dir):
        os.makedirs(out_dir)
        print("created output directory for meshes")

    if world == 1:
        fusion = sf.SobFusion(params)
    else:
        from .parallel import SlabFusion
        fusion = SlabFusion(params, dist)

    def whole_mesh(vol):
        if world == 1:
            return fusion.mc.run(vol)[0]
        got = fusion.gather_mesh(vol, dst=0)
        return got[0] if got is not None else None

    total, last_vertices = 0.0, 0
    for i in range(n_frames):
        depth = synthetic_depth(i, params) if o["synthetic"] > 0 else read_depth(depths[i])
        if masks and i < len(masks):                    # demo.cpp:304-308
            mask = read_mask(masks[i])
            if mask.shape != depth.shape:
                raise AppError("mask does not match the depth map")
            depth = np.where(mask != 0, depth, 0).astype(np.uint16)
        if rank == 0:
            print("--- FRAME NO. %d ---" % i)
        d = torch.from_numpy(depth.view(np.int16)).cuda().view(torch.uint16)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fusion(d)
        torch.cuda.synchronize()
        total += time.perf_counter() - t0
        if o["logger"] or o["viz"]:
            mesh = whole_mesh(fusion.phi_global)
            warped = whole_mesh(fusion.phi_global_psi_inv) if i >= 1 else None
            if rank == 0:
                last_vertices = int(mesh.shape[0])
                print("no. of point-normal pairs in the canonical model: %d" % last_vertices)
                if warped is not None:
                    print("no. of point-normal pairs in the canonical model warped to live: %d" % int(warped.shape[0]))
                if o["logger"]:
                    write_vtk(os.path.join(out_dir, "canonical_mesh_%06d.vtk" % i), mesh.cpu().numpy())
                    print("saved canonical_mesh_%06d.vtk" % i)
                    if warped is not None:
                        write_vtk(os.path.join(out_dir, "canonical_warped_to_live_mesh_%06d.vtk" % i), warped.cpu().numpy())
                        print("saved canonical_warped_to_live_mesh_%06d.vtk" % i)
    if o["json"] and rank == 0:
        print(json.dumps({"frames": n_frames, "seconds": total, "frames_per_s": n_frames / total if total > 0 else 0.0, "vertices": last_vertices,
                          "volume": list(params.volume_dims), "n_gpus": world}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except AppError as e:
        sys.stderr.write("error: %s. exiting...\n" % e)
        sys.exit(1)
