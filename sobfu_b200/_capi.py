"""ctypes binding of libsobfu_b200.so (include/sobfu_b200.h).  No compute happens in Python."""
import ctypes as C
import os

from .build import LIB as _DEFAULT_LIB

LIB = os.environ.get("SOBFU_B200_LIB", _DEFAULT_LIB)   # tuning aid: load an alternative build of the same library


class Sobfu200Error(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("dims", C.c_int * 3), ("voxel_size", C.c_float * 3), ("trunc_dist", C.c_float), ("eta", C.c_float),
                ("max_weight", C.c_float), ("verbosity", C.c_int), ("max_iter", C.c_int), ("s", C.c_int),
                ("max_update_norm", C.c_float), ("lambda_", C.c_float), ("alpha", C.c_float), ("w_reg", C.c_float)]


class SolveInfo(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("max_norm", C.c_float), ("max_idx_f", C.c_float),
                ("max_idx", C.c_longlong), ("loop_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_int)]


class IterLog(C.Structure):
    _fields_ = [("max_norm", C.c_float), ("max_idx_f", C.c_float), ("e_data", C.c_float), ("e_reg", C.c_float)]


_P, _I, _F, _Z = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_FP, _IP = C.POINTER(C.c_float), C.POINTER(C.c_int)

# name -> argtypes; every entry of include/sobfu_b200.h that returns int
SIGNATURES = {
    "sobfu_b200_set_stream": [_P],
    "sobfu_b200_solver_create": [C.POINTER(_P), C.POINTER(Params)],
    "sobfu_b200_solver_time_phases": [_P, _I, _FP],
    "sobfu_b200_debug_peer_ranges": [_I, _I, _I, _I, _I, _I, _I, _I, _I, _IP, _IP],
    "sobfu_b200_solver_get_trace": [_P, C.POINTER(C.c_ulonglong), _I, _IP],
    "sobfu_b200_read_depth_png": [C.c_char_p, _P, _I, _IP, _IP],
    "sobfu_b200_read_mask_png": [C.c_char_p, _P, _I, _IP, _IP],
    "sobfu_b200_write_depth_png": [C.c_char_p, _P, _I, _I],
    "sobfu_b200_write_vtk_mesh": [C.c_char_p, _P, C.c_longlong, _I],
    "sobfu_b200_debug_schedule": [_I, _I, _I, _I, _I, _IP, _IP, _IP, _I, _IP, _I, _IP, _IP],
    "sobfu_b200_solver_create_ex": [C.POINTER(_P), C.POINTER(Params), C.c_uint],
    "sobfu_b200_sobolev_taps_computed": [_I, _F, _FP],
    "sobfu_b200_solver_destroy": [_P],
    "sobfu_b200_solver_estimate_psi": [_P, _P, _P, _P, _P, _P, _P, C.POINTER(SolveInfo)],
    "sobfu_b200_solver_estimate_psi_host": [_P, _P, _P, _P, _P, _P, _P, C.POINTER(SolveInfo)],
    "sobfu_b200_solver_get_log": [_P, C.POINTER(IterLog), _I],
    "sobfu_b200_solver_get_taps": [_P, _FP],
    "sobfu_b200_solver_set_variant": [_P, _I],
    "sobfu_b200_solver_time_loop": [_P, _I, _FP, _FP, _FP],
    "sobfu_b200_sobolev_taps": [_I, _F, _FP],
    "sobfu_b200_init_identity": [_P, _I, _I, _I],
    "sobfu_b200_apply": [_P, _P, _P, _I, _I, _I],
    "sobfu_b200_estimate_inverse": [_P, _P, _I, _I, _I, _I],
    "sobfu_b200_clear_field": [_P, _I, _I, _I],
    "sobfu_b200_tsdf_gradient": [_P, _P, _I, _I, _I],
    "sobfu_b200_laplacian": [_P, _P, _I, _I, _I],
    "sobfu_b200_jacobian": [_P, _P, _I, _I, _I, _I],
    "sobfu_b200_potential_gradient": [_P, _P, _P, _P, _P, _F, _I, _I, _I],
    "sobfu_b200_sobolev_filter": [_P, _P, _FP, _I, _I, _I],
    "sobfu_b200_sobolev_filter_s": [_P, _P, _FP, _I, _I, _I, _I],
    "sobfu_b200_update_psi": [_P, _P, _P, _F, _I, _I, _I],
    "sobfu_b200_data_energy": [_P, _P, _I, _FP],
    "sobfu_b200_reg_energy": [_P, _I, _FP],
    "sobfu_b200_max_update_norm": [_P, _I, _FP, _FP, C.POINTER(C.c_longlong)],
    "sobfu_b200_debug_max_update_norm_cand": [_P, _I, _FP, _FP, C.POINTER(C.c_longlong)],
    "sobfu_b200_tsdf_clear": [_P, _I, _I, _I],
    "sobfu_b200_tsdf_init_sphere": [_P, _I, _I, _I, _FP, _F, _F, _FP, _F],
    "sobfu_b200_tsdf_init_box": [_P, _I, _I, _I, _FP, _F, _FP],
    "sobfu_b200_tsdf_init_ellipsoid": [_P, _I, _I, _I, _FP, _F, _FP],
    "sobfu_b200_tsdf_init_plane": [_P, _I, _I, _I, _FP, _F, _F],
    "sobfu_b200_tsdf_init_torus": [_P, _I, _I, _I, _FP, _F, _FP],
    "sobfu_b200_tsdf_fuse": [_P, _P, _I, _I, _I, _F],
    "sobfu_b200_tsdf_integrate": [_P, _Z, _I, _I, _P, _I, _I, _I, _FP, _F, _F, _FP, _FP, _F, _F, _F, _F],
    "sobfu_b200_depth_bilateral": [_P, _Z, _P, _Z, _I, _I, _I, _F, _F],
    "sobfu_b200_depth_truncate": [_P, _Z, _I, _I, _F],
    "sobfu_b200_compute_dists": [_P, _Z, _P, _Z, _I, _I, _F, _F, _F, _F],
    "sobfu_b200_marching_cubes": [_P, _I, _I, _I, _FP, _FP, _FP, _P, _P, _I, _IP, _P, _P, _P, _I, _IP],
    "sobfu_b200_marching_cubes_slab": [_P, _I, _I, _I, _I, _I, _I, _FP, _FP, _FP, _P, _P, _I, _IP, _P, _P, _P, _I, _IP],
    "sobfu_b200_comm_unique_id": [_P],
    "sobfu_b200_solver_attach_comm": [_P, _P, _I, _I],
    "sobfu_b200_solver_peer_export": [_P, _P],
    "sobfu_b200_solver_peer_attach": [_P, _P],
    "sobfu_b200_slab_range": [_I, _I, _I, _IP, _IP],
    "sobfu_b200_tail_window": [_I, _I, _I, _I, _IP, _IP, _IP],
}
OTHER_SYMBOLS = ["sobfu_b200_last_error", "sobfu_b200_version", "sobfu_b200_solver_workspace_bytes", "sobfu_b200_io_last_error",
                 "sobfu_b200_solver_tail_fallbacks"]

_lib = None


def lib():
    """Loads the shared library; raises loudly when it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise Sobfu200Error(
            "libsobfu_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python -m sobfu_b200.build`; sobfu_b200 has no CPU/PyTorch fallback." % LIB)
    L = C.CDLL(LIB)
    for name, args in SIGNATURES.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.sobfu_b200_last_error.restype = C.c_char_p
    L.sobfu_b200_version.restype = C.c_char_p
    L.sobfu_b200_io_last_error.restype = C.c_char_p
    L.sobfu_b200_solver_workspace_bytes.restype = C.c_size_t
    L.sobfu_b200_solver_workspace_bytes.argtypes = [_P]
    L.sobfu_b200_solver_tail_fallbacks.restype = C.c_int
    L.sobfu_b200_solver_tail_fallbacks.argtypes = [_P]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise Sobfu200Error("sobfu_b200 error %d: %s" % (rc, lib().sobfu_b200_last_error().decode()))


def fvec(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])
