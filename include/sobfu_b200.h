/*
 * sobfu_b200.h -- C ABI of the Blackwell-native SobolevFusion solver hot path (libsobfu_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * The header-compatible C++ shim in include/sobfu/ and include/kfusion/ (same class names and signatures as
 * dgrzech/sobfu) forwards to these entry points; sobfu_b200/ (Python, ctypes) binds the same symbols for the
 * parity tests and bench.py.  Each entry point cites the reference interface (file:line in dgrzech/sobfu)
 * it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SOBFU_B200_E* code otherwise; the message of the
 *     last failure on the calling thread is available from sobfu_b200_last_error().  (The reference prints
 *     and exit(0)s on any CUDA error, src/kfusion/device_memory.cpp:7-10; the C++ shim keeps that behaviour.)
 *   - pointers are DEVICE pointers on the current CUDA device unless the name ends in _host
 *   - layouts are the reference's: a volume of dims (X,Y,Z) is row-major with x fastest,
 *     idx = x + X*(y + Y*z) (test/deformation_field_test.cpp:102-104);
 *     TSDF voxels are float2 {tsdf, weight} (include/kfusion/internal.hpp:59-78);
 *     vector fields are float4 {x,y,z,0} and psi stores ABSOLUTE voxel coordinates
 *     (src/sobfu/cuda/vector_fields.cu:72-78); a Jacobian voxel is Mat4f = 4 x float4, rows 0..2 used
 *     (include/kfusion/internal.hpp:12-14)
 *   - work is issued on the stream set by sobfu_b200_set_stream (default: the legacy default stream, as in the
 *     reference) and every call is complete when it returns, like the reference's
 *   - one thread at a time per solver handle
 */
#ifndef SOBFU_B200_H
#define SOBFU_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOBFU_B200_OK 0
#define SOBFU_B200_EINVAL (-1)   /* bad argument (e.g. unsupported (s, lambda), see below) */
#define SOBFU_B200_ECUDA (-2)    /* a CUDA runtime/driver call failed */
#define SOBFU_B200_ENOMEM (-3)
#define SOBFU_B200_ECOMM (-4)    /* NCCL / peer-access failure */

typedef struct sobfu_b200_solver sobfu_b200_solver; /* opaque */

/* mirrors Params / SolverParams (include/sobfu/params.hpp:7-38, include/sobfu/solver.hpp:16-19) */
typedef struct {
    int   dims[3];        /* volume_dims */
    float voxel_size[3];  /* Params::voxel_sizes() */
    float trunc_dist, eta, max_weight;
    int   verbosity;      /* 0 silent, 1 log at iter 1, %50, last (solver.cu:132), 2 every iteration */
    int   max_iter;
    int   s;              /* Sobolev filter length: 7 is what the reference's kernels are compiled for (KERNEL_RADIUS 3, solver.cu:211);
                             3, 9 and 11 -- tabulated by the reference too (solver.cpp:160-251) -- run here as well (single GPU) */
    float max_update_norm;
    float lambda;         /* one of .05 .1 .2 .4 (solver.cpp:160-262); anything else -> SOBFU_B200_EINVAL */
    float alpha, w_reg;
} sobfu_b200_params;

typedef struct {
    int   iters;          /* iterations executed */
    int   converged;      /* 1 if max update norm <= max_update_norm stopped the loop (solver.cu:183) */
    float max_norm;       /* max update norm of the last executed iteration (Reductor::max_update_norm) */
    float max_idx_f;      /* its voxel index as the reference reports it: a float (reductor.cu:367) */
    long long max_idx;    /* the same voxel index, exact */
    float loop_ms;        /* device time of the gradient-descent loop (CUDA events on the solver stream) */
    float total_ms;       /* device time of the whole call incl. psi^-1 and the final warps */
    int   launches;       /* kernels launched by this call */
} sobfu_b200_solve_info;

/* per-iteration record; energies are only filled on logging iterations (else 0) */
typedef struct {
    float max_norm, max_idx_f, e_data, e_reg;
} sobfu_b200_iter_log;

const char *sobfu_b200_last_error(void);
const char *sobfu_b200_version(void);
/* stream used by the free functions and as the caller-visible stream of solver calls (cudaStream_t) */
int sobfu_b200_set_stream(void *cuda_stream);

/* ---- solver: sobfu::cuda::Solver (include/sobfu/solver.hpp:56-67, src/sobfu/solver.cpp:7-101) ---- */
int sobfu_b200_solver_create(sobfu_b200_solver **out, const sobfu_b200_params *p);
int sobfu_b200_solver_destroy(sobfu_b200_solver *s);
/* Solver::estimate_psi -> device::estimate_psi (src/sobfu/cuda/solver.cu:85-205).
 * reads phi_global, phi_n; overwrites phi_n_psi, phi_global_psi_inv, psi (warm start, in place) and psi_inv */
int sobfu_b200_solver_estimate_psi(sobfu_b200_solver *s, const void *phi_global, void *phi_global_psi_inv,
                                   const void *phi_n, void *phi_n_psi, void *psi, void *psi_inv,
                                   sobfu_b200_solve_info *info);
/* same call with HOST buffers: H2D of phi_global, phi_n, psi and D2H of phi_n_psi, phi_global_psi_inv, psi,
 * psi_inv happen inside (pinned staging owned by the handle).  Any output pointer may be NULL to skip its D2H. */
int sobfu_b200_solver_estimate_psi_host(sobfu_b200_solver *s, const void *phi_global_host,
                                        void *phi_global_psi_inv_host, const void *phi_n_host, void *phi_n_psi_host,
                                        void *psi_host, void *psi_inv_host, sobfu_b200_solve_info *info);
/* copies min(n, iters) records of the last solve */
int sobfu_b200_solver_get_log(sobfu_b200_solver *s, sobfu_b200_iter_log *out, int n);
/* the s normalised taps (decompose_sobolev_filter, src/sobfu/solver.cpp:160-262); room for 11 floats */
int sobfu_b200_solver_get_taps(sobfu_b200_solver *s, float *taps);
/* bytes of device scratch owned by the handle */
size_t sobfu_b200_solver_workspace_bytes(sobfu_b200_solver *s);
/* kernel variant: 0 = auto (fastest applicable), 1 = generic per-voxel kernels, 2 = TMA pipelines for both passes,
 * 4 = TMA pipelines with the round-1 pass A (gather4 fetches consumed in the step that issues them; the default consumes them one
 *     step later) */
int sobfu_b200_solver_set_variant(sobfu_b200_solver *s, int variant);
/* benchmarking aid: run `iters` gradient-descent iterations on the state left by the last estimate_psi
 * without convergence checks; returns device ms of pass A, pass B and the whole loop */
int sobfu_b200_solver_time_loop(sobfu_b200_solver *s, int iters, float *ms_pass_a, float *ms_pass_b, float *ms_loop);

int sobfu_b200_sobolev_taps(int s, float lambda, float *taps); /* solver.cpp:160-262: the reference's tables, unit sum */
/* the filter for ANY lambda > 0 and odd s in [3, 11]: (Id - lambda * Laplacian) S = delta on an s^3 grid
 * (the system the reference's unused get_3d_sobolev_filter builds, solver.cpp:107-158), separated into its dominant rank-1
 * factor and normalised to unit sum.  Reproduces the reference's tables to their printed digits where those are consistent. */
int sobfu_b200_sobolev_taps_computed(int s, float lambda, float *taps);
/* solver_create with options: SOBFU_B200_CREATE_COMPUTE_FILTER = use sobolev_taps_computed for a lambda that is not tabulated
 * (s must still be 7: the kernels are 7-tap like the reference's, solver.cu:211) instead of refusing it */
#define SOBFU_B200_CREATE_COMPUTE_FILTER 1u
int sobfu_b200_solver_create_ex(sobfu_b200_solver **out, const sobfu_b200_params *p, unsigned flags);

/* ---- deformation field: sobfu::cuda::DeformationField (include/sobfu/vector_fields.hpp:52-66) ---- */
int sobfu_b200_init_identity(void *psi, int X, int Y, int Z);                       /* vector_fields.cu:56-79 */
int sobfu_b200_apply(const void *phi, void *phi_warped, const void *psi, int X, int Y, int Z); /* :81-109 */
int sobfu_b200_estimate_inverse(const void *psi, void *psi_inv, int X, int Y, int Z, int iters); /* :111-138 (48) */
int sobfu_b200_clear_field(void *field4, int X, int Y, int Z);                      /* vector_fields.cu:28-50 */

/* ---- differentiators used directly by the reference's gtest harness ---- */
int sobfu_b200_tsdf_gradient(const void *phi, void *grad4, int X, int Y, int Z);    /* vector_fields.cu:149-208 */
int sobfu_b200_laplacian(const void *psi, void *L4, int X, int Y, int Z);           /* vector_fields.cu:278-337 */
int sobfu_b200_jacobian(const void *psi, void *J_mat4f, int X, int Y, int Z, int mode); /* :389-472 */
int sobfu_b200_potential_gradient(const void *phi_n_psi, const void *phi_global, const void *grad4, const void *L4,
                                  void *nabla_U4, float w_reg, int X, int Y, int Z);   /* solver.cu:15-47 */
int sobfu_b200_sobolev_filter(void *dst4, const void *src4, const float *taps7_host, int X, int Y, int Z); /* solver.cu:237-459 */
/* the same three sweeps for an odd number s <= 11 of taps (radius (s - 1) / 2; the reference compiles radius 3 only, solver.cu:211) */
int sobfu_b200_sobolev_filter_s(void *dst4, const void *src4, const float *taps_host, int s, int X, int Y, int Z);
int sobfu_b200_update_psi(void *psi, const void *nabla_U_S4, void *updates4, float alpha, int X, int Y, int Z); /* solver.cu:53-79 */

/* ---- Reductor (include/sobfu/reductor.hpp:24-50, src/sobfu/reductor.cpp:38-57) ---- */
int sobfu_b200_data_energy(const void *phi_global, const void *phi_n, int N, float *out);      /* reductor.cu:11-112 */
int sobfu_b200_reg_energy(const void *J_mat4f, int N, float *out);                              /* reductor.cu:114-214 */
int sobfu_b200_max_update_norm(const void *updates4, int N, float *value, float *index_f, long long *index); /* :342-456 */

/* ---- TSDF volume: kfusion::cuda::TsdfVolume (include/kfusion/cuda/tsdf_volume.hpp:17-92) ---- */
int sobfu_b200_tsdf_clear(void *vol, int X, int Y, int Z);                          /* tsdf_volume.cu:23-46 */
int sobfu_b200_tsdf_init_sphere(void *vol, int X, int Y, int Z, const float *voxel_size3, float trunc_dist,
                                float eta, const float *centre3, float radius);     /* tsdf_volume.cu:249-275 */
/* signed distance fields of primitives centred in the volume, weight 1 (TsdfVolume::initBox / initEllipsoid / initPlane /
 * initTorus, tsdf_volume.cpp:108-146 -> tsdf_volume.cu:181-247, 277-334): half extents b, semi-axes r, plane height z (metres
 * from the volume origin, not centred), torus (major radius, tube radius) */
int sobfu_b200_tsdf_init_box(void *vol, int X, int Y, int Z, const float *voxel_size3, float trunc_dist, const float *b3);
int sobfu_b200_tsdf_init_ellipsoid(void *vol, int X, int Y, int Z, const float *voxel_size3, float trunc_dist, const float *r3);
int sobfu_b200_tsdf_init_plane(void *vol, int X, int Y, int Z, const float *voxel_size3, float trunc_dist, float z);
int sobfu_b200_tsdf_init_torus(void *vol, int X, int Y, int Z, const float *voxel_size3, float trunc_dist, const float *t2);
int sobfu_b200_tsdf_fuse(void *phi_global, const void *phi_n_psi, int X, int Y, int Z, float max_weight); /* :103-130 */
/* vol2cam: R (row-major 3x3) and t; dists: float image with row pitch in bytes (tsdf_volume.cu:62-101,141-162) */
int sobfu_b200_tsdf_integrate(const void *dists, size_t pitch_bytes, int cols, int rows, void *vol, int X, int Y,
                              int Z, const float *voxel_size3, float trunc_dist, float eta, const float *R9,
                              const float *t3, float fx, float fy, float cx, float cy);

/* ---- depth pre-processing (include/kfusion/cuda/imgproc.hpp:11-23) ---- */
int sobfu_b200_depth_bilateral(const void *src_u16, size_t src_pitch, void *dst_u16, size_t dst_pitch, int cols,
                               int rows, int ksz, float sigma_spatial, float sigma_depth);  /* imgproc.cu:8-53 */
int sobfu_b200_depth_truncate(void *depth_u16, size_t pitch, int cols, int rows, float max_dist); /* imgproc.cu:60-77 */
int sobfu_b200_compute_dists(const void *depth_u16, size_t depth_pitch, void *dists_f32, size_t dists_pitch, int cols,
                             int rows, float fx, float fy, float cx, float cy);               /* imgproc.cu:233-254 */

/* ---- marching cubes: kfusion::cuda::MarchingCubes::run (include/kfusion/cuda/marching_cubes.hpp:19-56) ----
 * verts/normals: float4 per vertex (pose*v with y,z negated, w=1: marching_cubes.cu:273-276); output is ordered by
 * voxel index (deterministic, unlike the reference's atomics order).  occupied_* (optional, may be NULL) receive the
 * compacted voxel ids / cube indices / vertex counts.  *n_vertices and *n_voxels are HOST outputs. */
int sobfu_b200_marching_cubes(const void *vol, int X, int Y, int Z, const float *volume_size3, const float *R9,
                              const float *t3, void *verts4, void *normals4, int vertex_cap, int *n_vertices,
                              int *occupied_voxel, int *occupied_cube, int *occupied_nverts, int voxel_cap,
                              int *n_voxels);
/* z-slab form (marching cubes runs per slab): `vol_slab` holds planes [z0, z0 + nz_avail) of an X x Y x Z volume,
 * nz_avail = nz + 1 when the first plane of the upper neighbour's slab has been appended (every rank but the last), else nz.
 * Extracts the cells whose lower corner lies in [z0, z0 + nz); voxel ids and coordinates are GLOBAL, so the ranks' outputs
 * concatenated in rank order equal the single-GPU output of sobfu_b200_marching_cubes bit for bit.  volume_size3 is the size
 * of the whole volume.  The reference is single-GPU (marching_cubes.cpp:24-76); this replaces nothing there. */
int sobfu_b200_marching_cubes_slab(const void *vol_slab, int X, int Y, int Z, int z0, int nz, int nz_avail,
                                   const float *volume_size3, const float *R9, const float *t3, void *verts4, void *normals4,
                                   int vertex_cap, int *n_vertices, int *occupied_voxel, int *occupied_cube,
                                   int *occupied_nverts, int voxel_cap, int *n_voxels);

/* ---- multi-GPU (z-slab partition): one process per GPU ----
 * rank 0 obtains an id, the launcher broadcasts it (torch.distributed), every rank attaches. */
#define SOBFU_B200_COMM_ID_BYTES 128
int sobfu_b200_comm_unique_id(void *id128_host);
/* slab of `rank`: planes [z0, z0 + nz); needs Z % nranks == 0 and >= 4 planes per rank */
int sobfu_b200_slab_range(int Z, int rank, int nranks, int *z0, int *nz);
/* switches the solver (created for the GLOBAL dims) to slab mode: afterwards estimate_psi takes the rank's slab of
 * phi_global, phi_global_psi_inv, phi_n_psi, psi, psi_inv and the WHOLE phi_n (replicated: every rank integrates the
 * depth frame itself).  Per iteration: ONE psi halo exchange (4 planes; nabla_U on the halo planes is recomputed) with both
 * neighbours and a scalar MAX all-reduce over NCCL; per frame: one neighbour exchange of a window of psi and phi_global for
 * psi^-1 / phi_global o psi^-1 (all-gather only when a displacement leaves the window, see sobfu_b200_solver_tail_fallbacks). */
int sobfu_b200_solver_attach_comm(sobfu_b200_solver *s, const void *id128_host, int rank, int nranks);
/* slab mode: psi^-1 and phi_global o psi^-1 read a window of psi / phi_global (the rank's planes + up to 16 planes of either
 * neighbour, SOBFU_B200_TAIL_HALO overrides) instead of all-gathered volumes; a gather that leaves the window is detected on the
 * device and the step is repeated on the all-gathered volumes, so results do not depend on the bound.  Returns how many solves of
 * this handle took that fallback. */
int sobfu_b200_solver_tail_fallbacks(sobfu_b200_solver *s);
/* host-only: the planes [win_z0, win_z0 + win_nz) that window covers for `rank` of `nranks` (halo < 0: the default of 16 planes) */
int sobfu_b200_tail_window(int Z, int rank, int nranks, int halo, int *win_z0, int *win_nz, int *halo_used);

/* Peer mode (ranks on one NVLink / NVSwitch domain, e.g. the 8 GPUs of a B200 node): the per-iteration psi halo exchange
 * and the convergence test leave NCCL.  Pass B on the slab faces stores its new psi planes straight into the neighbours'
 * halo planes over NVLink (CUDA IPC mappings) and signals through counters in the neighbours' memory; every rank publishes
 * its per-iteration maximum into every rank's table (last CTA of pass B).  An iteration is then TWO kernels on one stream; the
 * work items next to a slab face are full-length z chunks issued first in both passes, so the halo planes and the
 * acknowledgements travel while the middle of the slab is computed.
 *   every rank: peer_export(block);  launcher: all-gather the blocks (rank-major);  every rank: peer_attach(all blocks).
 * Optional: without it (or with SOBFU_B200_NO_PEER set, or when attach fails) the solver keeps exchanging over NCCL.  The Python
 * launcher attaches it by default: measured on B200 at 256^3, 12709 vs 8416 iterations/s on 8 GPUs, 4878 vs 4778 on 2.
 * Results are bit-identical in both modes.  Replaces nothing in the reference (single GPU, solver.cu:85-205). */
#define SOBFU_B200_PEER_HANDLE_BYTES 128
int sobfu_b200_solver_peer_export(sobfu_b200_solver *s, void *handle_block_host);
int sobfu_b200_solver_peer_attach(sobfu_b200_solver *s, const void *all_handle_blocks_host);   /* NULL: detach */
/* measurement aid (slab mode over NCCL, after an estimate_psi; collective): mean milliseconds per iteration of
 * out5[0] A_mid, [1] wait for the psi halos + A_edge, [2] wait for the global maximum + B_edge, [3] B_mid, [4] the whole iteration */
int sobfu_b200_solver_time_phases(sobfu_b200_solver *s, int iters, float *out5);
/* host-only: the work decomposition of a launch of pass A (pass 0) or pass B (pass 1) over up to three z ranges of a slab of
 * X x Y x Zlocal voxels, planned for `sms` SMs: items[i] = {cta, x0, y0, zb, ze, face} in issue order.  For the schedule tests. */
int sobfu_b200_debug_schedule(int pass, int X, int Y, int Zlocal, int nranges, const int *lo, const int *hi, const int *face, int sms,
                              int *items6, int cap, int *n_items, int *grid);
/* host-only: how peer mode splits the local planes [lo, hi) of a launch into | lower face chunk | upper face chunk | middle |
 * (face chunks first, one z chunk each); ranges9 = {lo, hi, face} x 3 */
/* test aid: Reductor::max_update_norm through the running-candidate form the tiled pass B uses in the loop (ties on the NORM,
 * first in the reference's traversal order wins; reductor.cu:357-368) */
int sobfu_b200_debug_max_update_norm_cand(const void *updates_f4, int N, float *value, float *index_as_float, long long *index);
int sobfu_b200_debug_peer_ranges(int pass, int X, int Y, int Zlocal, int lo, int hi, int has_lo, int has_hi, int sms, int *ranges9, int *n_ranges);
/* measurement aid (peer mode, SOBFU_B200_TRACE=1): device-side timeline of the last estimate_psi, 8 words per launch in launch
 * order (pass A, pass B, ...): first CTA start / last CTA end [ns], sum / max ns waited for the maxima table, sum / max ns
 * waited for a neighbour's counter, CTAs, reserved */
int sobfu_b200_solver_get_trace(sobfu_b200_solver *s, unsigned long long *out8, int cap_launches, int *n_launches);

/* ---- file formats of the application layer (host only; src/apps/demo.cpp:237-246, 301-309) ----
 * 16-bit depth PNG (cv::imread(path, CV_LOAD_IMAGE_ANYDEPTH)), 8-bit mask PNG (cv::imread(path, CV_8U)), legacy-VTK mesh
 * (pcl::io::saveVTKFile).  out == NULL: only the image size is returned.  Errors: sobfu_b200_io_last_error(). */
int sobfu_b200_read_depth_png(const char *path, unsigned short *out, int capacity_pixels, int *cols, int *rows);
int sobfu_b200_read_mask_png(const char *path, unsigned char *out, int capacity_pixels, int *cols, int *rows);
int sobfu_b200_write_depth_png(const char *path, const unsigned short *depth, int cols, int rows);
int sobfu_b200_write_vtk_mesh(const char *path, const float *vertices, long long n_vertices, int stride_floats);
const char *sobfu_b200_io_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SOBFU_B200_H */
