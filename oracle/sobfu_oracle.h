/*
 * sobfu_oracle.h -- CPU restatement of the SobolevFusion solver hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for sobfu_b200.  It restates, in plain C, the arithmetic of the
 * reference's CUDA kernels (dgrzech/sobfu); every function cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libsobfu_b200.so) never links, loads or calls it.
 *
 * Numerics contract (reference build flags CMakeLists.txt:42-44 --ftz=true --prec-div=false
 * --prec-sqrt=false):
 *   - every thread that runs oracle code has MXCSR FTZ+DAZ set (orc_set_ftz) == CUDA .ftz
 *   - lerp is two fmaf (utils.hpp:33-36); float4 operators are un-fused mul/add (utils.hpp:245-275)
 *   - __fdividef(x, 2.f) == x * 0.5f exactly;  __fsqrt_rd emulated exactly (orc_sqrt_rd)
 *   - this file must be compiled with -ffp-contract=off -mfma
 *   - MUFU-approximate ops (__fdividef by a non power of two, sqrtf/powf under --prec-sqrt=false,
 *     __expf) cannot be reproduced bit-for-bit on a CPU: functions using them are marked
 *     "approx" and compared with a tolerance; everything else is bit-exact.
 *
 * Pinning: the reference ships no golden vectors for this path; the oracle is
 * pinned (a) against the reference's own asserting gtest cases, re-stated in tests/test_oracle_pins.py,
 * and (b) against dumps of the reference's own CUDA (oracle/_ref, built verbatim from /root/reference)
 * committed under tests/golden/ by oracle/make_golden.py.
 */
#ifndef SOBFU_ORACLE_H
#define SOBFU_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y; } orc_f2;
typedef struct { float x, y, z, w; } orc_f4;

typedef struct {
    int   iters;         /* iterations executed */
    float max_norm;      /* last max update norm */
    float max_idx;       /* index of the arg-max voxel as float (reductor.cu:367) */
    int   converged;
} orc_solve_result;

/* per-iteration log record (filled for every iteration when log != NULL) */
typedef struct {
    float max_norm, max_idx, e_data, e_reg;
} orc_iter_log;

void  orc_set_ftz(void);
float orc_sqrt_rd(float x);

/* solver.cpp:160-262 -- returns 0 on a known (s, lambda) pair, -1 otherwise */
int   orc_sobolev_taps(int s, float lambda, float *taps);

/* vector_fields.cu:64-79 */
void  orc_init_identity(orc_f4 *psi, int X, int Y, int Z);
/* vector_fields.cu:81-100 + utils.hpp:50-86 */
void  orc_apply(const orc_f2 *phi, orc_f2 *out, const orc_f4 *psi, int X, int Y, int Z);
/* vector_fields.cu:157-208 */
void  orc_tsdf_gradient(const orc_f2 *phi, orc_f4 *grad, int X, int Y, int Z);
/* vector_fields.cu:291-337 */
void  orc_laplacian(const orc_f4 *psi, orc_f4 *L, int X, int Y, int Z);
/* vector_fields.cu:415-472 ; J is 16 floats per voxel (Mat4f), 4th row left untouched */
void  orc_jacobian(const orc_f4 *psi, float *J, int X, int Y, int Z, int mode);
/* solver.cu:15-33 */
void  orc_potential_gradient(const orc_f2 *phi_n_psi, const orc_f2 *phi_global, const orc_f4 *grad,
                             const orc_f4 *L, orc_f4 *nabla_U, float w_reg, int N);
/* solver.cu:237-446 : dst = S*x src ; dst += S*y src ; dst += S*z src (clamp to edge) */
void  orc_sobolev_filter(orc_f4 *dst, const orc_f4 *src, const float *taps7, int X, int Y, int Z);
/* 2 * R + 1 taps (generalisation of the reference's compile-time KERNEL_RADIUS 3, solver.cu:211) */
void  orc_sobolev_filter_r(orc_f4 *dst, const orc_f4 *src, const float *taps, int R, int X, int Y, int Z);
/* solver.cu:53-69 */
void  orc_update_psi(orc_f4 *psi, const orc_f4 *nabla_U_S, orc_f4 *updates, float alpha, int N);
/* reductor.cu:342-456 + reductor.cpp:81-94 : value (sqrt_rd) and index-as-float of the arg max */
void  orc_max_update_norm(const orc_f4 *updates, int N, float *value, float *index);
/* reductor.cu:11-112 + reductor.cpp:38-43,68-79 (exact emulation of the reduction tree) */
float orc_data_energy(const orc_f2 *phi_global, const orc_f2 *phi_n, int N);
/* reductor.cu:114-214 + reductor.cpp:45-50 ; J = 16 floats per voxel */
float orc_reg_energy(const float *J, int N);
/* vector_fields.cu:111-138 + utils.hpp:124-164 ; psi_inv must hold the starting guess */
void  orc_estimate_inverse(const orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z, int iters);

/* solver.cu:85-205 : the whole estimate_psi pipeline. scratch is allocated internally.
 * log (may be NULL) must hold max_iter records. */
int   orc_estimate_psi_taps(const orc_f2 *phi_global, orc_f2 *phi_global_psi_inv, const orc_f2 *phi_n,
                            orc_f2 *phi_n_psi, orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z, int max_iter,
                            float max_update_norm, const float *taps7, float alpha, float w_reg, int log_energies,
                            orc_solve_result *res, orc_iter_log *log);
int   orc_estimate_psi(const orc_f2 *phi_global, orc_f2 *phi_global_psi_inv, const orc_f2 *phi_n,
                       orc_f2 *phi_n_psi, orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z,
                       int max_iter, float max_update_norm, int s, float lambda, float alpha, float w_reg,
                       int log_energies, orc_solve_result *res, orc_iter_log *log);

/* one solver iteration on caller-provided state (used by the cpu_baseline timing leg);
 * scratch = 4*N float4 (grad, L, nabla_U, nabla_U_S) + N float4 updates */
void  orc_solver_iteration_r(const orc_f2 *phi_global, const orc_f2 *phi_n, orc_f2 *phi_n_psi, orc_f4 *psi,
                             orc_f4 *scratch, const float *taps, int R, float alpha, float w_reg, int X, int Y, int Z,
                             float *max_norm, float *max_idx);
int   orc_estimate_psi_taps_r(const orc_f2 *phi_global, orc_f2 *phi_global_psi_inv, const orc_f2 *phi_n,
                              orc_f2 *phi_n_psi, orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z, int max_iter,
                              float max_update_norm, const float *taps, int R, float alpha, float w_reg, int log_energies,
                              orc_solve_result *res, orc_iter_log *log);
void  orc_solver_iteration(const orc_f2 *phi_global, const orc_f2 *phi_n, orc_f2 *phi_n_psi, orc_f4 *psi,
                           orc_f4 *scratch, const float *taps7, float alpha, float w_reg, int X, int Y, int Z,
                           float *max_norm, float *max_idx);

/* ---- secondary (per-frame) kernels ---- */
/* tsdf_volume.cu:23-46 */
void  orc_tsdf_clear(orc_f2 *vol, int N);
/* tsdf_volume.cu:249-275  (approx: powf/sqrtf/__fdividef) */
void  orc_tsdf_init_shape(orc_f2 *vol, int X, int Y, int Z, float vx, float vy, float vz, float trunc, int shape, float a, float b,
                          float c);   /* 0 box, 1 ellipsoid, 2 plane, 3 torus: tsdf_volume.cu:181-247, 277-334 */
void  orc_tsdf_init_sphere(orc_f2 *vol, int X, int Y, int Z, float vx, float vy, float vz, float trunc, float eta,
                           float cx, float cy, float cz, float radius);
/* tsdf_volume.cu:103-130 (approx: __fdividef) */
void  orc_tsdf_fuse(orc_f2 *phi_global, const orc_f2 *phi_n_psi, int N, float max_weight);
/* tsdf_volume.cu:62-101 + device.hpp:36-41 (approx: __fdividef).  R row-major 3x3, t 3 */
void  orc_tsdf_integrate(const float *dists, int cols, int rows, orc_f2 *vol, int X, int Y, int Z, float vx,
                         float vy, float vz, float trunc, float eta, const float *R, const float *t, float fx,
                         float fy, float cx, float cy);
/* imgproc.cu:8-53 (approx: __expf, '/') */
void  orc_bilateral(const unsigned short *src, unsigned short *dst, int cols, int rows, int ksz,
                    float sigma_spatial, float sigma_depth);
/* imgproc.cu:60-77 */
void  orc_truncate_depth(unsigned short *depth, int cols, int rows, float max_dist);
/* imgproc.cu:233-254 (approx: sqrtf) */
void  orc_compute_dists(const unsigned short *depth, float *dists, int cols, int rows, float fx, float fy,
                        float cx, float cy);
/* marching_cubes.cu:40-144 : occupied voxels (sorted by voxel index). returns count; arrays sized cap */
int   orc_mc_occupied(const orc_f2 *vol, int X, int Y, int Z, int *voxel_idx, int *cube_idx, int *num_verts,
                      int cap);
/* marching_cubes.cu:185-276 : triangles for the sorted occupied list. verts/normals: float4 per vertex.
 * returns number of vertices written */
int   orc_mc_triangles(const orc_f2 *vol, int X, int Y, int Z, float sx, float sy, float sz, const float *R,
                       const float *t, const int *voxel_idx, int count, orc_f4 *verts, orc_f4 *normals, int cap);
/* the Bourke tables as used by the oracle (numVerts[256], tri[256*16]) */
const int *orc_mc_num_verts_table(void);
const int *orc_mc_tri_table(void);

#ifdef __cplusplus
}
#endif
#endif
