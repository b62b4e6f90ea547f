/*
 * sobfu_oracle.c -- CPU restatement of the SobolevFusion solver hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sobfu_oracle.h).  Compile with
 *   gcc -O2 -fPIC -shared -fopenmp -mfma -ffp-contract=off -fno-fast-math
 * Citations are file:line in dgrzech/sobfu (mounted at /root/reference in the build container).
 */
#include "sobfu_oracle.h"

#include <math.h>
#include <pmmintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

#define IDX(x, y, z) ((size_t)(x) + (size_t)X * ((size_t)(y) + (size_t)Y * (size_t)(z)))

/* CUDA --ftz=true == x86 FTZ (results) + DAZ (inputs); per thread, so every omp region calls this */
void orc_set_ftz(void) {
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
}
/* every parallel region switches FTZ+DAZ on for its threads and restores the caller's MXCSR afterwards, so that
 * loading the oracle into a Python process does not change NumPy's arithmetic */
static inline unsigned orc_ftz_on(void) {
    const unsigned old = _mm_getcsr();
    _mm_setcsr(old | 0x8040u); /* FTZ (bit 15) | DAZ (bit 6) */
    return old;
}

/* __fsqrt_rd (utils.hpp:279-281): largest float r with r*r <= x */
float orc_sqrt_rd(float x) {
    if (!(x > 0.f)) return (x == 0.f) ? x : sqrtf(x);
    float r = sqrtf(x);
    if ((double)r * (double)r > (double)x) r = nextafterf(r, 0.f);
    return r;
}

/* utils.hpp:33-36 */
static inline float lerpf(float v0, float v1, float t) { return fmaf(t, v0, fmaf(-t, v1, v1)); }

/* ------------------------------------------------------------------------------------------------ */
/* solver.cpp:160-262 */
int orc_sobolev_taps(int s, float lambda, float *h) {
    int ok = 0;
    if (s == 3 && lambda == 0.1f) { h[0] = 0.06537f; h[1] = 0.99572f; h[2] = h[0]; ok = 1; }
    if (s == 7) {
        float a = 0, b = 0, c = 0, d = 0;
        if (lambda == 0.05f) { a = 0.00006f; b = 0.00015f; c = 0.03917f; d = 0.99846f; ok = 1; }
        if (lambda == 0.1f)  { a = 0.00030f; b = 0.00441f; c = 0.06571f; d = 0.99565f; ok = 1; }
        if (lambda == 0.2f)  { a = 0.00120f; b = 0.01094f; c = 0.10204f; d = 0.98941f; ok = 1; }
        if (lambda == 0.4f)  { a = 0.00169f; b = 0.01312f; c = 0.10927f; d = 0.98781f; ok = 1; }
        if (ok) { h[0] = a; h[1] = b; h[2] = c; h[3] = d; h[4] = c; h[5] = b; h[6] = a; }
    }
    if (s == 9) {
        if (lambda == 0.05f) {
            float t[9] = {0.000003f, 0.00006f, 0.00155f, 0.03917f, 0.99846f, 0.03917f, 0.00155f, 0.00006f, 0.000003f};
            memcpy(h, t, sizeof t); ok = 1;
        }
        if (lambda == 0.1f) {
            float t[9] = {0.00002f, 0.00030f, 0.00441f, 0.06571f, 0.99565f, 0.06571f, 0.00441f, 0.00030f, 0.00002f};
            memcpy(h, t, sizeof t); ok = 1;
        }
    }
    if (s == 11 && lambda == 0.1f) {
        float t[11] = {0.0000015f, 0.00002f, 0.00030f, 0.00441f, 0.06571f, 0.99565f,
                       0.06571f,   0.00441f, 0.00030f, 0.00002f, 0.0000015f};
        memcpy(h, t, sizeof t); ok = 1;
    }
    if (!ok) return -1;
    /* solver.cpp:253-261 : normalise to unit sum, fp32 left-to-right */
    float sum = 0.f;
    for (int i = 0; i < s; ++i) sum += h[i];
    for (int i = 0; i < s; ++i) h[i] /= sum;
    return 0;
}

/* vector_fields.cu:64-79 : idx += zstep is an exact float increment for z < 2^24 */
void orc_init_identity(orc_f4 *psi, int X, int Y, int Z) {
#pragma omp parallel for
    for (int z = 0; z < Z; ++z)
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x) {
                orc_f4 v = {(float)x, (float)y, (float)z, 0.f};
                psi[IDX(x, y, z)] = v;
            }
}

/* utils.hpp:50-86 */
static inline orc_f2 interp_tsdf(const orc_f2 *vol, float px, float py, float pz, int X, int Y, int Z) {
    float cx = fminf(fmaxf(0.f, px), (float)X - 1);
    float cy = fminf(fmaxf(0.f, py), (float)Y - 1);
    float cz = fminf(fmaxf(0.f, pz), (float)Z - 1);
    int gx = (int)floorf(cx), gy = (int)floorf(cy), gz = (int)floorf(cz);
    int x1 = gx + 1, y1 = gy + 1, z1 = gz + 1;
    if (cx == 0.f || cx == (float)X - 1) x1--;
    if (cy == 0.f || cy == (float)Y - 1) y1--;
    if (cz == 0.f || cz == (float)Z - 1) z1--;
    float a = cx - gx, b = cy - gy, c = cz - gz;
#define V(i, j, k) vol[IDX(i, j, k)].x
    float t = lerpf(lerpf(lerpf(V(x1, y1, z1), V(x1, y1, gz), c), lerpf(V(x1, gy, z1), V(x1, gy, gz), c), b),
                    lerpf(lerpf(V(gx, y1, z1), V(gx, y1, gz), c), lerpf(V(gx, gy, z1), V(gx, gy, gz), c), b), a);
#undef V
    orc_f2 r = {t, vol[IDX(gx, gy, gz)].y};
    return r;
}

/* vector_fields.cu:81-100 */
void orc_apply(const orc_f2 *phi, orc_f2 *out, const orc_f4 *psi, int X, int Y, int Z) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int z = 0; z < Z; ++z)
            for (int y = 0; y < Y; ++y)
                for (int x = 0; x < X; ++x) {
                    orc_f4 p = psi[IDX(x, y, z)];
                    out[IDX(x, y, z)] = interp_tsdf(phi, p.x, p.y, p.z, X, Y, Z);
                }
        _mm_setcsr(csr__);
    }
}

/* vector_fields.cu:157-208 : both taps collapse onto the in-range neighbour on boundary planes */
void orc_tsdf_gradient(const orc_f2 *phi, orc_f4 *grad, int X, int Y, int Z) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int z = 0; z < Z; ++z)
            for (int y = 0; y < Y; ++y)
                for (int x = 0; x < X; ++x) {
                    int x1 = x + 1, x2 = x - 1, y1 = y + 1, y2 = y - 1, z1 = z + 1, z2 = z - 1;
                    if (x == 0) x2 = x + 1; else if (x == X - 1) x1 = x - 1;
                    if (y == 0) y2 = y + 1; else if (y == Y - 1) y1 = y - 1;
                    if (z == 0) z2 = z + 1; else if (z == Z - 1) z1 = z - 1;
                    orc_f4 n;
                    n.x = (phi[IDX(x1, y, z)].x - phi[IDX(x2, y, z)].x) * 0.5f; /* __fdividef(.,2) */
                    n.y = (phi[IDX(x, y1, z)].x - phi[IDX(x, y2, z)].x) * 0.5f;
                    n.z = (phi[IDX(x, y, z1)].x - phi[IDX(x, y, z2)].x) * 0.5f;
                    n.w = 0.f;
                    grad[IDX(x, y, z)] = n;
                }
        _mm_setcsr(csr__);
    }
}

/* vector_fields.cu:291-337 : on a boundary plane both neighbours are the voxel itself */
void orc_laplacian(const orc_f4 *psi, orc_f4 *L, int X, int Y, int Z) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int z = 0; z < Z; ++z)
            for (int y = 0; y < Y; ++y)
                for (int x = 0; x < X; ++x) {
                    int x1 = x + 1, x2 = x - 1, y1 = y + 1, y2 = y - 1, z1 = z + 1, z2 = z - 1;
                    if (x == 0 || x == X - 1) x1 = x2 = x;
                    if (y == 0 || y == Y - 1) y1 = y2 = y;
                    if (z == 0 || z == Z - 1) z1 = z2 = z;
                    const orc_f4 c = psi[IDX(x, y, z)];
                    const orc_f4 a1 = psi[IDX(x1, y, z)], a2 = psi[IDX(x2, y, z)];
                    const orc_f4 b1 = psi[IDX(x, y1, z)], b2 = psi[IDX(x, y2, z)];
                    const orc_f4 c1 = psi[IDX(x, y, z1)], c2 = psi[IDX(x, y, z2)];
                    orc_f4 r;
                    /* vector_fields.cu:333-335, left to right with un-fused ops */
                    r.x = -1.f * (((((((-6.f * c.x) + a1.x) + a2.x) + b1.x) + b2.x) + c1.x) + c2.x);
                    r.y = -1.f * (((((((-6.f * c.y) + a1.y) + a2.y) + b1.y) + b2.y) + c1.y) + c2.y);
                    r.z = -1.f * (((((((-6.f * c.z) + a1.z) + a2.z) + b1.z) + b2.z) + c1.z) + c2.z);
                    r.w = 0.f;
                    L[IDX(x, y, z)] = r;
                }
        _mm_setcsr(csr__);
    }
}

/* vector_fields.cu:415-472 */
void orc_jacobian(const orc_f4 *psi, float *J, int X, int Y, int Z, int mode) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int z = 0; z < Z; ++z)
            for (int y = 0; y < Y; ++y)
                for (int x = 0; x < X; ++x) {
                    int x1 = x + 1, x2 = x - 1, y1 = y + 1, y2 = y - 1, z1 = z + 1, z2 = z - 1;
                    if (x == 0) x2 = x + 1; else if (x == X - 1) x1 = x - 1;
                    if (y == 0) y2 = y + 1; else if (y == Y - 1) y1 = y - 1;
                    if (z == 0) z2 = z + 1; else if (z == Z - 1) z1 = z - 1;
                    int nb[6][3] = {{x1, y, z}, {x2, y, z}, {x, y1, z}, {x, y2, z}, {x, y, z1}, {x, y, z2}};
                    float v[6][3];
                    for (int k = 0; k < 6; ++k) {
                        orc_f4 p = psi[IDX(nb[k][0], nb[k][1], nb[k][2])];
                        if (mode == 1) { /* get_displacement, vector_fields.cu:24-26 */
                            v[k][0] = p.x + -(float)nb[k][0];
                            v[k][1] = p.y + -(float)nb[k][1];
                            v[k][2] = p.z + -(float)nb[k][2];
                        } else {
                            v[k][0] = p.x; v[k][1] = p.y; v[k][2] = p.z;
                        }
                    }
                    float Jx[3], Jy[3], Jz[3];
                    for (int c = 0; c < 3; ++c) {
                        Jx[c] = (v[0][c] + -v[1][c]) * 0.5f;
                        Jy[c] = (v[2][c] + -v[3][c]) * 0.5f;
                        Jz[c] = (v[4][c] + -v[5][c]) * 0.5f;
                    }
                    float *o = J + 16 * IDX(x, y, z);
                    for (int r = 0; r < 3; ++r) { /* vector_fields.cu:465-468 */
                        o[4 * r + 0] = Jx[r]; o[4 * r + 1] = Jy[r]; o[4 * r + 2] = Jz[r]; o[4 * r + 3] = 0.f;
                    }
                }
        _mm_setcsr(csr__);
    }
}

/* solver.cu:15-33 */
void orc_potential_gradient(const orc_f2 *phi_n_psi, const orc_f2 *phi_global, const orc_f4 *grad,
                            const orc_f4 *L, orc_f4 *nabla_U, float w_reg, int N) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int i = 0; i < N; ++i) {
            float d = phi_n_psi[i].x - phi_global[i].x;
            orc_f4 r;
            r.x = (grad[i].x * d) + (L[i].x * w_reg);
            r.y = (grad[i].y * d) + (L[i].y * w_reg);
            r.z = (grad[i].z * d) + (L[i].z * w_reg);
            r.w = 0.f;
            nabla_U[i] = r;
        }
        _mm_setcsr(csr__);
    }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* solver.cu:237-446.  rows: dst = sum ; columns: dst += sum ; depth: dst += sum.  Taps S[KERNEL_RADIUS - j]
 * for j = -3..3 accumulated from 0 with un-fused mul/add; borders clamp to edge (:256,:263,:270). */
void orc_sobolev_filter(orc_f4 *dst, const orc_f4 *src, const float *S, int X, int Y, int Z) {
    orc_sobolev_filter_r(dst, src, S, 3, X, Y, Z);
}

/* the same three sweeps for a filter of 2 * R + 1 taps (the reference compiles KERNEL_RADIUS = 3 only, solver.cu:211; its tables
 * solver.cpp:160-251 also hold 3-, 9- and 11-tap filters): taps S[R - j], j = -R..R, same accumulation order, clamp to edge */
void orc_sobolev_filter_r(orc_f4 *dst, const orc_f4 *src, const float *S, int R, int X, int Y, int Z) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int z = 0; z < Z; ++z)
            for (int y = 0; y < Y; ++y)
                for (int x = 0; x < X; ++x) {
                    float sx[3] = {0.f, 0.f, 0.f}, sy[3] = {0.f, 0.f, 0.f}, sz[3] = {0.f, 0.f, 0.f};
                    for (int j = -R; j <= R; ++j) {
                        float s = S[R - j];
                        const orc_f4 a = src[IDX(clampi(x + j, 0, X - 1), y, z)];
                        const orc_f4 b = src[IDX(x, clampi(y + j, 0, Y - 1), z)];
                        const orc_f4 c = src[IDX(x, y, clampi(z + j, 0, Z - 1))];
                        sx[0] += a.x * s; sx[1] += a.y * s; sx[2] += a.z * s;
                        sy[0] += b.x * s; sy[1] += b.y * s; sy[2] += b.z * s;
                        sz[0] += c.x * s; sz[1] += c.y * s; sz[2] += c.z * s;
                    }
                    orc_f4 r;
                    r.x = (sx[0] + sy[0]) + sz[0];
                    r.y = (sx[1] + sy[1]) + sz[1];
                    r.z = (sx[2] + sy[2]) + sz[2];
                    r.w = 0.f;
                    dst[IDX(x, y, z)] = r;
                }
        _mm_setcsr(csr__);
    }
}

/* solver.cu:53-69 : psi.w is left untouched */
void orc_update_psi(orc_f4 *psi, const orc_f4 *g, orc_f4 *updates, float alpha, int N) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int i = 0; i < N; ++i) {
            orc_f4 u = {g[i].x * alpha, g[i].y * alpha, g[i].z * alpha, 0.f};
            updates[i] = u;
            psi[i].x -= u.x; psi[i].y -= u.y; psi[i].z -= u.z;
        }
        _mm_setcsr(csr__);
    }
}

/* precomp.cpp:20-43 with maxBlocks 65536 / maxThreads 512 (reductor.cpp:17) */
static void blocks_threads(int n, int *blocks, int *threads) {
    int t;
    if (n < 1024) {
        int x = (n + 1) / 2; --x; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16; t = x + 1;
    } else t = 512;
    int b = (n + (t * 2 - 1)) / (t * 2);
    if (b > 65536) b = 65536;
    *blocks = b; *threads = t;
}

static inline float norm_rd(orc_f4 v) { return orc_sqrt_rd(((v.x * v.x) + (v.y * v.y)) + (v.z * v.z)); }

/* reductor.cu:342-456 + reductor.cpp:81-94, traversal order reproduced so ties resolve identically */
void orc_max_update_norm(const orc_f4 *updates, int N, float *value, float *index) {
    int blocks, threads;
    blocks_threads(N, &blocks, &threads);
    float *bv = (float *)malloc(sizeof(float) * blocks), *bi = (float *)malloc(sizeof(float) * blocks);
    unsigned n = (unsigned)N, bs = (unsigned)threads, gridSize = bs * 2 * (unsigned)blocks;
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
        float *sv = (float *)malloc(sizeof(float) * bs), *si = (float *)malloc(sizeof(float) * bs);
#pragma omp for
        for (int b = 0; b < blocks; ++b) {
            for (unsigned tid = 0; tid < bs; ++tid) {
                float mx = 0.f, my = 0.f;
                unsigned i = (unsigned)b * bs * 2 + tid;
                while (i < n) {
                    float nv = norm_rd(updates[i]);
                    if (nv > mx) { mx = nv; my = (float)i; }
                    if (i + bs < n) {
                        float n2 = norm_rd(updates[i + bs]);
                        if (n2 > mx) { mx = n2; my = (float)i + bs; }
                    }
                    i += gridSize;
                }
                sv[tid] = mx; si[tid] = my;
            }
            for (unsigned s = bs / 2; s >= 1; s >>= 1) /* tid < s keeps its own unless strictly smaller */
                for (unsigned tid = 0; tid < s; ++tid)
                    if (sv[tid + s] > sv[tid]) { sv[tid] = sv[tid + s]; si[tid] = si[tid + s]; }
            bv[b] = sv[0]; bi[b] = si[0];
        }
        free(sv); free(si);
        _mm_setcsr(csr__);
    }
    float rv = 0.f, ri = 0.f;
    for (int b = 0; b < blocks; ++b)
        if (bv[b] > rv) { rv = bv[b]; ri = bi[b]; }
    free(bv); free(bi);
    *value = rv; *index = ri;
}

/* shared tree of reduce_data_kernel / reduce_reg_sobolev_kernel (reductor.cu:41-111): smem tree down to
 * 64, then lane sums with shuffle-down offsets 16..1; lane 0 holds the block sum */
static float block_tree(float *sd, unsigned bs) {
    for (unsigned s = bs / 2; s >= 64; s >>= 1)
        for (unsigned tid = 0; tid < s; ++tid) sd[tid] = sd[tid] + sd[tid + s];
    float lane[32];
    for (unsigned t = 0; t < 32; ++t) {
        lane[t] = (t < bs) ? sd[t] : 0.f;
        if (bs >= 64) lane[t] += sd[t + 32];
    }
    for (int off = 16; off > 0; off /= 2) {
        float nl[32];
        for (int t = 0; t < 32; ++t) nl[t] = lane[t] + ((t + off < 32) ? lane[t + off] : lane[t]);
        memcpy(lane, nl, sizeof lane);
    }
    return lane[0];
}

/* reductor.cu:11-112 ; mySum += d*d is contracted to an fma by nvcc (-fmad default) */
float orc_data_energy(const orc_f2 *pg, const orc_f2 *pn, int N) {
    int blocks, threads;
    blocks_threads(N, &blocks, &threads);
    unsigned n = (unsigned)N, bs = (unsigned)threads, gridSize = bs * 2 * (unsigned)blocks;
    float *out = (float *)malloc(sizeof(float) * blocks);
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
        float *sd = (float *)calloc(bs < 64 ? 64 : bs, sizeof(float));
#pragma omp for
        for (int b = 0; b < blocks; ++b) {
            for (unsigned tid = 0; tid < bs; ++tid) {
                float s = 0.f;
                unsigned i = (unsigned)b * bs * 2 + tid;
                while (i < n) {
                    float d = pg[i].x - pn[i].x;
                    s = fmaf(d, d, s);
                    if (i + bs < n) {
                        float e = pg[i + bs].x - pn[i + bs].x;
                        s = fmaf(e, e, s);
                    }
                    i += gridSize;
                }
                sd[tid] = s;
            }
            out[b] = block_tree(sd, bs);
        }
        free(sd);
        _mm_setcsr(csr__);
    }
    float r = 0.f;
    for (int b = 0; b < blocks; ++b) r += out[b]; /* reductor.cpp:68-79 */
    free(out);
    return 0.5f * r;
}

/* reductor.cu:114-214 */
float orc_reg_energy(const float *J, int N) {
    int blocks, threads;
    blocks_threads(N, &blocks, &threads);
    unsigned n = (unsigned)N, bs = (unsigned)threads, gridSize = bs * 2 * (unsigned)blocks;
    float *out = (float *)malloc(sizeof(float) * blocks);
#define NSQ(p) (((p)[0] * (p)[0] + (p)[1] * (p)[1]) + (p)[2] * (p)[2])
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
        float *sd = (float *)calloc(bs < 64 ? 64 : bs, sizeof(float));
#pragma omp for
        for (int b = 0; b < blocks; ++b) {
            for (unsigned tid = 0; tid < bs; ++tid) {
                float s = 0.f;
                unsigned i = (unsigned)b * bs * 2 + tid;
                while (i < n) {
                    const float *m = J + 16 * (size_t)i;
                    s += (NSQ(m) + NSQ(m + 4)) + NSQ(m + 8);
                    if (i + bs < n) {
                        const float *q = J + 16 * (size_t)(i + bs);
                        s += (NSQ(q) + NSQ(q + 4)) + NSQ(q + 8);
                    }
                    i += gridSize;
                }
                sd[tid] = s;
            }
            out[b] = block_tree(sd, bs);
        }
        free(sd);
        _mm_setcsr(csr__);
    }
#undef NSQ
    float r = 0.f;
    for (int b = 0; b < blocks; ++b) r += out[b];
    free(out);
    return 0.5f * r;
}

/* utils.hpp:124-164 with get_displacement (vector_fields.cu:24-26) */
static inline void interp_disp(const orc_f4 *psi, float px, float py, float pz, int X, int Y, int Z, float *o) {
    float cx = fminf(fmaxf(0.f, px), (float)X - 1);
    float cy = fminf(fmaxf(0.f, py), (float)Y - 1);
    float cz = fminf(fmaxf(0.f, pz), (float)Z - 1);
    int gx = (int)floorf(cx), gy = (int)floorf(cy), gz = (int)floorf(cz);
    int x1 = gx + 1, y1 = gy + 1, z1 = gz + 1;
    if (cx == 0.f || cx == (float)X - 1) x1--;
    if (cy == 0.f || cy == (float)Y - 1) y1--;
    if (cz == 0.f || cz == (float)Z - 1) z1--;
    float a = cx - gx, b = cy - gy, c = cz - gz;
    int xs[2] = {gx, x1}, ys[2] = {gy, y1}, zs[2] = {gz, z1};
    float d[2][2][2][3];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                orc_f4 p = psi[IDX(xs[i], ys[j], zs[k])];
                d[i][j][k][0] = p.x + -(float)xs[i];
                d[i][j][k][1] = p.y + -(float)ys[j];
                d[i][j][k][2] = p.z + -(float)zs[k];
            }
    for (int q = 0; q < 3; ++q)
        o[q] = lerpf(lerpf(lerpf(d[1][1][1][q], d[1][1][0][q], c), lerpf(d[1][0][1][q], d[1][0][0][q], c), b),
                     lerpf(lerpf(d[0][1][1][q], d[0][1][0][q], c), lerpf(d[0][0][1][q], d[0][0][0][q], c), b), a);
}

/* vector_fields.cu:111-138 : every voxel only reads psi and its own psi_inv value, so the 48 launches
 * are 48 independent fixed-point steps per voxel */
void orc_estimate_inverse(const orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z, int iters) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int z = 0; z < Z; ++z)
            for (int y = 0; y < Y; ++y)
                for (int x = 0; x < X; ++x) {
                    orc_f4 v = psi_inv[IDX(x, y, z)];
                    for (int it = 0; it < iters; ++it) {
                        float d[3];
                        interp_disp(psi, v.x, v.y, v.z, X, Y, Z, d);
                        v.x = (float)x + -(d[0] * 1.f);
                        v.y = (float)y + -(d[1] * 1.f);
                        v.z = (float)z + -(d[2] * 1.f);
                        v.w = 0.f;
                    }
                    psi_inv[IDX(x, y, z)] = v;
                }
        _mm_setcsr(csr__);
    }
}

void orc_solver_iteration(const orc_f2 *phi_global, const orc_f2 *phi_n, orc_f2 *phi_n_psi, orc_f4 *psi,
                          orc_f4 *scratch, const float *taps7, float alpha, float w_reg, int X, int Y, int Z,
                          float *max_norm, float *max_idx) {
    orc_solver_iteration_r(phi_global, phi_n, phi_n_psi, psi, scratch, taps7, 3, alpha, w_reg, X, Y, Z, max_norm, max_idx);
}
void orc_solver_iteration_r(const orc_f2 *phi_global, const orc_f2 *phi_n, orc_f2 *phi_n_psi, orc_f4 *psi,
                            orc_f4 *scratch, const float *taps, int R, float alpha, float w_reg, int X, int Y, int Z,
                            float *max_norm, float *max_idx) {
    size_t N = (size_t)X * Y * Z;
    orc_f4 *grad = scratch, *L = scratch + N, *nU = scratch + 2 * N, *nUS = scratch + 3 * N, *upd = scratch + 4 * N;
    orc_tsdf_gradient(phi_n_psi, grad, X, Y, Z);                                   /* solver.cu:120 */
    orc_laplacian(psi, L, X, Y, Z);                                                /* solver.cu:127 */
    orc_potential_gradient(phi_n_psi, phi_global, grad, L, nU, w_reg, (int)N);     /* solver.cu:149 */
    orc_sobolev_filter_r(nUS, nU, taps, R, X, Y, Z);                               /* solver.cu:155-160 */
    orc_update_psi(psi, nUS, upd, alpha, (int)N);                                  /* solver.cu:163 */
    orc_apply(phi_n, phi_n_psi, psi, X, Y, Z);                                     /* solver.cu:168 */
    orc_max_update_norm(upd, (int)N, max_norm, max_idx);                           /* solver.cu:172 */
}

/* solver.cu:85-205 */
int orc_estimate_psi(const orc_f2 *phi_global, orc_f2 *phi_global_psi_inv, const orc_f2 *phi_n,
                     orc_f2 *phi_n_psi, orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z, int max_iter,
                     float max_update_norm, int s, float lambda, float alpha, float w_reg, int log_energies,
                     orc_solve_result *res, orc_iter_log *log) {
    float taps[16];
    if (s != 7 || orc_sobolev_taps(s, lambda, taps) != 0) return -1; /* KERNEL_RADIUS is 3: solver.cu:211 */
    return orc_estimate_psi_taps(phi_global, phi_global_psi_inv, phi_n, phi_n_psi, psi, psi_inv, X, Y, Z, max_iter, max_update_norm,
                                 taps, alpha, w_reg, log_energies, res, log);
}

/* the same loop with the seven filter taps given explicitly (filters outside the reference's tables) */
int orc_estimate_psi_taps(const orc_f2 *phi_global, orc_f2 *phi_global_psi_inv, const orc_f2 *phi_n,
                          orc_f2 *phi_n_psi, orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z, int max_iter,
                          float max_update_norm, const float *taps, float alpha, float w_reg, int log_energies,
                          orc_solve_result *res, orc_iter_log *log) {
    return orc_estimate_psi_taps_r(phi_global, phi_global_psi_inv, phi_n, phi_n_psi, psi, psi_inv, X, Y, Z, max_iter, max_update_norm,
                                   taps, 3, alpha, w_reg, log_energies, res, log);
}
/* ... and with a filter of 2 * R + 1 explicit taps */
int orc_estimate_psi_taps_r(const orc_f2 *phi_global, orc_f2 *phi_global_psi_inv, const orc_f2 *phi_n,
                            orc_f2 *phi_n_psi, orc_f4 *psi, orc_f4 *psi_inv, int X, int Y, int Z, int max_iter,
                            float max_update_norm, const float *taps, int R, float alpha, float w_reg, int log_energies,
                            orc_solve_result *res, orc_iter_log *log) {
    size_t N = (size_t)X * Y * Z;
    orc_f4 *scratch = (orc_f4 *)malloc(sizeof(orc_f4) * 5 * N);
    float *J = log_energies ? (float *)malloc(sizeof(float) * 16 * N) : NULL;
    if (!scratch || (log_energies && !J)) { free(scratch); free(J); return -2; }

    orc_apply(phi_n, phi_n_psi, psi, X, Y, Z); /* solver.cu:106 */
    int iter = 1, converged = 0;
    float mv = 0.f, mi = 0.f;
    while (iter <= max_iter) {
        int do_log = (log_energies == 2) || (log_energies == 1 && (iter == 1 || iter % 50 == 0 || iter == max_iter));
        float e_data = 0.f, e_reg = 0.f;
        if (do_log) { /* solver.cu:124,132-142 */
            orc_jacobian(psi, J, X, Y, Z, 1);
            e_data = orc_data_energy(phi_global, phi_n_psi, (int)N);
            e_reg = orc_reg_energy(J, (int)N);
        }
        orc_solver_iteration_r(phi_global, phi_n, phi_n_psi, psi, scratch, taps, R, alpha, w_reg, X, Y, Z, &mv, &mi);
        if (log) { log[iter - 1].max_norm = mv; log[iter - 1].max_idx = mi; log[iter - 1].e_data = e_data; log[iter - 1].e_reg = e_reg; }
        if (mv <= max_update_norm) { converged = 1; break; } /* solver.cu:183 */
        iter++;
    }
    if (res) { res->iters = converged ? iter : max_iter; res->max_norm = mv; res->max_idx = mi; res->converged = converged; }
    orc_init_identity(psi_inv, X, Y, Z);                         /* solver.cu:196 */
    orc_estimate_inverse(psi, psi_inv, X, Y, Z, 48);             /* solver.cu:197 */
    orc_apply(phi_global, phi_global_psi_inv, psi_inv, X, Y, Z); /* solver.cu:199 */
    free(scratch); free(J);
    return 0;
}

/* ================================================================================================ */
/* secondary per-frame kernels                                                                       */

void orc_tsdf_clear(orc_f2 *vol, int N) { memset(vol, 0, sizeof(orc_f2) * (size_t)N); }

static inline orc_f2 pack_tsdf(float sdf, float trunc, float weight) {
    orc_f2 r;
    r.y = weight;
    if (sdf >= trunc) r.x = 1.f;
    else if (sdf <= -trunc) r.x = -1.f;
    else r.x = sdf / trunc; /* __fdividef: approx on the GPU */
    return r;
}

/* tsdf_volume.cu:249-275 : vc.z is a running float sum (vc += zstep) */
void orc_tsdf_init_sphere(orc_f2 *vol, int X, int Y, int Z, float vx, float vy, float vz, float trunc, float eta,
                          float cx, float cy, float cz, float radius) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x) {
                float px = fmaf((float)x, vx, vx * 0.5f), py = fmaf((float)y, vy, vy * 0.5f), pz = vz * 0.5f;
                for (int z = 0; z < Z; ++z, pz += vz) {
                    float dx = px - cx, dy = py - cy, dz = pz - cz;
                    float d = sqrtf(dx * dx + dy * dy + dz * dz);
                    float sdf = d - radius;
                    vol[IDX(x, y, z)] = pack_tsdf(sdf, trunc, (sdf > -eta) ? 1.f : 0.f);
                }
            }
        _mm_setcsr(csr__);
    }
}

/* tsdf_volume.cu:181-247, 277-334: box / ellipsoid / plane / torus, centred in the volume (the plane is not), weight 1.
 * shape 0..3; prm = half extents | semi-axes | (z) | (major radius, tube radius).  norm(float3) = sqrt(fma(x,x,fma(y,y,z*z)))
 * (temp_utils.hpp:33-35,86), norm(float2) = sqrt_rn(x*x + y*y) without contraction (utils.hpp:212-214). */
static inline float orc_norm3(float x, float y, float z) { return sqrtf(fmaf(x, x, fmaf(y, y, z * z))); }
static inline float orc_norm2(float x, float y) { return sqrtf(x * x + y * y); }
void orc_tsdf_init_shape(orc_f2 *vol, int X, int Y, int Z, float vx, float vy, float vz, float trunc, int shape, float a, float b,
                         float c) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x) {
                float cx = 0.f, cy = 0.f, cz = 0.f;
                if (shape != 2) { cx = (float)X / 2.f * vx; cy = (float)Y / 2.f * vy; cz = (float)Z / 2.f * vz; }
                const float px = fmaf((float)x, vx, vx * 0.5f) - cx, py = fmaf((float)y, vy, vy * 0.5f) - cy;
                float pz = vz * 0.5f - cz;
                for (int z = 0; z < Z; ++z, pz += vz) {
                    float sdf;
                    if (shape == 0) {
                        const float dx = fabsf(px) - a, dy = fabsf(py) - b, dz = fabsf(pz) - c;
                        sdf = fminf(fmaxf(dx, fmaxf(dy, dz)), 0.f) + orc_norm3(fmaxf(dx, 0.f), fmaxf(dy, 0.f), fmaxf(dz, 0.f));
                    } else if (shape == 1) {
                        const float k0 = orc_norm3(px / a, py / b, pz / c);
                        const float k1 = orc_norm3(px / (a * a), py / (b * b), pz / (c * c));
                        sdf = k0 * (k0 - 1.f) / k1;
                    } else if (shape == 2) {
                        sdf = pz - a;
                    } else {
                        const float qx = orc_norm2(px, pz) - a;
                        sdf = orc_norm2(qx, py) - b;
                    }
                    vol[IDX(x, y, z)] = pack_tsdf(sdf, trunc, 1.f);
                }
            }
        _mm_setcsr(csr__);
    }
}

/* tsdf_volume.cu:103-130 */
void orc_tsdf_fuse(orc_f2 *pg, const orc_f2 *pn, int N, float max_weight) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int i = 0; i < N; ++i) {
            orc_f2 t = pn[i];
            if (t.y == 0.f || (t.y == 1.f && (t.x == 0.f || t.x == -1.f))) continue;
            orc_f2 p = pg[i];
            orc_f2 r = {fmaf(p.y, p.x, t.x) / (p.y + 1.f), fminf(p.y + 1.f, max_weight)};
            pg[i] = r;
        }
        _mm_setcsr(csr__);
    }
}

/* tsdf_volume.cu:62-101, device.hpp:36-41,61-65, temp_utils.hpp:33-35 */
void orc_tsdf_integrate(const float *dists, int cols, int rows, orc_f2 *vol, int X, int Y, int Z, float vx,
                        float vy, float vz, float trunc, float eta, const float *R, const float *t, float fx,
                        float fy, float cx, float cy) {
#pragma omp parallel
    {
        const unsigned csr__ = orc_ftz_on();
#pragma omp for
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x) {
                float v[3] = {fmaf((float)x, vx, vx * 0.5f), fmaf((float)y, vy, vy * 0.5f), vz * 0.5f};
                float c[3];
                for (int r = 0; r < 3; ++r)
                    c[r] = fmaf(R[3 * r + 0], v[0], fmaf(R[3 * r + 1], v[1], R[3 * r + 2] * v[2])) + t[r];
                for (int z = 0; z < Z; ++z, c[2] += vz) { /* vc_cam += zstep: x,y gain +0.f */
                    float u = fmaf(fx, c[0] / c[2], cx), w = fmaf(fy, c[1] / c[2], cy);
                    if (u < 0 || w < 0 || u >= cols || w >= rows) continue;
                    float Dp = dists[(int)floorf(w) * cols + (int)floorf(u)]; /* point-sampled texture */
                    if (Dp <= 0.f || c[2] <= 0) continue;
                    float psdf = Dp - c[2];
                    vol[IDX(x, y, z)] = pack_tsdf(psdf, trunc, (psdf > -eta) ? 1.f : 0.f);
                }
            }
        _mm_setcsr(csr__);
    }
}

/* imgproc.cu:8-53 */
void orc_bilateral(const unsigned short *src, unsigned short *dst, int cols, int rows, int ksz, float sigma_spatial,
                   float sigma_depth) {
    sigma_depth *= 1000; /* imgproc.cu:44 */
    float ss = 0.5f / (sigma_spatial * sigma_spatial), sd = 0.5f / (sigma_depth * sigma_depth);
#pragma omp parallel for
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int value = src[y * cols + x];
            int tx = x - ksz / 2 + ksz, ty = y - ksz / 2 + ksz;
            if (tx > cols - 1) tx = cols - 1;
            if (ty > rows - 1) ty = rows - 1;
            float sum1 = 0, sum2 = 0;
            for (int cy = (y - ksz / 2 > 0 ? y - ksz / 2 : 0); cy < ty; ++cy)
                for (int cx = (x - ksz / 2 > 0 ? x - ksz / 2 : 0); cx < tx; ++cx) {
                    int depth = src[cy * cols + cx];
                    float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
                    float color2 = (float)((value - depth) * (value - depth));
                    float weight = expf(-(space2 * ss + color2 * sd));
                    sum1 += depth * weight;
                    sum2 += weight;
                }
            dst[y * cols + x] = (unsigned short)(int)rintf(sum1 / sum2);
        }
}

/* imgproc.cu:60-77 */
void orc_truncate_depth(unsigned short *depth, int cols, int rows, float max_dist) {
    unsigned short m = (unsigned short)(max_dist * 1000.f);
    for (int i = 0; i < cols * rows; ++i)
        if (depth[i] > m) depth[i] = 0;
}

/* imgproc.cu:233-254 */
void orc_compute_dists(const unsigned short *depth, float *dists, int cols, int rows, float fx, float fy, float cx,
                       float cy) {
    float fix = 1.f / fx, fiy = 1.f / fy;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float xl = (x - cx) * fix, yl = (y - cy) * fiy;
            float lambda = sqrtf(xl * xl + yl * yl + 1);
            dists[y * cols + x] = depth[y * cols + x] * lambda * 0.001f;
        }
}

/* ================================================================================================ */
/* marching cubes (marching_cubes.cu:40-276, tables marching_cubes.cpp:100-368)                       */
#include "mc_tables.h"

static int g_nv[256], g_tri[256 * 16], g_tables_ready = 0;
static void mc_tables(void) {
    if (g_tables_ready) return;
    for (int c = 0; c < 256; ++c) {
        const char *s = kMcTri[c];
        int n = 0;
        for (; s[n]; ++n) g_tri[c * 16 + n] = s[n] <= '9' ? s[n] - '0' : s[n] - 'a' + 10;
        g_nv[c] = n;
        for (int k = n; k < 16; ++k) g_tri[c * 16 + k] = -1;
    }
    g_tables_ready = 1;
}
const int *orc_mc_num_verts_table(void) { mc_tables(); return g_nv; }
const int *orc_mc_tri_table(void) { mc_tables(); return g_tri; }

/* computeCubeIndex, marching_cubes.cu:40-79 */
static int cube_index(const orc_f2 *vol, int x, int y, int z, int X, int Y, float *f) {
    const int ox[8] = {0, 1, 1, 0, 0, 1, 1, 0}, oy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, oz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    for (int k = 0; k < 8; ++k) {
        orc_f2 v = vol[IDX(x + ox[k], y + oy[k], z + oz[k])];
        if (v.y == 0.f) return 0;
        f[k] = v.x;
    }
    int c = 0;
    for (int k = 0; k < 8; ++k) c += (f[k] < 0.f) << k;
    return c;
}

/* OccupiedVoxels, marching_cubes.cu:81-144, in voxel-index order (the reference's order depends on its atomics) */
int orc_mc_occupied(const orc_f2 *vol, int X, int Y, int Z, int *voxel_idx, int *cube_idx, int *num_verts, int cap) {
    mc_tables();
    int n = 0;
    for (int z = 0; z < Z - 1; ++z)
        for (int y = 0; y < Y - 1; ++y)
            for (int x = 0; x < X - 1; ++x) {
                float f[8];
                int c = cube_index(vol, x, y, z, X, Y, f);
                int nv = (c == 0 || c == 255) ? 0 : g_nv[c];
                if (nv > 0) {
                    if (n < cap) { voxel_idx[n] = (int)IDX(x, y, z); cube_idx[n] = c; num_verts[n] = nv; }
                    ++n;
                }
            }
    return n;
}

/* TrianglesGenerator, marching_cubes.cu:185-276 (approx: '/', rsqrt on the GPU) */
int orc_mc_triangles(const orc_f2 *vol, int X, int Y, int Z, float sx, float sy, float sz, const float *R, const float *t,
                     const int *voxel_idx, int count, orc_f4 *verts, orc_f4 *normals, int cap) {
    mc_tables();
    const float cs[3] = {sx / X, sy / Y, sz / Z};
    const int ox[8] = {0, 1, 1, 0, 0, 1, 1, 0}, oy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, oz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    const int e0[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, e1[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
    int out = 0;
    for (int i = 0; i < count; ++i) {
        int v = voxel_idx[i], z = v / (X * Y), y = (v - z * X * Y) / X, x = v - z * X * Y - y * X;
        float f[8], p[8][3], vl[12][3];
        int c = cube_index(vol, x, y, z, X, Y, f);
        for (int k = 0; k < 8; ++k) {
            p[k][0] = ((float)(x + ox[k]) + 0.5f) * cs[0]; p[k][1] = ((float)(y + oy[k]) + 0.5f) * cs[1]; p[k][2] = ((float)(z + oz[k]) + 0.5f) * cs[2];
        }
        for (int e = 0; e < 12; ++e) {
            float tt = (0.f - f[e0[e]]) / (f[e1[e]] - f[e0[e]] + 1e-15f);
            for (int q = 0; q < 3; ++q) vl[e][q] = fmaf(tt, p[e1[e]][q] - p[e0[e]][q], p[e0[e]][q]);
        }
        for (int k = 0; k < g_nv[c]; k += 3) {
            const float *p1 = vl[g_tri[c * 16 + k]], *p2 = vl[g_tri[c * 16 + k + 1]], *p3 = vl[g_tri[c * 16 + k + 2]];
            float a[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]}, b[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
            float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            float inv = 1.f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const float *ps[3] = {p1, p2, p3};
            for (int q = 0; q < 3; ++q, ++out) {
                if (out >= cap) continue;
                float w[3];
                for (int r = 0; r < 3; ++r) w[r] = fmaf(R[3 * r], ps[q][0], fmaf(R[3 * r + 1], ps[q][1], R[3 * r + 2] * ps[q][2])) + t[r];
                orc_f4 vv = {w[0], -w[1], -w[2], 1.f}, nn = {n[0] * inv, -n[1] * inv, -n[2] * inv, 1.f};
                verts[out] = vv;
                if (normals) normals[out] = nn;
            }
        }
    }
    return out;
}
