// solver_generic.cu -- shape-agnostic (any X,Y,Z) kernels of the gradient-descent loop: one thread per voxel.
// They define the loop's data flow and serve as the in-library baseline for the tiled/TMA kernels, which must
// produce identical bits.  Reference semantics cited per kernel (file:line in dgrzech/sobfu).
#include "solver_kernels.cuh"

namespace sb {

namespace {
constexpr int BX = 32, BY = 4, BZ = 2;   // 256 threads, x fastest

SB_DEV bool voxel_of_thread(const Dims d, int &x, int &y, int &z) {
    x = blockIdx.x * BX + threadIdx.x;
    y = blockIdx.y * BY + threadIdx.y;
    z = blockIdx.z * BZ + threadIdx.z;
    return x < d.X && y < d.Y && z < d.Z;
}
inline dim3 grid_for(const Dims d) { return dim3((d.X + BX - 1) / BX, (d.Y + BY - 1) / BY, (d.Z + BZ - 1) / BZ); }

// ---- prologue: AoS -> planes -------------------------------------------------------------------------------
// psi / phi_global: the owned slab; phi_n: the whole volume (replicated on every rank) -> pn plane + gather4 atlas
__global__ void unpack_kernel(const float4 *__restrict__ psi, const float2 *__restrict__ phi_global,
                              const float2 *__restrict__ phi_n, LoopArgs a) {
    const size_t nl = (size_t)a.d.X * a.d.Y * a.d.Z, ng = (size_t)a.dg.X * a.dg.Y * a.dg.Z;
    const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = first; i < nl; i += stride) {
        const float4 p = psi[i];
        a.px[i] = p.x; a.py[i] = p.y; a.pz[i] = p.z;
        const_cast<float *>(a.pg)[i] = phi_global[i].x;
    }
    const int XY = a.dg.X * a.dg.Y;
    for (size_t i = first; i < ng; i += stride) {
        const float v = phi_n[i].x;
        const_cast<float *>(a.pn)[i] = v;
        if (a.pn_surf) {   // same value into the gather4 atlas
            const int z = (int)(i / XY), r = (int)(i - (size_t)z * XY), y = r / a.dg.X, x = r - y * a.dg.X;
            surf2Dwrite(v, a.pn_surf, (int)sizeof(float) * ((z & a.amask) * a.dg.X + x), (z >> a.ashift) * a.dg.Y + y);
        }
    }
}

// phi_n o psi before the first iteration (solver.cu:106 -> apply_kernel, vector_fields.cu:81-100), including the two
// halo planes when they exist in the global volume
__global__ void initial_warp_kernel(LoopArgs a) {
    const long XY = (long)a.d.X * a.d.Y;
    const long hl = min(PSI_HALO, a.z0), hh = min(PSI_HALO, a.dg.Z - (a.z0 + a.d.Z));    // halo planes that exist in the volume
    const long lo = -XY * hl, hi = XY * (a.d.Z + hh);
    for (long i = lo + (long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long)gridDim.x * blockDim.x) {
        const TriCoord t = tri_coord(a.px[i], a.py[i], a.pz[i], a.dg);
        a.w[i] = sample_scalar<1>(a.pn, t, a.dg);
    }
}

// ---- pass A: nabla_U = (phi_n_psi - phi_global) * grad(phi_n_psi) + w_reg * L(psi) ---------------------------
//   grad : TsdfDifferentiator::operator(), vector_fields.cu:157-208 (axis term 0 on that axis' boundary planes)
//   L    : SecondOrderDifferentiator::laplacian, vector_fields.cu:291-337 (both neighbours = self on a boundary)
//   comb : calculate_potential_gradient_kernel, solver.cu:15-33
// When `log` is set it also accumulates the data energy (reductor.cu:11-112) and the regulariser energy of the
// displacement Jacobian (vector_fields.cu:415-472 mode 1 + reductor.cu:114-214) in double precision.
SB_DEV float lap_comp(const float *__restrict__ p, size_t i, size_t ixa, size_t ixb, size_t iya, size_t iyb,
                      size_t iza, size_t izb) {
    float v = mul(p[i], -6.f);
    v = add(v, p[ixa]); v = add(v, p[ixb]);
    v = add(v, p[iya]); v = add(v, p[iyb]);
    v = add(v, p[iza]); v = add(v, p[izb]);
    return mul(v, -1.f);
}

// sum over the three rows of the displacement Jacobian (Differentiator mode 1, vector_fields.cu:415-472) of their squared norms
// at local voxel (x, y, z): get_displacement subtracts the neighbour's own coordinates (vector_fields.cu:24-26); the term of
// Reductor::reg_energy_sobolev (reductor.cu:114-214), in its association (row0 + row1) + row2
SB_DEV float jacobian_rows_nsq(const LoopArgs &a, int x, int y, int z) {
    const Dims d = a.d;
    const size_t sy = (size_t)d.X, sz = (size_t)d.X * d.Y;
    const size_t i = x + sy * y + sz * z;
    const int zg = a.z0 + z;
    const bool z_lo = (zg == 0), z_hi = (zg == a.dg.Z - 1);
    const size_t gxa = (x == d.X - 1) ? i - 1 : i + 1, gxb = (x == 0) ? i + 1 : i - 1;
    const size_t gya = (y == d.Y - 1) ? i - sy : i + sy, gyb = (y == 0) ? i + sy : i - sy;
    const size_t gza = z_hi ? i - sz : i + sz, gzb = z_lo ? i + sz : i - sz;
    const float *P[3] = {a.px, a.py, a.pz};
    const int cxa = (x == d.X - 1) ? x - 1 : x + 1, cxb = (x == 0) ? x + 1 : x - 1;
    const int cya = (y == d.Y - 1) ? y - 1 : y + 1, cyb = (y == 0) ? y + 1 : y - 1;
    const int cza = z_hi ? zg - 1 : zg + 1, czb = z_lo ? zg + 1 : zg - 1;
    float rows = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ox_a = (c == 0) ? (float)cxa : (c == 1 ? (float)y : (float)zg);
        const float ox_b = (c == 0) ? (float)cxb : (c == 1 ? (float)y : (float)zg);
        const float oy_a = (c == 0) ? (float)x : (c == 1 ? (float)cya : (float)zg);
        const float oy_b = (c == 0) ? (float)x : (c == 1 ? (float)cyb : (float)zg);
        const float oz_a = (c == 0) ? (float)x : (c == 1 ? (float)y : (float)cza);
        const float oz_b = (c == 0) ? (float)x : (c == 1 ? (float)y : (float)czb);
        const float jx = mul(sub(sub(P[c][gxa], ox_a), sub(P[c][gxb], ox_b)), 0.5f);
        const float jy = mul(sub(sub(P[c][gya], oy_a), sub(P[c][gyb], oy_b)), 0.5f);
        const float jz = mul(sub(sub(P[c][gza], oz_a), sub(P[c][gzb], oz_b)), 0.5f);
        const float nsq = add(add(mul(jx, jx), mul(jy, jy)), mul(jz, jz));   // norm_sq, utils.hpp:283-285
        rows = (c == 0) ? nsq : add(rows, nsq);
    }
    return rows;
}

// ---- the two logged energies with the reference's fp32 summation (single GPU, N >= 1024) ---------------------------------
// Reductor::data_energy / reg_energy_sobolev (reductor.cpp:38-50): reduce6-style kernels (reductor.cu:11-112, 114-214) of
// `blocks` x `bs` threads -- thread t of block b sums elements b*2*bs + t (+ bs), stride 2*bs*blocks, the data term contracted
// to an fma by nvcc -- a shared-memory tree down to 64, the last two steps in registers with shuffle-down, and the block
// results added up serially in block order on the host (final_reduce, reductor.cpp:68-79).  Same geometry, same order here, so
// the logged energies match the reference's digit for digit; the serial sum runs in one thread of a second launch.
__global__ void __launch_bounds__(512) energy_tree_kernel(LoopArgs a, unsigned n, unsigned grid_size, float *__restrict__ partial) {
    __shared__ float sdat[512], sreg[512];
    const unsigned bs = blockDim.x, tid = threadIdx.x;
    const int X = a.d.X, XY = a.d.X * a.d.Y;
    float sd = 0.f, sr = 0.f;
    for (unsigned i = blockIdx.x * bs * 2u + tid; i < n; i += grid_size) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned e = i + (h ? bs : 0u);
            if (e < n) {
                const float df = sub(a.pg[e], a.w[e]);               // phi_global.x - (phi_n o psi).x, reductor.cu:26-35
                sd = __fmaf_rn(df, df, sd);
                const int z = (int)(e / (unsigned)XY), r = (int)(e - (unsigned)z * (unsigned)XY), y = r / X, x = r - y * X;
                sr = add(sr, jacobian_rows_nsq(a, x, y, z));
            }
        }
    }
    sdat[tid] = sd; sreg[tid] = sr;
    __syncthreads();
    for (unsigned s = bs / 2; s >= 64; s >>= 1) {
        if (tid < s) { sdat[tid] = sd = add(sd, sdat[tid + s]); sreg[tid] = sr = add(sr, sreg[tid + s]); }
        __syncthreads();
    }
    if (tid < 32) {
        sd = add(sd, sdat[tid + 32]); sr = add(sr, sreg[tid + 32]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sd = add(sd, __shfl_down_sync(0xffffffffu, sd, off));
            sr = add(sr, __shfl_down_sync(0xffffffffu, sr, off));
        }
        if (tid == 0) { partial[blockIdx.x] = sd; partial[gridDim.x + blockIdx.x] = sr; }
    }
}
__global__ void energy_final_kernel(const float *__restrict__ partial, unsigned blocks, double *e_data, double *e_reg) {
    if (threadIdx.x >= 2) return;
    const float *p = partial + threadIdx.x * blocks;
    float r = 0.f;
    for (unsigned b = 0; b < blocks; ++b) r = add(r, p[b]);
    *(threadIdx.x ? e_reg : e_data) = (double)r;
}

template <bool LOG>
__global__ void __launch_bounds__(BX *BY *BZ) pass_a_generic_kernel(LoopArgs a, int it) {
    if (loop_finished(a, it)) {
        if (a.check && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0 &&
            threadIdx.z == 0 && !a.state->converged) {
            a.state->iters = it;
            a.state->converged = 1;
        }
        return;
    }
    int x, y, z;
    const bool in = voxel_of_thread(a.d, x, y, z);
    double ed = 0.0, er = 0.0;
    if (in) {
        const Dims d = a.d;
        const size_t sy = (size_t)d.X, sz = (size_t)d.X * d.Y;
        const size_t i = x + sy * y + sz * z;
        const int zg = a.z0 + z;                  // boundary rules apply on the global faces only
        const bool z_lo = (zg == 0), z_hi = (zg == a.dg.Z - 1);
        const bool bx = (x == 0 || x == d.X - 1), by = (y == 0 || y == d.Y - 1), bz = z_lo || z_hi;
        // neighbour indices for the central differences (mirror onto the in-range neighbour at a boundary)
        const size_t gxa = (x == d.X - 1) ? i - 1 : i + 1, gxb = (x == 0) ? i + 1 : i - 1;
        const size_t gya = (y == d.Y - 1) ? i - sy : i + sy, gyb = (y == 0) ? i + sy : i - sy;
        const size_t gza = z_hi ? i - sz : i + sz, gzb = z_lo ? i + sz : i - sz;
        // neighbour indices for the Laplacian (self on a boundary plane)
        const size_t lxa = bx ? i : i + 1, lxb = bx ? i : i - 1;
        const size_t lya = by ? i : i + sy, lyb = by ? i : i - sy;
        const size_t lza = bz ? i : i + sz, lzb = bz ? i : i - sz;

        const float wv = a.w[i];
        const float diff = sub(wv, a.pg[i]);
        // d.X == 1 would make gxa/gxb point outside; volumes are at least 2 voxels wide on every axis
        const float nx = mul(sub(a.w[gxa], a.w[gxb]), 0.5f);   // __fdividef(., 2.f)
        const float ny = mul(sub(a.w[gya], a.w[gyb]), 0.5f);
        const float nz = mul(sub(a.w[gza], a.w[gzb]), 0.5f);
        const float Lx = lap_comp(a.px, i, lxa, lxb, lya, lyb, lza, lzb);
        const float Ly = lap_comp(a.py, i, lxa, lxb, lya, lyb, lza, lzb);
        const float Lz = lap_comp(a.pz, i, lxa, lxb, lya, lyb, lza, lzb);
        const float ux = add(mul(nx, diff), mul(Lx, a.w_reg));
        const float uy = add(mul(ny, diff), mul(Ly, a.w_reg));
        const float uz = add(mul(nz, diff), mul(Lz, a.w_reg));

        // store with a replicated halo (clamp-to-edge of the filter, solver.cu:256,263,270)
        const GLayout gl = a.gl;
        const size_t o = gl.at(x, y, z);
        a.gx[o] = ux; a.gy[o] = uy; a.gz[o] = uz;
        // every boundary voxel replicates itself three cells outwards along the axes it bounds
#define SB_HALO(cond, stride)                                                                                   \
    if (cond) {                                                                                                 \
        _Pragma("unroll") for (int k = 1; k <= 3; ++k) {                                                        \
            a.gx[o + (stride) * k] = ux; a.gy[o + (stride) * k] = uy; a.gz[o + (stride) * k] = uz;              \
        }                                                                                                       \
    }
        SB_HALO(x == 0, -1L) SB_HALO(x == d.X - 1, 1L)
        SB_HALO(y == 0, -(long)gl.PX) SB_HALO(y == d.Y - 1, (long)gl.PX)
        SB_HALO(z_lo, -(long)gl.plane) SB_HALO(z_hi, (long)gl.plane)
#undef SB_HALO

        if (LOG) {
            ed = (double)diff * (double)diff;
            const float rows = jacobian_rows_nsq(a, x, y, z);
            er = (double)rows;
        }
    }
    if (LOG) {
        __shared__ double sd[BX * BY * BZ / 32], sr[BX * BY * BZ / 32];
        const int tid = threadIdx.x + BX * (threadIdx.y + BY * threadIdx.z);
        ed = warp_sum_f64(ed); er = warp_sum_f64(er);
        if ((tid & 31) == 0) { sd[tid >> 5] = ed; sr[tid >> 5] = er; }
        __syncthreads();
        if (tid == 0) {
            double td = 0.0, tr = 0.0;
            for (int k = 0; k < BX * BY * BZ / 32; ++k) { td += sd[k]; tr += sr[k]; }
            atomicAdd(&a.e_data[it], td);
            atomicAdd(&a.e_reg[it], tr);
        }
    }
}

// ---- pass B: Sobolev filter + psi update + max-norm partials + re-warp ---------------------------------------
//   filter : convolution_{rows,columns,depth}_kernel, solver.cu:237-446:  (S*x g + S*y g) + S*z g, taps
//            S[KERNEL_RADIUS - j], j = -3..3, each sum started from 0 with un-fused mul/add
//   update : update_psi_kernel, solver.cu:53-69
//   max    : reduce_max_kernel + final_reduce_max, reductor.cu:342-456, reductor.cpp:81-94
//   warp   : apply_kernel, vector_fields.cu:81-100 (solver.cu:168)
SB_DEV float tap7(const float *__restrict__ g, size_t o, long stride, const float (&S)[MAX_TAPS]) {
    float s = 0.f;
#pragma unroll
    for (int j = -3; j <= 3; ++j) s = add(s, mul(S[3 - j], g[o + (long)j * stride]));
    return s;
}

// 2 * R + 1 taps for R != 3 (the reference's tables also hold 3-, 9- and 11-tap filters, solver.cpp:160-251; its kernels are
// compiled for 7, solver.cu:211): same order of operations, taps S[R - j] for j = -R..R, clamp to edge by clamping the
// coordinate (the replicated halo of the nabla_U planes is 3 deep).  Single GPU only.
SB_DEV float tap_r(const float *__restrict__ g, const GLayout gl, int x, int y, int z, int axis, const Dims d, const float *S, int R) {
    float s = 0.f;
    for (int j = -R; j <= R; ++j) {
        const int xx = axis == 0 ? min(max(x + j, 0), d.X - 1) : x, yy = axis == 1 ? min(max(y + j, 0), d.Y - 1) : y;
        const int zz = axis == 2 ? min(max(z + j, 0), d.Z - 1) : z;
        s = add(s, mul(S[R - j], g[gl.at(xx, yy, zz)]));
    }
    return s;
}

__global__ void __launch_bounds__(BX *BY *BZ) pass_b_generic_kernel(LoopArgs a, int it) {
    if (loop_finished(a, it)) {
        if (a.check && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0 &&
            !a.state->converged) {
            a.state->iters = it;
            a.state->converged = 1;
        }
        return;
    }
    int x, y, z;
    const bool in = voxel_of_thread(a.d, x, y, z);
    unsigned long long key = 0ull;
    if (in) {
        const Dims d = a.d;
        const size_t i = x + (size_t)d.X * y + (size_t)d.X * d.Y * z;
        const GLayout gl = a.gl;
        const size_t o = gl.at(x, y, z);
        const long sy = gl.PX, sz = (long)gl.plane;
        float fx, fy, fz;
        if (a.radius == 3) {
            fx = add(add(tap7(a.gx, o, 1, a.S), tap7(a.gx, o, sy, a.S)), tap7(a.gx, o, sz, a.S));
            fy = add(add(tap7(a.gy, o, 1, a.S), tap7(a.gy, o, sy, a.S)), tap7(a.gy, o, sz, a.S));
            fz = add(add(tap7(a.gz, o, 1, a.S), tap7(a.gz, o, sy, a.S)), tap7(a.gz, o, sz, a.S));
        } else {
            const int R = a.radius;
            fx = add(add(tap_r(a.gx, gl, x, y, z, 0, d, a.S, R), tap_r(a.gx, gl, x, y, z, 1, d, a.S, R)), tap_r(a.gx, gl, x, y, z, 2, d, a.S, R));
            fy = add(add(tap_r(a.gy, gl, x, y, z, 0, d, a.S, R), tap_r(a.gy, gl, x, y, z, 1, d, a.S, R)), tap_r(a.gy, gl, x, y, z, 2, d, a.S, R));
            fz = add(add(tap_r(a.gz, gl, x, y, z, 0, d, a.S, R), tap_r(a.gz, gl, x, y, z, 1, d, a.S, R)), tap_r(a.gz, gl, x, y, z, 2, d, a.S, R));
        }
        const float ux = mul(fx, a.alpha), uy = mul(fy, a.alpha), uz = mul(fz, a.alpha);
        const float npx = sub(a.px[i], ux), npy = sub(a.py[i], uy), npz = sub(a.pz[i], uz);
        a.px[i] = npx; a.py[i] = npy; a.pz[i] = npz;
        const float nsq = add(add(mul(ux, ux), mul(uy, uy)), mul(uz, uz));
        const unsigned ig = (unsigned)(i + (size_t)a.z0 * d.X * d.Y);     // global voxel index
        // the reference compares norms (__fsqrt_rd of the sum of squares, utils.hpp:279-281): ties are ties of the norm
        const float nr = __fsqrt_rd(nsq);
        key = nr > 0.f ? (((unsigned long long)__float_as_uint(nr) << 32) | (unsigned long long)(0xffffffffu - rank_of(ig, a.rm))) : 0ull;
        const TriCoord t = tri_coord(npx, npy, npz, a.dg);
        a.w[i] = sample_scalar<1>(a.pn, t, a.dg);
    }
    __shared__ unsigned long long sk[BX * BY * BZ / 32];
    const int tid = threadIdx.x + BX * (threadIdx.y + BY * threadIdx.z);
    key = warp_max_u64(key);
    if ((tid & 31) == 0) sk[tid >> 5] = key;
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = 0ull;
        for (int k = 0; k < BX * BY * BZ / 32; ++k) m = sk[k] > m ? sk[k] : m;
        atomicMax(&a.maxkey[it], m);
    }
}

// ---- epilogue: planes -> AoS ---------------------------------------------------------------------------------
// psi.w is preserved (update_psi_kernel never writes it); phi_n_psi = {warped tsdf, weight of the floor voxel}
// (utils.hpp:83).
// WARP: the warped plane a.w is not current (the tiled loop keeps phi_n o psi on chip): sample it here (apply_kernel,
// vector_fields.cu:81-100) instead of a separate pass that writes a.w and a read of it
template <bool WARP>
__global__ void pack_kernel(float4 *__restrict__ psi, float2 *__restrict__ phi_n_psi, const float2 *__restrict__ phi_n,
                            LoopArgs a) {
    const size_t n = (size_t)a.d.X * a.d.Y * a.d.Z;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = a.px[i], y = a.py[i], z = a.pz[i];
        float4 p = psi[i];
        p.x = x; p.y = y; p.z = z;
        psi[i] = p;
        const TriCoord t = tri_coord(x, y, z, a.dg);
        const float wgt = phi_n[(size_t)t.gx + (size_t)a.dg.X * ((size_t)t.gy + (size_t)a.dg.Y * t.gz)].y;
        phi_n_psi[i] = make_float2(WARP ? sample_scalar<1>(a.pn, t, a.dg) : a.w[i], wgt);
    }
}
}  // namespace

static int stream_grid(size_t n) {
    size_t b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : b);
}

void launch_unpack(const float4 *psi, const float2 *phi_global, const float2 *phi_n, const LoopArgs &a, cudaStream_t st) {
    const size_t n = (size_t)a.dg.X * a.dg.Y * a.dg.Z;
    unpack_kernel<<<stream_grid(n), 256, 0, st>>>(psi, phi_global, phi_n, a);
}
void launch_initial_warp(const LoopArgs &a, cudaStream_t st) {
    const size_t n = (size_t)a.d.X * a.d.Y * (a.d.Z + 2 * PSI_HALO);
    initial_warp_kernel<<<stream_grid(n), 256, 0, st>>>(a);
}
void launch_pass_a_generic(const LoopArgs &a, int it, int log, cudaStream_t st) {
    // log == 2: the energies of this iteration come from launch_energy_trees (the reference's fp32 summation order)
    if (log == 1) pass_a_generic_kernel<true><<<grid_for(a.d), dim3(BX, BY, BZ), 0, st>>>(a, it);
    else pass_a_generic_kernel<false><<<grid_for(a.d), dim3(BX, BY, BZ), 0, st>>>(a, it);
}
// reduction sizing of the reference (precomp.cpp:20-43 with 65536 blocks / 512 threads, reductor.cpp:17) for n >= 1024
unsigned energy_tree_blocks(size_t n) {
    const size_t b = (n + 1023) / 1024;
    return (unsigned)(b > 65536 ? 65536 : b);
}
// e_data[it], e_reg[it] <- the reference's fp32 sums (not yet halved); `partial` holds 2 * energy_tree_blocks(n) floats; the
// warped plane a.w must be current
void launch_energy_trees(const LoopArgs &a, int it, float *partial, cudaStream_t st) {
    const size_t n = (size_t)a.d.X * a.d.Y * a.d.Z;
    const unsigned blocks = energy_tree_blocks(n);
    energy_tree_kernel<<<blocks, 512, 0, st>>>(a, (unsigned)n, 1024u * blocks, partial);
    energy_final_kernel<<<1, 32, 0, st>>>(partial, blocks, a.e_data + it, a.e_reg + it);
}
void launch_pass_b_generic(const LoopArgs &a, int it, cudaStream_t st) {
    pass_b_generic_kernel<<<grid_for(a.d), dim3(BX, BY, BZ), 0, st>>>(a, it);
}
void launch_pack(float4 *psi, float2 *phi_n_psi, const float2 *phi_n, const LoopArgs &a, bool warp, cudaStream_t st) {
    const size_t n = (size_t)a.d.X * a.d.Y * a.d.Z;
    if (warp) pack_kernel<true><<<stream_grid(n), 256, 0, st>>>(psi, phi_n_psi, phi_n, a);
    else pack_kernel<false><<<stream_grid(n), 256, 0, st>>>(psi, phi_n_psi, phi_n, a);
}

}  // namespace sb
