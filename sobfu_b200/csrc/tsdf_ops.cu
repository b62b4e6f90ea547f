// tsdf_ops.cu -- per-frame voxel kernels around the solver: TSDF clear / analytic init / projective integration /
// running-average fusion, and the depth-image preparation (bilateral, truncation, ray lengths).
//
// These are the secondary rows of SURVEY.md section 8 (a17-a20).  Unlike the solver core they use the GPU's
// approximate units exactly as the reference does (__fdividef, __expf, sqrtf/powf under --prec-sqrt=false
// --prec-div=false), so this file must be compiled with the reference's numerics flags and plain float
// expressions are written in the same association as the reference's so that nvcc contracts them identically.
#include "solver_kernels.cuh"

namespace sb {
namespace {

constexpr int ZCHUNK = 32;   // z-extent per thread of the column-marching kernels

__global__ void tsdf_clear_kernel(float4 *__restrict__ v4, size_t n4, float2 *__restrict__ tail, size_t ntail) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        v4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = make_float2(0.f, 0.f);
}

SB_DEV float2 pack_tsdf(float sdf, float trunc, float weight) {
    // tsdf_volume.cu:268-274 / :92-98
    if (sdf >= trunc) return make_float2(1.f, weight);
    if (sdf <= -trunc) return make_float2(-1.f, weight);
    return make_float2(__fdividef(sdf, trunc), weight);
}

// init_sphere_kernel, tsdf_volume.cu:249-275.  The reference marches z per (x,y) column with `vc += zstep`, i.e.
// vc.z is a running float sum; each thread here replays that sum up to its chunk so the bits are identical while
// the grid also parallelises over z.
__global__ void init_sphere_kernel(float2 *__restrict__ vol, Dims d, float3 vs, float trunc, float eta, float3 centre,
                                   float radius) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z0 = blockIdx.z * ZCHUNK;
    if (x >= d.X || y >= d.Y) return;
    const float vx = x * vs.x + vs.x / 2.f, vy = y * vs.y + vs.y / 2.f;
    float vz = vs.z / 2.f;
    for (int i = 0; i < z0; ++i) vz += vs.z;
    const int z1 = min(z0 + ZCHUNK, d.Z);
    float2 *p = vol + x + (size_t)d.X * (y + (size_t)d.Y * z0);
    for (int i = z0; i < z1; ++i, vz += vs.z, p += (size_t)d.X * d.Y) {
        const float dist = sqrtf(powf(vx - centre.x, 2) + powf(vy - centre.y, 2) + powf(vz - centre.z, 2));
        const float sdf = dist - radius;
        const float weight = (sdf > -eta) ? 1.f : 0.f;
        *p = pack_tsdf(sdf, trunc, weight);
    }
}

// init_{box,ellipsoid,plane,torus}_kernel, tsdf_volume.cu:181-247, 277-334: signed distance fields of primitives centred in the
// volume (the plane is not centred), weight 1 everywhere.  Same running sum along z as the reference (vc += zstep), replayed
// per z chunk; expressions are written as in the reference so that the compiler contracts the same multiply-adds.
enum { SHAPE_BOX = 0, SHAPE_ELLIPSOID = 1, SHAPE_PLANE = 2, SHAPE_TORUS = 3 };
SB_DEV float norm3(float x, float y, float z) { return sqrtf(__fmaf_rn(x, x, __fmaf_rn(y, y, z * z))); }      // temp_utils.hpp:33-35,86
SB_DEV float norm2(float x, float y) { return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); }      // utils.hpp:212-214
template <int SHAPE>
__global__ void init_shape_kernel(float2 *__restrict__ vol, Dims d, float3 vs, float trunc, float3 prm) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z0 = blockIdx.z * ZCHUNK;
    if (x >= d.X || y >= d.Y) return;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (SHAPE != SHAPE_PLANE) { cx = d.X / 2.f * vs.x; cy = d.Y / 2.f * vs.y; cz = d.Z / 2.f * vs.z; }
    const float vx = (x * vs.x + vs.x / 2.f) - cx, vy = (y * vs.y + vs.y / 2.f) - cy;
    float vz = (vs.z / 2.f) - cz;
    for (int i = 0; i < z0; ++i) vz += vs.z;
    const int z1 = min(z0 + ZCHUNK, d.Z);
    float2 *p = vol + x + (size_t)d.X * (y + (size_t)d.Y * z0);
    for (int i = z0; i < z1; ++i, vz += vs.z, p += (size_t)d.X * d.Y) {
        float sdf;
        if (SHAPE == SHAPE_BOX) {
            const float dx = fabs(vx) - prm.x, dy = fabs(vy) - prm.y, dz = fabs(vz) - prm.z;
            sdf = fmin(fmax(dx, fmax(dy, dz)), 0.f) + norm3(fmax(dx, 0.f), fmax(dy, 0.f), fmax(dz, 0.f));
        } else if (SHAPE == SHAPE_ELLIPSOID) {
            const float k0 = norm3(vx / prm.x, vy / prm.y, vz / prm.z);
            const float k1 = norm3(vx / (prm.x * prm.x), vy / (prm.y * prm.y), vz / (prm.z * prm.z));
            sdf = k0 * (k0 - 1.f) / k1;
        } else if (SHAPE == SHAPE_PLANE) {
            sdf = vz - prm.x;
        } else {
            const float qx = norm2(vx, vz) - prm.x, qy = vy;
            sdf = norm2(qx, qy) - prm.y;
        }
        *p = pack_tsdf(sdf, trunc, 1.f);
    }
}

// TsdfIntegrator::operator()(phi_global, phi_n_psi), tsdf_volume.cu:103-130
__global__ void tsdf_fuse_kernel(float2 *__restrict__ pg, const float2 *__restrict__ pn, size_t n, float max_weight) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 t = pn[i];
        if (t.y == 0.f || (t.y == 1.f && (t.x == 0.f || t.x == -1.f))) continue;
        const float2 prev = pg[i];
        const float tsdf_new = __fdividef(__fmaf_rn(prev.y, prev.x, t.x), prev.y + 1.f);
        const float weight_new = fminf(prev.y + 1.f, max_weight);
        pg[i] = make_float2(tsdf_new, weight_new);
    }
}

struct Aff { float R[9]; float t[3]; };

// TsdfIntegrator::operator()(volume), tsdf_volume.cu:62-101; Projector, device.hpp:36-41; Aff3f * v, device.hpp:61-65
// with dot() of temp_utils.hpp:33-35.  The depth "texture" is point sampled: texel (floor(u), floor(v)).
__global__ void tsdf_integrate_kernel(const float *__restrict__ dists, size_t pitch, int cols, int rows,
                                      float2 *__restrict__ vol, Dims d, float3 vs, float trunc, float eta, Aff aff, float fx,
                                      float fy, float cx, float cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z0 = blockIdx.z * ZCHUNK;
    if (x >= d.X || y >= d.Y) return;
    const float vx = x * vs.x + vs.x / 2.f, vy = y * vs.y + vs.y / 2.f, vz = vs.z / 2.f;
    const float camx = __fmaf_rn(aff.R[0], vx, __fmaf_rn(aff.R[1], vy, aff.R[2] * vz)) + aff.t[0];
    const float camy = __fmaf_rn(aff.R[3], vx, __fmaf_rn(aff.R[4], vy, aff.R[5] * vz)) + aff.t[1];
    float camz = __fmaf_rn(aff.R[6], vx, __fmaf_rn(aff.R[7], vy, aff.R[8] * vz)) + aff.t[2];
    for (int i = 0; i < z0; ++i) camz += vs.z;   // replay `vc_cam += zstep` (x and y only ever gain +0.f)
    const int z1 = min(z0 + ZCHUNK, d.Z);
    float2 *p = vol + x + (size_t)d.X * (y + (size_t)d.Y * z0);
    for (int i = z0; i < z1; ++i, camz += vs.z, p += (size_t)d.X * d.Y) {
        const float u = __fmaf_rn(fx, __fdividef(camx, camz), cx);
        const float v = __fmaf_rn(fy, __fdividef(camy, camz), cy);
        if (u < 0 || v < 0 || u >= cols || v >= rows) continue;
        const float Dp = __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(dists) + (size_t)(int)v * pitch) + (int)u);
        if (Dp <= 0.f || camz <= 0) continue;
        const float psdf = Dp - camz;
        const float weight = (psdf > -eta) ? 1.f : 0.f;
        *p = pack_tsdf(psdf, trunc, weight);
    }
}

// bilateral_kernel, imgproc.cu:8-53 (window [x-k/2, min(x-k/2+k, cols-1)) -- note the exclusive, clamped upper end)
__global__ void bilateral_kernel(const unsigned short *__restrict__ src, size_t sp, unsigned short *__restrict__ dst, size_t dp,
                                 int cols, int rows, int ksz, float ss, float sd) {
    const int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= cols || y >= rows) return;
    auto at = [&](int r, int c) { return (int)*(reinterpret_cast<const unsigned short *>(reinterpret_cast<const char *>(src) + (size_t)r * sp) + c); };
    const int value = at(y, x);
    const int tx = min(x - ksz / 2 + ksz, cols - 1), ty = min(y - ksz / 2 + ksz, rows - 1);
    float sum1 = 0, sum2 = 0;
    for (int cy = max(y - ksz / 2, 0); cy < ty; ++cy)
        for (int cx = max(x - ksz / 2, 0); cx < tx; ++cx) {
            const int depth = at(cy, cx);
            const float space2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
            const float color2 = (value - depth) * (value - depth);
            const float weight = __expf(-(space2 * ss + color2 * sd));
            sum1 += depth * weight;
            sum2 += weight;
        }
    *(reinterpret_cast<unsigned short *>(reinterpret_cast<char *>(dst) + (size_t)y * dp) + x) = __float2int_rn(sum1 / sum2);
}

// truncate_depth_kernel, imgproc.cu:60-77
__global__ void truncate_kernel(unsigned short *__restrict__ depth, size_t pitch, int cols, int rows, unsigned short max_mm) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < cols && y < rows) {
        unsigned short *p = reinterpret_cast<unsigned short *>(reinterpret_cast<char *>(depth) + (size_t)y * pitch) + x;
        if (*p > max_mm) *p = 0;
    }
}

// compute_dists_kernel, imgproc.cu:233-254.  The reference's guard is `x < cols || y < rows` (:237), which lets
// threads outside the image write past the row; inside the image the result is the same, and we do not write outside.
__global__ void dists_kernel(const unsigned short *__restrict__ depth, size_t dp, float *__restrict__ dists, size_t fp, int cols,
                             int rows, float2 finv, float2 c) {
    const int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x < cols && y < rows) {
        const float xl = (x - c.x) * finv.x;
        const float yl = (y - c.y) * finv.y;
        const float lambda = sqrtf(xl * xl + yl * yl + 1);
        const unsigned short dv = *(reinterpret_cast<const unsigned short *>(reinterpret_cast<const char *>(depth) + (size_t)y * dp) + x);
        *(reinterpret_cast<float *>(reinterpret_cast<char *>(dists) + (size_t)y * fp) + x) = dv * lambda * 0.001f;
    }
}
}  // namespace

static int sgrid(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b ? b : 1)); }

void launch_tsdf_clear(float2 *vol, size_t n, cudaStream_t st) {
    // cudaMalloc'd volumes are 256 B aligned: clear with 128-bit stores, odd tail (n odd) with one float2 store
    if ((reinterpret_cast<uintptr_t>(vol) & 15) == 0)
        tsdf_clear_kernel<<<sgrid(n / 2), 256, 0, st>>>(reinterpret_cast<float4 *>(vol), n / 2, vol + (n / 2) * 2, n & 1);
    else
        cudaMemsetAsync(vol, 0, n * sizeof(float2), st);
}
void launch_tsdf_init_sphere(float2 *vol, Dims d, float3 vs, float trunc, float eta, float3 c, float r, cudaStream_t st) {
    dim3 block(32, 8), grid((d.X + 31) / 32, (d.Y + 7) / 8, (d.Z + ZCHUNK - 1) / ZCHUNK);
    init_sphere_kernel<<<grid, block, 0, st>>>(vol, d, vs, trunc, eta, c, r);
}
void launch_tsdf_init_shape(float2 *vol, Dims d, float3 vs, float trunc, int shape, float3 prm, cudaStream_t st) {
    dim3 block(32, 8), grid((d.X + 31) / 32, (d.Y + 7) / 8, (d.Z + ZCHUNK - 1) / ZCHUNK);
    switch (shape) {
        case SHAPE_BOX: init_shape_kernel<SHAPE_BOX><<<grid, block, 0, st>>>(vol, d, vs, trunc, prm); break;
        case SHAPE_ELLIPSOID: init_shape_kernel<SHAPE_ELLIPSOID><<<grid, block, 0, st>>>(vol, d, vs, trunc, prm); break;
        case SHAPE_PLANE: init_shape_kernel<SHAPE_PLANE><<<grid, block, 0, st>>>(vol, d, vs, trunc, prm); break;
        default: init_shape_kernel<SHAPE_TORUS><<<grid, block, 0, st>>>(vol, d, vs, trunc, prm); break;
    }
}
void launch_tsdf_fuse(float2 *pg, const float2 *pn, size_t n, float max_weight, cudaStream_t st) {
    tsdf_fuse_kernel<<<sgrid(n), 256, 0, st>>>(pg, pn, n, max_weight);
}
void launch_tsdf_integrate(const float *dists, size_t pitch, int cols, int rows, float2 *vol, Dims d, float3 vs, float trunc,
                           float eta, const float *R, const float *t, float fx, float fy, float cx, float cy, cudaStream_t st) {
    Aff a;
    for (int i = 0; i < 9; ++i) a.R[i] = R[i];
    for (int i = 0; i < 3; ++i) a.t[i] = t[i];
    dim3 block(32, 8), grid((d.X + 31) / 32, (d.Y + 7) / 8, (d.Z + ZCHUNK - 1) / ZCHUNK);
    tsdf_integrate_kernel<<<grid, block, 0, st>>>(dists, pitch, cols, rows, vol, d, vs, trunc, eta, a, fx, fy, cx, cy);
}
void launch_bilateral(const unsigned short *src, size_t sp, unsigned short *dst, size_t dp, int cols, int rows, int ksz, float ss,
                      float sd, cudaStream_t st) {
    dim3 block(32, 8), grid((cols + 31) / 32, (rows + 7) / 8);
    bilateral_kernel<<<grid, block, 0, st>>>(src, sp, dst, dp, cols, rows, ksz, ss, sd);
}
void launch_truncate(unsigned short *depth, size_t pitch, int cols, int rows, unsigned short max_mm, cudaStream_t st) {
    dim3 block(32, 8), grid((cols + 31) / 32, (rows + 7) / 8);
    truncate_kernel<<<grid, block, 0, st>>>(depth, pitch, cols, rows, max_mm);
}
void launch_dists(const unsigned short *depth, size_t dp, float *dists, size_t fp, int cols, int rows, float fix, float fiy, float cx,
                  float cy, cudaStream_t st) {
    dim3 block(32, 8), grid((cols + 31) / 32, (rows + 7) / 8);
    dists_kernel<<<grid, block, 0, st>>>(depth, dp, dists, fp, cols, rows, make_float2(fix, fiy), make_float2(cx, cy));
}

}  // namespace sb
