// marching_cubes.cu -- zero level set of a TSDF volume as triangles: kfusion::cuda::MarchingCubes::run
// (src/kfusion/marching_cubes.cpp:24-76 + src/kfusion/cuda/marching_cubes.cu of the reference).
//
// Same per-voxel arithmetic (cube index: any zero-weight corner -> no triangles, bit k = f[k] < 0, marching_cubes.cu:40-79;
// edge interpolation t = (0 - f0) / (f1 - f0 + 1e-15f), flat normals, pose transform, y and z negated, :185-276), different
// schedule: the reference compacts occupied voxels with warp ballots + a global atomic counter (order depends on the
// hardware schedule) and scans with thrust; here every voxel writes (occupied, #vertices) as one 64-bit word, ONE
// cub exclusive scan yields both the voxel slot and the vertex offset, and the triangles come out ordered by voxel index:
// the output is deterministic and identical across runs, devices and slab partitions.
//
// z-slab mode: a rank extracts the cells whose lower corner lies in its planes [z0, z0 + nz) from its slab
// plus ONE plane of the upper neighbour; voxel ids and vertex coordinates use the global z, so the ranks' outputs
// concatenated in rank order are bit-identical to the single-GPU output.
#include <cub/device/device_scan.cuh>

#include <mutex>
#include <string>

#include "mc_tables.h"
#include "solver_kernels.cuh"

namespace sb {
namespace {

__constant__ unsigned char c_num_verts[256];
__constant__ signed char c_tri[256 * 16];

__device__ __forceinline__ int cube_index(const float2 *__restrict__ vol, int x, int y, int z, const Dims d, float (&f)[8]) {
    const size_t sy = d.X, sz = (size_t)d.X * d.Y, o = x + sy * y + sz * z;
    const size_t off[8] = {0, 1, 1 + sy, sy, sz, 1 + sz, 1 + sy + sz, sy + sz};   // marching_cubes.cu:42-63
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float2 v = __ldg(vol + o + off[k]);
        if (v.y == 0.f) return 0;
        f[k] = v.x;
    }
    int c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) c |= (f[k] < 0.f) ? (1 << k) : 0;
    return c;
}

// pass 1: (1 << 32 | numVerts) for occupied voxels, 0 otherwise.  `d` = planes available at `vol` (the slab + its halo plane),
// `nz` = planes whose cells this call owns (local z in [0, nz))
__global__ void classify_kernel(const float2 *__restrict__ vol, Dims d, int nz, unsigned long long *__restrict__ counts) {
    const size_t n = (size_t)d.X * d.Y * nz;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(i / ((size_t)d.X * d.Y)), r = (int)(i - (size_t)z * d.X * d.Y), y = r / d.X, x = r - y * d.X;
        unsigned long long v = 0ull;
        if (x + 1 < d.X && y + 1 < d.Y && z + 1 < d.Z) {
            float f[8];
            const int c = cube_index(vol, x, y, z, d, f);
            const int nv = (c == 0 || c == 255) ? 0 : c_num_verts[c];
            if (nv > 0) v = (1ull << 32) | (unsigned)nv;
        }
        counts[i] = v;
    }
}

struct Pose { float R[9]; float t[3]; };

__device__ __forceinline__ float3 interp(float3 p0, float3 p1, float f0, float f1) {   // vertex_interp, marching_cubes.cu:195-201
    const float t = (0.f - f0) / (f1 - f0 + 1e-15f);
    return make_float3(p0.x + t * (p1.x - p0.x), p0.y + t * (p1.y - p0.y), p0.z + t * (p1.z - p0.z));
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, a.z * b.z)); }

// pass 2: triangles of the occupied voxels, at the offsets of the scan
__global__ void triangles_kernel(const float2 *__restrict__ vol, Dims d, int nz, int z0, float3 cell, Pose pose,
                                 const unsigned long long *__restrict__ counts, const unsigned long long *__restrict__ scan,
                                 float4 *__restrict__ verts, float4 *__restrict__ normals, int vertex_cap, int *__restrict__ occ_voxel,
                                 int *__restrict__ occ_cube, int *__restrict__ occ_nverts, int voxel_cap) {
    const size_t n = (size_t)d.X * d.Y * nz;
    const size_t id0 = (size_t)z0 * d.X * d.Y;        // global voxel id of local voxel 0
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long c = counts[i];
        if (c == 0ull) continue;
        const unsigned long long s = scan[i];
        const int slot = (int)(s >> 32), voff = (int)(s & 0xffffffffull), nv = (int)(c & 0xffffffffull);
        if (slot >= voxel_cap) continue;
        const int z = (int)(i / ((size_t)d.X * d.Y)), r = (int)(i - (size_t)z * d.X * d.Y), y = r / d.X, x = r - y * d.X;
        float f[8];
        const int cube = cube_index(vol, x, y, z, d, f);
        if (occ_voxel) { occ_voxel[slot] = (int)(i + id0); occ_cube[slot] = cube; occ_nverts[slot] = nv; }
        if (!verts) continue;
        // get_node_coo, marching_cubes.cu:185-193: centre of the voxel
        float3 v[8];
        const int dx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, dy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, dz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float3 c3 = make_float3((float)(x + dx[k]), (float)(y + dy[k]), (float)(z0 + z + dz[k]));
            c3.x += 0.5f; c3.y += 0.5f; c3.z += 0.5f;
            c3.x *= cell.x; c3.y *= cell.y; c3.z *= cell.z;
            v[k] = c3;
        }
        const int e0[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, e1[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};   // :220-231
        float3 vl[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) vl[e] = interp(v[e0[e]], v[e1[e]], f[e0[e]], f[e1[e]]);
        for (int k = 0; k < nv; k += 3) {
            const int i1 = c_tri[cube * 16 + k], i2 = c_tri[cube * 16 + k + 1], i3 = c_tri[cube * 16 + k + 2];
            const float3 p1 = vl[i1], p2 = vl[i2], p3 = vl[i3];
            const float3 a = make_float3(p3.x - p1.x, p3.y - p1.y, p3.z - p1.z), b = make_float3(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z);
            float3 nr = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);   // cross(v3 - v1, v2 - v1)
            const float inv = rsqrtf(dot3(nr, nr));
            nr = make_float3(nr.x * inv, nr.y * inv, nr.z * inv);
            const float3 ps[3] = {p1, p2, p3};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int o = voff + k + q;
                if (o >= vertex_cap) continue;
                const float3 p = ps[q];
                const float3 w = make_float3(dot3(make_float3(pose.R[0], pose.R[1], pose.R[2]), p) + pose.t[0],
                                             dot3(make_float3(pose.R[3], pose.R[4], pose.R[5]), p) + pose.t[1],
                                             dot3(make_float3(pose.R[6], pose.R[7], pose.R[8]), p) + pose.t[2]);
                verts[o] = make_float4(w.x, -w.y, -w.z, 1.f);        // store_point, marching_cubes.cu:273-276
                if (normals) normals[o] = make_float4(nr.x, -nr.y, -nr.z, 1.f);
            }
        }
    }
}

// the tables live in __constant__ memory, i.e. per device: uploaded once per device the library is used on
bool upload_tables(std::string &err) {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return true;
    unsigned char nv[256];
    signed char tri[256 * 16];
    for (int c = 0; c < 256; ++c) {
        const char *s = kMcTri[c];
        int n = 0;
        for (; s[n]; ++n) tri[c * 16 + n] = (signed char)(s[n] <= '9' ? s[n] - '0' : s[n] - 'a' + 10);
        nv[c] = (unsigned char)n;
        for (int k = n; k < 16; ++k) tri[c * 16 + k] = -1;
    }
    if (cudaMemcpyToSymbol(c_num_verts, nv, sizeof nv) != cudaSuccess || cudaMemcpyToSymbol(c_tri, tri, sizeof tri) != cudaSuccess) {
        err = std::string("marching cubes tables: ") + cudaGetErrorString(cudaGetLastError());
        return false;
    }
    if (dev >= 0 && dev < 64) done[dev] = true;
    return true;
}

}  // namespace

// scratch (counts + scan + cub temp), kept between calls and grown on demand: a cudaMalloc/cudaFree pair per frame would
// serialise the device twice per extraction.  One set per (host thread, device): no lock is held across the extraction, and a
// thread that moves between devices keeps (and re-uses) the scratch of each of them.
struct McScratch {
    unsigned long long *buf = nullptr;
    size_t buf_elems = 0;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
};
static thread_local McScratch g_mc_dev[64];

int marching_cubes_run(const float2 *vol, Dims dg, int z0, int nz, int nz_avail, float3 size, const float *R, const float *t,
                       float4 *verts, float4 *normals, int vertex_cap, int *n_vertices, int *occ_voxel, int *occ_cube,
                       int *occ_nverts, int voxel_cap, int *n_voxels, cudaStream_t st, std::string &err) {
    if (!upload_tables(err)) return -2;
    const Dims d{dg.X, dg.Y, nz_avail};               // what is addressable at `vol`
    const size_t n = (size_t)d.X * d.Y * nz;
    if (n > 0x7fffffffull) { err = "marching cubes: more than 2^31 voxels in one call"; return -1; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) { err = "marching cubes: device ordinal out of range"; return -1; }
    McScratch &g_mc = g_mc_dev[dev];
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n, st);
    if (g_mc.buf_elems < 2 * n || g_mc.tmp_bytes < tmp_bytes) {
        cudaFree(g_mc.buf); cudaFree(g_mc.tmp);
        g_mc.buf = nullptr; g_mc.tmp = nullptr; g_mc.buf_elems = 0; g_mc.tmp_bytes = 0;
        if (cudaMalloc(&g_mc.buf, 2 * n * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&g_mc.tmp, tmp_bytes) != cudaSuccess) {
            err = std::string("marching cubes scratch: ") + cudaGetErrorString(cudaGetLastError());
            cudaFree(g_mc.buf);
            g_mc.buf = nullptr; g_mc.tmp = nullptr;
            return -3;
        }
        g_mc.buf_elems = 2 * n;
        g_mc.tmp_bytes = tmp_bytes;
    }
    unsigned long long *counts = g_mc.buf, *scan = g_mc.buf + n;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    classify_kernel<<<grid, 256, 0, st>>>(vol, d, nz, counts);
    cub::DeviceScan::ExclusiveSum(g_mc.tmp, tmp_bytes, counts, scan, (int)n, st);
    unsigned long long last[2] = {0, 0};
    cudaMemcpyAsync(&last[0], counts + n - 1, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&last[1], scan + n - 1, 8, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    int rc = 0;
    if (e == cudaSuccess) {
        const unsigned long long total = last[0] + last[1];
        int nvox = (int)(total >> 32), nvert = (int)(total & 0xffffffffull);
        if (voxel_cap <= 0) voxel_cap = vertex_cap > 0 ? vertex_cap / 3 : nvox;   // marching_cubes.cpp:34
        Pose pose;
        for (int i = 0; i < 9; ++i) pose.R[i] = R[i];
        for (int i = 0; i < 3; ++i) pose.t[i] = t[i];
        // cell size as generateTriangles computes it (marching_cubes.cu:292-294): fp32 division on the host, GLOBAL dims
        const float3 cell = make_float3(size.x / dg.X, size.y / dg.Y, size.z / dg.Z);
        if (nvox > 0 && (verts || occ_voxel))
            triangles_kernel<<<grid, 256, 0, st>>>(vol, d, nz, z0, cell, pose, counts, scan, verts, normals, vertex_cap, occ_voxel,
                                                   occ_cube, occ_nverts, voxel_cap);
        e = cudaStreamSynchronize(st);
        if (nvox > voxel_cap) nvox = voxel_cap;
        if (n_voxels) *n_voxels = nvox;
        *n_vertices = nvert;
    }
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
        err = std::string("marching cubes: ") + cudaGetErrorString(e);
        rc = -2;
    }
    return rc;
}

}  // namespace sb
