// pass_a_tiled.cu -- pass A of the solver iteration for x-extents that are a multiple of 4:
//     nabla_U = (phi_n o psi - phi_global) * grad(phi_n o psi) + w_reg * L(psi)
// (TsdfDifferentiator vector_fields.cu:157-208, SecondOrderDifferentiator::laplacian :291-337,
//  calculate_potential_gradient_kernel solver.cu:15-33 of the reference -- one kernel instead of three + Jacobian).
//
// Mapping to the hardware: each thread owns 4 consecutive x (one 128-bit load/store per plane and array), a warp owns
// one or two full 128 B..512 B row segments, and the block marches along z keeping the z-1 / z / z+1 values of psi
// (3 planes) and of the warped TSDF in registers, so every input byte is requested from L2/HBM once per block:
//   - z neighbours: register window, - x neighbours: warp shuffles (edge lanes: one scalar load),
//   - y neighbours: re-read of the neighbouring warp's row, served by L1.
// Algorithmic traffic: R psi 12 + R w 4 + R phi_global 4 + W nabla_U 12 = 32 B/voxel (the reference layouts would be 48).
// The arithmetic is the reference's, operation for operation (see common.cuh); outputs are bit-identical to
// pass_a_generic_kernel.
#include "solver_kernels.cuh"

namespace sb {
namespace {

constexpr int NWARP = 8;

SB_DEV float4 ld4(const float *__restrict__ p, size_t i) { return *reinterpret_cast<const float4 *>(p + i); }
SB_DEV void st4(float *__restrict__ p, size_t i, float4 v) { *reinterpret_cast<float4 *>(p + i) = v; }
SB_DEV float comp(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// x-1 / x+1 neighbours of the 4 voxels of a thread: inner ones from its own registers, outer ones from the
// neighbouring lane (same row) or, at the edge of the warp's row segment, from global memory.
template <int LX>
SB_DEV void x_neighbours(const float4 c, const float *__restrict__ p, size_t i, int lx, bool has_left, bool has_right,
                         float (&xm)[4], float (&xp)[4]) {
    float l = __shfl_up_sync(0xffffffffu, c.w, 1);
    float r = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lx == 0) l = has_left ? __ldg(p + i - 1) : c.x;
    if (lx == LX - 1) r = has_right ? __ldg(p + i + 4) : c.w;
    xm[0] = l;   xm[1] = c.x; xm[2] = c.y; xm[3] = c.z;
    xp[0] = c.y; xp[1] = c.z; xp[2] = c.w; xp[3] = r;
}

template <int LX>
__global__ void __launch_bounds__(NWARP * 32) pass_a_tiled_kernel(LoopArgs a, int it, int zchunk) {
    if (loop_finished(a, it)) {
        if (a.check && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && !a.state->converged) {
            a.state->iters = it;
            a.state->converged = 1;
        }
        return;
    }
    constexpr int RW = 32 / LX;              // rows per warp
    const Dims d = a.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int x0r = (blockIdx.x * LX + lx) * 4;
    const int yr = blockIdx.y * (NWARP * RW) + warp * RW + ly;
    const bool active = x0r < d.X && yr < d.Y;
    // inactive threads shadow the last valid position so that loads stay in range and shuffles stay convergent
    const int x0 = min(x0r, d.X - 4), y = min(yr, d.Y - 1);
    const int zb = blockIdx.z * zchunk, ze = min(zb + zchunk, d.Z);
    if (zb >= ze) return;

    const size_t sy = (size_t)d.X, sz = (size_t)d.X * d.Y;
    const size_t row = (size_t)x0 + sy * y;
    const size_t rym = (size_t)x0 + sy * max(y - 1, 0), ryp = (size_t)x0 + sy * min(y + 1, d.Y - 1);
    const bool y_lo = (y == 0), y_hi = (y == d.Y - 1), by = y_lo || y_hi;
    const bool has_left = x0 > 0, has_right = x0 + 4 < d.X;
    const float *__restrict__ P[3] = {a.px, a.py, a.pz};
    float *__restrict__ G[3] = {a.gx, a.gy, a.gz};
    const GLayout gl = a.gl;

    // register window over z: index 0 = z-1, 1 = z, 2 = z+1
    float4 pw[3][3], ww[3];
    {
        const size_t o0 = row + sz * max(zb - 1, 0), o1 = row + sz * zb;
#pragma unroll
        for (int c = 0; c < 3; ++c) { pw[c][0] = ld4(P[c], o0); pw[c][1] = ld4(P[c], o1); }
        ww[0] = ld4(a.w, o0); ww[1] = ld4(a.w, o1);
    }
    for (int z = zb; z < ze; ++z) {
        const size_t oz = sz * z;
        const size_t on = row + sz * min(z + 1, d.Z - 1);
        // issue every load of this step up front
#pragma unroll
        for (int c = 0; c < 3; ++c) pw[c][2] = ld4(P[c], on);
        ww[2] = ld4(a.w, on);
        const float4 g4 = ld4(a.pg, row + oz);
        float4 pym[3], pyp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { pym[c] = ld4(P[c], rym + oz); pyp[c] = ld4(P[c], ryp + oz); }
        const float4 wym = ld4(a.w, rym + oz), wyp = ld4(a.w, ryp + oz);

        const bool z_lo = (z == 0), z_hi = (z == d.Z - 1), bz = z_lo || z_hi;
        float wxm[4], wxp[4];
        x_neighbours<LX>(ww[1], a.w, row + oz, lx, has_left, has_right, wxm, wxp);
        float nx[4], ny[4], nz[4], df[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool x_lo = (x0 + j == 0), x_hi = (x0 + j == d.X - 1);
            // central differences; on a boundary plane both taps are the single in-range neighbour (-> +0)
            const float a1 = x_hi ? wxm[j] : wxp[j], a2 = x_lo ? wxp[j] : wxm[j];
            const float b1 = y_hi ? comp(wym, j) : comp(wyp, j), b2 = y_lo ? comp(wyp, j) : comp(wym, j);
            const float c1 = z_hi ? comp(ww[0], j) : comp(ww[2], j), c2 = z_lo ? comp(ww[2], j) : comp(ww[0], j);
            nx[j] = mul(sub(a1, a2), 0.5f);
            ny[j] = mul(sub(b1, b2), 0.5f);
            nz[j] = mul(sub(c1, c2), 0.5f);
            df[j] = sub(comp(ww[1], j), comp(g4, j));
        }
        const size_t o = gl.at(x0, y, z);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float pxm[4], pxp[4];
            x_neighbours<LX>(pw[c][1], P[c], row + oz, lx, has_left, has_right, pxm, pxp);
            float u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool bx = (x0 + j == 0) || (x0 + j == d.X - 1);
                const float ctr = comp(pw[c][1], j);
                float v = mul(ctr, -6.f);
                v = add(v, bx ? ctr : pxp[j]);
                v = add(v, bx ? ctr : pxm[j]);
                v = add(v, by ? ctr : comp(pyp[c], j));
                v = add(v, by ? ctr : comp(pym[c], j));
                v = add(v, bz ? ctr : comp(pw[c][2], j));
                v = add(v, bz ? ctr : comp(pw[c][0], j));
                const float L = mul(v, -1.f);
                const float n = (c == 0) ? nx[j] : (c == 1 ? ny[j] : nz[j]);
                u[j] = add(mul(n, df[j]), mul(L, a.w_reg));
            }
            if (active) {
                const float4 uv = make_float4(u[0], u[1], u[2], u[3]);
                st4(G[c], o, uv);
                // replicated halo of 3 (clamp-to-edge of the filter, solver.cu:256,263,270)
                if (x0 == 0) st4(G[c], o - 4, make_float4(u[0], u[0], u[0], u[0]));
                if (x0 + 4 == d.X) st4(G[c], o + 4, make_float4(u[3], u[3], u[3], u[3]));
                if (y_lo) { st4(G[c], o - gl.PX, uv); st4(G[c], o - 2 * (size_t)gl.PX, uv); st4(G[c], o - 3 * (size_t)gl.PX, uv); }
                if (y_hi) { st4(G[c], o + gl.PX, uv); st4(G[c], o + 2 * (size_t)gl.PX, uv); st4(G[c], o + 3 * (size_t)gl.PX, uv); }
                if (z_lo) { st4(G[c], o - gl.plane, uv); st4(G[c], o - 2 * gl.plane, uv); st4(G[c], o - 3 * gl.plane, uv); }
                if (z_hi) { st4(G[c], o + gl.plane, uv); st4(G[c], o + 2 * gl.plane, uv); st4(G[c], o + 3 * gl.plane, uv); }
            }
            pw[c][0] = pw[c][1]; pw[c][1] = pw[c][2];
        }
        ww[0] = ww[1]; ww[1] = ww[2];
    }
}

}  // namespace

// 16 B vector accesses need X % 4 == 0; the TMA tensor maps of pass B need the same (row pitch multiple of 16 B)
bool tiled_supported(const Dims d) { return d.X % 4 == 0 && d.X >= 32 && d.Y >= 8 && d.Z >= 8; }

void launch_pass_a_tiled(const LoopArgs &a, int it, int log, cudaStream_t st) {
    if (log) { launch_pass_a_generic(a, it, 1, st); return; }   // logging iterations (rare) also accumulate the energies
    const Dims d = a.d;
    const bool wide = d.X % 128 == 0;
    const int LX = wide ? 32 : 16, RW = 32 / LX;
    const int gx = (d.X + 4 * LX - 1) / (4 * LX), gy = (d.Y + NWARP * RW - 1) / (NWARP * RW);
    // z chunks: enough blocks for ~3 resident CTAs on each of the 148 SMs, chunks of at least 16 planes
    int nz = (148 * 3 + gx * gy - 1) / (gx * gy);
    if (nz < 1) nz = 1;
    int zchunk = (d.Z + nz - 1) / nz;
    if (zchunk < 16) zchunk = d.Z < 16 ? d.Z : 16;
    nz = (d.Z + zchunk - 1) / zchunk;
    dim3 grid(gx, gy, nz);
    if (wide) pass_a_tiled_kernel<32><<<grid, NWARP * 32, 0, st>>>(a, it, zchunk);
    else pass_a_tiled_kernel<16><<<grid, NWARP * 32, 0, st>>>(a, it, zchunk);
}

}  // namespace sb
