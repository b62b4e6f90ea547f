// solver_tiled.cu -- the two kernels of a solver iteration as persistent, TMA-fed, z-marching pipelines (sm_100a).
//
//   pass A   w = (phi_n o psi).x                                   (apply_kernel vector_fields.cu:81-100, solver.cu:106,168)
//            nabla_U = (w - phi_global) * grad(w) + w_reg * L(psi)
//            (TsdfDifferentiator vector_fields.cu:157-208, laplacian :291-337, potential gradient solver.cu:15-33)
//   pass B   nabla_U_S = S *x nabla_U + S *y nabla_U + S *z nabla_U   (solver.cu:237-446, 7 taps, clamp to edge)
//            psi -= alpha * nabla_U_S ; arg-max partials of |alpha * nabla_U_S|   (solver.cu:53-69, reductor.cu:342-456)
// The warp of the live TSDF is computed by its consumer (pass A, tile + a one-voxel cross halo) instead of being
// written by pass B and read back: the warped volume never touches HBM inside the loop.
//
// Mapping to the hardware (both kernels)
//   * persistent CTAs walk (x-tile, y-tile, z-chunk) work items; the planes an item needs form one continuous stream
//   * one elected thread feeds a ring of shared-memory stages with cp.async.bulk.tensor.3d (TMA); consumers wait on
//     "full" mbarriers (complete_tx) and hand stages back through "empty" mbarriers (one arrival per warp), so warps
//     drift independently up to the ring depth -- no __syncthreads in the steady state
//   * every thread owns 4 consecutive x (128-bit shared and global accesses); x/y neighbours come from the staged
//     tile + halo, z neighbours from the neighbouring stages (pass A) or a 7-deep register window (pass B)
//   * pass B needs no border cases: pass A stores nabla_U with a replicated halo (clamp to edge == plain TMA box);
//     pass A reads unpadded planes: out-of-range box elements are zero-filled by TMA and never used, because on a
//     boundary plane the reference's stencils substitute in-range values (see below)
// Algorithmic HBM traffic: A 32 B/voxel (psi 12, phi_global 4, phi_n gathers 4, nabla_U 12), B 36 B/voxel
// (nabla_U 12, psi 12 + 12); the reference's layouts and kernel split would need 48 + 64.  Results are bit-identical to the
// generic kernels and to the oracle (tests/test_parity_gpu.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "solver_kernels.cuh"
#include "tma_utils.cuh"

namespace sb {

// Programmatic dependent launch: the two kernels of an iteration alternate on one stream, and every kernel boundary costs ~5 us of
// launch latency + CTA scheduling (device-side timeline, profiles/r2_peer_trace_n2.json) -- 3 % of an iteration at 256^3 on one
// GPU, 15 % on a 32-plane slab.  Each CTA allows the next launch to start at once (griddepcontrol.launch_dependents): its CTAs are
// scheduled as the SMs drain, run their prologue (barrier init) and then block in griddepcontrol.wait until this grid has
// completed and flushed -- same ordering as a plain stream launch, minus the exposed latency.  Kernels launched without the
// attribute (or behind an event wait) see both instructions as no-ops.
SB_DEVI void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SB_DEVI void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args &&...args) {
    static const bool off = getenv("SOBFU_B200_PDL") == nullptr;   // opt-in: measured neutral at 256^3 on one GPU (profiles/r2_tuning_log.md)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = off ? 0 : 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct TmaMaps {
    CUtensorMap g[3];      // nabla_U components (padded), pass B input
    CUtensorMap in[3];     // psi x/y/z planes, pass A input (box with cross halo)
    CUtensorMap pb_psi[3]; // psi x/y/z planes, pass B input (tile without halo)
};

namespace {

// work items = (x tile, y tile, z chunk) over up to three z ranges of the slab, issued range by range
struct Sched {
    int tiles_x, tiles_y, nitems;
    int nr;
    int zlo[3], zhi[3], nz[3], zchunk[3], face[3];
};

SB_DEVI float c4(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// (phi_n o psi).x at one voxel: interpolate_tsdf of utils.hpp:50-86 on the phi_n.x plane
SB_DEVI float warp_sample(const float *__restrict__ pn, float px, float py, float pz, const Dims d, int X, int XY) {
    const TriCoord t = tri_coord(px, py, pz, d);
    const int r00 = t.gy * X + t.gz * XY, r10 = t.y1 * X + t.gz * XY;
    const int r01 = t.gy * X + t.z1 * XY, r11 = t.y1 * X + t.z1 * XY;
    return tri_lerp(__ldg(pn + r11 + t.x1), __ldg(pn + r10 + t.x1), __ldg(pn + r01 + t.x1), __ldg(pn + r00 + t.x1),
                    __ldg(pn + r11 + t.gx), __ldg(pn + r10 + t.gx), __ldg(pn + r01 + t.gx), __ldg(pn + r00 + t.gx), t);
}

// Same value through the texture unit: two gather4 fetches (planes gz and z1 of the atlas) return the 2 x 2 x 2 footprint,
// with no address arithmetic and no load instructions in the SM's LSU.  The sample is written in two halves so that a thread
// can put the gathers of several samples in flight and do other work before it touches the first result.
//   issue : tri_coord() of utils.hpp:50-75 without integer detours -- floor stays a float, slice -> atlas tile by exact fp32
//           arithmetic (every value is an integer < 2^24) -- then the two fetches
//   finish: x1 / y1 equal gx / gy on the first and last planes (utils.hpp:61-72); the neighbouring texel returned by the gather
//           is then replaced by the base texel, so the value is bit-identical to warp_sample() for every input
struct TexSample {
    float4 lo, hi;       // texels of slices gz and z1: .w (i,j) .z (i+1,j) .x (i,j+1) .y (i+1,j+1)
    float a, b, c;       // fractional weights
    bool sx, sy;
};
// the fractional weights and the two "upper index not advanced" flags of utils.hpp:61-75 alone (no fetch)
SB_DEVI void tex_weights(TexSample &s, float px, float py, float pz, const Dims d) {
    const float mx = (float)d.X - 1.f, my = (float)d.Y - 1.f, mz = (float)d.Z - 1.f;
    const float cx = fminf(fmaxf(0.f, px), mx), cy = fminf(fmaxf(0.f, py), my), cz = fminf(fmaxf(0.f, pz), mz);
    s.sx = (cx == 0.f || cx == mx);
    s.sy = (cy == 0.f || cy == my);
    s.a = __fsub_rn(cx, floorf(cx)); s.b = __fsub_rn(cy, floorf(cy)); s.c = __fsub_rn(cz, floorf(cz));
}
// the two gather4 fetches alone: their 8 result registers are all a sample in flight needs (the software-pipelined pass A
// recomputes the weights from psi when it consumes them a step later)
SB_DEVI void tex_fetch(float4 &lo, float4 &hi, cudaTextureObject_t tex, int ashift, int amask, float px, float py, float pz, const Dims d) {
    const float mx = (float)d.X - 1.f, my = (float)d.Y - 1.f, mz = (float)d.Z - 1.f;
    const float cx = fminf(fmaxf(0.f, px), mx), cy = fminf(fmaxf(0.f, py), my), cz = fminf(fmaxf(0.f, pz), mz);
    const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
    const float kxf = (float)(amask + 1), ikx = __int_as_float((127 - ashift) << 23);      // kx = 2^ashift and 1 / kx
    const float u = fx + 1.f, v = fy + 1.f;                           // footprint (gx, gx+1) x (gy, gy+1)
    const float z1 = (cz == 0.f || cz == mz) ? fz : fz + 1.f;
    const float r0 = floorf(fz * ikx), r1 = floorf(z1 * ikx);         // atlas row of the slice; column = z - row * kx
    lo = tex2Dgather<float4>(tex, __fmaf_rn(__fmaf_rn(-r0, kxf, fz), (float)d.X, u), __fmaf_rn(r0, (float)d.Y, v), 0);
    hi = tex2Dgather<float4>(tex, __fmaf_rn(__fmaf_rn(-r1, kxf, z1), (float)d.X, u), __fmaf_rn(r1, (float)d.Y, v), 0);
}
SB_DEVI void tex_issue(TexSample &s, cudaTextureObject_t tex, int ashift, int amask, float px, float py, float pz, const Dims d) {
    tex_weights(s, px, py, pz, d);
    tex_fetch(s.lo, s.hi, tex, ashift, amask, px, py, pz, d);
}
SB_DEVI float tex_finish(const TexSample &s) {
    const float v000 = s.lo.w, v100 = s.sx ? s.lo.w : s.lo.z;
    const float v010 = s.sy ? v000 : s.lo.x, v110 = s.sy ? v100 : (s.sx ? s.lo.x : s.lo.y);
    const float v001 = s.hi.w, v101 = s.sx ? s.hi.w : s.hi.z;
    const float v011 = s.sy ? v001 : s.hi.x, v111 = s.sy ? v101 : (s.sx ? s.hi.x : s.hi.y);
    return lerp(lerp(lerp(v111, v110, s.c), lerp(v101, v100, s.c), s.b), lerp(lerp(v011, v010, s.c), lerp(v001, v000, s.c), s.b), s.a);
}
SB_DEVI float warp_sample_tex(cudaTextureObject_t tex, int ashift, int amask, float px, float py, float pz, const Dims d) {
    TexSample s;
    tex_issue(s, tex, ashift, amask, px, py, pz, d);
    return tex_finish(s);
}

// work item -> tile origin, plane range [zb, ze) and face tag.  Host and device: the same function drives the kernels and the
// schedule checks of the CPU tests (sobfu_b200_debug_schedule).
__host__ __device__ __forceinline__ void locate_item(const Sched &sc, int item, int TX, int TY, int &x0t, int &y0t, int &zb, int &ze, int &face) {
    const int xy = sc.tiles_x * sc.tiles_y;
    int tz = item / xy;
    const int rem = item - tz * xy;
    const int tyi = rem / sc.tiles_x;
    x0t = (rem - tyi * sc.tiles_x) * TX;
    y0t = tyi * TY;
    // which range? (no dynamic indexing: the schedule stays in registers)
    const bool r1 = tz >= sc.nz[0], r2 = tz >= sc.nz[0] + sc.nz[1];
    tz -= r2 ? sc.nz[0] + sc.nz[1] : (r1 ? sc.nz[0] : 0);
    const int chunk = r2 ? sc.zchunk[2] : (r1 ? sc.zchunk[1] : sc.zchunk[0]);
    zb = (r2 ? sc.zlo[2] : (r1 ? sc.zlo[1] : sc.zlo[0])) + tz * chunk;
    const int zend = r2 ? sc.zhi[2] : (r1 ? sc.zhi[1] : sc.zhi[0]);
    ze = zb + chunk < zend ? zb + chunk : zend;
    face = r2 ? sc.face[2] : (r1 ? sc.face[1] : sc.face[0]);
}

// position of a CTA in its plane stream: work item -> tile origin and plane range [p, p_last]
template <int TX, int TY, int LO, int HI>
struct Stream {
    int item, p, p_last, x0t, y0t, zb, ze, face;
    SB_DEVI void open(int it, const Sched &sc, int Z) {
        item = it;
        if (item >= sc.nitems) return;
        locate_item(sc, item, TX, TY, x0t, y0t, zb, ze, face);
        (void)Z;
        p = zb - LO;
        p_last = ze - 1 + HI;
    }
    SB_DEVI bool valid(const Sched &sc) const { return item < sc.nitems; }
    SB_DEVI void next(const Sched &sc, int Z) {
        if (++p > p_last) open(item + (int)gridDim.x, sc, Z);
    }
};

// =============================================================================================================
// pass B
// =============================================================================================================
namespace pb {
constexpr int LX = 16, RW = 32 / LX, NW = 12;     // 16 lanes x 4 voxels per row, 2 rows per warp, 12 warps
constexpr int TX = 4 * LX, TY = NW * RW;          // 64 x 24 outputs per plane
constexpr int SX = TX + 8, SY = TY + 6;           // staged box 4|64|4 floats x 3|24|3 rows
constexpr int NSTAGE = 6;                         // planes q-3..q live, two in flight
#ifndef PB_SCATTER_Z
#define PB_SCATTER_Z 1
#endif
#ifndef PB_PF_AHEAD
#define PB_PF_AHEAD 4                             // 8 measured slower (0.1528 vs 0.1444 ms at 256^3)
#endif
#ifndef PB_BACKOFF_NS
#define PB_BACKOFF_NS 128                         // 400 measured the same
#endif
constexpr int PF_AHEAD = PB_PF_AHEAD;             // L2 prefetch distance (planes) ahead of the shared-memory fill
constexpr int COMP_BYTES = ((SX * SY * 4 + 127) / 128) * 128;
constexpr int STAGE_BYTES = 3 * COMP_BYTES;
constexpr int NPSI = 3;                           // psi ring: centre plane of this step + the next two
constexpr int PSI_COMP_BYTES = TX * TY * 4;       // one component of the tile, no halo
constexpr int PSI_STAGE_BYTES = 3 * PSI_COMP_BYTES;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + NPSI * PSI_STAGE_BYTES + 128;
constexpr unsigned TX_BYTES = 3u * SX * SY * 4u;

// PEER: the peer-mode instantiation (face-tagged items, stores into the neighbours, counters, publication); the plain one
// carries none of that code
template <bool PEER>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
    pass_b_tma_kernel(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapy,
                      const __grid_constant__ CUtensorMap mapz, const __grid_constant__ CUtensorMap mpx,
                      const __grid_constant__ CUtensorMap mpy, const __grid_constant__ CUtensorMap mpz, LoopArgs a, int it,
                      Sched sc) {
    pdl_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 127u) & ~127u;
    __shared__ unsigned long long bars[2 * NSTAGE];
    __shared__ unsigned long long skey[NW];
    const unsigned full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[NSTAGE]);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NW); }
        mbar_fence_init();
    }
    pdl_wait();        // everything below reads what the previous launches wrote
    bool fin;
    if (PEER && threadIdx.x == 0) trace_begin(a.trace);
    if (PEER && a.peer_n > 0) {      // peer mode: one warp waits for the maxima every rank published (lane r: rank r), the block follows
        __shared__ int s_fin;
        if (threadIdx.x < 32) {
            const unsigned long long t0 = a.trace ? global_timer_ns() : 0ull;
            const bool f = loop_finished_peer(a, it);
            if (threadIdx.x == 0) { s_fin = f ? 1 : 0; trace_wait(a.trace, 2, t0); }
        }
        __syncthreads();
        fin = s_fin != 0;
    } else {
        fin = loop_finished(a, it);
    }
    if (fin) {
        if (a.check && blockIdx.x == 0 && threadIdx.x == 0 && !a.state->converged) {
            a.state->iters = it;
            a.state->converged = 1;
        }
        return;
    }
    const Dims d = a.d;
    const int X = d.X, XY = d.X * d.Y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lx = lane % LX, ty = warp * RW + lane / LX;
    const unsigned own_off = (unsigned)(((ty + 3) * SX + 4 * lx + 4) * 4);
    const unsigned psi0 = smem + NSTAGE * STAGE_BYTES, psi_own = (unsigned)((ty * TX + 4 * lx) * 4);
    __syncthreads();   // barrier init (above) visible to the block

    typedef Stream<TX, TY, 3, 3> St;
    if (warp == NW) {
        // ---- producer warp: one lane streams nabla_U planes into the ring and prefetches further ahead into L2 ----
        if (lane != 0) return;
        St pr, pf;                         // pf runs PF_AHEAD planes ahead of pr and only prefetches into L2
        pr.open(blockIdx.x, sc, d.Z);
        pf.open(blockIdx.x, sc, d.Z);
        auto prefetch = [&]() {
            if (!pf.valid(sc)) return;
            tma_prefetch_3d(&mapx, pf.x0t, pf.y0t, pf.p + 3);
            tma_prefetch_3d(&mapy, pf.x0t, pf.y0t, pf.p + 3);
            tma_prefetch_3d(&mapz, pf.x0t, pf.y0t, pf.p + 3);
            // psi of plane r is needed when r is the centre, i.e. 3 planes after nabla_U(r) arrives; pull it into L2 a few
            // steps before that (the psi maps carry the TX x TY tile without halo)
            const int r = pf.p - 3;
            if (r >= pf.zb) {
                tma_prefetch_3d(&mpx, pf.x0t, pf.y0t, r + PSI_HALO);
                tma_prefetch_3d(&mpy, pf.x0t, pf.y0t, r + PSI_HALO);
                tma_prefetch_3d(&mpz, pf.x0t, pf.y0t, r + PSI_HALO);
            }
            pf.next(sc, d.Z);
        };
        for (int k = 0; k < PF_AHEAD; ++k) prefetch();
        for (unsigned qi = 0; pr.valid(sc); ++qi) {
            prefetch();
            const unsigned slot = qi % NSTAGE, n = qi / NSTAGE;
            // every warp released the previous plane of this slot; polling with a 128 ns back-off leaves the issue slots of this
            // scheduler to its three consumer warps (-2 % kernel time)
            if (n > 0) mbar_wait_backoff(empty0 + 8 * slot, (n - 1) & 1u, (unsigned)PB_BACKOFF_NS);
            const unsigned dst = smem + slot * STAGE_BYTES, bar = full0 + 8 * slot;
            // psi of the plane that becomes the centre with this nabla_U plane rides on the same barrier; its slot (qi % 3) was
            // last read in the step whose end released this nabla_U slot, so the wait above covers it too
            const bool with_psi = pr.p - 3 >= pr.zb;
            mbar_expect_tx(bar, TX_BYTES + (with_psi ? 3u * PSI_COMP_BYTES : 0u));
            // padded coordinates: plane z sits at z + 3; the box origin (x0t, y0t) is interior (x0t - 4, y0t - 3)
            tma_load_3d(dst, &mapx, bar, pr.x0t, pr.y0t, pr.p + 3);
            tma_load_3d(dst + COMP_BYTES, &mapy, bar, pr.x0t, pr.y0t, pr.p + 3);
            tma_load_3d(dst + 2 * COMP_BYTES, &mapz, bar, pr.x0t, pr.y0t, pr.p + 3);
            if (with_psi) {
                const unsigned pd = psi0 + (qi % NPSI) * PSI_STAGE_BYTES;
                tma_load_3d(pd, &mpx, bar, pr.x0t, pr.y0t, pr.p - 3 + PSI_HALO);
                tma_load_3d(pd + PSI_COMP_BYTES, &mpy, bar, pr.x0t, pr.y0t, pr.p - 3 + PSI_HALO);
                tma_load_3d(pd + 2 * PSI_COMP_BYTES, &mpz, bar, pr.x0t, pr.y0t, pr.p - 3 + PSI_HALO);
            }
            pr.next(sc, d.Z);
        }
        return;
    }

    // ---- consumers ----
    float *__restrict__ P[3] = {a.px, a.py, a.pz};
    unsigned q = 0;                       // planes consumed so far
    MaxCand best{0u, 0u, 0u};
    const unsigned zoff = (unsigned)a.z0 * (unsigned)XY;
#if PB_SCATTER_Z
    // z taps in scatter form: acc[i] accumulates the z sum of the OUTPUT plane (newest plane - 6 + i); a plane that arrives adds its
    // seven tap products to the seven sums in flight, in the reference's order (input planes ascending = k = -3 .. 3, each sum started
    // from 0).  The taps are symmetric (S[i] == S[6 - i] bit for bit: tables and computed filters alike), so a value needs 4 products
    // instead of 7 -- 36 FMUL less per thread and plane of ~870 instructions; the centre quad is re-read from the ring instead.
    float4 acc[7][3];
#else
    float4 win[7][3];                     // nabla_U of this thread's 4 voxels at planes c-3 .. c+3
#endif
    int ack_seen = 0;                     // faces whose acknowledgement this CTA has already waited for
    St cs;
    for (cs.open(blockIdx.x, sc, d.Z); cs.valid(sc); cs.open(cs.item + (int)gridDim.x, sc, d.Z)) {
        const int x0 = cs.x0t + 4 * lx, y = cs.y0t + ty;
        const bool active = x0 < X && y < d.Y;
        const int row = min(x0, X - 4) + X * min(y, d.Y - 1);
        const bool face_item = PEER && a.push && cs.face != 0;
        if (face_item && !(ack_seen & cs.face)) {   // peer mode: the neighbour has read the halo planes this item is about to overwrite
            if (tid == 0) {
                const unsigned long long t0 = a.trace ? global_timer_ns() : 0ull;
                peer_wait_ge(a.my_ack + (cs.face - 1), a.expect_ack, a.peer_error);
                trace_wait(a.trace, 4, t0);
            }
            asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
            ack_seen |= cs.face;
        }
        for (int p = cs.p; p <= cs.p_last; ++p) {
            const unsigned slot = q % NSTAGE;
            const int zc = p - 3;         // centre plane whose window is complete with plane p
            const int o = row + XY * zc;
            const unsigned pslot = psi0 + (q % NPSI) * PSI_STAGE_BYTES + psi_own;
            mbar_wait(full0 + 8 * slot, (q / NSTAGE) & 1u);
            ++q;
            const unsigned sp = smem + slot * STAGE_BYTES + own_off;
#if PB_SCATTER_Z
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 g = lds4(sp + c * COMP_BYTES);
                float pr[4][4];           // pr[t][j] = S[t] * g_j, t = 0 .. 3 (S[6 - t] == S[t])
#pragma unroll
                for (int t = 0; t < 4; ++t) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) pr[t][j] = mul(a.S[t], c4(g, j));
                }
                // output plane p-3+i takes tap S[3 - k] with k = p - (p-3+i) = 3 - i, i.e. S[i]: acc[i] += S[i] * g (i = 6 starts a sum)
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const int t = i < 3 ? i : 6 - i;          // acc[i] is the shifted acc[i + 1] of the previous step
                    acc[i][c] = make_float4(add(acc[i + 1][c].x, pr[t][0]), add(acc[i + 1][c].y, pr[t][1]), add(acc[i + 1][c].z, pr[t][2]),
                                            add(acc[i + 1][c].w, pr[t][3]));
                }
                acc[6][c] = make_float4(add(0.f, pr[0][0]), add(0.f, pr[0][1]), add(0.f, pr[0][2]), add(0.f, pr[0][3]));
            }
#else
#pragma unroll
            for (int k = 0; k < 6; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) win[k][c] = win[k + 1][c];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) win[6][c] = lds4(sp + c * COMP_BYTES);
#endif
            if (zc >= cs.zb) {
                const unsigned sc0 = smem + ((q - 4u) % NSTAGE) * STAGE_BYTES + own_off;   // stage of the centre plane
                // psi of the centre plane arrived with this step's nabla_U plane (a plain LDG here exposed its latency: the
                // register budget of 13 warps/SM leaves no room to issue it early)
                float4 psi4[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) psi4[c] = lds4(pslot + c * PSI_COMP_BYTES);
                float np[3][4], nsq[4];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned s0 = sc0 + c * COMP_BYTES;
#if PB_SCATTER_Z
                    const float4 L = lds4(s0 - 16), R = lds4(s0 + 16), C = lds4(s0);
                    float fz[4] = {acc[0][c].x, acc[0][c].y, acc[0][c].z, acc[0][c].w};     // complete with this plane
#else
                    const float4 L = lds4(s0 - 16), R = lds4(s0 + 16), C = win[3][c];
                    float fz[4] = {0.f, 0.f, 0.f, 0.f};
#endif
                    const float v[12] = {L.x, L.y, L.z, L.w, C.x, C.y, C.z, C.w, R.x, R.y, R.z, R.w};
                    float fx[4] = {0.f, 0.f, 0.f, 0.f}, fy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k = -3; k <= 3; ++k) {
                        const float s = a.S[3 - k];
                        const float4 yk = (k == 0) ? C : lds4(s0 + k * (SX * 4));
#if !PB_SCATTER_Z
                        const float4 zk = win[3 + k][c];
#endif
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            fx[j] = add(fx[j], mul(s, v[4 + j + k]));
                            fy[j] = add(fy[j], mul(s, c4(yk, j)));
#if !PB_SCATTER_Z
                            fz[j] = add(fz[j], mul(s, c4(zk, j)));
#endif
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float f = add(add(fx[j], fy[j]), fz[j]);
                        const float u = mul(f, a.alpha);
                        np[c][j] = sub(c4(psi4[c], j), u);
                        nsq[j] = (c == 0) ? mul(u, u) : add(nsq[j], mul(u, u));
                    }
                }
                if (active) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        *reinterpret_cast<float4 *>(P[c] + o) = make_float4(np[c][0], np[c][1], np[c][2], np[c][3]);
                    // peer mode: the planes next to a slab face also go straight into the neighbour's halo planes over NVLink
                    if (PEER && a.peer_hi[0] && zc >= d.Z - PSI_HALO) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            *reinterpret_cast<float4 *>(a.peer_hi[c] + o) = make_float4(np[c][0], np[c][1], np[c][2], np[c][3]);
                    }
                    if (PEER && a.peer_lo[0] && zc < PSI_HALO) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            *reinterpret_cast<float4 *>(a.peer_lo[c] + o) = make_float4(np[c][0], np[c][1], np[c][2], np[c][3]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) max_cand_update(best, nsq[j], (unsigned)(o + j) + zoff, a.rm);   // global voxel index
                }
            }
            // hand the stage of plane q-4 (0-based: the centre plane just used) back to the producer
            __syncwarp();
            if (lane == 0 && q >= 4u) mbar_arrive(empty0 + 8 * ((q - 4u) % NSTAGE));
        }
        if (face_item) {   // the planes of this item are in the neighbour's halo: count the item there
            // every consumer's stores to the neighbour are issued (block barrier), then ONE thread's system-scope fence orders them
            // -- cumulatively -- before the counter the neighbour polls; the other warps go on with the next item meanwhile
            asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
            if (tid == 0) {
                __threadfence_system();
                if (cs.face == 1 && a.cnt_lo) atomicAdd_system(a.cnt_lo, 1ull);
                if (cs.face == 2 && a.cnt_hi) atomicAdd_system(a.cnt_hi, 1ull);
            }
        }
    }
    const unsigned long long bkey = warp_max_u64(max_cand_key(best, a.rm));
    if (lane == 0) skey[warp] = bkey;
    asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");   // consumers only (the producer warp has left)
    if (tid == 0) {
        unsigned long long m = 0ull;
#pragma unroll
        for (int k = 0; k < NW; ++k) m = skey[k] > m ? skey[k] : m;
        atomicMax(&a.maxkey[it], m);
        if (PEER && a.tickets) {   // peer mode: the last CTA of the launch publishes this rank's maximum into every rank's table
            __threadfence();
            if (atomicAdd(&a.tickets[it], 1u) == gridDim.x - 1u) {
                // the word carries everything its readers need (value + valid bit): relaxed stores, all in flight together
                const unsigned long long v = atomicMax(&a.maxkey[it], 0ull) | PEER_VALID;
                for (int r = 0; r < a.peer_n; ++r)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.pub[r] + (size_t)it * a.peer_n + a.my_rank), "l"(v) : "memory");
            }
        }
        if (PEER) trace_end(a.trace);
    }
}
}  // namespace pb

// =============================================================================================================
// pass A
// =============================================================================================================
#ifndef PA_NW
#define PA_NW 4         // warps per CTA
#endif
#ifndef PA_CTAS
#define PA_CTAS 4       // CTAs per SM
#endif
#ifndef PA_LX
#define PA_LX 8         // lanes per tile row (4 voxels each)
#endif

SB_DEVI void sts4(unsigned saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
SB_DEVI void sts1(unsigned saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }

// ---- the stencil half of pass A, shared by its kernels: one thread = 4 consecutive voxels (a "quad") of one tile row ----
struct Quad {
    int x0, y;                        // volume coordinates of the first voxel
    bool active, x_lo, x_hi, y_lo, y_hi;
};
// w_reg * laplacian(psi) at the centre plane (vector_fields.cu:291-337); on a boundary plane both neighbours of that axis are
// the voxel itself.  sC / sM: shared addresses of the quad in the staged psi planes zc and zc-1 (component stride ARR bytes, row
// pitch SX floats), Zpl: the quad of psi at zc+1; bz: zc is the first or last plane of the VOLUME
template <int SX, int ARR>
SB_DEVI void laplacian_quad(float (&Lw)[3][4], const Quad &qd, bool bz, unsigned sC, unsigned sM, const float4 (&Zpl)[3], float w_reg) {
    const bool by = qd.y_lo || qd.y_hi, edge = by || bz;   // y / z faces are rare: their selects live in a branch interior threads skip
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const unsigned p0 = sC + c * ARR;
        const float4 C = lds4(p0);
        float4 Ym = lds4(p0 - SX * 4), Yp = lds4(p0 + SX * 4), Zp = Zpl[c], Zm = lds4(sM + c * ARR);
        const float xl = lds1(p0 - 4), xr = lds1(p0 + 16);
        // only the first / last voxel of a row can sit on an x face (X % 4 == 0)
        const float xm[4] = {qd.x_lo ? C.x : xl, C.x, C.y, qd.x_hi ? C.w : C.z}, xp[4] = {qd.x_lo ? C.x : C.y, C.z, C.w, qd.x_hi ? C.w : xr};
        if (edge) {
            if (by) Yp = Ym = C;
            if (bz) Zp = Zm = C;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = mul(c4(C, j), -6.f);
            v = add(v, xp[j]);
            v = add(v, xm[j]);
            v = add(v, c4(Yp, j));
            v = add(v, c4(Ym, j));
            v = add(v, c4(Zp, j));
            v = add(v, c4(Zm, j));
            Lw[c][j] = mul(mul(v, -1.f), w_reg);
        }
    }
}
// central differences of the warped TSDF at the centre plane (vector_fields.cu:157-208; both taps on the in-range neighbour at
// a boundary -> +0), nabla_U = (w - phi_global) * grad(w) + w_reg * L (solver.cu:15-33), stored with the replicated halo of 3
// that turns the filter's clamp to edge into plain loads (solver.cu:256,263,270).  w0: shared address of the quad in the warped
// plane zc (row pitch SX floats); wm / wc / wp: the quad of w at zc-1, zc, zc+1; zc is a LOCAL plane
template <int SX>
SB_DEVI void gradient_store_quad(const LoopArgs &a, const Quad &qd, int zc, bool z_lo, bool z_hi, unsigned w0, float4 wm, float4 wc, float4 wp,
                                 float4 g4, const float (&Lw)[3][4]) {
    const bool edge = qd.y_lo || qd.y_hi || z_lo || z_hi;
    float nx[4], ny[4], nz[4], df[4];
    {
        const float4 C = wc, Ym = lds4(w0 - SX * 4), Yp = lds4(w0 + SX * 4);
        const float xl = lds1(w0 - 4), xr = lds1(w0 + 16);
        const float xm[4] = {qd.x_lo ? C.y : xl, C.x, C.y, C.z}, xp[4] = {C.y, C.z, C.w, qd.x_hi ? C.z : xr};
        float4 Y1 = Yp, Y2 = Ym, Z1 = wp, Z2 = wm;
        if (edge) {
            if (qd.y_hi) Y1 = Ym;
            if (qd.y_lo) Y2 = Yp;
            if (z_hi) Z1 = wm;
            if (z_lo) Z2 = wp;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            nx[j] = mul(sub(xp[j], xm[j]), 0.5f);      // __fdividef(., 2.f)
            ny[j] = mul(sub(c4(Y1, j), c4(Y2, j)), 0.5f);
            nz[j] = mul(sub(c4(Z1, j), c4(Z2, j)), 0.5f);
            df[j] = sub(c4(C, j), c4(g4, j));
        }
    }
    const GLayout gl = a.gl;
    const int X = a.d.X;
    const size_t o = gl.at(min(qd.x0, X - 4), min(qd.y, a.d.Y - 1), zc);
    float *__restrict__ G[3] = {a.gx, a.gy, a.gz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float n = (c == 0) ? nx[j] : (c == 1 ? ny[j] : nz[j]);
            u[j] = add(mul(n, df[j]), Lw[c][j]);
        }
        if (qd.active) {
            float *__restrict__ g = G[c];
            const float4 uv = make_float4(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<float4 *>(g + o) = uv;
            if (qd.x0 == 0) *reinterpret_cast<float4 *>(g + o - 4) = make_float4(u[0], u[0], u[0], u[0]);
            if (qd.x0 + 4 == X) *reinterpret_cast<float4 *>(g + o + 4) = make_float4(u[3], u[3], u[3], u[3]);
            if (qd.y_lo) {
#pragma unroll
                for (int k = 1; k <= 3; ++k) *reinterpret_cast<float4 *>(g + o - (size_t)k * gl.PX) = uv;
            }
            if (qd.y_hi) {
#pragma unroll
                for (int k = 1; k <= 3; ++k) *reinterpret_cast<float4 *>(g + o + (size_t)k * gl.PX) = uv;
            }
            if (z_lo) {
#pragma unroll
                for (int k = 1; k <= 3; ++k) *reinterpret_cast<float4 *>(g + o - (size_t)k * gl.plane) = uv;
            }
            if (z_hi) {
#pragma unroll
                for (int k = 1; k <= 3; ++k) *reinterpret_cast<float4 *>(g + o + (size_t)k * gl.plane) = uv;
            }
        }
    }
}

namespace pa {
constexpr int LX = PA_LX, RW = 32 / LX, NW = PA_NW;
constexpr int NTHREADS = NW * 32;
constexpr int TX = 4 * LX, TY = NW * RW;          // outputs per plane
constexpr int SX = TX + 8, SY = TY + 2;           // staged box 4|TX|4 floats x 1|TY|1 rows
constexpr int AHEAD = 1;                          // planes in flight ahead of the one being consumed (2 measured slower)
constexpr int NSTAGE = 3 + AHEAD;                 // planes p-2, p-1, p live + AHEAD in flight (+ L2 prefetch)
constexpr int PF_AHEAD = 4;                       // L2 prefetch distance ahead of the shared-memory fill
constexpr int ARR_BYTES = ((SX * SY * 4 + 127) / 128) * 128;
constexpr int STAGE_BYTES = 3 * ARR_BYTES;        // psi x, y, z
constexpr int WBUF_BYTES = ARR_BYTES;             // one plane of warped TSDF (tile + cross halo), same geometry
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 3 * WBUF_BYTES + 128;
constexpr unsigned TX_BYTES = 3u * SX * SY * 4u;
constexpr int NHALO = 2 * TX + 2 * TY;            // cross halo cells of a plane: rows y0-1, y0+TY and columns x0-1, x0+TX
static_assert(NHALO <= NTHREADS, "one halo cell per thread");

// One CTA = NW warps marching a TX x TY column along z, all in lockstep (one block barrier per plane).  There is no producer
// warp: after the barrier of step q the stage of plane q-3 is free by construction, and thread 0 refills it with plane q+1
// (no "empty" barriers; a CTA of whole consumer warps lets PA_CTAS CTAs/SM keep 128 registers per thread).
// Step for plane p (centre zc = p-1):
//   w_reg * laplacian of psi at zc (psi planes zc-1, zc, zc+1 are read from the ring)
//   w(p) = phi_n o psi at this thread's quad and at its cross-halo cell -> shared (triple buffer) | barrier |
//   central differences of w at zc, nabla_U -> global (+ replicated halo)
// Tried and measured slower (profiles/r1_tuning_log.md): a producer warp (caps the CTA at 96 registers), all gathers of a
// step in flight before the first use / across the stencil (more instructions + spills, same stall time), bigger CTAs.
template <bool TEX, bool PEER>
__global__ void __launch_bounds__(NTHREADS, PA_CTAS)
    pass_a_tma_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                      const __grid_constant__ CUtensorMap m2, LoopArgs a, int it, Sched sc) {
    pdl_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 127u) & ~127u;
    const unsigned wbuf0 = smem + NSTAGE * STAGE_BYTES;
    __shared__ unsigned long long bars[NSTAGE];
    const unsigned full0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(full0 + 8 * s, 1);
        mbar_fence_init();
    }
    pdl_wait();        // everything below reads what the previous launches wrote
    if (a.a_uses_max ? loop_finished(a, it) : (a.check && a.state->converged)) {
        if (a.check && blockIdx.x == 0 && threadIdx.x == 0 && !a.state->converged) {
            a.state->iters = it;
            a.state->converged = 1;
        }
        return;
    }
    if (PEER && threadIdx.x == 0) trace_begin(a.trace);

    const Dims d = a.d, dg = a.dg;           // local slab extent / global volume
    const int X = d.X, XY = d.X * d.Y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();   // barrier init (above) visible to the block

    typedef Stream<TX, TY, 1, 1> St;
    St pr;                                    // position of the feeder (thread 0) in the CTA's plane stream
    unsigned qi = 0;
    int halo_seen = 0;                        // faces whose halo counter thread 0 has already waited for
    auto feed = [&]() {
        if (!pr.valid(sc)) return;
        if (PEER && a.wait_halo && pr.face != 0 && !(halo_seen & pr.face)) {
            // peer mode: the first item of this CTA that reads the halo planes of that face -- the neighbour's pass B of the
            // previous iteration has stored into them once the counter says so
            const unsigned long long want = pr.face == 1 ? a.expect_lo : a.expect_hi;
            if (want) {
                const unsigned long long t0 = a.trace ? global_timer_ns() : 0ull;
                peer_wait_ge(a.my_cnt + (pr.face - 1), want, a.peer_error);
                trace_wait(a.trace, 4, t0);
            }
            asm volatile("fence.proxy.async;" ::: "memory");     // the planes are read by TMA (async proxy)
            halo_seen |= pr.face;
        }
        const unsigned slot = qi % NSTAGE, dst = smem + slot * STAGE_BYTES, bar = full0 + 8 * slot;
        mbar_expect_tx(bar, TX_BYTES);
        // the box starts 4 floats / 1 row before the tile (out-of-range elements arrive as zeros); the psi planes are
        // allocated with PSI_HALO halo planes on either side, so local plane p is plane p + PSI_HALO of the tensor map
        tma_load_3d(dst, &m0, bar, pr.x0t - 4, pr.y0t - 1, pr.p + PSI_HALO);
        tma_load_3d(dst + ARR_BYTES, &m1, bar, pr.x0t - 4, pr.y0t - 1, pr.p + PSI_HALO);
        tma_load_3d(dst + 2 * ARR_BYTES, &m2, bar, pr.x0t - 4, pr.y0t - 1, pr.p + PSI_HALO);
        if (pr.p + PF_AHEAD <= pr.p_last) {         // L2 prefetch further down the same column
            tma_prefetch_3d(&m0, pr.x0t - 4, pr.y0t - 1, pr.p + PF_AHEAD + PSI_HALO);
            tma_prefetch_3d(&m1, pr.x0t - 4, pr.y0t - 1, pr.p + PF_AHEAD + PSI_HALO);
            tma_prefetch_3d(&m2, pr.x0t - 4, pr.y0t - 1, pr.p + PF_AHEAD + PSI_HALO);
        }
        ++qi;
        pr.next(sc, d.Z);
    };
    if (tid == 0) {
        pr.open(blockIdx.x, sc, d.Z);
        for (int k = 0; k < AHEAD; ++k) feed();
    }

    const int lx = lane % LX, ty = warp * RW + lane / LX;
    const unsigned own_off = (unsigned)(((ty + 1) * SX + 4 * lx + 4) * 4);
    // cross-halo cell served by this thread (threads 0 .. NHALO-1): position inside the staged box
    int hx = -1, hy = -1;
    if (tid < TX) { hx = 4 + tid; hy = 0; }
    else if (tid < 2 * TX) { hx = 4 + tid - TX; hy = TY + 1; }
    else if (tid < 2 * TX + TY) { hx = 3; hy = 1 + tid - 2 * TX; }
    else if (tid < NHALO) { hx = 4 + TX; hy = 1 + tid - 2 * TX - TY; }
    const unsigned halo_off = (unsigned)((hy * SX + hx) * 4);
    const float *__restrict__ pn = a.pn;
    unsigned q = 0;
    St cs;
    for (cs.open(blockIdx.x, sc, d.Z); cs.valid(sc); cs.open(cs.item + (int)gridDim.x, sc, d.Z)) {
        Quad qd;
        qd.x0 = cs.x0t + 4 * lx; qd.y = cs.y0t + ty;
        qd.active = qd.x0 < X && qd.y < d.Y;
        const int row = min(qd.x0, X - 4) + X * min(qd.y, d.Y - 1);
        qd.y_lo = (qd.y == 0); qd.y_hi = (qd.y == d.Y - 1);
        qd.x_lo = (qd.x0 == 0); qd.x_hi = (qd.x0 + 4 == X);
        const int hgx = cs.x0t - 4 + hx, hgy = cs.y0t - 1 + hy;       // volume coordinates of the halo cell
        const bool halo_on = hx >= 0 && hgx >= 0 && hgx < X && hgy >= 0 && hgy < d.Y;
        float4 wm = make_float4(0.f, 0.f, 0.f, 0.f), wc = wm;        // this thread's quad of w at planes z-1 and z
        for (int p = cs.p; p <= cs.p_last; ++p) {
            const unsigned slot = q % NSTAGE;
            const int zc = p - 1;             // plane that becomes the centre when plane p has arrived
            const bool centre_on = zc >= cs.zb;
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (centre_on) g4 = *reinterpret_cast<const float4 *>(a.pg + row + XY * zc);
            mbar_wait(full0 + 8 * slot, (q / NSTAGE) & 1u);
            const unsigned wcur = wbuf0 + (q % 3u) * WBUF_BYTES;             // warped plane p   (written now)
            const unsigned wctr = wbuf0 + ((q + 2u) % 3u) * WBUF_BYTES;      // warped plane p-1 (written in the last step)
            ++q;
            const unsigned stP = smem + slot * STAGE_BYTES;                                  // plane p   (z+1)
            const unsigned sC = smem + ((q + NSTAGE - 2u) % NSTAGE) * STAGE_BYTES + own_off; // plane p-1 (centre)
            const unsigned sM = smem + ((q + NSTAGE - 3u) % NSTAGE) * STAGE_BYTES + own_off; // plane p-2 (z-1)
            float4 zp[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) zp[k] = lds4(stP + own_off + k * ARR_BYTES);
            const bool plane_on = (a.z0 + p >= 0 && a.z0 + p < dg.Z);        // planes outside the volume are never used
            const bool z_lo = (a.z0 + zc == 0), z_hi = (a.z0 + zc == dg.Z - 1);   // global faces only

            float Lw[3][4];
            if (centre_on) laplacian_quad<SX, ARR_BYTES>(Lw, qd, z_lo || z_hi, sC, sM, zp, a.w_reg);

            // ---- warp of plane p: own quad + this thread's cross-halo cell ----
            float4 wp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (plane_on) {
                if (qd.active) {
                    if (TEX) {
                        wp.x = warp_sample_tex(a.pn_tex, a.ashift, a.amask, zp[0].x, zp[1].x, zp[2].x, dg);
                        wp.y = warp_sample_tex(a.pn_tex, a.ashift, a.amask, zp[0].y, zp[1].y, zp[2].y, dg);
                        wp.z = warp_sample_tex(a.pn_tex, a.ashift, a.amask, zp[0].z, zp[1].z, zp[2].z, dg);
                        wp.w = warp_sample_tex(a.pn_tex, a.ashift, a.amask, zp[0].w, zp[1].w, zp[2].w, dg);
                    } else {
                        wp.x = warp_sample(pn, zp[0].x, zp[1].x, zp[2].x, dg, X, XY);
                        wp.y = warp_sample(pn, zp[0].y, zp[1].y, zp[2].y, dg, X, XY);
                        wp.z = warp_sample(pn, zp[0].z, zp[1].z, zp[2].z, dg, X, XY);
                        wp.w = warp_sample(pn, zp[0].w, zp[1].w, zp[2].w, dg, X, XY);
                    }
                }
                if (halo_on) {
                    const float hxv = lds1(stP + halo_off), hyv = lds1(stP + halo_off + ARR_BYTES), hzv = lds1(stP + halo_off + 2 * ARR_BYTES);
                    sts1(wcur + halo_off, TEX ? warp_sample_tex(a.pn_tex, a.ashift, a.amask, hxv, hyv, hzv, dg) : warp_sample(pn, hxv, hyv, hzv, dg, X, XY));
                }
                sts4(wcur + own_off, wp);
            }
            __syncthreads();                  // warped plane p (and p-1) visible to the CTA; ring stage of plane p-3 is free
            if (tid == 0) feed();

            if (centre_on) gradient_store_quad<SX>(a, qd, zc, z_lo, z_hi, wctr + own_off, wm, wc, wp, g4, Lw);
            wm = wc; wc = wp;
        }
        // peer mode: every plane of this item has been staged (thread 0 waited for the last one itself), so the item no longer
        // reads the halo planes -- tell the neighbour on that face that it may overwrite them
        if (PEER && a.wait_halo && cs.face != 0 && tid == 0) {
            if (cs.face == 1 && a.ack_lo) atomicAdd_system(a.ack_lo, 1ull);
            if (cs.face == 2 && a.ack_hi) atomicAdd_system(a.ack_hi, 1ull);
        }
    }
    if (PEER && tid == 0) trace_end(a.trace);
}
}  // namespace pa

// =============================================================================================================
// pass A, software-pipelined across planes (the default)
// =============================================================================================================
// Same arithmetic and the same outputs as pa::pass_a_tma_kernel.  There the two gather4 fetches of a sample are consumed right
// after they are issued: five dependent texture round trips per thread and plane, 16 warps per SM -- issue slots are used 51 % of
// the time (profiles/r1_ncu_tma_v9_summary.md).  Here the fetches of plane p are issued in step p and consumed in step p+1: a whole
// step of stencil arithmetic (and the other warps' steps) sits between a fetch and its first use.  Only the 8 result registers of
// a fetch stay live across the step (5 samples per thread: 40 registers); the interpolation weights are recomputed from psi,
// which is still in the ring.  Price: the centre plane trails the newest plane by two instead of one (one more pipeline-fill step
// per work item, one more ring stage).  This is the default pass A; the kernel above stays for volumes without a gather atlas and
// as variant 4.
#ifndef PA2_CTAS
#define PA2_CTAS 4      // measured at 256^3: 4 CTAs/SM (128 registers) 0.182 ms, 3 (168) 0.199, 2 (210) 0.249; the round-1 kernel 0.192
#endif
#ifndef PA2_AHEAD
#define PA2_AHEAD 2     // planes requested ahead of the one being consumed: with 1 the TMA latency of every plane was exposed (21 % of the
#endif                  // stall samples sat in the wait for the `full` barrier, profiles/r2_ncu_pa2_*.txt)
namespace pa2 {
constexpr int LX = pa::LX, RW = pa::RW, NW = pa::NW, NTHREADS = pa::NTHREADS;
constexpr int TX = pa::TX, TY = pa::TY, SX = pa::SX, SY = pa::SY;     // same tile: the tensor maps are shared
// ring: when thread 0 refills after the barrier of step p, planes p-2, p-1 and p are still needed (they are p-3 .. p-1 of the
// next step) and AHEAD planes are in flight or about to be: the load goes into the stage of plane p-3, last read in this step
constexpr int AHEAD = PA2_AHEAD, NSTAGE = 3 + AHEAD;
#ifndef PA2_PF_AHEAD
#define PA2_PF_AHEAD 4       // 2, 6 and 8 measured the same (0.1695 - 0.1712 ms)
#endif
constexpr int NWB = 2;       // warped planes p-1 (written in this step) and p-2 (its x / y neighbours read in this step); the block
                             // barrier at the end of the step separates the reads of a buffer from its next writes
constexpr int PF_AHEAD = PA2_PF_AHEAD;       // L2 prefetch distance ahead of the shared-memory fill
constexpr int ARR_BYTES = pa::ARR_BYTES, STAGE_BYTES = 3 * ARR_BYTES;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + NWB * ARR_BYTES + 128;
constexpr unsigned TX_BYTES = pa::TX_BYTES;
constexpr int NHALO = pa::NHALO;

template <bool PEER>
__global__ void __launch_bounds__(NTHREADS, PA2_CTAS)
    pass_a_pipe_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                       const __grid_constant__ CUtensorMap m2, LoopArgs a, int it, Sched sc) {
    pdl_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 127u) & ~127u;
    const unsigned wbuf0 = smem + NSTAGE * STAGE_BYTES;
    __shared__ unsigned long long bars[NSTAGE];
    const unsigned full0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(full0 + 8 * s, 1);
        mbar_fence_init();
    }
    pdl_wait();
    if (a.a_uses_max ? loop_finished(a, it) : (a.check && a.state->converged)) {
        if (a.check && blockIdx.x == 0 && threadIdx.x == 0 && !a.state->converged) {
            a.state->iters = it;
            a.state->converged = 1;
        }
        return;
    }
    if (PEER && threadIdx.x == 0) trace_begin(a.trace);

    const Dims d = a.d, dg = a.dg;
    const int X = d.X, XY = d.X * d.Y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();

    // planes zb-1 .. ze+1 of an item: the last step only drains the pipeline (its plane is staged but not used)
    typedef Stream<TX, TY, 1, 2> St;
    St pr;
    unsigned qi = 0;
    int halo_seen = 0;
    auto feed = [&]() {
        if (!pr.valid(sc)) return;
        if (PEER && a.wait_halo && pr.face != 0 && !(halo_seen & pr.face)) {
            const unsigned long long want = pr.face == 1 ? a.expect_lo : a.expect_hi;
            if (want) {
                const unsigned long long t0 = a.trace ? global_timer_ns() : 0ull;
                peer_wait_ge(a.my_cnt + (pr.face - 1), want, a.peer_error);
                trace_wait(a.trace, 4, t0);
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            halo_seen |= pr.face;
        }
        const unsigned slot = qi % NSTAGE, dst = smem + slot * STAGE_BYTES, bar = full0 + 8 * slot;
        mbar_expect_tx(bar, TX_BYTES);
        tma_load_3d(dst, &m0, bar, pr.x0t - 4, pr.y0t - 1, pr.p + PSI_HALO);
        tma_load_3d(dst + ARR_BYTES, &m1, bar, pr.x0t - 4, pr.y0t - 1, pr.p + PSI_HALO);
        tma_load_3d(dst + 2 * ARR_BYTES, &m2, bar, pr.x0t - 4, pr.y0t - 1, pr.p + PSI_HALO);
        if (pr.p + PF_AHEAD < pr.p_last) {
            tma_prefetch_3d(&m0, pr.x0t - 4, pr.y0t - 1, pr.p + PF_AHEAD + PSI_HALO);
            tma_prefetch_3d(&m1, pr.x0t - 4, pr.y0t - 1, pr.p + PF_AHEAD + PSI_HALO);
            tma_prefetch_3d(&m2, pr.x0t - 4, pr.y0t - 1, pr.p + PF_AHEAD + PSI_HALO);
        }
        ++qi;
        pr.next(sc, d.Z);
    };
    if (tid == 0) {
        pr.open(blockIdx.x, sc, d.Z);
        for (int k = 0; k < AHEAD; ++k) feed();
    }

    const int lx = lane % LX, ty = warp * RW + lane / LX;
    const unsigned own_off = (unsigned)(((ty + 1) * SX + 4 * lx + 4) * 4);
    int hx = -1, hy = -1;         // cross-halo cell served by this thread (threads 0 .. NHALO-1), as in pa.  (Spreading the cells evenly
                                  // over the four warps -- 24 lanes each -- measured slower: 0.1729 vs 0.1697 ms; all four warps then run
                                  // the halo path.)
    if (tid < TX) { hx = 4 + tid; hy = 0; }
    else if (tid < 2 * TX) { hx = 4 + tid - TX; hy = TY + 1; }
    else if (tid < 2 * TX + TY) { hx = 3; hy = 1 + tid - 2 * TX; }
    else if (tid < NHALO) { hx = 4 + TX; hy = 1 + tid - 2 * TX - TY; }
    const unsigned halo_off = (unsigned)((hy * SX + hx) * 4);
    unsigned q = 0;
    St cs;
    for (cs.open(blockIdx.x, sc, d.Z); cs.valid(sc); cs.open(cs.item + (int)gridDim.x, sc, d.Z)) {
        Quad qd;
        qd.x0 = cs.x0t + 4 * lx; qd.y = cs.y0t + ty;
        qd.active = qd.x0 < X && qd.y < d.Y;
        const int row = min(qd.x0, X - 4) + X * min(qd.y, d.Y - 1);
        qd.y_lo = (qd.y == 0); qd.y_hi = (qd.y == d.Y - 1);
        qd.x_lo = (qd.x0 == 0); qd.x_hi = (qd.x0 + 4 == X);
        const int hgx = cs.x0t - 4 + hx, hgy = cs.y0t - 1 + hy;
        const bool halo_on = hx >= 0 && hgx >= 0 && hgx < X && hgy >= 0 && hgy < d.Y;
        float4 wm = make_float4(0.f, 0.f, 0.f, 0.f), wc = wm;        // the quad of w at planes p-3 and p-2
        float4 lo[4], hi[4], hlo, hhi;                                // gathers in flight: the quad and the halo cell at plane p-1
        bool pend = false;                                            // ... if that plane lies inside the volume
#pragma unroll
        for (int j = 0; j < 4; ++j) lo[j] = hi[j] = wm;
        hlo = hhi = wm;
        for (int p = cs.p; p <= cs.p_last; ++p) {
            const unsigned slot = q % NSTAGE;
            const int zc = p - 2;             // centre: its z+1 neighbour w(p-1) is completed in this step
            const bool centre_on = zc >= cs.zb;
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (centre_on) g4 = *reinterpret_cast<const float4 *>(a.pg + row + XY * zc);
            mbar_wait(full0 + 8 * slot, (q / NSTAGE) & 1u);
            ++q;
            const unsigned stP = smem + slot * STAGE_BYTES;                                       // plane p
            const unsigned sP1 = smem + ((q + NSTAGE - 2u) % NSTAGE) * STAGE_BYTES;               // plane p-1
            const unsigned sC = smem + ((q + NSTAGE - 3u) % NSTAGE) * STAGE_BYTES + own_off;      // plane p-2 (centre)
            const unsigned sM = smem + ((q + 2u * NSTAGE - 4u) % NSTAGE) * STAGE_BYTES + own_off; // plane p-3
            const unsigned wnew = wbuf0 + ((q + 1u) % NWB) * ARR_BYTES;                           // warped plane p-1 (completed now)
            const unsigned wctr = wbuf0 + (q % NWB) * ARR_BYTES;                                  // warped plane p-2 (completed in the last step)

            // ---- 1. the gathers of plane p-1 have had a whole step to arrive: interpolate (weights recomputed from psi(p-1)) ----
            float4 z1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) z1[k] = lds4(sP1 + own_off + k * ARR_BYTES);
            float4 wp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pend) {
                if (qd.active) {
                    float w4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        TexSample s;
                        s.lo = lo[j]; s.hi = hi[j];
                        tex_weights(s, c4(z1[0], j), c4(z1[1], j), c4(z1[2], j), dg);
                        w4[j] = tex_finish(s);
                    }
                    wp = make_float4(w4[0], w4[1], w4[2], w4[3]);
                }
                if (halo_on) {
                    TexSample s;
                    s.lo = hlo; s.hi = hhi;
                    tex_weights(s, lds1(sP1 + halo_off), lds1(sP1 + halo_off + ARR_BYTES), lds1(sP1 + halo_off + 2 * ARR_BYTES), dg);
                    sts1(wnew + halo_off, tex_finish(s));
                }
                sts4(wnew + own_off, wp);
            }
            // ---- 2. issue the gathers of plane p (consumed in the next step) ----
            pend = (a.z0 + p >= 0 && a.z0 + p < dg.Z) && p < cs.p_last;
            if (pend) {
                if (qd.active) {
                    float4 zp[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) zp[k] = lds4(stP + own_off + k * ARR_BYTES);
#pragma unroll
                    for (int j = 0; j < 4; ++j) tex_fetch(lo[j], hi[j], a.pn_tex, a.ashift, a.amask, c4(zp[0], j), c4(zp[1], j), c4(zp[2], j), dg);
                }
                if (halo_on)
                    tex_fetch(hlo, hhi, a.pn_tex, a.ashift, a.amask, lds1(stP + halo_off), lds1(stP + halo_off + ARR_BYTES), lds1(stP + halo_off + 2 * ARR_BYTES), dg);
            }
            // ---- 3. nabla_U of the centre plane p-2 ----
            if (centre_on) {
                const bool z_lo = (a.z0 + zc == 0), z_hi = (a.z0 + zc == dg.Z - 1);   // global faces only
                float Lw[3][4];
                laplacian_quad<SX, ARR_BYTES>(Lw, qd, z_lo || z_hi, sC, sM, z1, a.w_reg);
                gradient_store_quad<SX>(a, qd, zc, z_lo, z_hi, wctr + own_off, wm, wc, wp, g4, Lw);
            }
            __syncthreads();                  // warped plane p-1 visible to the CTA (and p-2 no longer read); ring stage of plane p-3 is free
            if (tid == 0) feed();
            wm = wc; wc = wp;
        }
        if (PEER && a.wait_halo && cs.face != 0 && tid == 0) {
            if (cs.face == 1 && a.ack_lo) atomicAdd_system(a.ack_lo, 1ull);
            if (cs.face == 2 && a.ack_hi) atomicAdd_system(a.ack_hi, 1ull);
        }
    }
    if (PEER && tid == 0) trace_end(a.trace);
}
}  // namespace pa2

int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// (Measured in round 2 and NOT adopted: scoring a candidate by the busiest SM -- sum over its resident CTAs -- instead of the busiest CTA
// picks longer chunks (256^3: 8 x 32 planes instead of 9 x 29) and pass A ran 9 % slower: with 1024 items on 592 CTAs a quarter of the
// CTAs has one item instead of two and the SMs run half empty for the second half of the kernel.  Even work per CTA matters more than
// fewer pipeline prologues.)
// Number of z chunks per range.  The CTAs take items round robin (item b, b + G, ...), ranges in the given order; a chunk of
// c planes costs c + halo_planes * halo_cost plane-steps.  Ranges of fewer than 16 planes stay one chunk; for the others the
// chunk count that minimises the busiest CTA's load -- given the items already placed before it -- is taken (chunks of >= 16
// planes -- 8 for ranges under 64 planes or when 16-plane chunks cannot fill the CTAs -- so that the pipeline prologue stays a small share).
Sched make_sched(const Dims d, const ZRanges &zr, int TX, int TY, int halo_planes, double halo_cost, int ctas, double *worst_load = nullptr) {
    Sched s;
    s.tiles_x = (d.X + TX - 1) / TX;
    s.tiles_y = (d.Y + TY - 1) / TY;
    const int xy = s.tiles_x * s.tiles_y;
    s.nr = zr.n;
    s.nitems = 0;
    for (int r = 0; r < 3; ++r) { s.zlo[r] = s.zhi[r] = 0; s.nz[r] = 0; s.zchunk[r] = 1; s.face[r] = 0; }
    std::vector<double> load(ctas, 0.0), trial(ctas);
    for (int r = 0; r < zr.n && r < MAX_ZRANGES; ++r) {
        const int Z = zr.hi[r] - zr.lo[r];
        s.zlo[r] = zr.lo[r]; s.zhi[r] = zr.hi[r]; s.face[r] = zr.face[r];
        if (Z <= 0) continue;
        int best_nz = 1;
        double best_score = 1e300;
        // chunks of >= 16 planes keep the pipeline prologue a small share; thin slabs (many ranks) and small volumes (16-plane
        // chunks would leave CTAs without work, e.g. 128^3) may go down to 8: there parallelism matters more
        const int min_chunk = (Z < 64 || xy * ((Z + 15) / 16) < ctas) ? 8 : 16;
        for (int nz = 1; nz <= 64; ++nz) {
            const int chunk = (Z + nz - 1) / nz;
            if (chunk < min_chunk && nz > 1) break;
            if (zr.face[r] != 0 && nz > 1) break;      // a face range is ONE chunk: every rank counts tiles_x * tiles_y items per face
            const int nch = (Z + chunk - 1) / chunk;
            trial = load;
            int item = s.nitems;
            for (int k = 0; k < nch; ++k) {
                const int planes = (k + 1 < nch ? chunk : Z - chunk * (nch - 1));
                const double cost = planes + halo_planes * halo_cost + 1.0;      // + 1: fixed cost of opening an item
                for (int t = 0; t < xy; ++t, ++item) trial[item % ctas] += cost;
            }
            double worst = 0.0;
            for (int b = 0; b < ctas; ++b) worst = trial[b] > worst ? trial[b] : worst;
            if (worst < best_score - 1e-9) { best_score = worst; best_nz = nz; }
        }
        s.zchunk[r] = (Z + best_nz - 1) / best_nz;
        s.nz[r] = (Z + s.zchunk[r] - 1) / s.zchunk[r];
        int item = s.nitems;
        for (int k = 0; k < s.nz[r]; ++k) {
            const int planes = (k + 1 < s.nz[r] ? s.zchunk[r] : Z - s.zchunk[r] * (s.nz[r] - 1));
            for (int t = 0; t < xy; ++t, ++item) load[item % ctas] += planes + halo_planes * halo_cost + 1.0;
        }
        s.nitems += xy * s.nz[r];
    }
    if (worst_load) {
        double w = 0.0;
        for (int b = 0; b < ctas; ++b) w = load[b] > w ? load[b] : w;
        *worst_load = w;
    }
    return s;
}

// the schedule depends on the shape and the ranges only: computed once per distinct launch geometry (the solver thread only)
Sched cached_sched(const Dims d, const ZRanges &zr, int TX, int TY, int halo_planes, double halo_cost, int ctas) {
    struct Key { int v[16]; };
    static std::vector<std::pair<Key, Sched>> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Key k{};
    k.v[0] = d.X; k.v[1] = d.Y; k.v[2] = d.Z; k.v[3] = TX; k.v[4] = TY; k.v[5] = ctas; k.v[6] = zr.n;
    for (int r = 0; r < MAX_ZRANGES; ++r) { k.v[7 + 3 * r] = r < zr.n ? zr.lo[r] : 0; k.v[8 + 3 * r] = r < zr.n ? zr.hi[r] : 0; k.v[9 + 3 * r] = r < zr.n ? zr.face[r] : 0; }
    for (auto &e : cache)
        if (!memcmp(&e.first, &k, sizeof k)) return e.second;
    const Sched sc = make_sched(d, zr, TX, TY, halo_planes, halo_cost, ctas);
    if (cache.size() > 64) cache.clear();
    cache.emplace_back(k, sc);
    return sc;
}

LaunchInfo launch_info(const Sched &sc, int grid) {
    LaunchInfo li{grid, {0, 0, 0}};
    for (int r = 0; r < 3; ++r)
        if (sc.face[r] >= 0 && sc.face[r] < 3) li.face_items[sc.face[r]] += sc.tiles_x * sc.tiles_y * sc.nz[r];
    return li;
}

}  // namespace

// Peer mode: the local planes [lo, hi) of a launch as | lower face chunk [lo, lo+c) | upper face chunk [hi-c, hi) | middle |.
// The face chunks are ordinary full-length work items (no extra pipeline prologue for a handful of planes), exactly one z chunk
// each and issued FIRST: in pass B they carry the planes the neighbours need (stored into the neighbour's halo planes and counted
// there long before the middle of the slab is done), in pass A they are the only items that read halo planes (so they can
// acknowledge early, and the neighbour's next pass B never waits).  c is the face chunk that minimises the busiest CTA's load
// under the static round-robin assignment (>= 4: the planes a neighbour needs / the planes that read halo planes).
ZRanges plan_peer_ranges(const Dims d, int pass, int lo, int hi, bool has_lo, bool has_hi, int sms) {
    if (sms <= 0) sms = sm_count();
    const int TX = pass ? pb::TX : pa::TX, TY = pass ? pb::TY : pa::TY;
    const int ctas = pass ? sms : PA_CTAS * sms;
    const int halo_planes = pass ? 6 : 2;
    const double halo_cost = pass ? 0.35 : 0.5;
    const int nfaces = (has_lo ? 1 : 0) + (has_hi ? 1 : 0), Z = hi - lo;
    ZRanges best{1, {lo, 0, 0}, {hi, 0, 0}, {0, 0, 0}};
    if (nfaces == 0) return best;
    double best_w = 1e300;
    // pass B: up to 64 planes (a limit of 32 kept the face chunk of a 128-plane slab with ONE face at 21 planes next to 54-plane middle
    // chunks: 93 us against 80 us for three equal chunks); pass A keeps the limit its multi-GPU measurements were taken with
    const int climit = pass ? 64 : 32;
    const int cmax = Z / nfaces < climit ? Z / nfaces : climit;
    for (int c = 4; c <= cmax; ++c) {
        ZRanges zr{0, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        int mlo = lo, mhi = hi;
        if (has_lo) { zr.lo[zr.n] = lo; zr.hi[zr.n] = lo + c; zr.face[zr.n] = 1; ++zr.n; mlo = lo + c; }
        if (has_hi) { zr.lo[zr.n] = hi - c; zr.hi[zr.n] = hi; zr.face[zr.n] = 2; ++zr.n; mhi = hi - c; }
        if (mhi > mlo) { zr.lo[zr.n] = mlo; zr.hi[zr.n] = mhi; zr.face[zr.n] = 0; ++zr.n; }
        double w = 0.0;
        make_sched(d, zr, TX, TY, halo_planes, halo_cost, ctas, &w);
        if (w < best_w - 1e-9) { best_w = w; best = zr; }
    }
    return best;
}

// 16 B vector accesses and TMA row pitches need X % 4 == 0
bool tiled_supported(const Dims d) { return d.X % 4 == 0 && d.X >= 32 && d.Y >= 8 && d.Z >= 8; }

TmaMaps *tma_maps_create(const LoopArgs &a) {
    if (!get_tensor_map_encoder()) return nullptr;
    TmaMaps *m = new TmaMaps();
    float *g[3] = {a.gx, a.gy, a.gz};
    const size_t XYp = (size_t)a.d.X * a.d.Y;
    const float *in[3] = {a.px - PSI_HALO * XYp, a.py - PSI_HALO * XYp, a.pz - PSI_HALO * XYp};   // allocations start PSI_HALO planes before local plane 0
    bool ok = true;
    for (int c = 0; c < 3; ++c) ok = ok && encode_map_3d(&m->g[c], g[c], a.gl.PX, a.gl.PY, a.gl.PZ, pb::SX, pb::SY);
    for (int c = 0; c < 3; ++c) ok = ok && encode_map_3d(&m->in[c], in[c], a.d.X, a.d.Y, a.d.Z + 2 * PSI_HALO, pa::SX, pa::SY);
    for (int c = 0; c < 3; ++c) ok = ok && encode_map_3d(&m->pb_psi[c], in[c], a.d.X, a.d.Y, a.d.Z + 2 * PSI_HALO, pb::TX, pb::TY);
    ok = ok && cudaFuncSetAttribute(pb::pass_b_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb::SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pb::pass_b_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb::SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pa::pass_a_tma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa::SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pa::pass_a_tma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa::SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pa::pass_a_tma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa::SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pa::pass_a_tma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa::SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pa2::pass_a_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa2::SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pa2::pass_a_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa2::SMEM_BYTES) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "sobfu_b200: TMA tensor maps unavailable; using the generic kernels\n");
        cudaGetLastError();
        delete m;
        return nullptr;
    }
    return m;
}

void tma_maps_destroy(TmaMaps *m) { delete m; }

LaunchInfo launch_pass_b_tma(const LoopArgs &a, const TmaMaps *m, int it, const ZRanges &zr, cudaStream_t st) {
    const int ctas = sm_count();
    const Sched sc = cached_sched(a.d, zr, pb::TX, pb::TY, 6, 0.35, ctas);
    if (sc.nitems == 0) return LaunchInfo{0, {0, 0, 0}};
    const int grid = sc.nitems < ctas ? sc.nitems : ctas;
    if (a.peer_n > 0) launch_pdl(pb::pass_b_tma_kernel<true>, grid, (pb::NW + 1) * 32, pb::SMEM_BYTES, st, m->g[0], m->g[1], m->g[2], m->pb_psi[0], m->pb_psi[1], m->pb_psi[2], a, it, sc);
    else launch_pdl(pb::pass_b_tma_kernel<false>, grid, (pb::NW + 1) * 32, pb::SMEM_BYTES, st, m->g[0], m->g[1], m->g[2], m->pb_psi[0], m->pb_psi[1], m->pb_psi[2], a, it, sc);
    return launch_info(sc, grid);
}

// alternative pass A kernels are selected per solver (sobfu_b200_solver_set_variant) through this per-thread switch, set by the host
// loop before it launches; 0 = the default kernels
static thread_local int g_pass_a_variant = 0;
void set_pass_a_variant(int v) { g_pass_a_variant = v; }

LaunchInfo launch_pass_a_tma(const LoopArgs &a, const TmaMaps *m, int it, int log, const ZRanges &zr, cudaStream_t st) {
    if (log) {   // logging iterations (rare): materialise the warped plane, then the generic kernel that also sums the energies
        launch_initial_warp(a, st);
        launch_pass_a_generic(a, it, log, st);
        return LaunchInfo{0, {0, 0, 0}};
    }
    const bool peer = a.peer_n > 0 && a.wait_halo;
    if (g_pass_a_variant != 4 && a.pn_tex) {      // default: software-pipelined gathers (one more pipeline-fill step per item: 3 halo steps)
        const int pctas = PA2_CTAS * sm_count();
        const Sched psc = cached_sched(a.d, zr, pa2::TX, pa2::TY, 3, 0.6, pctas);
        if (psc.nitems == 0) return LaunchInfo{0, {0, 0, 0}};
        const int pgrid = psc.nitems < pctas ? psc.nitems : pctas;
        if (peer) launch_pdl(pa2::pass_a_pipe_kernel<true>, pgrid, pa2::NTHREADS, pa2::SMEM_BYTES, st, m->in[0], m->in[1], m->in[2], a, it, psc);
        else launch_pdl(pa2::pass_a_pipe_kernel<false>, pgrid, pa2::NTHREADS, pa2::SMEM_BYTES, st, m->in[0], m->in[1], m->in[2], a, it, psc);
        return launch_info(psc, pgrid);
    }
    const int ctas = PA_CTAS * sm_count();
    const Sched sc = cached_sched(a.d, zr, pa::TX, pa::TY, 2, 0.5, ctas);
    if (sc.nitems == 0) return LaunchInfo{0, {0, 0, 0}};
    const int grid = sc.nitems < ctas ? sc.nitems : ctas;
    if (a.pn_tex && peer) launch_pdl(pa::pass_a_tma_kernel<true, true>, grid, pa::NTHREADS, pa::SMEM_BYTES, st, m->in[0], m->in[1], m->in[2], a, it, sc);
    else if (a.pn_tex) launch_pdl(pa::pass_a_tma_kernel<true, false>, grid, pa::NTHREADS, pa::SMEM_BYTES, st, m->in[0], m->in[1], m->in[2], a, it, sc);
    else if (peer) launch_pdl(pa::pass_a_tma_kernel<false, true>, grid, pa::NTHREADS, pa::SMEM_BYTES, st, m->in[0], m->in[1], m->in[2], a, it, sc);
    else launch_pdl(pa::pass_a_tma_kernel<false, false>, grid, pa::NTHREADS, pa::SMEM_BYTES, st, m->in[0], m->in[1], m->in[2], a, it, sc);
    return launch_info(sc, grid);
}

}  // namespace sb

// Host-only view of plan_peer_ranges (tests/test_schedule_cpu.py): ranges[9] = {lo, hi, face} x 3
extern "C" int sobfu_b200_debug_peer_ranges(int pass, int X, int Y, int Zlocal, int lo, int hi, int has_lo, int has_hi, int sms, int *ranges, int *n_ranges) {
    using namespace sb;
    if (!ranges || !n_ranges || (pass != 0 && pass != 1) || hi - lo < 8 || sms < 1) return -1;
    const ZRanges zr = plan_peer_ranges(Dims{X, Y, Zlocal}, pass, lo, hi, has_lo != 0, has_hi != 0, sms);
    *n_ranges = zr.n;
    for (int r = 0; r < MAX_ZRANGES; ++r) { ranges[3 * r] = zr.lo[r]; ranges[3 * r + 1] = zr.hi[r]; ranges[3 * r + 2] = zr.face[r]; }
    return 0;
}

// Host-only view of the work decomposition of a launch (no GPU needed): the items in issue order with the CTA that takes them
// under the static round-robin assignment.  pass 0 = A, 1 = B; `sms` = number of SMs to plan for (148 on B200).
// items: [cap][6] = {cta, x0, y0, zb, ze, face}.  Used by tests/test_schedule_cpu.py to check that every plane of every range is
// covered exactly once for the shapes the multi-GPU runs use.
extern "C" int sobfu_b200_debug_schedule(int pass, int X, int Y, int Zlocal, int nranges, const int *lo, const int *hi, const int *face, int sms,
                                         int *items, int cap, int *n_items, int *grid) {
    using namespace sb;
    if (!lo || !hi || !face || !n_items || !grid || nranges < 1 || nranges > MAX_ZRANGES || sms < 1 || pass < 0 || pass > 2) return -1;
    ZRanges zr;
    zr.n = nranges;
    for (int r = 0; r < MAX_ZRANGES; ++r) { zr.lo[r] = r < nranges ? lo[r] : 0; zr.hi[r] = r < nranges ? hi[r] : 0; zr.face[r] = r < nranges ? face[r] : 0; }
    // pass 0: pass A without pipelined gathers (variant 4), 1: pass B, 2: pass A (default kernel: one more pipeline-fill step per item)
    const int TX = pass == 1 ? pb::TX : pa::TX, TY = pass == 1 ? pb::TY : pa::TY;
    const int ctas = pass == 1 ? sms : (pass == 2 ? PA2_CTAS : PA_CTAS) * sms;
    const Sched sc = pass == 1 ? make_sched(Dims{X, Y, Zlocal}, zr, TX, TY, 6, 0.35, ctas)
                               : (pass == 2 ? make_sched(Dims{X, Y, Zlocal}, zr, TX, TY, 3, 0.6, ctas) : make_sched(Dims{X, Y, Zlocal}, zr, TX, TY, 2, 0.5, ctas));
    *n_items = sc.nitems;
    *grid = sc.nitems < ctas ? sc.nitems : ctas;
    for (int i = 0; i < sc.nitems && i < cap; ++i) {
        int x0, y0, zb, ze, f;
        locate_item(sc, i, TX, TY, x0, y0, zb, ze, f);
        int *o = items + 6 * i;
        o[0] = *grid ? i % *grid : 0; o[1] = x0; o[2] = y0; o[3] = zb; o[4] = ze; o[5] = f;
    }
    return 0;
}

