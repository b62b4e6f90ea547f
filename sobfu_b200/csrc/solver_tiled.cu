// solver_tiled.cu -- pass B of the solver iteration as a persistent, TMA-fed, z-marching kernel:
//     nabla_U_S = S *x nabla_U + S *y nabla_U + S *z nabla_U           (solver.cu:237-446, 7 taps, clamp to edge)
//     psi      -= alpha * nabla_U_S ; partial arg-max of |alpha * nabla_U_S|   (solver.cu:53-69, reductor.cu:342-456)
//     (phi_n o psi).x re-warped with the new psi                              (vector_fields.cu:81-100, solver.cu:168)
// -- five of the reference's launches (3 convolutions, update, apply) plus its reduction, in one pass over HBM.
//
// Mapping to the hardware
//   * one persistent CTA per SM walks (x-tile, y-tile, z-chunk) work items; inside an item it marches along z
//   * nabla_U planes (3 components, tile + halo of 3 in x and y) are brought into a ring of shared-memory stages by
//     TMA (cp.async.bulk.tensor.3d, one elected thread, mbarrier complete_tx); the halo needs no special cases
//     because pass A stores nabla_U with a replicated border (clamp to edge == plain loads)
//   * each thread owns 4 consecutive x: x taps = two extra LDS.128, y taps = six LDS.128, z taps = a 7-deep register
//     window that is refilled with one LDS.128 per component and step
//   * psi is read and written once (LDG/STG.128), the warp gathers phi_n.x from a 4 B/voxel plane
//   * the max update norm is reduced warp-shuffle -> shared -> one 64-bit atomicMax per CTA
// Algorithmic traffic: R nabla_U 12 + R psi 12 + W psi 12 + R phi_n 4 + W w 4 = 44 B/voxel (reference layouts: 64).
// Arithmetic and summation order are the reference's (taps S[3-j], j=-3..3 from 0; (x + y) + z); results are
// bit-identical to pass_b_generic_kernel.
#include <cuda.h>

#include <cstdio>

#include "solver_kernels.cuh"

namespace sb {

struct TmaMaps {
    CUtensorMap m[3];
};

namespace {

// ---- tile configuration -------------------------------------------------------------------------------------
constexpr int LX = 16;                    // lanes along x per row -> 64 voxels
constexpr int RW = 32 / LX;               // rows per warp
constexpr int NW = 12;                    // warps per CTA
constexpr int TX = 4 * LX, TY = NW * RW;  // 64 x 24 outputs per plane
constexpr int SX = TX + 8, SY = TY + 6;   // staged box: 4|64|4 floats wide (16 B aligned own quads), 3|24|3 rows
constexpr int NSTAGE = 6;                 // planes z .. z+3 live, two in flight
constexpr int COMP_BYTES = ((SX * SY * 4 + 127) / 128) * 128;
constexpr int STAGE_BYTES = 3 * COMP_BYTES;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 128;
constexpr unsigned TX_BYTES = 3u * SX * SY * 4u;

SB_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
SB_DEV void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SB_DEV void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SB_DEV void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
SB_DEV void tma_load_3d(unsigned dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            dst),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// explicit shared-window loads (32-bit shared addresses; element offsets in floats)
SB_DEV float4 lds4(unsigned saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
SB_DEV float c4(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

struct Sched {
    int tiles_x, tiles_y, nz, zchunk, nitems;
};

__global__ void __launch_bounds__(NW * 32, 1)
    pass_b_tma_kernel(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapy,
                      const __grid_constant__ CUtensorMap mapz, LoopArgs a, int it, Sched sc) {
    if (loop_finished(a, it)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned smem = (smem_u32(smem_raw) + 127u) & ~127u;   // TMA destinations need 128 B alignment
    __shared__ unsigned long long full[NSTAGE];
    __shared__ unsigned long long skey[NW];

    const Dims d = a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lx = lane % LX, ly = lane / LX;
    const int ty = warp * RW + ly;            // row inside the tile
    const unsigned own_off = (unsigned)(((ty + 3) * SX + 4 * lx + 4) * 4);   // byte offset of the thread's own quad in a component
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const size_t sy = (size_t)d.X, sz = (size_t)d.X * d.Y;
    float *__restrict__ P[3] = {a.px, a.py, a.pz};
    unsigned q_issue = 0, q_wait = 0;         // running plane counters (slot = q % NSTAGE, parity = (q / NSTAGE) & 1)
    unsigned long long best = 0ull;

    for (int item = blockIdx.x; item < sc.nitems; item += gridDim.x) {
        const int tz = item / (sc.tiles_x * sc.tiles_y), rem = item - tz * sc.tiles_x * sc.tiles_y;
        const int tyi = rem / sc.tiles_x, txi = rem - tyi * sc.tiles_x;
        const int x0t = txi * TX, y0t = tyi * TY;
        const int zb = tz * sc.zchunk, ze = min(zb + sc.zchunk, d.Z);
        const int x0 = x0t + 4 * lx, y = y0t + ty;
        const bool active = x0 < d.X && y < d.Y;
        const size_t row = (size_t)min(x0, d.X - 4) + sy * min(y, d.Y - 1);
        // planes zb-3 .. ze+2 stream through the ring; padded plane index = z + 3, box origin = (x0t, y0t) in padded
        // coordinates, i.e. interior (x0t - 4, y0t - 3)
        const int p_first = zb - 3, p_last = ze + 2;
        auto issue = [&](int p) {
            const unsigned slot = q_issue % NSTAGE;
            const unsigned dst = smem + slot * STAGE_BYTES;
            mbar_expect_tx(&full[slot], TX_BYTES);
            tma_load_3d(dst, &mapx, &full[slot], x0t, y0t, p + 3);
            tma_load_3d(dst + COMP_BYTES, &mapy, &full[slot], x0t, y0t, p + 3);
            tma_load_3d(dst + 2 * COMP_BYTES, &mapz, &full[slot], x0t, y0t, p + 3);
        };
        __syncthreads();                      // every thread is done with the previous item's stages
        if (tid == 0) {
            issue(p_first); ++q_issue;
            if (p_first + 1 <= p_last) { issue(p_first + 1); ++q_issue; }
        } else {
            q_issue += (p_first + 1 <= p_last) ? 2 : 1;
        }

        float4 win[7][3];                     // nabla_U of this thread's 4 voxels at planes c-3 .. c+3
        for (int p = p_first; p <= p_last; ++p) {
            __syncthreads();                  // all reads of the stage that plane p+2 will overwrite are finished
            if (p + 2 <= p_last) {
                if (tid == 0) issue(p + 2);
                ++q_issue;
            }
            const unsigned slot = q_wait % NSTAGE, par = (q_wait / NSTAGE) & 1u;
            ++q_wait;
            mbar_wait(&full[slot], par);
            const unsigned sp = smem + slot * STAGE_BYTES + own_off;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) win[k][c] = win[k + 1][c];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) win[6][c] = lds4(sp + c * COMP_BYTES);
            const int zc = p - 3;             // centre plane whose window is now complete
            if (zc < zb) continue;

            // stage that holds the centre plane: it was the (q_wait-1-3)-th plane
            const unsigned cslot = (q_wait - 4u) % NSTAGE;
            const unsigned sc0 = smem + cslot * STAGE_BYTES + own_off;
            const size_t o = row + sz * zc;
            float4 psi4[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) psi4[c] = *reinterpret_cast<const float4 *>(P[c] + o);

            float np[3][4], nsq[4];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned s0 = sc0 + c * COMP_BYTES;
                const float4 L = lds4(s0 - 16), R = lds4(s0 + 16), C = win[3][c];
                const float v[12] = {L.x, L.y, L.z, L.w, C.x, C.y, C.z, C.w, R.x, R.y, R.z, R.w};
                float fx[4] = {0.f, 0.f, 0.f, 0.f}, fy[4] = {0.f, 0.f, 0.f, 0.f}, fz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = -3; k <= 3; ++k) {
                    const float s = a.S[3 - k];
                    const float4 yk = (k == 0) ? C : lds4(s0 + k * (SX * 4));
                    const float4 zk = win[3 + k][c];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        fx[j] = add(fx[j], mul(s, v[4 + j + k]));
                        fy[j] = add(fy[j], mul(s, c4(yk, j)));
                        fz[j] = add(fz[j], mul(s, c4(zk, j)));
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
                float wv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const TriCoord t = tri_coord(np[0][j], np[1][j], np[2][j], d);
                    wv[j] = sample_scalar<1>(a.pn, t, d);
                    const unsigned long long key = ((unsigned long long)__float_as_uint(nsq[j]) << 32) |
                                                   (unsigned long long)(0xffffffffu - rank_of((unsigned)(o + j), a.rm));
                    best = key > best ? key : best;
                }
                *reinterpret_cast<float4 *>(a.w + o) = make_float4(wv[0], wv[1], wv[2], wv[3]);
            }
        }
    }
    best = warp_max_u64(best);
    if (lane == 0) skey[warp] = best;
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = 0ull;
#pragma unroll
        for (int k = 0; k < NW; ++k) m = skey[k] > m ? skey[k] : m;
        atomicMax(&a.maxkey[it], m);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

Sched make_sched(const Dims d, int ctas) {
    Sched s;
    s.tiles_x = (d.X + TX - 1) / TX;
    s.tiles_y = (d.Y + TY - 1) / TY;
    const int xy = s.tiles_x * s.tiles_y;
    // pick the number of z chunks so that the item count fills whole rounds of `ctas` CTAs (chunks >= 16 planes)
    int best_nz = 1;
    double best_score = -1.0;
    for (int nz = 1; nz <= d.Z / 8 && nz <= 64; ++nz) {
        const int chunk = (d.Z + nz - 1) / nz;
        if (chunk < 16 && nz > 1) break;
        const int n = xy * ((d.Z + chunk - 1) / chunk);
        const int rounds = (n + ctas - 1) / ctas;
        const double balance = (double)n / ((double)rounds * ctas);
        const double overlap = (double)chunk / (chunk + 6.0 * 0.35);   // the 6 window-fill steps are ~1/3 of a full step
        const double score = balance * overlap;
        if (score > best_score) { best_score = score; best_nz = nz; }
    }
    s.zchunk = (d.Z + best_nz - 1) / best_nz;
    s.nz = (d.Z + s.zchunk - 1) / s.zchunk;
    s.nitems = xy * s.nz;
    return s;
}

}  // namespace

TmaMaps *tma_maps_create(const LoopArgs &a) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return nullptr;
    TmaMaps *m = new TmaMaps();
    float *base[3] = {a.gx, a.gy, a.gz};
    for (int c = 0; c < 3; ++c) {
        const cuuint64_t dims[3] = {(cuuint64_t)a.gl.PX, (cuuint64_t)a.gl.PY, (cuuint64_t)a.gl.PZ};
        const cuuint64_t strides[2] = {(cuuint64_t)a.gl.PX * 4, (cuuint64_t)a.gl.plane * 4};
        const cuuint32_t box[3] = {SX, SY, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&m->m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base[c], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            fprintf(stderr, "sobfu_b200: cuTensorMapEncodeTiled failed (%d); falling back to the generic kernels\n", (int)r);
            delete m;
            return nullptr;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(pass_b_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
            delete m;
            return nullptr;
        }
        attr_set = true;
    }
    return m;
}

void tma_maps_destroy(TmaMaps *m) { delete m; }

void launch_pass_b_tma(const LoopArgs &a, const TmaMaps *m, int it, cudaStream_t st) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const Sched sc = make_sched(a.d, sms);
    const int grid = sc.nitems < sms ? sc.nitems : sms;
    pass_b_tma_kernel<<<grid, NW * 32, SMEM_BYTES, st>>>(m->m[0], m->m[1], m->m[2], a, it, sc);
}

}  // namespace sb
