// solver_kernels.cuh -- device-side state and kernel launch prototypes of the solver loop.
//
// Workspace layout in HBM (DESIGN.md "Data layout"): the loop never touches the reference's float4 / float2
// AoS volumes.  Per voxel it keeps
//     psi      3 float planes  (px, py, pz)                     12 B   read by A, read+written by B
//     w        1 float plane   (phi_n o psi).x                   4 B   read by A, written by B
//     pg, pn   2 float planes  phi_global.x, phi_n.x            8 B   pg read by A, pn gathered by B
//     g        3 float planes  nabla_U, padded by a replicated halo of 3 (4 along x for 16 B alignment)
// = 36 B/voxel of scratch against the reference's 240 B.
#pragma once
#include "common.cuh"

namespace sb {

// halo depths (planes) of the slab-local arrays: one iteration depends on psi within 4 planes (1 for the stencils of pass A
// + 3 for the filter of pass B), so with 4 halo planes of psi a rank can compute nabla_U on its own 3 halo planes itself
// and the iteration needs ONE exchange (psi) instead of two (nabla_U and psi)
constexpr int MAX_TAPS = 11;           // longest filter the reference tabulates (solver.cpp:160-251: s = 3, 7, 9, 11)
constexpr int PSI_HALO = 4;
constexpr int PG_HALO = 3;

// padded layout of the nabla_U planes
struct GLayout {
    int PX, PY, PZ;        // padded extents: X+8, Y+6, Z+6
    size_t plane;          // PX*PY
    size_t total;          // PX*PY*PZ floats per component
    __host__ __device__ size_t at(int x, int y, int z) const {
        return (size_t)(x + 4) + (size_t)PX * ((size_t)(y + 3) + (size_t)PY * (size_t)(z + 3));
    }
};

struct LoopState {            // lives in device memory, one per solver
    int converged;            // sticky: set once an iteration's max update norm <= threshold
    int iters;                // number of iterations executed when converged was set
    int pad[2];
};

struct LoopArgs {
    // planes
    float *px, *py, *pz, *w;
    const float *pg, *pn;
    float *gx, *gy, *gz;
    // z-slab decomposition: `d` is the LOCAL extent (X, Y, owned planes), `dg` the global volume and z0
    // the global z of local plane 0.  px/py/pz/w point at local plane 0 and carry PSI_HALO halo planes on either side,
    // pg PG_HALO; pn is the whole volume (the warp gathers anywhere); nabla_U has 3 halo planes (GLayout).
    // Boundary rules (axis term dropped / clamp to edge) apply on GLOBAL faces only.  Single GPU: dg == d, z0 == 0.
    Dims d, dg;
    int z0;
    GLayout gl;
    // parameters: the filter has 2 * radius + 1 taps; the tiled kernels are built for radius 3 (the reference's KERNEL_RADIUS,
    // solver.cu:211), other radii run the generic kernels
    float S[MAX_TAPS];
    int radius;
    float alpha, w_reg, thr;
    // convergence / logging
    LoopState *state;
    unsigned long long *maxkey;   // [max_iter]
    double *e_data, *e_reg;       // [max_iter] (sums, not yet halved)
    RankMap rm;
    int check;                    // 0: time_loop mode (no convergence logic)
    int a_uses_max;               // 0: pass A may run before the global maximum of the previous iteration is known
                                  //    (overlapped slab mode): it then only honours the sticky flag, which pass B raises
    // phi_n.x as a 2-D texture atlas (slice z at tile (z & amask, z >> ashift) of X x Y texels) for gather4 fetches of
    // the trilinear footprint; 0 when unavailable (then phi_n.x is gathered from the pn plane with plain loads)
    cudaTextureObject_t pn_tex;
    cudaSurfaceObject_t pn_surf;
    int ashift, amask;
    // ---- peer mode (slab ranks on one NVLink / NVSwitch domain; the neighbours' buffers are mapped through CUDA IPC) ----
    // Pass B stores the psi planes a neighbour needs straight into that neighbour's halo planes (the halo exchange is part
    // of the kernel that produces the data) and counts its finished CTAs in the neighbour's control block; pass A of the
    // next iteration waits on its own counters before it touches the halo planes.  The per-iteration maxima are published
    // by every rank into every rank's `allmax` table: the convergence test of pass B reads all of them (bit 63 = valid).
    float *peer_lo[3], *peer_hi[3];          // neighbour's psi components, offset so that [row + XY * zc] (my local zc) is its halo voxel; null: none
    unsigned long long *cnt_lo, *cnt_hi;     // the neighbours' counters this rank bumps
    const unsigned long long *my_cnt;        // own counters: [0] fed by the lower neighbour, [1] by the upper one
    unsigned long long expect_lo, expect_hi; // pass A (wait_halo): proceed once my_cnt[k] >= expect
    // flow control the other way: a rank may overwrite a neighbour's halo planes (pass B of iteration i) only after that
    // neighbour's pass A of iteration i has read them -- pass A on the faces counts its CTAs in the neighbours' `consumed`
    unsigned long long *ack_lo, *ack_hi;     // the neighbours' `consumed` counters this rank bumps at the end of pass A (faces)
    const unsigned long long *my_ack;        // own: [0] bumped by the lower neighbour, [1] by the upper one
    unsigned long long expect_ack;           // pass B (push): proceed once my_ack[k] >= expect_ack
    const unsigned long long *allmax;        // own table [iteration][rank]
    unsigned long long *peer_error;          // own control block: set when a wait timed out
    int peer_n;                              // > 0: peer mode with this many ranks
    // Peer mode runs an iteration as TWO launches: work items carry a face tag (ZRanges::face).  Pass A handles the middle of
    // the slab first and the items that read halo planes last (they wait for the neighbours' counters and acknowledge when
    // they are done); pass B handles the face items first (they wait for the acknowledgements, store into the neighbours and
    // bump their counters as soon as the item is finished -- the halo travels while the middle of the slab is computed).
    int push;                                // pass B: face items store to the neighbours and count there (per ITEM)
    int wait_halo;                           // pass A: face items wait on my_cnt / acknowledge in ack_lo / ack_hi (per ITEM)
    // publication of this rank's maximum of iteration `it` by the last CTA of pass B (ticket), into every rank's table
    unsigned int *tickets;                   // [max_iter], zeroed per solve; null: no publication (time_loop mode / NCCL mode)
    unsigned long long *pub[16];             // every rank's allmax table (own included)
    int my_rank;
    // measurement aid (SOBFU_B200_TRACE=1): 8 words per launch, see trace_* below; null otherwise
    unsigned long long *trace;
};

// control block of one rank in peer mode, exported through CUDA IPC; `allmax[max_iter * nranks]` follows the header
struct PeerCtl {
    unsigned long long halo_cnt[2];
    unsigned long long consumed[2];
    unsigned long long error;
    unsigned long long pad[3];
};
constexpr unsigned long long PEER_VALID = 1ull << 63;
constexpr int MAX_PEERS = 16;

SB_DEV unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SB_DEV unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spins until *p >= want (counters) -- bounded: after 4 s the error word of the control block is set and the caller goes on
// (the host reports SOBFU_B200_ECOMM), so a lost peer cannot hang the GPU
// (not inlined: the spin loop stays out of the kernels' steady-state loops, whose code size matters)
static __device__ __noinline__ unsigned long long peer_wait_ge(const unsigned long long *p, unsigned long long want, unsigned long long *err) {
    unsigned long long v = ld_acquire_sys(p);
    if (v >= want) return v;
    if (err && *reinterpret_cast<volatile unsigned long long *>(err)) return v;   // an earlier wait already gave up: drain quickly
    const unsigned long long t0 = global_timer_ns();
    while ((v = ld_acquire_sys(p)) < want) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > 4000000000ull) { if (err) *err = 1ull; break; }
    }
    return v;
}

// same wait by the lanes r < n of ONE warp, lane r on p[r]: the n polls are in flight together (n serial acquire loads of a
// table the peers write cost ~1 us each, at the start of every CTA of pass B); returns the maximum of the values, bit 63 cleared
static __device__ __noinline__ unsigned long long peer_wait_all_max(const unsigned long long *p, int n, unsigned long long *err) {
    const int lane = threadIdx.x & 31;
    unsigned long long v = PEER_VALID;
    if (lane < n) {
        v = ld_acquire_sys(p + lane);
        if (v < PEER_VALID && !(err && *reinterpret_cast<volatile unsigned long long *>(err))) {
            const unsigned long long t0 = global_timer_ns();
            while ((v = ld_acquire_sys(p + lane)) < PEER_VALID) {
                __nanosleep(32);
                if (global_timer_ns() - t0 > 4000000000ull) { if (err) *err = 1ull; break; }
            }
        }
    }
    v &= ~PEER_VALID;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

// trace slot of one launch (device-side timeline, %globaltimer): [0] max(~start) [1] max(end) [2] sum / [3] max ns the CTAs waited
// for the maxima table, [4] sum / [5] max ns they waited for a neighbour's counter, [6] CTAs, [7] max(~time the first wait ended)
SB_DEV void trace_begin(unsigned long long *tr) {
    if (tr) { atomicMax(tr + 0, ~global_timer_ns()); atomicAdd(tr + 6, 1ull); }
}
SB_DEV void trace_end(unsigned long long *tr) {
    if (tr) atomicMax(tr + 1, global_timer_ns());
}
SB_DEV void trace_wait(unsigned long long *tr, int k, unsigned long long t0) {
    if (tr) { const unsigned long long dt = global_timer_ns() - t0; atomicAdd(tr + k, dt); atomicMax(tr + k + 1, dt); }
}

// z ranges (local planes) a launch works on: the whole slab, or the planes next to / away from the slab faces
// face: 0 = no neighbour involved, 1 = next to the lower neighbour, 2 = next to the upper one (peer mode only).  Work items
// are issued in the order of the ranges.
constexpr int MAX_ZRANGES = 3;
struct ZRanges {
    int n;
    int lo[MAX_ZRANGES], hi[MAX_ZRANGES];
    int face[MAX_ZRANGES];
};
// what a launch did: grid size and the number of work items per face tag (the units of the peer counters)
struct LaunchInfo {
    int grid;
    int face_items[3];
};

// decision shared by every block of iteration `it` (0-based): has the loop already ended?
SB_DEV bool loop_finished(const LoopArgs &a, int it) {
    if (!a.check) return false;
    if (a.state->converged) return true;
    if (it == 0) return false;
    return __uint_as_float((unsigned)(a.maxkey[it - 1] >> 32)) <= a.thr;      // the key carries the norm (common.cuh MaxCand)
}
// same decision in peer mode: the maximum of iteration it-1 over all ranks, from the table the ranks publish into (ONE
// whole warp per block calls this -- lane r polls rank r's entry -- and broadcasts the result)
SB_DEV bool loop_finished_peer(const LoopArgs &a, int it) {
    if (!a.check) return false;
    if (a.state->converged) return true;
    if (it == 0) return false;
    const unsigned long long m = peer_wait_all_max(a.allmax + (size_t)(it - 1) * a.peer_n, a.peer_n, a.peer_error);
    return __uint_as_float((unsigned)(m >> 32)) <= a.thr;
}

void launch_unpack(const float4 *psi, const float2 *phi_global, const float2 *phi_n, const LoopArgs &a, cudaStream_t st);
// planes [z0, z0 + nz) of the volume that an array covers, and where a gather outside of it is reported (see field_ops.cu)
struct ZWindow {
    int z0, nz;
    int *overflow;
};
void launch_estimate_inverse_slab(const float4 *psi_win, float4 *psi_inv_local, Dims dg, int z0, int nzl, int iters, ZWindow w, cudaStream_t st);
void launch_apply_slab(const float2 *phi_win, float2 *out_local, const float4 *psi_local, Dims dg, int nzl, ZWindow w, cudaStream_t st);
void launch_initial_warp(const LoopArgs &a, cudaStream_t st);
void launch_pass_a_generic(const LoopArgs &a, int it, int log, cudaStream_t st);   // log: 0 none, 1 energies in double (atomics), 2 energies elsewhere
unsigned energy_tree_blocks(size_t n);
void launch_energy_trees(const LoopArgs &a, int it, float *partial, cudaStream_t st);
void launch_pass_b_generic(const LoopArgs &a, int it, cudaStream_t st);
void launch_pack(float4 *psi, float2 *phi_n_psi, const float2 *phi_n, const LoopArgs &a, bool warp, cudaStream_t st);   // warp: a.w is stale, sample here

// tiled kernels (pass_a_tiled.cu / pass_b_tma.cu); return false when the shape is not supported
bool tiled_supported(const Dims d);
struct TmaMaps;   // opaque: CUtensorMaps of the nabla_U components (pass B) and of the psi / w planes (pass A)
TmaMaps *tma_maps_create(const LoopArgs &a);
void tma_maps_destroy(TmaMaps *m);
LaunchInfo launch_pass_b_tma(const LoopArgs &a, const TmaMaps *m, int it, const ZRanges &zr, cudaStream_t st);
void set_pass_a_variant(int v);   // 0: default kernel (software-pipelined gathers), 4: the kernel without them; per host thread
LaunchInfo launch_pass_a_tma(const LoopArgs &a, const TmaMaps *m, int it, int log, const ZRanges &zr, cudaStream_t st);   // grid 0: generic path
// peer mode: the planes [lo, hi) of a launch (pass 0 = A, 1 = B) as | lower face chunk | upper face chunk | middle |
ZRanges plan_peer_ranges(const Dims d, int pass, int lo, int hi, bool has_lo, bool has_hi, int sms = 0);

// free-standing field kernels (field_ops.cu)
void launch_init_identity(float4 *psi, Dims d, cudaStream_t st);
void launch_apply(const float2 *phi, float2 *out, const float4 *psi, Dims d, cudaStream_t st);
void launch_estimate_inverse(const float4 *psi, float4 *psi_inv, Dims d, int iters, bool from_identity, cudaStream_t st);

}  // namespace sb
