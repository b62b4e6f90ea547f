// common.cuh -- shared device helpers and error plumbing for libsobfu_b200 (sm_100a only).
//
// Numerics contract (DESIGN.md "Numerics"): results must equal the reference CUDA's bit for bit on the solver
// core.  The reference builds with --ftz=true --prec-div=false --prec-sqrt=false (CMakeLists.txt:42-44) and its
// float4 operators are un-fused __fmul_rn/__fadd_rn (include/sobfu/cuda/utils.hpp:245-275) while lerp is two
// FMAs (utils.hpp:33-36).  We therefore (a) compile with --ftz=true and (b) spell every float op on the solver
// core through the wrappers below so that nvcc can neither contract nor re-associate them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

#define SB_DEV __device__ __forceinline__

SB_DEV float mul(float a, float b) { return __fmul_rn(a, b); }
SB_DEV float add(float a, float b) { return __fadd_rn(a, b); }
SB_DEV float sub(float a, float b) { return __fsub_rn(a, b); }
// utils.hpp:33-36  lerp(v0, v1, t) = fma(t, v0, fma(-t, v1, v1))
SB_DEV float lerp(float v0, float v1, float t) { return __fmaf_rn(t, v0, __fmaf_rn(-t, v1, v1)); }

struct Dims {
    int X, Y, Z;
};

// Trilinear sampling geometry of utils.hpp:50-86 / :124-164: clamp, floor, upper index (not advanced when the
// clamped coordinate sits exactly on the first or last plane) and the fractional weights.
struct TriCoord {
    int gx, gy, gz, x1, y1, z1;
    float a, b, c;
};
SB_DEV TriCoord tri_coord(float px, float py, float pz, const Dims d) {
    TriCoord t;
    const float mx = (float)d.X - 1.f, my = (float)d.Y - 1.f, mz = (float)d.Z - 1.f;
    const float cx = fminf(fmaxf(0.f, px), mx);
    const float cy = fminf(fmaxf(0.f, py), my);
    const float cz = fminf(fmaxf(0.f, pz), mz);
    t.gx = __float2int_rd(cx);
    t.gy = __float2int_rd(cy);
    t.gz = __float2int_rd(cz);
    t.x1 = t.gx + ((cx == 0.f || cx == mx) ? 0 : 1);
    t.y1 = t.gy + ((cy == 0.f || cy == my) ? 0 : 1);
    t.z1 = t.gz + ((cz == 0.f || cz == mz) ? 0 : 1);
    t.a = __fsub_rn(cx, (float)t.gx);
    t.b = __fsub_rn(cy, (float)t.gy);
    t.c = __fsub_rn(cz, (float)t.gz);
    return t;
}

// nesting order of utils.hpp:78-82: innermost along z, then y, then x; the "+1" corner is v0 of each lerp
SB_DEV float tri_lerp(float v111, float v110, float v101, float v100, float v011, float v010, float v001, float v000,
                      const TriCoord &t) {
    return lerp(lerp(lerp(v111, v110, t.c), lerp(v101, v100, t.c), t.b),
                lerp(lerp(v011, v010, t.c), lerp(v001, v000, t.c), t.b), t.a);
}

// trilinear sample of a scalar volume stored with element stride `S` floats (S=1: plane, S=2: float2 .x)
template <int S>
SB_DEV float sample_scalar(const float *__restrict__ v, const TriCoord &t, const Dims d) {
    const size_t sy = (size_t)d.X, sz = (size_t)d.X * d.Y;
    const size_t o00 = sy * t.gy + sz * t.gz, o10 = sy * t.y1 + sz * t.gz, o01 = sy * t.gy + sz * t.z1,
                 o11 = sy * t.y1 + sz * t.z1;
    const float v000 = __ldg(v + S * (o00 + t.gx)), v100 = __ldg(v + S * (o00 + t.x1));
    const float v010 = __ldg(v + S * (o10 + t.gx)), v110 = __ldg(v + S * (o10 + t.x1));
    const float v001 = __ldg(v + S * (o01 + t.gx)), v101 = __ldg(v + S * (o01 + t.x1));
    const float v011 = __ldg(v + S * (o11 + t.gx)), v111 = __ldg(v + S * (o11 + t.x1));
    return tri_lerp(v111, v110, v101, v100, v011, v010, v001, v000, t);
}

// 64-bit max over a warp
SB_DEV unsigned long long warp_max_u64(unsigned long long k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other > k ? other : k;
    }
    return k;
}
SB_DEV double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Traversal order of the reference's arg-max reduction (reductor.cu:342-456 + reductor.cpp:81-94):
// ties keep the FIRST candidate met: per thread in (pass, half) order, then through the shared-memory tree, where
// slot tid keeps its own value against slot tid+s for s = bs/2 .. 1 -- i.e. among equal values the thread whose
// bit-reversed tid is smallest survives -- and finally the lowest block on the CPU.  rank_of() maps a voxel index to
// that order: (block, bitrev(tid), pass, half).
struct RankMap {
    unsigned bs;        // threads per block of the reference reduction (512 for N >= 1024), a power of two
    unsigned bits;      // log2(bs)
    unsigned grid;      // bs * 2 * blocks
    unsigned npass;     // ceil(N / grid)
};
SB_DEV unsigned rank_of(unsigned idx, const RankMap m) {
    const unsigned pass = idx / m.grid, r = idx - pass * m.grid;
    const unsigned b = r / (2 * m.bs), q = r - b * 2 * m.bs;
    const unsigned half = q >= m.bs ? 1u : 0u, t = q - half * m.bs;
    const unsigned tr = m.bits ? (__brev(t) >> (32u - m.bits)) : 0u;
    return ((b * m.bs + tr) * m.npass + pass) * 2u + half;
}

// Running arg-max candidate of a thread under the reference's rule (reductor.cu:357-368): NORMS are compared, i.e.
// __fsqrt_rd(sum of squares), and among equal norms the first element in traversal order (rank_of) wins -- two different sums
// of squares that round down to the same square root ARE a tie.  The square root is only taken for the few candidates that
// can matter: lo_bits is the smallest sum of squares whose norm equals the current best (r * r rounded up), so one integer
// compare rejects everything below it exactly.
struct MaxCand {
    unsigned norm_bits, lo_bits, idx;
};
SB_DEV void max_cand_update(MaxCand &m, float nsq, unsigned idx, const RankMap rm) {
    const unsigned b = __float_as_uint(nsq);
    if (b >= m.lo_bits && b != 0u) {
        const float r = __fsqrt_rd(nsq);
        const unsigned rb = __float_as_uint(r);
        if (rb > m.norm_bits) { m.norm_bits = rb; m.lo_bits = __float_as_uint(__fmul_ru(r, r)); m.idx = idx; }
        else if (rb == m.norm_bits && rank_of(idx, rm) < rank_of(m.idx, rm)) m.idx = idx;
    }
}
// 64-bit key of a candidate: norm in the high word, inverted traversal rank in the low word (max of keys = the reference's winner)
SB_DEV unsigned long long max_cand_key(const MaxCand &m, const RankMap rm) {
    return m.norm_bits ? (((unsigned long long)m.norm_bits << 32) | (unsigned long long)(0xffffffffu - rank_of(m.idx, rm))) : 0ull;
}

}  // namespace sb
