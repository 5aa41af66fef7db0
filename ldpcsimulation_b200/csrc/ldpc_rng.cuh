// ldpc_rng.cuh -- counter-based channel / perturbation noise for the decode kernels.
//
// Replaces the reference's libc random() + Box-Muller macro (inc/rand.h:6-20, seeded with
// time(0) at src/decodeMinSum.cpp:187), whose stream->sample mapping is compiler dependent.
// Here every sample is a pure function of (seed, frame id, sample index, row, stream), so a
// frame can be regenerated anywhere -- on another GPU, or on the CPU by the test oracle.
//
//   Philox4x32-10, key = seed, counter = (block, row<<2 | stream, frame_lo, frame_hi)
//   one block -> four u32 -> two Box-Muller pairs -> four N(0,1) samples (fp32)
//
// Every floating-point step is an explicitly rounded single operation (__fmaf_rn & co.), so
// nvcc cannot contract or reorder it and a host implementation using fmaf() matches bit for bit.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace ldpc {

enum : uint32_t { STREAM_CHANNEL = 0, STREAM_DECODER = 1 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// (a, b) -> (n0, n1): angle from the top 24 bits of a, radius from all 32 bits of b.
__device__ __forceinline__ void box_muller_pair(uint32_t a, uint32_t b, float &n0, float &n1)
{
    const int a24 = (int)(a >> 8);
    const int k = (a24 + (1 << 21)) >> 22;                  // nearest quarter turn
    const int rem = a24 - (k << 22);                        // |rem| <= 2^21
    float phi = __fmul_rn(__int2float_rn(rem), 5.9604644775390625e-08f);
    phi = __fmul_rn(phi, 6.2831855f);
    const float z = __fmul_rn(phi, phi);
    float ps = __fmaf_rn(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = __fmaf_rn(ps, z, -1.6666654611e-1f);
    const float s = __fmaf_rn(__fmul_rn(ps, z), phi, phi);
    float pc = __fmaf_rn(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = __fmaf_rn(pc, z, 4.166664568298827e-2f);
    const float c = __fmaf_rn(__fmul_rn(pc, z), z, __fmaf_rn(-0.5f, z, 1.0f));
    // quadrant k: (cos, sin) = (c, s), (-s, c), (-c, -s), (s, -c); written without branches (a switch diverges
    // four ways inside a warp); negation is a sign-bit flip, exactly what unary minus does
    const bool swp = (k & 1) != 0;
    const uint32_t cneg = ((uint32_t)(k + 1) & 2u) << 30, sneg = ((uint32_t)k & 2u) << 30;
    const float cs = __uint_as_float(__float_as_uint(swp ? s : c) ^ cneg);
    const float sn = __uint_as_float(__float_as_uint(swp ? c : s) ^ sneg);
    const unsigned long long v = (unsigned long long)b + 1ull;      // [1, 2^32]
    int e = 63 - __clzll((long long)v);
    const uint32_t top = (uint32_t)((v << (63 - e)) >> 40);         // 24 bits, leading one set
    float m = __fmul_rn(__uint2float_rn(top), 1.1920928955078125e-07f);
    e -= 32;
    if (m > 1.41421356f) { m = __fmul_rn(m, 0.5f); e += 1; }
    const float x = __fadd_rn(m, -1.0f);
    const float zz = __fmul_rn(x, x);
    float p = 7.0376836292e-2f;
    p = __fmaf_rn(p, x, -1.1514610310e-1f); p = __fmaf_rn(p, x, 1.1676998740e-1f);
    p = __fmaf_rn(p, x, -1.2420140846e-1f); p = __fmaf_rn(p, x, 1.4249322787e-1f);
    p = __fmaf_rn(p, x, -1.6668057665e-1f); p = __fmaf_rn(p, x, 2.0000714765e-1f);
    p = __fmaf_rn(p, x, -2.4999993993e-1f); p = __fmaf_rn(p, x, 3.3333331174e-1f);
    const float fe = __int2float_rn(e);
    float yv = __fmul_rn(__fmul_rn(x, zz), p);
    yv = __fmaf_rn(-2.12194440e-4f, fe, yv);
    yv = __fmaf_rn(-0.5f, zz, yv);
    float ln = __fadd_rn(x, yv);
    ln = __fmaf_rn(0.693359375f, fe, ln);
    float t = __fmul_rn(-2.0f, ln);
    if (t < 0.0f) t = 0.0f;
    const float rad = __fsqrt_rn(t);
    n0 = __fmul_rn(rad, cs); n1 = __fmul_rn(rad, sn);
}

// Fast variant (LDPC_GPU_CHANNEL_FAST): the same Philox words through the SFU approximations (MUFU.LG2 / SQRT / SIN / COS,
// relative error about 2^-21).  Deterministic on a given GPU architecture, seed-addressable like the reference variant, but
// not reproducible on a CPU: ldpc_gpu_channel_dump is then the only source of the samples.  Same resolution: radius from all
// 32 bits of b (|n| <= 6.66), angle from the top 24 bits of a.  About 7 instead of 35 instructions per sample.
__device__ __forceinline__ void box_muller_pair_fast(uint32_t a, uint32_t b, float &n0, float &n1)
{
    const float u = __fmaf_rn(__uint2float_rn(b), 2.3283064365386963e-10f, 2.3283064365386963e-10f);   // (b + 1) / 2^32 in (0, 1]
    float l2, rad, sn, cs;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(__fmul_rn(l2, -1.3862943611198906f)));           // sqrt(-2 ln u)
    const float phi = __fmul_rn(__uint2float_rn(a >> 8), 3.7450703e-07f);                               // 2 pi a24 / 2^24
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(phi));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(phi));
    n0 = __fmul_rn(rad, cs); n1 = __fmul_rn(rad, sn);
}

__device__ __forceinline__ void normal4_fast(uint64_t seed, uint64_t frame, uint32_t block, uint32_t row, uint32_t stream, float n[4])
{
    uint32_t r[4];
    philox4x32_10(block, (row << 2) | (stream & 3u), (uint32_t)frame, (uint32_t)(frame >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    box_muller_pair_fast(r[0], r[1], n[0], n[1]);
    box_muller_pair_fast(r[2], r[3], n[2], n[3]);
}

__device__ __forceinline__ void normal4(uint64_t seed, uint64_t frame, uint32_t block, uint32_t row, uint32_t stream, float n[4])
{
    uint32_t r[4];
    philox4x32_10(block, (row << 2) | (stream & 3u), (uint32_t)frame, (uint32_t)(frame >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    box_muller_pair(r[0], r[1], n[0], n[1]);
    box_muller_pair(r[2], r[3], n[2], n[3]);
}

// ranu() shape (inc/rand.h:12-13) on 31 bits of each Philox word
__device__ __forceinline__ void uniform4(uint64_t seed, uint64_t frame, uint32_t block, uint32_t row, uint32_t stream, double u[4])
{
    uint32_t r[4];
    philox4x32_10(block, (row << 2) | (stream & 3u), (uint32_t)frame, (uint32_t)(frame >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
    for (int q = 0; q < 4; q++) u[q] = __ddiv_rn(__dadd_rn(1.0, (double)(r[q] >> 1)), 2147483649.0);
}

} // namespace ldpc
