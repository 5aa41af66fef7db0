// ldpc_ms_h2.cuh -- min-sum family, two frames per thread in one __half2 (LDPC_GPU_PREC_F16X2).
//
// The scheduled kernel (ldpc_ms_fast.cuh) with every message word holding the same edge of TWO
// frames as a packed pair of binary16 values.  HMNMX2 / HADD2 / HSET2 are single instructions on
// sm_100a, so a frame pair costs about what one fp32 frame costs, in instructions, shared-memory
// wavefronts and registers alike.
//
// This is NOT the reference's arithmetic (its messages are doubles and grow without bound,
// src/decodeMinSum.cpp:177-178): messages are rounded to 11 significant bits and v2c is clamped to
// +-H2_CLAMP so that a variable's sum of up to 1 + dv terms stays far inside binary16's range.  It is a
// labelled throughput instantiation: decisions equal the fp64 path's on frames that converge,
// BER / FER agree within Monte-Carlo confidence intervals (tests/test_gpu_parity.py,
// tests/test_ber_statistics.py); the parity instantiation is fp64 and the headline number is fp32.
#pragma once
#include <cuda_fp16.h>
#include "ldpc_ms_fast.cuh"

namespace ldpc {

#define LDPC_H2_CLAMP 512.0f

LDPC_DEVINL uint32_t h2_bits(__half2 x) { return *reinterpret_cast<uint32_t *>(&x); }
LDPC_DEVINL __half2 h2_from(uint32_t x) { return *reinterpret_cast<__half2 *>(&x); }

static inline size_t ms_h2_smem_bytes(const CodeDev &c)
{
    return ((size_t)32 + 4 * ((size_t)c.dvN + c.N) + 8 * (size_t)((c.N + 31) / 32) + 15) & ~(size_t)15;
}

template <int DC, int DV, int NFIX, int NT_MAX, int MINB>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_h2_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = NFIX, nwords = (N + 31) >> 5, nblk = (N + 3) >> 2, NG = DC / 4;
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);                 // [2]
    uint32_t *msg = reinterpret_cast<uint32_t *>(smem_raw + 32);                  // [DV*N] packed pairs
    uint32_t *yq = msg + DV * N;                                                  // [N]
    uint32_t *dbits = yq + N;                                                     // [2][nwords]
    unsigned char *msgb = reinterpret_cast<unsigned char *>(msg);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, M = c.M;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    const __half2 INF2 = __floats2half2_rn(65504.0f, 65504.0f);
    const __half2 inv_alpha2 = __float2half2_rn(p.inv_alpha_f), delta2 = __float2half2_rn((float)p.delta);
    const __half2 zero2 = __float2half2_rn(0.0f), cpos = __float2half2_rn(LDPC_H2_CLAMP), cneg = __float2half2_rn(-LDPC_H2_CLAMP);
    const long long npairs = (io.n_frames + 1) / 2;
    CtaTotals tot; tot.clear();

    for (long long pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const long long fa = 2 * pr, fb = (2 * pr + 1 < io.n_frames) ? 2 * pr + 1 : 2 * pr;    // a dead lane replays frame fa, unreported
        const bool live_b = 2 * pr + 1 < io.n_frames;
        const uint8_t *cwa = codeword_row(io, c, fa), *cwb = codeword_row(io, c, fb);
        if (tid < 2) { fs[tid].uncoded = 0; fs[tid].errors = 0; fs[tid].flag = 0; }
        for (int w = tid; w < 2 * nwords; w += nt) dbits[w] = 0u;
        __syncthreads();
        // ---- channel front end for both frames -------------------------------------------------
        int unca = 0, uncb = 0;
        for (int b = tid; b < nblk; b += nt) {
            double ya[4], yb[4];
            raw_samples4(io, p, c, fa, cwa, b, ya);
            raw_samples4(io, p, c, fb, cwb, b, yb);
            const uint2 cc = __ldg(reinterpret_cast<const uint2 *>(c.col_of_var) + b);
            uint32_t niba = 0, nibb = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                float va, vb;
                if (fcond) { va = condition_ms_guarded(ya[q], p, qflags); vb = condition_ms_guarded(yb[q], p, qflags); }
                else {
                    double da = ya[q], db = yb[q];
                    if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) { da = quantize_ms(da, p); db = quantize_ms(db, p); }
                    if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) { da = fmin(fmax(da, -p.Ymax), p.Ymax); db = fmin(fmax(db, -p.Ymax), p.Ymax); }
                    va = (float)da; vb = (float)db;
                }
                const bool na = !(va > 0.0f), nb = !(vb > 0.0f);
                const uint32_t w = h2_bits(__floats2half2_rn(va, vb));
                const int col = (int)(((q < 2 ? cc.x : cc.y) >> (16 * (q & 1))) & 0xffffu);
                yq[col] = w;
#pragma unroll
                for (int s = 0; s < DV; s++) msg[s * N + col] = w;
                unca += (int)(na != ((cwa ? cwa[i] : 0) != 0)); uncb += (int)(nb != ((cwb ? cwb[i] : 0) != 0));
                niba |= (uint32_t)na << q; nibb |= (uint32_t)nb << q;
                if (io.out_soft && p.T == 0) {
                    if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)fa * N + i] = va; if (live_b) ((double *)io.out_soft)[(size_t)fb * N + i] = vb; }
                    else { ((float *)io.out_soft)[(size_t)fa * N + i] = va; if (live_b) ((float *)io.out_soft)[(size_t)fb * N + i] = vb; }
                }
            }
            if (niba) atomicOr(&dbits[(4 * b) >> 5], niba << ((4 * b) & 31));
            if (nibb) atomicOr(&dbits[nwords + ((4 * b) >> 5)], nibb << ((4 * b) & 31));
        }
        for (int o = 16; o; o >>= 1) { unca += __shfl_xor_sync(0xffffffffu, unca, o); uncb += __shfl_xor_sync(0xffffffffu, uncb, o); }
        if (lane == 0) { if (unca) atomicAdd(&fs[0].uncoded, unca); if (uncb) atomicAdd(&fs[1].uncoded, uncb); }
        __syncthreads();

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            if (last) for (int w = tid; w < 2 * nwords; w += nt) dbits[w] = 0u;
            // ---- check-node phase (row j, both frames) ---------------------------------------------
            for (int j = tid; j < M; j += nt) {
                uint32_t v[DC];
                __half2 m1 = INF2, m2 = INF2;
                uint32_t par = 0u;
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const uint4 w = __ldg(&c.sched[(size_t)g * M + j]);
                    const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int k = g * 4 + q;
                        v[k] = *reinterpret_cast<const uint32_t *>(msgb + off[q]);
                        const __half2 a = __habs2(h2_from(v[k]));
                        m2 = __hmin2(m2, __hmax2(m1, a));
                        m1 = __hmin2(m1, a);
                        par ^= v[k];
                    }
                }
                __half2 o1 = m1, o2 = m2;
                if (normalized) { o1 = __hmul2(o1, inv_alpha2); o2 = __hmul2(o2, inv_alpha2); }
                if (offset) { o1 = __hmax2(__hsub2(o1, delta2), zero2); o2 = __hmax2(__hsub2(o2, delta2), zero2); }
                const uint32_t psign = par & 0x80008000u;
                const uint32_t s1 = h2_bits(o1) ^ psign, s2 = h2_bits(o2) ^ psign;
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const uint4 w = __ldg(&c.sched[(size_t)g * M + j]);
                    const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int k = g * 4 + q;
                        const uint32_t eq = __heq2_mask(__habs2(h2_from(v[k])), m1);      // 0xffff per half where this edge is the row minimum
                        const uint32_t sel = (s1 & ~eq) | (s2 & eq);
                        *reinterpret_cast<uint32_t *>(msgb + off[q]) = sel ^ (v[k] & 0x80008000u);
                    }
                }
            }
            __syncthreads();
            // ---- variable-node phase (storage columns, both frames) ----------------------------------
            for (int col = tid; col < N; col += nt) {
                __half2 cm[DV];
                __half2 sum = h2_from(yq[col]);
#pragma unroll
                for (int s = 0; s < DV; s++) { cm[s] = h2_from(msg[s * N + col]); sum = __hadd2(sum, cm[s]); }
#pragma unroll
                for (int s = 0; s < DV; s++) msg[s * N + col] = h2_bits(__hmin2(__hmax2(__hsub2(sum, cm[s]), cneg), cpos));
                if (last) {
                    const int i = (int)__ldg(&c.var_of_col[col]);
                    const float sa = __low2float(sum), sb = __high2float(sum);
                    if (!(sa > 0.0f)) atomicOr(&dbits[i >> 5], 1u << (i & 31));
                    if (!(sb > 0.0f)) atomicOr(&dbits[nwords + (i >> 5)], 1u << (i & 31));
                    if (io.out_soft) {
                        if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)fa * N + i] = sa; if (live_b) ((double *)io.out_soft)[(size_t)fb * N + i] = sb; }
                        else { ((float *)io.out_soft)[(size_t)fa * N + i] = sa; if (live_b) ((float *)io.out_soft)[(size_t)fb * N + i] = sb; }
                    }
                }
            }
            __syncthreads();
        }
        finish_frame(c, p, io, fa, cwa, dbits, &fs[0], p.T, -1, 0, 0, 1, -1, tot);
        if (live_b) finish_frame(c, p, io, fb, cwb, dbits + nwords, &fs[1], p.T, -1, 0, 0, 1, -1, tot);
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
