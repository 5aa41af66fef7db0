// ldpc_ms_quad.cuh -- fp32 min-sum family for SMALL codes (M <= 512 rows, row weight <= 8: the (3,6) PEG code of BASELINE
// configs[0]), several frames per thread.
//
// A small code leaves a thread-per-row kernel with six edges of work between two CTA barriers, and every edge pays its own
// index load, address computation and 4-byte shared-memory access (ms_fast_kernel: 3.4 Gbit/s at T = 50 on PEGReg504x1008).
// Here a CTA owns a TILE of FI = 4 NV frames and every shared-memory word is a float4 holding the same value of four frames:
//
//   S4[i]   a-posteriori sum of variable i      msg4[s*N + i]   c2v on slot s of variable i      y4[i]   channel value
//
//   * check thread = row j for all FI frames: the row's c2v stay in registers between iterations (ms_rc_kernel's structure),
//     v2c = S - c2v is rebuilt from the published sums -- the reference's own expression (src/decodeMinSum.cpp:452-476), same
//     operands, same order, so the results are bit-identical to ms_fast_kernel<float> -- and the row's edge positions are
//     loaded once per kernel, not once per iteration;
//   * one 16-byte access moves four frames: a quarter of the LSU instructions and address arithmetic per edge-frame;
//   * variable thread = variable i for all FI frames: S = y + sum_s c2v in nlist order, one float4 store.
//
// Arithmetic contract: that of the fp32 instantiation (ldpc_ms_fast.cuh); tests/test_gpu_parity.py compares the two kernels
// bit for bit (LDPC_GPU_NO_QUAD selects ms_fast_kernel).
#pragma once
#include "ldpc_ms_fast.cuh"

namespace ldpc {

template <int NV>
static inline size_t ms_quad_smem_bytes(const CodeDev &c)
{
    const size_t FI = 4 * NV;
    return (16 * FI + 4 * FI * ((size_t)c.dvN + 2 * (size_t)c.N) + 4 * FI * (size_t)((c.N + 31) / 32) + 15) & ~(size_t)15;
}

LDPC_DEVINL float4 q4_sub(const float4 a, const float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
LDPC_DEVINL float4 q4_add(const float4 a, const float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
LDPC_DEVINL float &q4_at(float4 &a, const int q) { return q == 0 ? a.x : q == 1 ? a.y : q == 2 ? a.z : a.w; }
LDPC_DEVINL float q4_get(const float4 &a, const int q) { return q == 0 ? a.x : q == 1 ? a.y : q == 2 ? a.z : a.w; }

// DC / DV: compile-time bounds of the row / column weight; REGC: every row has weight DC.  NV: float4 vectors per value (FI = 4 NV).
template <int DC, int DV, bool REGC, int NV, int NT_MAX, int MINB>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_quad_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    constexpr int FI = 4 * NV;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);                  // [FI]
    float4 *msg4 = reinterpret_cast<float4 *>(smem_raw + 16 * FI);                  // [dvN][NV]
    const int N = c.N, M = c.M, nwords = (N + 31) >> 5, npad = nwords << 5, nblk = (N + 3) >> 2;
    float4 *S4 = msg4 + (size_t)c.dvN * NV;                                         // [N][NV]
    float4 *y4 = S4 + (size_t)N * NV;                                               // [N][NV]
    uint32_t *dbits = reinterpret_cast<uint32_t *>(y4 + (size_t)N * NV);            // [FI][nwords]
    float *yf = reinterpret_cast<float *>(y4), *Sf = reinterpret_cast<float *>(S4);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const float INF = real_inf<float>();
    const float alpha = (float)p.alpha, delta = (float)p.delta;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    const long long ntiles = (io.n_frames + FI - 1) / FI;
    CtaTotals tot; tot.clear();

    // this thread's row: (variable << 16 | message position) of every edge, kept for the whole kernel (dvN <= 65535)
    const bool has_row = tid < M;
    const int deg = has_row ? (REGC ? DC : (int)c.cn_deg[tid]) : 0;
    uint32_t edge[DC];
    {
        constexpr int VPL = 8, NG = (DC + VPL - 1) / VPL;
        const uint4 *cnv = reinterpret_cast<const uint4 *>(c.cn_pos);
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const uint4 w = has_row ? __ldg(&cnv[(size_t)g * M + tid]) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int q = 0; q < VPL; q++) {
                const int k = g * VPL + q;
                if (k < DC) edge[k] = (k < deg) ? ((__ldg(&c.cn_var[(size_t)k * M + tid]) << 16) | (uint32_t)IdxVec<uint16_t>::get(w, q)) : 0u;
            }
        }
    }
    float4 v[DC][NV];                                                                // c2v of the previous iteration, this thread's row

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long f0 = tile * FI;
        if (tid < FI) { fs[tid].uncoded = 0; fs[tid].errors = 0; fs[tid].flag = 0; }
        for (int w = tid; w < FI * nwords; w += nt) dbits[w] = 0u;
        __syncthreads();
        // ---- channel front end, (block of 4 samples, frame lane) per thread; S = y, c2v = 0 (src/decodeMinSum.cpp:214-240, :364-370)
        for (int t = tid; t < nblk * FI; t += nt) {
            const int b = t / FI, fl = t % FI;
            const long long f = (f0 + fl < io.n_frames) ? f0 + fl : io.n_frames - 1;     // dead lanes replay the last frame, unreported
            const uint8_t *cw = codeword_row(io, c, f);
            double s4[4];
            raw_samples4(io, p, c, f, cw, b, s4);
            uint32_t nib = 0; int unc = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                float vr;
                if (fcond) vr = condition_ms_guarded(s4[q], p, qflags);
                else {
                    double d = s4[q];
                    if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) d = quantize_ms(d, p);
                    if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) { if (d > p.Ymax) d = p.Ymax; if (d < -p.Ymax) d = -p.Ymax; }
                    vr = (float)d;
                }
                const bool rneg = !(vr > 0.0f);
                yf[(size_t)i * FI + fl] = vr; Sf[(size_t)i * FI + fl] = vr;
                const int cb = cw ? cw[i] : 0;
                unc += (int)(rneg != (cb != 0));
                nib |= (uint32_t)rneg << q;
                if (io.out_soft && p.T == 0 && f0 + fl < io.n_frames) {
                    if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i] = (double)vr;
                    else ((float *)io.out_soft)[(size_t)f * N + i] = vr;
                }
            }
            if (nib) atomicOr(&dbits[fl * nwords + ((4 * b) >> 5)], nib << ((4 * b) & 31));
            if (unc) atomicAdd(&fs[fl].uncoded, unc);
        }
#pragma unroll
        for (int k = 0; k < DC; k++)
#pragma unroll
            for (int h = 0; h < NV; h++) v[k][h] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        __syncthreads();

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            // ---- check-node phase: src/decodeMinSum.cpp:410-450 (+ :494-515), one row x FI frames per thread ----
            if (has_row) {
                float m1[FI], m2[FI]; uint32_t sg[FI];
#pragma unroll
                for (int q = 0; q < FI; q++) { m1[q] = INF; m2[q] = INF; sg[q] = 0u; }
#pragma unroll
                for (int k = 0; k < DC; k++) if (REGC || k < deg) {
                    const float4 *sp = S4 + (size_t)(edge[k] >> 16) * NV;
#pragma unroll
                    for (int h = 0; h < NV; h++) {
                        v[k][h] = q4_sub(sp[h], v[k][h]);                               // v2c = sum - c2v
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float x = q4_get(v[k][h], q), a = fabsf(x);
                            m2[4 * h + q] = fminf(m2[4 * h + q], fmaxf(m1[4 * h + q], a));
                            m1[4 * h + q] = fminf(m1[4 * h + q], a);
                            sg[4 * h + q] ^= __float_as_uint(x);
                        }
                    }
                }
                float s1[FI], s2[FI];
#pragma unroll
                for (int q = 0; q < FI; q++) {
                    float o1 = m1[q], o2 = m2[q];
                    if (normalized) { o1 = o1 / alpha; o2 = o2 / alpha; }
                    if (offset) { o1 = o1 - delta; o1 = (o1 > 0) ? o1 : 0.0f; o2 = o2 - delta; o2 = (o2 > 0) ? o2 : 0.0f; }
                    s1[q] = SignOps<float>::presign(o1, sg[q]); s2[q] = SignOps<float>::presign(o2, sg[q]);
                }
#pragma unroll
                for (int k = 0; k < DC; k++) if (REGC || k < deg) {
                    float4 *mp = msg4 + (size_t)(edge[k] & 0xffffu) * NV;
#pragma unroll
                    for (int h = 0; h < NV; h++) {
                        float4 o;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float x = q4_get(v[k][h], q);
                            const float sel = (fabsf(x) == m1[4 * h + q]) ? s2[4 * h + q] : s1[4 * h + q];
                            q4_at(o, q) = SignOps<float>::apply(sel, x);
                        }
                        v[k][h] = o; mp[h] = o;
                    }
                }
            }
            __syncthreads();
            // ---- variable-node phase: src/decodeMinSum.cpp:452-476, one variable x FI frames per thread ----
            for (int i0 = tid; i0 < npad; i0 += nt) {
                float4 sum[NV];
                const bool valid = i0 < N;
                if (valid) {
                    const int dv = (int)c.vn_deg[i0];
#pragma unroll
                    for (int h = 0; h < NV; h++) sum[h] = y4[(size_t)i0 * NV + h];
#pragma unroll
                    for (int s = 0; s < DV; s++) if (s < dv) {
#pragma unroll
                        for (int h = 0; h < NV; h++) sum[h] = q4_add(sum[h], msg4[((size_t)s * N + i0) * NV + h]);
                    }
#pragma unroll
                    for (int h = 0; h < NV; h++) S4[(size_t)i0 * NV + h] = sum[h];
                }
                if (last) {
#pragma unroll
                    for (int h = 0; h < NV; h++)
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float x = valid ? q4_get(sum[h], q) : 1.0f;
                            const unsigned bal = __ballot_sync(0xffffffffu, !(x > 0.0f));
                            if (lane == 0) dbits[(4 * h + q) * nwords + (i0 >> 5)] = bal;
                            if (valid && io.out_soft && f0 + 4 * h + q < io.n_frames) {
                                if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)(f0 + 4 * h + q) * N + i0] = (double)x;
                                else ((float *)io.out_soft)[(size_t)(f0 + 4 * h + q) * N + i0] = x;
                            }
                        }
                }
            }
            __syncthreads();
        }
        for (int fl = 0; fl < FI; fl++) {
            if (f0 + fl >= io.n_frames) break;                                       // uniform: dead lanes are not reported
            const uint8_t *cw = codeword_row(io, c, f0 + fl);
            finish_frame(c, p, io, f0 + fl, cw, dbits + fl * nwords, &fs[fl], p.T, -1, 0, 0, 1, -1, tot);
        }
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
