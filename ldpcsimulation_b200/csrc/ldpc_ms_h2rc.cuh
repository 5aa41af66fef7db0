// ldpc_ms_h2rc.cuh -- LDPC_GPU_PREC_F16X2 on the register-resident structure of ldpc_ms_rc.cuh.
//
// Same arithmetic as ms_h2_kernel (ldpc_ms_h2.cuh: every message word holds the same edge of TWO frames as
// a binary16 pair, v2c clamped to +-512; a labelled throughput instantiation, not the reference's
// arithmetic) and bit-identical results to it (tests/test_gpu_parity.py), organised like ms_rc_kernel:
//   * the row thread keeps its 32 c2v words in registers from one iteration to the next, the variable
//     phase publishes only the a-posteriori sums S = y + sum c2v, and the check thread rebuilds
//     v2c = clamp(S - c2v) itself: 3 shared-memory accesses per edge-iteration instead of 4;
//   * HMNMX2.XORSIGN (min.xorsign.abs.f16x2) carries the row's sign products in min1 and does the clamp
//     (sign(v) min(|v|, 512)) in one instruction; (min1, min2) advance two edges at a time (VHMNMX, the
//     three-input binary16x2 minimum);
//   * the select (|v| == min1 ? min2 : min1) is bits(min1) + bits(min2) - bits(min(|v|, min2)) on both halves
//     at once (no carry or borrow crosses the halves: every half-word stays within [0, 0xf7fe]).
#pragma once
#include "ldpc_ms_h2.cuh"
#include "ldpc_ms_rc.cuh"

namespace ldpc {

LDPC_DEVINL uint32_t h2_min_xorsign_abs(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

static inline size_t ms_h2rc_smem_bytes(const CodeDev &c)
{
    return ((size_t)32 + 4 * ((size_t)c.dvN + 3 * (size_t)c.N) + 16 * (size_t)((c.N + 31) / 32) + 16 + 15) & ~(size_t)15;
}

template <int DC, int DV, int NFIX, int NT_MAX, int MINB>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_h2rc_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = NFIX, nwords = (N + 31) >> 5, nblk = (N + 3) >> 2, NG = DC / 4, NB = N * 4;
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);                 // [2]
    uint32_t *msg = reinterpret_cast<uint32_t *>(smem_raw + 32);                  // [DV*N] c2v pairs
    uint32_t *S = msg + DV * N;                                                   // [N] a-posteriori sums
    uint32_t *yq = S + N;                                                         // [N]
    uint32_t *dbits = yq + N;                                                     // [2][nwords]
    unsigned char *msgb = reinterpret_cast<unsigned char *>(msg);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, M = c.M;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    const uint32_t INF2 = h2_bits(__floats2half2_rn(65504.0f, 65504.0f));
    const uint32_t inv_alpha2 = h2_bits(__float2half2_rn(p.inv_alpha_f)), one2 = h2_bits(__float2half2_rn(1.0f));
    const uint32_t delta2 = h2_bits(__float2half2_rn((float)p.delta)), clamp2 = h2_bits(__float2half2_rn(LDPC_H2_CLAMP));
    const bool has_row = tid < M;
    const int slot = has_row ? (int)__ldg(&c.row_slot[tid]) : 0;
    const long long npairs = (io.n_frames + 1) / 2;
    CtaTotals tot; tot.clear();
    uint32_t v[DC];                                                               // this thread's row: c2v of the previous iteration

    // channel front end for a frame pair: gen() stages block b of the pair (fa, fb) into ybuf / rnext / unc_next
    uint32_t *ybuf = dbits + 2 * nwords;                                          // [N]
    uint32_t *rnext = ybuf + N;                                                   // [2][nwords]
    int *unc_next = reinterpret_cast<int *>(rnext + 2 * nwords);                  // [2]
    // one frame of the pair at a time (the generator's live state is the register budget of this kernel:
    // the 32 c2v words stay live across it); frame a's four samples wait as binary16 in two registers
    auto gen_one = [&](long long f, const uint8_t *cw, int b, bool live, uint32_t &nib, int &unc, __half (&h)[4]) {
        double y4[4];
        raw_samples4(io, p, c, f, cw, b, y4);
        nib = 0; unc = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = 4 * b + q;
            float vf;
            if (fcond) vf = condition_ms_guarded(y4[q], p, qflags);
            else {
                double d = y4[q];
                if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) d = quantize_ms(d, p);
                if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) d = fmin(fmax(d, -p.Ymax), p.Ymax);
                vf = (float)d;
            }
            const bool ng = !(vf > 0.0f);
            h[q] = __float2half_rn(vf);
            unc += (int)(ng != ((cw ? cw[i] : 0) != 0));
            nib |= (uint32_t)ng << q;
            if (io.out_soft && p.T == 0 && live) {                                // T = 0: the conditioned samples themselves (fp32, as ms_h2_kernel)
                if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i] = vf;
                else ((float *)io.out_soft)[(size_t)f * N + i] = vf;
            }
        }
    };
    auto gen = [&](long long fa, long long fb, const uint8_t *cwa, const uint8_t *cwb, int b) {
        __half ha[4], hb[4];
        uint32_t niba, nibb; int unca, uncb;
        gen_one(fa, cwa, b, true, niba, unca, ha);
        gen_one(fb, cwb, b, fb != fa, nibb, uncb, hb);
        const uint2 cc = __ldg(reinterpret_cast<const uint2 *>(c.col_of_var) + b);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int col = (int)(((q < 2 ? cc.x : cc.y) >> (16 * (q & 1))) & 0xffffu);
            ybuf[col] = h2_bits(__halves2half2(ha[q], hb[q]));
        }
        if (niba) atomicOr(&rnext[(4 * b) >> 5], niba << ((4 * b) & 31));
        if (nibb) atomicOr(&rnext[nwords + ((4 * b) >> 5)], nibb << ((4 * b) & 31));
        if (unca) atomicAdd(&unc_next[0], unca);
        if (uncb) atomicAdd(&unc_next[1], uncb);
    };
    auto pair_frames = [&](long long pr, long long &fa, long long &fb) { fa = 2 * pr; fb = (2 * pr + 1 < io.n_frames) ? 2 * pr + 1 : 2 * pr; };

    if ((long long)blockIdx.x < npairs) {                                         // the CTA's first pair
        for (int w = tid; w < 2 * nwords; w += nt) rnext[w] = 0u;
        if (tid < 2) unc_next[tid] = 0;
        __syncthreads();
        long long fa, fb; pair_frames(blockIdx.x, fa, fb);
        const uint8_t *cwa = codeword_row(io, c, fa), *cwb = codeword_row(io, c, fb);
        for (int b = tid; b < nblk; b += nt) gen(fa, fb, cwa, cwb, b);
        __syncthreads();
    }

    for (long long pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        long long fa, fb; pair_frames(pr, fa, fb);                                // a dead lane replays frame fa, unreported
        const bool live_b = 2 * pr + 1 < io.n_frames;
        const uint8_t *cwa = codeword_row(io, c, fa), *cwb = codeword_row(io, c, fb);
        // install the staged pair: S = y, c2v = 0
        if (tid < 2) { fs[tid].uncoded = unc_next[tid]; fs[tid].errors = 0; fs[tid].flag = 0; }
        for (int w = tid; w < 2 * nwords; w += nt) dbits[w] = rnext[w];
        for (int cp = tid; cp < N / 2; cp += nt) {
            const uint2 w2 = *reinterpret_cast<const uint2 *>(&ybuf[2 * cp]);
            *reinterpret_cast<uint2 *>(&yq[2 * cp]) = w2;
            *reinterpret_cast<uint2 *>(&S[2 * cp]) = w2;
        }
#pragma unroll
        for (int k = 0; k < DC; k++) v[k] = 0u;
        __syncthreads();
        for (int w = tid; w < 2 * nwords; w += nt) rnext[w] = 0u;
        if (tid < 2) unc_next[tid] = 0;
        const long long prn = pr + gridDim.x;
        const bool have_next = prn < npairs;
        long long fna = 0, fnb = 0; if (have_next) pair_frames(prn, fna, fnb);
        const uint8_t *cwna = have_next ? codeword_row(io, c, fna) : nullptr, *cwnb = have_next ? codeword_row(io, c, fnb) : nullptr;
        const int gen_done = 0;                                                 // the next pair is generated after this one's iterations: generating it
                                                                                // inside the loop (after the variable phase or after the check row) spills the 32 c2v words (measured 22-24 vs 34 Gbit/s)
        if (p.T == 0) __syncthreads();

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            if (last) for (int w = tid; w < 2 * nwords; w += nt) dbits[w] = 0u;
            // ---- check-node phase: one row per thread, both frames -----------------------------------
            if (has_row) {
                uint32_t off[DC];
                uint32_t m1 = INF2, m2 = INF2;
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const uint4 w = __ldg(&c.sched[(size_t)g * M + tid]);
                    off[4 * g] = w.x; off[4 * g + 1] = w.y; off[4 * g + 2] = w.z; off[4 * g + 3] = w.w;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int k = g * 4 + q;
                        uint32_t so;
                        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(so) : "r"((uint32_t)slot), "r"((uint32_t)-NB), "r"(off[k]));
                        const __half2 s = h2_from(*reinterpret_cast<const uint32_t *>(msgb + DV * NB + so));
                        v[k] = h2_min_xorsign_abs(h2_bits(__hsub2(s, h2_from(v[k]))), clamp2);      // v2c = clamp(sum - c2v)
                    }
#pragma unroll
                    for (int q = 0; q < 4; q += 2) {
                        const uint32_t a = v[g * 4 + q], b = v[g * 4 + q + 1];
                        const uint32_t lo = h2_min_xorsign_abs(a, b);
                        const __half2 hi = __hmax2(__habs2(h2_from(a)), __habs2(h2_from(b)));
                        const __half2 t = __hmax2(__habs2(h2_from(m1)), __habs2(h2_from(lo)));
                        m2 = h2_bits(__hmin2(__hmin2(t, h2_from(m2)), hi));
                        m1 = h2_min_xorsign_abs(m1, lo);
                    }
                }
                const uint32_t sg = m1 & 0x80008000u;                              // the rows' sign products (both frames)
                m1 &= 0x7fff7fffu;
                const uint32_t K = m1 + m2;                                        // per half-word, no carry: <= 2 * 0x7bff
                const uint32_t mult = (normalized ? inv_alpha2 : one2) ^ sg;
#pragma unroll
                for (int k = 0; k < DC; k++) {
                    const uint32_t t = h2_bits(__hmin2(__habs2(h2_from(v[k])), h2_from(m2)));
                    uint32_t rb;
                    asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rb) : "r"(t), "r"(K));                 // K - bits(t), half-word-wise
                    uint32_t o = h2_bits(__hmul2(h2_from(rb), h2_from(mult)));
                    if (offset) o = h2_bits(__hsub2(h2_from(o), h2_from(h2_min_xorsign_abs(o, delta2))));   // sgn(o) max(|o| - delta, 0)
                    o ^= v[k] & 0x80008000u;
                    v[k] = o;                                                      // c2v, kept for the next iteration
                    *reinterpret_cast<uint32_t *>(msgb + off[k]) = o;
                }
            }
            __syncthreads();
            // ---- variable-node phase: S = y + sum_s c2v, two adjacent storage columns per thread --------
            constexpr int VN_ROUNDS = (N / 2 + NT_MAX - 1) / NT_MAX;               // the host launches exactly NT_MAX threads
#pragma unroll
            for (int rr = 0; rr < VN_ROUNDS; rr++) {
                const int cp = tid + rr * NT_MAX;
                if (cp >= N / 2) break;
                const int col = 2 * cp;
                const uint2 y2 = *reinterpret_cast<const uint2 *>(&yq[col]);
                __half2 s0 = h2_from(y2.x), s1 = h2_from(y2.y);
#pragma unroll
                for (int s = 0; s < DV; s++) {
                    const uint2 cm = *reinterpret_cast<const uint2 *>(&msg[s * N + col]);
                    s0 = __hadd2(s0, h2_from(cm.x)); s1 = __hadd2(s1, h2_from(cm.y));
                }
                *reinterpret_cast<uint2 *>(&S[col]) = make_uint2(h2_bits(s0), h2_bits(s1));
                if (last) {
                    const unsigned vv = __ldg(reinterpret_cast<const unsigned *>(c.var_of_col) + cp);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int i = h ? (int)(vv >> 16) : (int)(vv & 0xffffu);
                        const __half2 sum = h ? s1 : s0;
                        const float sa = __low2float(sum), sb = __high2float(sum);
                        if (!(sa > 0.0f)) atomicOr(&dbits[i >> 5], 1u << (i & 31));
                        if (!(sb > 0.0f)) atomicOr(&dbits[nwords + (i >> 5)], 1u << (i & 31));
                        if (io.out_soft) {
                            if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)fa * N + i] = sa; if (live_b) ((double *)io.out_soft)[(size_t)fb * N + i] = sb; }
                            else { ((float *)io.out_soft)[(size_t)fa * N + i] = sa; if (live_b) ((float *)io.out_soft)[(size_t)fb * N + i] = sb; }
                        }
                    }
                }
            }
            __syncthreads();
        }
        finish_frame(c, p, io, fa, cwa, dbits, &fs[0], p.T, -1, 0, 0, 1, -1, tot);
        if (live_b) finish_frame(c, p, io, fb, cwb, dbits + nwords, &fs[1], p.T, -1, 0, 0, 1, -1, tot);
        if (have_next) for (int b = gen_done + tid; b < nblk; b += nt) gen(fna, fnb, cwna, cwnb, b);      // what T iterations did not cover
        __syncthreads();
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
