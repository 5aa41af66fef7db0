// ldpc_common.cuh -- pieces every decode kernel shares: the channel front end (a2/a3), the
// per-frame epilogue and the accounting of a19 (src/decodeMinSum.cpp:270-288).
#pragma once
#include <math.h>
#include <cuda_fp16.h>
#include "ldpc_types.cuh"
#include "ldpc_rng.cuh"

namespace ldpc {

#define LDPC_DEVINL __device__ __forceinline__

// sgn(0) = +1: src/decodeMinSum.cpp:518-523 (MS, BP, DD-BMP)
LDPC_DEVINL bool neg_ge(double x) { return !(x >= 0.0); }
LDPC_DEVINL bool neg_ge(float x)  { return !(x >= 0.0f); }
// sgn(0) = -1: src/decodeGDBF.cpp:495-501 (GDBF family, NGDBFhw)
LDPC_DEVINL bool neg_gt(double x) { return !(x > 0.0); }

LDPC_DEVINL double absr(double x) { return fabs(x); }
LDPC_DEVINL float  absr(float x)  { return fabsf(x); }
template <typename Real> LDPC_DEVINL Real real_inf();
template <> LDPC_DEVINL double real_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> LDPC_DEVINL float  real_inf<float>()  { return __int_as_float(0x7f800000); }

// a3: MS / DD-BMP quantiser, src/decodeMinSum.cpp:480-489
LDPC_DEVINL double quantize_ms(double x, const DecParams &p)
{
    const double s = neg_ge(x) ? -1.0 : 1.0;
    if (fabs(x) > p.Ymax) return s * p.Ymax;
    // dividing by a power of two is exact, so it may be a multiplication by the (exact) reciprocal
    const double t = fabs(x) * p.ms_Nq1;
    double q = s * (floor(p.ms_inv_twoY != 0.0 ? t * p.ms_inv_twoY : t / p.ms_twoY) + 0.0) * p.ms_step;
    if (q == 0.0) q = s * p.ms_step;
    return q;
}
// fp32 front end of the min-sum family, used by the fp32 instantiation when its samples are fp32
// anyway (Philox channel or LDPC_GPU_DT_F32 input): the same clip / floor / no-zero-level rule as
// quantize_ms (src/decodeMinSum.cpp:480-489, :224-229), evaluated in fp32.  A sample within one fp32
// ulp of a level boundary may land one level away from the fp64 rule; the fp64 instantiation, and
// the fp32 one on fp64 input, keep the double arithmetic.
LDPC_DEVINL float condition_ms_f32(float y, const DecParams &p, uint32_t qflags)
{
    float v = y;
    if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) {
        const float a = fabsf(y);
        // level * step: exact in fp32 when the step is a dyadic fp32 number; otherwise the product is rounded once, from double,
        // as (float)(level * step) of the double rule is
        const float lev = fmaxf(floorf(a * p.ms_scale_f), 1.0f);
        float q = p.ms_step_dyadic ? lev * p.ms_step_f : (float)__dmul_rn((double)lev, p.ms_step);
        q = (a > p.Ymax_f) ? p.Ymax_f : q;
        v = (y >= 0.0f) ? q : -q;
    }
    if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) v = fminf(fmaxf(v, -p.Ymax_f), p.Ymax_f);
    return v;
}

// The same conditioning of a DOUBLE sample, guaranteed to return what the double rule (quantize_ms + clamp) returns: the
// fp32 evaluation is used wherever it provably lands on the same level, i.e. when floor()'s argument is further than 2^-13
// from an integer (the fp32 error of |y| (Nq-1)/(2 Ymax) is below 3e-5 for Q <= 8) and |y| is not within 1e-5 of Ymax; the
// few samples inside those bands (about 3 in 10^4) are redone in double.  So the fp32 / binary16 instantiations and the fp64
// parity instantiation decode the very same quantised samples.
LDPC_DEVINL float condition_ms_guarded(double yd, const DecParams &p, uint32_t qflags)
{
    const float yf = (float)yd;
    if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) {
        const float a = fabsf(yf), t = a * p.ms_scale_f, fl = floorf(t);
        const bool risky = (t - fl) < 1.220703125e-4f || (fl + 1.0f - t) < 1.220703125e-4f || fabsf(a - p.Ymax_f) < 1e-5f || p.Q > 8;
        if (risky) {
            double d = quantize_ms(yd, p);
            if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) d = fmin(fmax(d, -p.Ymax), p.Ymax);
            return (float)d;
        }
    }
    return condition_ms_f32(yf, p, qflags);
}

// a3: GDBF quantiser, src/decodeGDBF.cpp:488-493
LDPC_DEVINL double quantize_gdbf(double x, const DecParams &p)
{
    const double s = neg_gt(x) ? -1.0 : 1.0;
    return s * floor((fabs(x) * p.g_qmax) / p.g_twol + 0.5) * p.g_step;
}
// a3/a17: NGDBFhw quantize(double)+pack(), src/NGDBFhw.cpp:639-663 (NQ = 5)
LDPC_DEVINL int hw_pack(double ymod, const DecParams &p)
{
    const bool neg = neg_gt(ymod);
    const int k = (int)floor(fabs(ymod) * p.hw_NL / p.hw_two_lmax);
    return (k & 31) | (neg ? 16 : 0);
}
// a17: unpack(), src/NGDBFhw.cpp:665-677
LDPC_DEVINL int hw_unpack(int code)
{
    const int mag = ((code << 1) & 31) | 1;
    return (code & 16) ? -mag : mag;
}

// Which codeword row does frame f carry?  decode_batch: row f of `codeword`; simulate: row
// (frame id mod n_cw) of the table (data.enc is read cyclically, src/decodeMinSum.cpp:195-211).
LDPC_DEVINL const uint8_t *codeword_row(const FrameIO &io, const CodeDev &c, long long f)
{
    if (io.codeword) return io.codeword + (size_t)f * c.N;
    if (io.cw_table && io.n_cw > 0) return io.cw_table + (size_t)((unsigned long long)(io.frame_begin + f) % (unsigned long long)io.n_cw) * c.N;
    return nullptr;
}

// LDPC_GPU_CHANNEL_FAST: y = x (1 + sigma n) in fp32, one FFMA (the sign flip is exact)
LDPC_DEVINL float fast_channel_sample(const DecParams &p, const uint8_t *cw, int i, int N, float n)
{
    const float y = __fmaf_rn(p.sigma_f, n, 1.0f);
    return (cw && i < N && cw[i]) ? -y : y;
}

// LDPC_GPU_DT_QP: the four Q-bit level codes of block b of a frame whose packed words start at w -> conditioned values (what
// quantize() makes of the samples).  STAGED: w points into shared memory (the frame was staged by a bulk copy), else global.
template <bool STAGED>
LDPC_DEVINL void packed_levels4_at(const uint32_t *w, const DecParams &p, const int b, double y[4])
{
    const int Q = p.Q;
    const uint32_t bit = (uint32_t)(4 * b * Q), wi = bit >> 5, sh = bit & 31u;
    const uint32_t lo = STAGED ? w[wi] : __ldg(w + wi), hi = (sh + 4 * Q > 32) ? (STAGED ? w[wi + 1] : __ldg(w + wi + 1)) : 0u;
    const uint32_t v = __funnelshift_r(lo, hi, sh);
    const uint32_t mmask = (1u << (Q - 1)) - 1u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t code = (v >> (q * Q)) & ((1u << Q) - 1u);
        const uint32_t mag = code & mmask;
        const double a = (mag == mmask) ? p.Ymax : __dmul_rn((double)(mag + 1u), p.ms_step);
        y[q] = (code >> (Q - 1)) ? -a : a;
    }
}
LDPC_DEVINL void packed_levels4(const FrameIO &io, const DecParams &p, const int N, const long long f, const int b, double y[4])
{
    packed_levels4_at<false>(reinterpret_cast<const uint32_t *>(io.y) + (size_t)f * (((size_t)N * p.Q) >> 5), p, b, y);   // N*Q is a multiple of 32
}

// ---- 1-D bulk copies global -> shared through the TMA unit (cp.async.bulk, completion on an mbarrier) ---------------------
// Used to stage the NEXT frames' packed samples while the current ones iterate (ldpc_ms_x2.cuh): the copy costs one thread two
// instructions and no registers, and its DRAM latency disappears behind the iterations.  Addresses and sizes: multiples of 16 B.
LDPC_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
LDPC_DEVINL void mbar_init(uint64_t *bar, const int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
LDPC_DEVINL void mbar_expect_tx(uint64_t *bar, const uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
LDPC_DEVINL void tma_load_1d(void *dst, const void *src, const uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
LDPC_DEVINL void mbar_wait(uint64_t *bar, const uint32_t parity)
{
    asm volatile("{\n .reg .pred P1;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra DONE;\n bra LAB_WAIT;\n DONE:\n }"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// a2: four raw channel samples y = x(1 + sigma n) of block b (src/decodeMinSum.cpp:216), from the
// caller's array or from the Philox channel.
// FULL = false (bit-flipping kernels, whose register budget decides their occupancy): without the bit-packed level format and the
// fast channel, which exist for the message-passing family only (the host refuses them elsewhere).
template <bool FULL = true>
LDPC_DEVINL void raw_samples4(const FrameIO &io, const DecParams &p, const CodeDev &c, long long f, const uint8_t *cw, int b, double y[4])
{
    const int i0 = 4 * b;
    if (FULL && io.y && io.y_dtype == LDPC_GPU_DT_QP) {
        packed_levels4(io, p, c.N, f, b, y);
    } else if (io.y) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + q;
            if (i < c.N) {
                if (io.y_dtype == LDPC_GPU_DT_Q8) {                 // quantiser levels: the conditioned value itself (the launch clears the quantiser flags)
                    const int k = (int)((const signed char *)io.y)[(size_t)f * c.N + i];
                    y[q] = (k >= 32) ? p.Ymax : (k <= -32) ? -p.Ymax : __dmul_rn((double)k, p.ms_step);
                } else
                y[q] = (io.y_dtype == LDPC_GPU_DT_F64) ? ((const double *)io.y)[(size_t)f * c.N + i]
                     : (io.y_dtype == LDPC_GPU_DT_F32) ? (double)((const float *)io.y)[(size_t)f * c.N + i]
                                                       : (double)__half2float(((const __half *)io.y)[(size_t)f * c.N + i]);
            }
            else y[q] = 1.0;
        }
    } else if (FULL && p.channel_mode == LDPC_GPU_CHANNEL_FAST) {
        float n[4];
        normal4_fast(io.seed, (unsigned long long)(io.frame_begin + f), (uint32_t)b, 0u, STREAM_CHANNEL, n);
#pragma unroll
        for (int q = 0; q < 4; q++) y[q] = (double)fast_channel_sample(p, cw, i0 + q, c.N, n[q]);
    } else {
        float n[4];
        normal4(io.seed, (unsigned long long)(io.frame_begin + f), (uint32_t)b, 0u, STREAM_CHANNEL, n);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + q;
            const double x = (cw && i < c.N && cw[i]) ? -1.0 : 1.0;
            y[q] = __dmul_rn(x, __dadd_rn(1.0, __dmul_rn(p.sigma, (double)n[q])));
        }
    }
}

// Per-CTA running totals, flushed once per launch.  They live in (static) shared memory and are
// touched by thread 0 only: as a register struct they cost 16 registers in every thread of the CTA
// for the whole kernel (profiles/r1_summary.md).
struct CtaTotals {
    unsigned long long *v;
    LDPC_DEVINL void clear() {
        __shared__ unsigned long long s_totals[CNT_N];
        v = s_totals;
        if (threadIdx.x < CNT_N) s_totals[threadIdx.x] = 0ull;
        __syncthreads();
    }
    LDPC_DEVINL void flush(unsigned long long *g) {
        if (!g) return;
        for (int q = 0; q < CNT_N; q++) if (v[q]) atomicAdd(&g[q], v[q]);
    }
};

// Shared scratch every kernel reserves at the front of its dynamic shared memory.
struct FrameScratch {
    int uncoded;      // increments of uncodedErrors for the current frame
    int errors;       // Hamming distance to the codeword
    int flag;         // block-wide boolean scratch
    int pad;
};

// End of frame: count decision errors against the codeword (countDecisionErrors,
// src/decodeMinSum.cpp:382-393), emit the per-frame outputs and do the accounting of
// src/decodeMinSum.cpp:270-288.  `dbits` holds the hard decisions, bit i = 1 <-> d_i = -1.
LDPC_DEVINL bool syndrome_ok(const CodeDev &c, const uint32_t *dbits);

// `satisfied` < 0 means "not computed": the syndrome of the decisions is then evaluated here, and only
// when somebody consumes it (a flags output, or a frame in error for the undetected-error counter).
// Must be called by every thread of the CTA; ends with a barrier.
LDPC_DEVINL void finish_frame(const CodeDev &c, const DecParams &p, const FrameIO &io, long long f, const uint8_t *cw,
                              const uint32_t *dbits, FrameScratch *fs, int it, int satisfied, int smoothed,
                              int smoothing_used, int phases, int errors_override, CtaTotals &tot)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int npad = (c.N + 31) & ~31;
    const size_t bpf = (size_t)(c.N + 7) >> 3;
    int local_err = 0;
    if (!cw && !io.out_bits) {
        // all-zero codeword, no decision output (the throughput entry's common case): errors = popcount of the decision
        // words, one word per thread instead of one bit per thread (bits past N are never set)
        for (int w = tid; w < (npad >> 5); w += nt) local_err += __popc(dbits[w]);
        for (int o = 16; o; o >>= 1) local_err += __shfl_xor_sync(0xffffffffu, local_err, o);
    } else
    for (int i0 = tid; i0 < npad; i0 += nt) {
        const bool valid = i0 < c.N;
        const uint32_t word = dbits[i0 >> 5];
        const int bit = (word >> lane) & 1;
        const int cb = (valid && cw) ? cw[i0] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, valid && (bit != cb));
        if (lane == 0) local_err += __popc(bal);
        if (io.out_bits && valid && (lane & 7) == 0) io.out_bits[(size_t)f * bpf + (i0 >> 3)] = (uint8_t)((word >> lane) & 0xffu);
    }
    if (lane == 0 && local_err) atomicAdd(&fs->errors, local_err);
    __syncthreads();
    if (satisfied < 0) {
        const int e_all = errors_override >= 0 ? errors_override : fs->errors;       // same value in every thread
        satisfied = (io.out_flags || e_all > 0) ? (int)syndrome_ok(c, dbits) : 0;
    }
    if (tid == 0) {
        const int e = errors_override >= 0 ? errors_override : fs->errors;
        if (io.out_iters)  io.out_iters[f] = it;
        if (io.out_errors) io.out_errors[f] = e;
        if (io.out_flags)  io.out_flags[f] = (uint8_t)((satisfied ? 1 : 0) | (smoothed ? 2 : 0) | ((phases & 15) << 4));
        if (e > 0) {
            tot.v[CNT_ERRORS] += (unsigned long long)e; tot.v[CNT_WORDERRS] += 1ull;
            if (io.ew_hist) atomicAdd(&io.ew_hist[e - 1], 1ull);
            if (satisfied) tot.v[CNT_UNDETECTED] += 1ull;
        }
        tot.v[CNT_UNCODED] += (unsigned long long)fs->uncoded;
        tot.v[CNT_WORDS] += 1ull; tot.v[CNT_BITS] += (unsigned long long)c.N; tot.v[CNT_ITERS] += (unsigned long long)it;
        tot.v[CNT_SMOOTH] += (unsigned long long)smoothing_used;
        if (io.it_hist && it < p.iter_hist_len) atomicAdd(&io.it_hist[it], 1ull);
        if (io.ph_hist && (p.flags & LDPC_GPU_F_REDECODE) && p.kind == LDPC_GPU_KIND_GDBF) atomicAdd(&io.ph_hist[phases - 1], 1ull);
    }
    __syncthreads();
}

// All parity checks satisfied by the decisions in dbits?  Every thread must call; returns the
// same value in every thread.
LDPC_DEVINL bool syndrome_ok(const CodeDev &c, const uint32_t *dbits)
{
    int bad = 0;
    for (int j = threadIdx.x; j < c.M; j += blockDim.x) {
        const int deg = c.cn_deg[j];
        unsigned par = 0;
        for (int k = 0; k < deg; k++) {
            const uint32_t i = c.cn_var[(size_t)k * c.M + j];
            par ^= dbits[i >> 5] >> (i & 31);
        }
        bad |= (int)(par & 1u);
    }
    return __syncthreads_or(bad) == 0;
}

} // namespace ldpc
