// ldpc_bf_kernels.cuh -- bit-flipping decoders: the GDBF / NGDBF family (src/decodeGDBF.cpp,
// src/RNGDBF.cpp) and the all-integer NGDBFhw model (src/NGDBFhw.cpp).
//
// One CTA owns one frame at a time; per-frame state on chip:
//     yq[N] (Real)  theta[N] (Real)  shape[N] (Real, noiseShaping)  emet[N] (Real, sequential+adaptation)
//     dsum[N] (int, smoothing)   d[N], r[N] (int8 +-1)   c2s[M] (int8 +-1)   dbits[N/32]
// Each thread handles four consecutive variables per step, which is one Philox block of
// perturbation noise (or four consecutive entries of the caller's noise row).
// The syndrome is frozen during a flip step (the reference flips d in place but reads only c2s,
// src/decodeGDBF.cpp:536-621), so flipping all variables in parallel is equivalent.
#pragma once
#include "ldpc_common.cuh"

namespace ldpc {

template <typename Real>
struct GdbfSmem {
    FrameScratch *fs; Real *yq, *theta, *shape, *emet; int *dsum; signed char *d, *r, *c2s; uint32_t *dbits; double *red; int *redi;
};

// per-frame arrays (shared memory, or a per-CTA slice of the HBM workspace when GSTATE)
template <typename Real>
static inline size_t gdbf_state_bytes(const CodeDev &c)
{
    size_t n = sizeof(Real) * 4 * (size_t)c.N + 4 * (size_t)c.N + 2 * (size_t)c.N + (size_t)c.M + 16;
    return (n + 15) & ~(size_t)15;
}
template <typename Real>
static inline size_t gdbf_smem_bytes(const CodeDev &c, bool gstate = false)
{
    size_t n = 16 + 4 * (size_t)((c.N + 31) / 32) + 16;
    n = (n + 15) & ~(size_t)15;
    n += 8 * 32 + 4 * 32;                 // per-warp reduction scratch
    n = (n + 15) & ~(size_t)15;
    if (!gstate) n += gdbf_state_bytes<Real>(c);
    return n;
}

template <typename Real>
LDPC_DEVINL GdbfSmem<Real> gdbf_carve(unsigned char *raw, unsigned char *state, const CodeDev &c)
{
    GdbfSmem<Real> s;
    s.fs = reinterpret_cast<FrameScratch *>(raw);
    size_t off = 16;
    s.dbits = reinterpret_cast<uint32_t *>(raw + off);
    off += 4 * (size_t)((c.N + 31) / 32) + 16; off = (off + 15) & ~(size_t)15;
    s.red = reinterpret_cast<double *>(raw + off);
    s.redi = reinterpret_cast<int *>(raw + off + 8 * 32);
    off += 8 * 32 + 4 * 32; off = (off + 15) & ~(size_t)15;
    if (!state) state = raw + off;
    s.yq = reinterpret_cast<Real *>(state);
    s.theta = s.yq + c.N; s.shape = s.theta + c.N; s.emet = s.shape + c.N;
    s.dsum = reinterpret_cast<int *>(s.emet + c.N);
    s.d = reinterpret_cast<signed char *>(s.dsum + c.N);
    s.r = s.d + c.N; s.c2s = s.r + c.N;
    return s;
}

// Pack d (+-1 bytes) into dbits (bit = 1 <-> d = -1).  Ends with a barrier.
LDPC_DEVINL void pack_decisions(const CodeDev &c, const signed char *d, uint32_t *dbits)
{
    const int npad = (c.N + 31) & ~31, lane = threadIdx.x & 31;
    for (int i0 = threadIdx.x; i0 < npad; i0 += blockDim.x) {
        const unsigned bal = __ballot_sync(0xffffffffu, i0 < c.N && d[i0] < 0);
        if (lane == 0) dbits[i0 >> 5] = bal;
    }
    __syncthreads();
}

// evaluateObjectiveFunction (src/decodeGDBF.cpp:624-633).  The reference accumulates N+M terms
// front to back in one double; the comparison f1 >= f2 that follows is an exact floating-point
// test, so the sum is kept sequential (thread 0) instead of a tree.  Mode switching is not on the
// throughput path.  Ends with a barrier; result valid in every thread.
template <typename Real>
LDPC_DEVINL double gdbf_objective(const CodeDev &c, const GdbfSmem<Real> &s)
{
    if (threadIdx.x == 0) {
        double f = 0;
        for (int i = 0; i < c.N; i++) f += (double)((Real)s.d[i] * s.yq[i]);
        for (int j = 0; j < c.M; j++) f += (double)s.c2s[j];
        s.red[0] = f;
    }
    __syncthreads();
    const double f = s.red[0];
    __syncthreads();
    return f;
}

template <typename Real, bool GSTATE>
__global__ void gdbf_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GdbfSmem<Real> s = gdbf_carve<Real>(smem_raw, GSTATE ? io.workspace + (size_t)blockIdx.x * io.ws_stride : nullptr, c);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int N = c.N, M = c.M, nblk = (N + 3) >> 2, T = p.T, W = p.windowsize;
    const uint32_t fl = p.flags;
    const bool redecode = (fl & LDPC_GPU_F_REDECODE) != 0;
    const int maxphase = redecode ? (p.maxphase > 0 ? p.maxphase : 1) : 1;
    const Real theta0 = (Real)p.theta, lambda = (Real)p.lambda;
    const Real INF = real_inf<Real>();
    CtaTotals tot; tot.clear();

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        const unsigned long long fid = (unsigned long long)(io.frame_begin + f);
        if (tid == 0) { s.fs->uncoded = 0; s.fs->errors = 0; s.fs->flag = 0; }
        __syncthreads();
        // ---- channel front end: src/decodeGDBF.cpp:251-274 / src/RNGDBF.cpp:251-275
        int unc = 0;
        for (int b = tid; b < nblk; b += nt) {
            double y4[4];
            raw_samples4<false>(io, p, c, f, cw, b, y4);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                double v = y4[q];
                if (fl & LDPC_GPU_F_SATURATE_SAMPLES) if (fabs(v) > p.Ymax) v *= p.Ymax / fabs(v);
                const bool rneg = !(v > 0);
                if (fl & LDPC_GPU_F_QUANTIZE_SAMPLES) v = quantize_gdbf(v, p);
                s.yq[i] = (Real)v;
                const int cb = cw ? cw[i] : 0;
                unc += (int)(rneg != (cb != 0));
                s.r[i] = rneg ? -1 : 1; s.d[i] = rneg ? -1 : 1;
                s.dsum[i] = 0; s.shape[i] = (Real)0; s.theta[i] = theta0;
            }
        }
        for (int o = 16; o; o >>= 1) unc += __shfl_xor_sync(0xffffffffu, unc, o);
        if (lane == 0 && unc) atomicAdd(&s.fs->uncoded, unc);
        __syncthreads();

        const Real noiseSigma = (Real)p.noiseSigma;
        int it = 0, total_it = 0, phase = 0, satisfied = 1, smoothed = 0, smoothing_used = 0;
        long long row = 0;                                            // next row of the noise array / Philox row
        while (phase < maxphase) {                                    // src/RNGDBF.cpp:280-400
            if (redecode) { for (int i = tid; i < N; i += nt) { s.d[i] = s.r[i]; s.dsum[i] = 0; } }
            if (fl & LDPC_GPU_F_THRESHOLDADAPTATION) for (int i = tid; i < N; i += nt) s.theta[i] = theta0;
            int mu = (fl & LDPC_GPU_F_SEQUENTIALMODE) ? 0 : 1;
            __syncthreads();
            double f1 = 0, f2 = 0;
            for (it = 0; it < T; it++) {
                // checkNodeUpdates, src/decodeGDBF.cpp:517-534
                int bad = 0;
                for (int j = tid; j < M; j += nt) {
                    const int deg = c.cn_deg[j];
                    int ng = 0;
                    for (int k = 0; k < deg; k++) ng ^= (s.d[c.cn_var[(size_t)k * M + j]] < 0);
                    s.c2s[j] = ng ? -1 : 1;
                    bad |= ng;
                }
                satisfied = (__syncthreads_or(bad) == 0);
                if (satisfied) break;                                 // :305-306
                if ((fl & LDPC_GPU_F_MODESWITCHING) && it > p.Tswitch) f1 = gdbf_objective<Real>(c, s);

                const long long row_pert = (fl & LDPC_GPU_F_ADDNOISE) ? row++ : -1;
                const long long row_unif = (fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) ? row++ : -1;
                // symNodeUpdates, src/decodeGDBF.cpp:536-621
                Real best = INF; int besti = -1;
                for (int b = tid; b < nblk; b += nt) {
                    double pert4[4] = {0, 0, 0, 0}, unif4[4] = {0, 0, 0, 0};
                    if (row_pert >= 0) {                              // :318-333
                        if (io.noise) {
#pragma unroll
                            for (int q = 0; q < 4; q++) if (4 * b + q < N) pert4[q] = io.noise[((size_t)f * io.noise_rows + row_pert) * N + 4 * b + q];
                        } else if (fl & LDPC_GPU_F_UNIFORMNOISE) uniform4(io.seed, fid, (uint32_t)b, (uint32_t)(row_pert + io.noise_row_base), STREAM_DECODER, pert4);
                        else { float n4[4]; normal4(io.seed, fid, (uint32_t)b, (uint32_t)(row_pert + io.noise_row_base), STREAM_DECODER, n4);
#pragma unroll
                               for (int q = 0; q < 4; q++) pert4[q] = (double)n4[q]; }
                    }
                    if (row_unif >= 0) {
                        if (io.noise) {
#pragma unroll
                            for (int q = 0; q < 4; q++) if (4 * b + q < N) unif4[q] = io.noise[((size_t)f * io.noise_rows + row_unif) * N + 4 * b + q];
                        } else uniform4(io.seed, fid, (uint32_t)b, (uint32_t)(row_unif + io.noise_row_base), STREAM_DECODER, unif4);
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int i = 4 * b + q;
                        if (i >= N) break;
                        const int deg = c.vn_deg[i];
                        const signed char di = s.d[i];
                        Real E = (di > 0) ? s.yq[i] : -s.yq[i];                    // d[i]*y[i]
                        Real wgt = (Real)1;
                        if (fl & LDPC_GPU_F_WEIGHTSYNDROMES)
                            wgt = redecode ? (Real)(p.alpha * p.Ymax / (double)deg)   // src/RNGDBF.cpp:566
                                           : (Real)p.alpha;                           // src/decodeGDBF.cpp:550
                        for (int sl = 0; sl < deg; sl++) E += (s.c2s[c.vn_chk[(size_t)sl * N + i]] > 0) ? wgt : -wgt;
                        if (row_pert >= 0) {
                            Real smp = (fl & LDPC_GPU_F_UNIFORMNOISE) ? (Real)(p.uni_scale * (pert4[q] - 0.5))
                                                                      : (Real)(p.noiseSigma * pert4[q]);
                            if (fl & LDPC_GPU_F_NOISESHAPING) { const Real prev = s.shape[i]; s.shape[i] = smp; smp = smp - prev; }
                            E += smp;
                        }
                        bool flip = false;
                        if (fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) {               // :561-597
                            const double val = ((double)(-E + s.theta[i])) / (double)noiseSigma;
                            const double pcdf = 0.5 * erfc(-val * 0.70710678118654752440);
                            const double lv[8] = { 0, 0.0625, 0.125, 0.25, 0.34375, 0.4106, 0.68359, 1 };
                            double md = 1; int mi = 0;
#pragma unroll
                            for (int l = 0; l < 8; l++) { double t = lv[l] - pcdf; t = t * t; if (t < md) { md = t; mi = l; } }
                            if (unif4[q] < lv[mi]) { flip = true; s.d[i] = -di; }
                        } else {
                            if (mu == 1 && E < s.theta[i]) { flip = true; s.d[i] = -di; }
                            if (mu == 0) { s.emet[i] = E; if (E < best) { best = E; besti = i; } }
                        }
                        if ((fl & LDPC_GPU_F_THRESHOLDADAPTATION) && !(mu == 0 && !(fl & LDPC_GPU_F_QUANTIZEPROBABILITIES)))
                            if (!flip) s.theta[i] *= lambda;                      // :612-617
                    }
                }
                if (mu == 0 && !(fl & LDPC_GPU_F_QUANTIZEPROBABILITIES)) {
                    // sequential mode: flip the first strict minimum (:604-610, :619-620)
                    for (int o = 16; o; o >>= 1) {
                        const Real ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                        if (oi >= 0 && (besti < 0 || ob < best || (ob == best && oi < besti))) { best = ob; besti = oi; }
                    }
                    if (lane == 0) { s.red[warp] = (double)best; s.redi[warp] = besti; }
                    __syncthreads();
                    if (tid == 0) {
                        double bb = s.red[0]; int bi = s.redi[0];
                        for (int wv = 1; wv < nwarps; wv++) { const double ob = s.red[wv]; const int oi = s.redi[wv];
                            if (oi >= 0 && (bi < 0 || ob < bb || (ob == bb && oi < bi))) { bb = ob; bi = oi; } }
                        if (bi >= 0) s.d[bi] = -s.d[bi];
                        if (fl & LDPC_GPU_F_THRESHOLDADAPTATION) {
                            // `flip` is raised at every new running minimum of E (:604-610), so theta_i is left
                            // alone exactly at the prefix minima and multiplied by lambda elsewhere
                            Real run = INF;
                            for (int i = 0; i < N; i++) { const Real e = s.emet[i]; if (e < run) run = e; else s.theta[i] *= lambda; }
                        }
                    }
                }
                __syncthreads();
                if ((fl & LDPC_GPU_F_MODESWITCHING) && it > p.Tswitch) {          // :338-346
                    f2 = gdbf_objective<Real>(c, s);
                    if (f1 >= f2) mu = 0;
                }
                if ((fl & LDPC_GPU_F_OUTPUTSMOOTHING) && it > T - W)               // :348-354
                    for (int i = tid; i < N; i += nt) s.dsum[i] += s.d[i];
            }
            if (fl & LDPC_GPU_F_OUTPUTSMOOTHING) {
                __syncthreads();
                if (!satisfied) { for (int i = tid; i < N; i += nt) s.d[i] = (s.dsum[i] > 0) ? 1 : -1; smoothed = 1; }   // :358-367
                else smoothed = 0;
                if (it > T - W) smoothing_used++;                                  // :371-374
            }
            total_it += it; phase++;
            __syncthreads();
            if (!redecode || satisfied) break;                                    // src/RNGDBF.cpp:398-399
        }
        pack_decisions(c, s.d, s.dbits);
        finish_frame(c, p, io, f, cw, s.dbits, s.fs, total_it, satisfied, smoothed, smoothing_used, phase, -1, tot);
    }
    if (tid == 0) tot.flush(io.counters);
}

// ---------------------------------------------------------------------------------------------
// gdbf_par_kernel: the parallel-flipping members of the family (everything except -D sequentialmode
// and -D modeswitching), restructured like hw_kernel below: decisions and syndromes are bit-packed,
// and because the syndrome is linear in d over GF(2) it is computed in full once per phase and then
// updated by the bits that flip (atomicXor into a toggle mask folded in between iterations) instead of
// being recomputed from d through E gathers every iteration (checkNodeUpdates, decodeGDBF.cpp:517-534).
// The flip metric is accumulated in the reference's order (d*y, then +w*s_j in nlist order, then the
// perturbation, :541-556) and reads the syndromes of the iteration's start, so the fp64 instantiation
// stays bit-exact.  Graph (variable side) cached in shared memory as uint16 once per CTA.
// ---------------------------------------------------------------------------------------------
template <typename Real>
static inline size_t gdbf_par_smem_bytes(const CodeDev &c)
{
    const size_t nwords = (size_t)(c.N + 31) / 32, mwords = (size_t)(c.M + 31) / 32;
    size_t n = 16 + sizeof(Real) * 3 * (size_t)c.N + 4 * (size_t)c.N + 4 * (2 * nwords + 2 * mwords) + 2 * (size_t)c.dv_max * c.N + 64;
    return (n + 15) & ~(size_t)15;
}

template <typename Real> struct alignas(16) GdbfPack4 { Real x[4]; };

template <typename Real>
__global__ void gdbf_par_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = c.N, M = c.M, nwords = (N + 31) >> 5, mwords = (M + 31) >> 5, nblk = (N + 3) >> 2, dvm = c.dv_max;
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);
    Real *yq = reinterpret_cast<Real *>(smem_raw + 16);
    Real *theta = yq + N, *shape = theta + N;
    int *dsum = reinterpret_cast<int *>(shape + N);
    uint32_t *dbits = reinterpret_cast<uint32_t *>(dsum + N);          // 1 <-> d = -1
    uint32_t *rbits = dbits + nwords;
    uint32_t *syn = rbits + nwords;                                     // 1 <-> s_j = -1 (unsatisfied)
    uint32_t *tog = syn + mwords;
    uint16_t *chk = reinterpret_cast<uint16_t *>(tog + mwords);         // [dv_max][N]
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, T = p.T, W = p.windowsize;
    const uint32_t fl = p.flags;
    const bool redecode = (fl & LDPC_GPU_F_REDECODE) != 0;
    const int maxphase = redecode ? (p.maxphase > 0 ? p.maxphase : 1) : 1;
    const Real theta0 = (Real)p.theta, lambda = (Real)p.lambda, noiseSigma = (Real)p.noiseSigma;
    // regular codes: the column weight and the syndrome weight are per-launch constants (the weight of -D redecode
    // is a double division per variable and iteration otherwise, src/RNGDBF.cpp:566)
    const int reg_dv = c.regular_dv;
    const Real wgt_reg = !(fl & LDPC_GPU_F_WEIGHTSYNDROMES) ? (Real)1
                         : (redecode ? (Real)(p.alpha * p.Ymax / (double)(reg_dv > 0 ? reg_dv : 1)) : (Real)p.alpha);
    const bool vec_ok = reg_dv > 0 && (N & 3) == 0 && (reinterpret_cast<size_t>(chk) & 7) == 0;
    CtaTotals tot; tot.clear();

    for (int e = tid; e < dvm * N; e += nt) chk[e] = (uint16_t)c.vn_chk[e];
    __syncthreads();

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        const unsigned long long fid = (unsigned long long)(io.frame_begin + f);
        if (tid == 0) { fs->uncoded = 0; fs->errors = 0; fs->flag = 0; }
        for (int w = tid; w < nwords; w += nt) rbits[w] = 0u;
        __syncthreads();
        // ---- channel front end: src/decodeGDBF.cpp:251-274 / src/RNGDBF.cpp:251-275
        int unc = 0;
        for (int b = tid; b < nblk; b += nt) {
            double y4[4];
            raw_samples4<false>(io, p, c, f, cw, b, y4);
            uint32_t rn = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                double v = y4[q];
                if (fl & LDPC_GPU_F_SATURATE_SAMPLES) if (fabs(v) > p.Ymax) v *= p.Ymax / fabs(v);
                const bool rneg = !(v > 0);
                if (fl & LDPC_GPU_F_QUANTIZE_SAMPLES) v = quantize_gdbf(v, p);
                yq[i] = (Real)v;
                unc += (int)(rneg != ((cw ? cw[i] : 0) != 0));
                rn |= (uint32_t)rneg << q;
                dsum[i] = 0; shape[i] = (Real)0; theta[i] = theta0;
            }
            if (rn) atomicOr(&rbits[(4 * b) >> 5], rn << ((4 * b) & 31));
        }
        for (int o = 16; o; o >>= 1) unc += __shfl_xor_sync(0xffffffffu, unc, o);
        if (lane == 0 && unc) atomicAdd(&fs->uncoded, unc);
        __syncthreads();

        int it = 0, total_it = 0, phase = 0, satisfied = 1, smoothed = 0, smoothing_used = 0;
        long long row = 0;
        while (phase < maxphase) {                                    // src/RNGDBF.cpp:280-400
            for (int w = tid; w < nwords; w += nt) dbits[w] = rbits[w];
            for (int w = tid; w < mwords; w += nt) tog[w] = 0u;
            if (redecode || phase == 0) for (int i = tid; i < N; i += nt) { dsum[i] = 0; if (fl & LDPC_GPU_F_THRESHOLDADAPTATION) theta[i] = theta0; }
            __syncthreads();
            for (int j0 = tid; j0 < (mwords << 5); j0 += nt) {        // full syndrome of the starting decisions
                unsigned par = 0;
                if (j0 < M) {
                    const int deg = c.cn_deg[j0];
                    for (int k = 0; k < deg; k++) { const uint32_t i = c.cn_var[(size_t)k * M + j0]; par ^= dbits[i >> 5] >> (i & 31); }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, (par & 1u) != 0);
                if (lane == 0) syn[j0 >> 5] = bal;
            }
            __syncthreads();
            for (it = 0; it < T; it++) {
                unsigned any = 0;
                for (int w = 0; w < mwords; w++) any |= syn[w];
                satisfied = (any == 0);
                if (satisfied) break;                                 // :305-306
                const long long row_pert = (fl & LDPC_GPU_F_ADDNOISE) ? row++ : -1;
                const long long row_unif = (fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) ? row++ : -1;
                const bool smooth_now = (fl & LDPC_GPU_F_OUTPUTSMOOTHING) && it > T - W;
                for (int b = tid; b < nblk; b += nt) {                // symNodeUpdates, :536-621, four variables per thread
                    double pert4[4] = {0, 0, 0, 0}, unif4[4] = {0, 0, 0, 0};
                    if (row_pert >= 0) {                              // :318-333
                        if (io.noise) {
#pragma unroll
                            for (int q = 0; q < 4; q++) if (4 * b + q < N) pert4[q] = io.noise[((size_t)f * io.noise_rows + row_pert) * N + 4 * b + q];
                        } else if (fl & LDPC_GPU_F_UNIFORMNOISE) uniform4(io.seed, fid, (uint32_t)b, (uint32_t)(row_pert + io.noise_row_base), STREAM_DECODER, pert4);
                        else { float n4[4]; normal4(io.seed, fid, (uint32_t)b, (uint32_t)(row_pert + io.noise_row_base), STREAM_DECODER, n4);
#pragma unroll
                               for (int q = 0; q < 4; q++) pert4[q] = (double)n4[q]; }
                    }
                    if (row_unif >= 0) {
                        if (io.noise) {
#pragma unroll
                            for (int q = 0; q < 4; q++) if (4 * b + q < N) unif4[q] = io.noise[((size_t)f * io.noise_rows + row_unif) * N + 4 * b + q];
                        } else uniform4(io.seed, fid, (uint32_t)b, (uint32_t)(row_unif + io.noise_row_base), STREAM_DECODER, unif4);
                    }
                    const uint32_t dword = dbits[(4 * b) >> 5];
                    uint32_t flipmask = 0;
                    if (vec_ok) {
                        // regular code, N % 4 == 0: the four variables of the thread move as vectors (one 16-byte access
                        // for y and theta, one 8-byte access per slot for the four check indices) instead of scalar accesses
                        // at a 16-byte lane stride (4-way bank conflicts: 42 % of the wavefronts were replays)
                        const int i0 = 4 * b;
                        const uint32_t d4 = (dword >> (i0 & 31)) & 15u;
                        const GdbfPack4<Real> y4v = *reinterpret_cast<const GdbfPack4<Real> *>(&yq[i0]);
                        GdbfPack4<Real> th4 = *reinterpret_cast<const GdbfPack4<Real> *>(&theta[i0]);
                        Real E[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) E[q] = ((d4 >> q) & 1u) ? -y4v.x[q] : y4v.x[q];     // d[i]*y[i]
                        for (int sl = 0; sl < reg_dv; sl++) {                                        // nlist order, per variable
                            const uint2 c4 = *reinterpret_cast<const uint2 *>(&chk[sl * N + i0]);
                            const int j4[4] = { (int)(c4.x & 0xffffu), (int)(c4.x >> 16), (int)(c4.y & 0xffffu), (int)(c4.y >> 16) };
#pragma unroll
                            for (int q = 0; q < 4; q++) E[q] += ((syn[j4[q] >> 5] >> (j4[q] & 31)) & 1u) ? -wgt_reg : wgt_reg;
                        }
                        bool th_dirty = false;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int i = i0 + q;
                            const bool dneg = (d4 >> q) & 1u;
                            if (row_pert >= 0) {
                                Real smp = (fl & LDPC_GPU_F_UNIFORMNOISE) ? (Real)(p.uni_scale * (pert4[q] - 0.5)) : (Real)(p.noiseSigma * pert4[q]);
                                if (fl & LDPC_GPU_F_NOISESHAPING) { const Real prev = shape[i]; shape[i] = smp; smp = smp - prev; }
                                E[q] += smp;
                            }
                            bool flip;
                            if (fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) {               // :561-597
                                const double val = ((double)(-E[q] + th4.x[q])) / (double)noiseSigma;
                                const double pcdf = 0.5 * erfc(-val * 0.70710678118654752440);
                                const double lv[8] = { 0, 0.0625, 0.125, 0.25, 0.34375, 0.4106, 0.68359, 1 };
                                double md = 1; int mi = 0;
#pragma unroll
                                for (int l = 0; l < 8; l++) { double t = lv[l] - pcdf; t = t * t; if (t < md) { md = t; mi = l; } }
                                flip = unif4[q] < lv[mi];
                            } else flip = E[q] < th4.x[q];                               // mu == 1
                            if (flip) {
                                flipmask |= 1u << q;
                                for (int sl = 0; sl < reg_dv; sl++) { const int j = chk[sl * N + i]; atomicXor(&tog[j >> 5], 1u << (j & 31)); }
                            } else if (fl & LDPC_GPU_F_THRESHOLDADAPTATION) { th4.x[q] *= lambda; th_dirty = true; }   // :612-617
                            if (smooth_now) dsum[i] += (dneg != flip) ? -1 : 1;                   // :348-354, d after the flip
                        }
                        if (th_dirty) *reinterpret_cast<GdbfPack4<Real> *>(&theta[i0]) = th4;
                        if (flipmask) atomicXor(&dbits[i0 >> 5], flipmask << (i0 & 31));
                        continue;
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int i = 4 * b + q;
                        if (i >= N) break;
                        const int deg = reg_dv > 0 ? reg_dv : (int)c.vn_deg[i];
                        const bool dneg = (dword >> ((4 * b + q) & 31)) & 1u;
                        Real E = dneg ? -yq[i] : yq[i];                              // d[i]*y[i]
                        Real wgt = wgt_reg;
                        if (reg_dv <= 0 && (fl & LDPC_GPU_F_WEIGHTSYNDROMES))
                            wgt = redecode ? (Real)(p.alpha * p.Ymax / (double)deg) : (Real)p.alpha;   // RNGDBF.cpp:566 / decodeGDBF.cpp:550
                        for (int sl = 0; sl < deg; sl++) { const int j = chk[sl * N + i]; E += ((syn[j >> 5] >> (j & 31)) & 1u) ? -wgt : wgt; }
                        if (row_pert >= 0) {
                            Real smp = (fl & LDPC_GPU_F_UNIFORMNOISE) ? (Real)(p.uni_scale * (pert4[q] - 0.5)) : (Real)(p.noiseSigma * pert4[q]);
                            if (fl & LDPC_GPU_F_NOISESHAPING) { const Real prev = shape[i]; shape[i] = smp; smp = smp - prev; }
                            E += smp;
                        }
                        bool flip;
                        if (fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) {               // :561-597
                            const double val = ((double)(-E + theta[i])) / (double)noiseSigma;
                            const double pcdf = 0.5 * erfc(-val * 0.70710678118654752440);
                            const double lv[8] = { 0, 0.0625, 0.125, 0.25, 0.34375, 0.4106, 0.68359, 1 };
                            double md = 1; int mi = 0;
#pragma unroll
                            for (int l = 0; l < 8; l++) { double t = lv[l] - pcdf; t = t * t; if (t < md) { md = t; mi = l; } }
                            flip = unif4[q] < lv[mi];
                        } else flip = E < theta[i];                                  // mu == 1
                        if (flip) {
                            flipmask |= 1u << q;
                            for (int sl = 0; sl < deg; sl++) { const int j = chk[sl * N + i]; atomicXor(&tog[j >> 5], 1u << (j & 31)); }
                        } else if (fl & LDPC_GPU_F_THRESHOLDADAPTATION) theta[i] *= lambda;   // :612-617
                        if (smooth_now) dsum[i] += (dneg != flip) ? -1 : 1;                   // :348-354, d after the flip
                    }
                    if (flipmask) atomicXor(&dbits[(4 * b) >> 5], flipmask << ((4 * b) & 31));
                }
                __syncthreads();
                for (int w = tid; w < mwords; w += nt) { syn[w] ^= tog[w]; tog[w] = 0u; }
                __syncthreads();
            }
            if (fl & LDPC_GPU_F_OUTPUTSMOOTHING) {
                if (!satisfied) {                                                     // :358-367
                    for (int w = tid; w < nwords; w += nt) dbits[w] = 0u;
                    __syncthreads();
                    for (int i = tid; i < N; i += nt) if (!(dsum[i] > 0)) atomicOr(&dbits[i >> 5], 1u << (i & 31));
                    smoothed = 1;
                } else smoothed = 0;
                if (it > T - W) smoothing_used++;                                     // :371-374
            }
            total_it += it; phase++;
            __syncthreads();
            if (!redecode || satisfied) break;                                        // src/RNGDBF.cpp:398-399
        }
        finish_frame(c, p, io, f, cw, dbits, fs, total_it, satisfied, smoothed, smoothing_used, phase, -1, tot);
    }
    if (tid == 0) tot.flush(io.counters);
}

// ---------------------------------------------------------------------------------------------
// NGDBFhw: src/NGDBFhw.cpp.  All-integer flip metric on 5-bit sign-magnitude samples; the 2648-entry
// per-frame noise buffer is read through a window that slides by one entry per iteration.
//
// Bit-packed state: decisions d (1 bit per variable, one word per warp of variables) and syndromes s
// (1 bit per check, M/32 words).  The reference recomputes every syndrome from d at the top of each
// iteration (checkNodeUpdates, :546-563); the syndrome is linear in d over GF(2), so here it is computed
// in full once per phase and then UPDATED: a variable that flips toggles the syndromes of its checks
// (atomicXor into a toggle mask that is folded in between iterations).  Same values, a few hundred
// toggles instead of E gathers per iteration.  The flip metric reads the syndromes of the iteration's
// start, as the reference does (its symNodeUpdates only reads `syndrome`, :565-593).
// Samples and noise are unpacked to their odd-integer values (+-(2k+1), unpack() :665-677) once per frame.
// ---------------------------------------------------------------------------------------------
static inline size_t hw_smem_bytes(const CodeDev &c)
{
    const size_t nwords = (size_t)(c.N + 31) / 32, mwords = (size_t)(c.M + 31) / 32;
    size_t n = 16 + 4 * (3 * nwords + 3 * mwords) + 32 * mwords + 2 * (size_t)c.dv_max * c.N + (size_t)c.N + LDPC_GPU_HW_QBUF + 64;
    return (n + 15) & ~(size_t)15;
}

// DV > 0: every variable has exactly DV checks (the slot loop unrolls); DV = 0: per-variable weights.
template <int DV>
__global__ void __launch_bounds__(384, 3) hw_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = c.N, M = c.M, nwords = (N + 31) >> 5, mwords = (M + 31) >> 5, npad = nwords << 5, nblk = (N + 3) >> 2;
    const int T = p.T, QB = LDPC_GPU_HW_QBUF, dvm = c.dv_max;
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);
    uint32_t *dbits = reinterpret_cast<uint32_t *>(smem_raw + 16);      // current decisions, 1 <-> d = 1
    uint32_t *rbits = dbits + nwords;                                   // received hard decisions
    uint32_t *cbits = rbits + nwords;                                   // codeword bits
    uint32_t *syn = cbits + nwords;                                     // syndrome bits, 1 <-> unsatisfied
    uint32_t *tog = syn + mwords;                                       // [2][mwords] toggles of the running iteration, by iteration parity
    uint8_t *synb = reinterpret_cast<uint8_t *>(tog + 2 * mwords);      // [32 mwords] the syndrome bits again, one byte each: the flip metric's operand
    uint16_t *chk = reinterpret_cast<uint16_t *>(synb + 32 * mwords);   // [dv_max][N] check of slot s of variable i
    signed char *yval = reinterpret_cast<signed char *>(chk + (size_t)dvm * N);   // unpack(y')   in [-31, 31], odd
    signed char *qval = yval + N;                                       // unpack(q')
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int maxPhases = p.maxphase > 0 ? p.maxphase : 1;
    CtaTotals tot; tot.clear();

    for (int e = tid; e < dvm * N; e += nt) chk[e] = (uint16_t)c.vn_chk[e];      // once per CTA: the graph, variable side
    __syncthreads();

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        const unsigned long long fid = (unsigned long long)(io.frame_begin + f);
        if (tid == 0) { fs->uncoded = 0; fs->errors = 0; fs->flag = 0; }
        for (int w = tid; w < nwords; w += nt) { rbits[w] = 0u; cbits[w] = 0u; }
        __syncthreads();
        int unc = 0;
        for (int b = tid; b < nblk; b += nt) {                        // :218-237
            double y4[4];
            raw_samples4<false>(io, p, c, f, cw, b, y4);
            uint32_t rn = 0, cn = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                double v = y4[q];
                if (fabs(v) > p.Ymax) v *= p.Ymax / fabs(v);
                const bool rneg = !(v > 0);
                const int cb = cw ? cw[i] : 0;
                unc += (int)(rneg && cb);                             // r*c < 0 with c in {0,1} (:141,230)
                rn |= (uint32_t)rneg << q; cn |= (uint32_t)(cb != 0) << q;
                yval[i] = (signed char)hw_unpack(hw_pack(v / p.hw_two_w, p));
            }
            if (rn) atomicOr(&rbits[(4 * b) >> 5], rn << ((4 * b) & 31));
            if (cn) atomicOr(&cbits[(4 * b) >> 5], cn << ((4 * b) & 31));
        }
        for (int b = tid; b < (QB + 3) / 4; b += nt) {                // :239-252
            double n4[4];
            if (io.noise) {
#pragma unroll
                for (int q = 0; q < 4; q++) n4[q] = (4 * b + q < QB) ? io.noise[(size_t)f * QB + 4 * b + q] : 0.0;
            } else { float nf[4]; normal4(io.seed, fid, (uint32_t)b, (uint32_t)io.noise_row_base, STREAM_DECODER, nf);
#pragma unroll
                     for (int q = 0; q < 4; q++) n4[q] = (double)nf[q]; }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= QB) break;
                const double qv = p.noiseSigma * n4[q];
                double qm = ((qv - p.theta0) / p.hw_two_w - 1.0);
                if (qm > p.hw_lmax) qm = p.hw_lmax; else if (qm < -p.hw_lmax) qm = -p.hw_lmax;
                qval[i] = (signed char)hw_unpack(hw_pack(qm, p));
            }
        }
        for (int o = 16; o; o >>= 1) unc += __shfl_xor_sync(0xffffffffu, unc, o);
        if (lane == 0 && unc) atomicAdd(&fs->uncoded, unc);
        __syncthreads();

        int qpointer = io.qpointer0 ? io.qpointer0[f] : 0;
        qpointer = ((qpointer % (QB - N)) + (QB - N)) % (QB - N);      // the window start lives in [0, QBUF - N) (:356-358); host batches are validated
        int leastIterations = T, leastErrors = N, satisfied = 1, it = 0;
        for (int phase = 0; phase < maxPhases; phase++) {             // :280-373
            for (int w = tid; w < nwords; w += nt) dbits[w] = rbits[w];
            for (int w = tid; w < 2 * mwords; w += nt) tog[w] = 0u;
            __syncthreads();
            // full syndrome of the starting decisions (checkNodeUpdates, :546-563), once per phase
            unsigned mine = 0;
            for (int j0 = tid; j0 < (mwords << 5); j0 += nt) {
                unsigned par = 0;
                if (j0 < M) {
                    const int deg = c.cn_deg[j0];
                    for (int k = 0; k < deg; k++) { const uint32_t i = c.cn_var[(size_t)k * M + j0]; par ^= dbits[i >> 5] >> (i & 31); }
                }
                par &= 1u; mine |= par;
                synb[j0] = (uint8_t)par;
                const unsigned bal = __ballot_sync(0xffffffffu, par != 0);
                if (lane == 0) syn[j0 >> 5] = bal;
            }
            int any = __syncthreads_or((int)mine);                    // some check unsatisfied (the same value in every thread)
            for (it = 0; it < T; it++) {
                satisfied = (any == 0);
                if (satisfied) break;                                 // :297-299
                uint32_t *togw = tog + (it & 1) * mwords;
                // symNodeUpdates :565-593, variables of a warp share one decision word.  Per variable: DV index loads, DV byte loads
                // (the r2bh capture of the bit-packed form: 147 warp-instructions per variable-iteration, a quarter of them the per-slot
                // shift / mask of the syndrome word and the runtime slot loop)
                for (int i0 = tid; i0 < npad; i0 += nt) {
                    const bool valid = i0 < N;
                    const uint32_t dw = dbits[i0 >> 5];
                    bool nd = (dw >> lane) & 1u;
                    if (valid) {
                        const int deg = DV > 0 ? DV : (int)c.vn_deg[i0];
                        const int d01 = (int)nd;
                        int unsat = 0;
                        if (DV > 0) {
#pragma unroll
                            for (int sl = 0; sl < (DV > 0 ? DV : 1); sl++) unsat += (int)synb[chk[sl * N + i0]];
                        } else
                            for (int sl = 0; sl < deg; sl++) unsat += (int)synb[chk[sl * N + i0]];
                        const int E = (1 - 2 * d01) * (int)yval[i0] + (deg - unsat) * p.hw_Smult + (int)qval[i0 + qpointer];
                        if (E <= p.hw_theta) {
                            nd = !nd;
                            for (int sl = 0; sl < deg; sl++) { const int j = chk[sl * N + i0]; atomicXor(&togw[j >> 5], 1u << (j & 31)); }
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, valid && nd);
                    if (lane == 0) dbits[i0 >> 5] = bal;
                }
                qpointer++;                                           // :356-358
                if (qpointer >= QB - N) qpointer = 0;
                __syncthreads();
                // fold the iteration's toggles into both syndrome forms; the other parity's toggle words (consumed one iteration ago) are cleared
                unsigned left = 0;
                for (int j0 = tid; j0 < (mwords << 5); j0 += nt) {
                    const unsigned b = (unsigned)synb[j0] ^ ((togw[j0 >> 5] >> (j0 & 31)) & 1u);
                    synb[j0] = (uint8_t)b; left |= b;
                }
                for (int w = tid; w < mwords; w += nt) { syn[w] ^= togw[w]; tog[((it + 1) & 1) * mwords + w] = 0u; }
                any = __syncthreads_or((int)left);
            }
            // countDecisionErrors against c in {0,1} (:362-372)
            int le = 0;
            for (int w = tid; w < nwords; w += nt) le += __popc(dbits[w] ^ cbits[w]);
            if (tid == 0) fs->errors = 0;
            __syncthreads();
            for (int o = 16; o; o >>= 1) le += __shfl_xor_sync(0xffffffffu, le, o);
            if (lane == 0 && le) atomicAdd(&fs->errors, le);
            __syncthreads();
            const int newErrors = fs->errors;
            if (newErrors < leastErrors) leastErrors = newErrors;
            if (it < leastIterations) leastIterations = it;
            __syncthreads();
        }
        if (tid == 0) fs->errors = 0;
        __syncthreads();
        finish_frame(c, p, io, f, cw, dbits, fs, leastIterations, satisfied, 0, 0, maxPhases, leastErrors, tot);
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
