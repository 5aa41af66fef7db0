// ldpc_sc_kernel.cuh -- LDPC_GPU_KIND_NGDBF_SC: the NGDBF decoder as the reference's SystemC model runs it
// (/root/reference/SystemC/NGDBF/inc/nodes.h:76-138 symnode, :167-201 checknode,
// inc/decoder.h:183-254 top level, inc/ldpcsim.h:85-119 quantiser).  Differences from C_implementations' decodeGDBF that
// this kernel reproduces:
//   * uniform Nq = 2^Q level quantiser that includes the end points +-Ymax (thresholds at the mid points), applied to the
//     channel samples, to every noise sample and to the local threshold at each comparison;
//   * per-node threshold adapted BOTH ways: theta /= lambda on a flip, theta *= lambda otherwise (nodes.h:121-129);
//   * syndrome weight w = alpha * Ymax / dv (nodes.h:59);
//   * ONE Gaussian per clock, quantised, shifted from node to node (decoder.h:187, nodes.h:107-111): node i at iteration t
//     reads the sample node i-1 read at iteration t-1.  Here the chain is a per-frame window q[0 .. N+T] of quantised samples
//     and node i reads q[(N-1-i) + (t-1)] (the model's chain runs on across frames; like NGDBFhw's carried qpointer that is a
//     sequential dependence between frames, so the window restarts per frame);
//   * smoothing over the decisions of the last W (= 32 in the model) iterations, applied whenever the iteration count reaches
//     T, also when the stop signal rises in that same clock (decoder.h:236-246);
//   * one clock of pipeline delay: the top level reads the decisions and stop signals the nodes wrote in the previous clock,
//     so the reported word is x_{t-1} where t is the clock in which `finished` rises; `it` = t - 1 flip steps.
// No runnable reference exists for this tree (SystemC is absent): parity is UNPINNED, the checker is the C
// restatement (sc_frame in the test oracle), which this kernel equals bit for bit (all arithmetic is IEEE double in the
// restatement's order; the library is built -fmad=false).
#pragma once
#include "ldpc_common.cuh"

namespace ldpc {

static inline size_t sc_smem_bytes(const CodeDev &c, int T, int Q)
{
    const size_t nwords = (size_t)(c.N + 31) / 32, mwords = (size_t)(c.M + 31) / 32;
    size_t n = 16 + 8 * (2 * (size_t)c.N + (size_t)c.N + T + 4) + 8 * 2 * ((size_t)1 << Q) + 4 * (size_t)c.N + 4 * (nwords + mwords) + 64;
    return (n + 15) & ~(size_t)15;
}

// quantize() of inc/ldpcsim.h:98-119: the value of the last threshold below Y (thresholds ascend, so a count)
LDPC_DEVINL double sc_quantize(const double *thr, const double *val, int Nq, double Y)
{
    int lo = 0, hi = Nq - 1;                      // number of thresholds with Y > thr[i]: thresholds are strictly increasing
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (Y > thr[mid]) lo = mid + 1; else hi = mid; }
    return val[lo];
}

__global__ void sc_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = c.N, M = c.M, nwords = (N + 31) >> 5, mwords = (M + 31) >> 5, npad = nwords << 5, nblk = (N + 3) >> 2;
    const int T = p.T, W = p.windowsize, Nq = 1 << p.Q, QL = N + T + 1, dvm = c.dv_max;
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);
    double *r = reinterpret_cast<double *>(smem_raw + 16);             // quantised channel values
    double *thl = r + N;                                               // local thresholds
    double *qn = thl + N;                                              // [N + T + 1 (+3)] quantised noise window
    double *thr = qn + (QL + 3);                                       // [Nq - 1] thresholds, [Nq] values
    double *val = thr + Nq;
    int *updown = reinterpret_cast<int *>(val + Nq);
    uint32_t *xbits = reinterpret_cast<uint32_t *>(updown + N);        // 1 <-> x = -1
    uint32_t *syn = xbits + nwords;                                    // 1 <-> check product -1
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const double lambda = p.lambda, step = 2.0 * p.Ymax / ((double)Nq - 1.0);
    CtaTotals tot; tot.clear();
    for (int i = tid; i < Nq; i += nt) {                               // initializeQuantization(), inc/ldpcsim.h:85-96
        if (i < Nq - 1) { thr[i] = -p.Ymax * ((double)Nq - 2.0) / ((double)Nq - 1.0) + (double)i * step; val[i] = -p.Ymax + (double)i * step; }
        else val[i] = p.Ymax;
    }
    __syncthreads();

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        const unsigned long long fid = (unsigned long long)(io.frame_begin + f);
        if (tid == 0) { fs->uncoded = 0; fs->errors = 0; fs->flag = 0; }
        for (int w = tid; w < nwords; w += nt) xbits[w] = 0u;
        __syncthreads();
        int unc = 0;
        for (int b = tid; b < nblk; b += nt) {                         // decoder.h:212, nodes.h:80-97 (reset behaviour)
            double y4[4];
            raw_samples4<false>(io, p, c, f, cw, b, y4);
            uint32_t xn = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                const double rv = sc_quantize(thr, val, Nq, y4[q]);
                const bool neg = !(rv > 0);
                r[i] = rv; thl[i] = p.theta; updown[i] = 0;
                unc += (int)(neg != ((cw ? cw[i] : 0) != 0));
                xn |= (uint32_t)neg << q;
            }
            if (xn) atomicOr(&xbits[(4 * b) >> 5], xn << ((4 * b) & 31));
        }
        for (int b = tid; b < (QL + 3) / 4; b += nt) {                 // decoder.h:187: quantize(sigma * rann()), one per clock
            double n4[4];
            if (io.noise) {
#pragma unroll
                for (int q = 0; q < 4; q++) n4[q] = (4 * b + q < QL) ? io.noise[(size_t)f * io.noise_rows + 4 * b + q] : 0.0;
            } else { float nf[4]; normal4(io.seed, fid, (uint32_t)b, (uint32_t)io.noise_row_base, STREAM_DECODER, nf);
#pragma unroll
                     for (int q = 0; q < 4; q++) n4[q] = (double)nf[q]; }
#pragma unroll
            for (int q = 0; q < 4; q++) qn[4 * b + q] = sc_quantize(thr, val, Nq, p.noiseSigma * n4[q]);
        }
        for (int o = 16; o; o >>= 1) unc += __shfl_xor_sync(0xffffffffu, unc, o);
        if (lane == 0 && unc) atomicAdd(&fs->uncoded, unc);
        __syncthreads();

        int t = 0, all_stop = 0;
        for (t = 1; ; t++) {
            // checknodes: products of the messages written in the previous clock (nodes.h:167-201)
            for (int j0 = tid; j0 < (mwords << 5); j0 += nt) {
                unsigned par = 0;
                if (j0 < M) {
                    const int deg = c.cn_deg[j0];
                    for (int k = 0; k < deg; k++) { const uint32_t i = c.cn_var[(size_t)k * M + j0]; par ^= xbits[i >> 5] >> (i & 31); }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, (par & 1u) != 0);
                if (lane == 0) syn[j0 >> 5] = bal;
            }
            __syncthreads();
            unsigned any = 0;
            for (int w = 0; w < mwords; w++) any |= syn[w];
            all_stop = (any == 0);
            // top level, clock t: smoothing window, stop rule (decoder.h:218-249)
            if (t > T - W) for (int i = tid; i < N; i += nt) updown[i] += ((xbits[i >> 5] >> (i & 31)) & 1u) ? -1 : 1;
            if (all_stop || t > T) break;
            // symnodes, clock t (nodes.h:104-137)
            for (int i0 = tid; i0 < npad; i0 += nt) {
                const bool valid = i0 < N;
                bool neg = (xbits[i0 >> 5] >> lane) & 1u;
                if (valid) {
                    const int deg = (int)c.vn_deg[i0];
                    const double wgt = p.alpha * p.Ymax / (double)deg;                  // nodes.h:59
                    double E = (neg ? -1.0 : 1.0) * r[i0] + qn[(N - 1 - i0) + (t - 1)];
                    for (int sl = 0; sl < deg; sl++) { const uint32_t j = c.vn_chk[(size_t)sl * N + i0]; E += wgt * (((syn[j >> 5] >> (j & 31)) & 1u) ? -1.0 : 1.0); }
                    if (E < sc_quantize(thr, val, Nq, thl[i0])) { thl[i0] = thl[i0] / lambda; neg = !neg; }
                    else thl[i0] = thl[i0] * lambda;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, valid && neg);
                __syncwarp();
                if (lane == 0) xbits[i0 >> 5] = bal;
            }
            __syncthreads();
        }
        __syncthreads();
        const int smoothed = t >= T;                                    // decoder.h:236
        if (smoothed) {
            for (int i0 = tid; i0 < npad; i0 += nt) {
                const unsigned bal = __ballot_sync(0xffffffffu, i0 < N && !(updown[i0] > 0));
                if (lane == 0) xbits[i0 >> 5] = bal;
            }
        }
        __syncthreads();
        finish_frame(c, p, io, f, cw, xbits, fs, t - 1, all_stop, smoothed, smoothed, 1, -1, tot);
    }
    if (tid == 0) tot.flush(io.counters);
    (void)dvm;
}

} // namespace ldpc
