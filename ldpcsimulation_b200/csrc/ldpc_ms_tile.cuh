// ldpc_ms_tile.cuh -- min-sum family for codes whose messages do not fit one SM (DVB-S2 N=64800,
// E=226799: 3.9 MB of fp32 message traffic per frame-iteration).  This path is HBM-bound, so the
// layout is chosen for the memory system rather than for shared memory:
//
//   one CTA owns a TILE of FI frames (FI * sizeof(Real) = LDPC_TILE_BYTES, one 128-byte line) and keeps
//       msg[(s*N + i) * FI + fl]     yq[i * FI + fl]          fl = frame lane inside the tile
//   in its slice of an HBM workspace.  A thread is (node, fl); a warp is 32/FI consecutive nodes x FI
//   frames.  A variable warp therefore touches 128 contiguous bytes per slot, and a check warp's
//   gather touches whole 32-byte sectors (the one-frame-per-CTA GSTATE kernel used 4 bytes of every
//   sector it fetched: profiles/r1_summary.md).  Algorithmic traffic per frame-iteration is the
//   minimum (4E + N) * sizeof(Real) bytes: every message read once and written once
//   per phase.  The row's indices are loaded once per thread as 16-byte vectors and are shared by the
//   FI lanes of the row (same address -> one transaction).
//
// Arithmetic and order are those of ms_fast_kernel / mp_kernel (the reference's); decisions, counters
// and a-posteriori sums are bit-identical to the single-frame kernels in both precisions.
#pragma once
#include "ldpc_ms_fast.cuh"
#include "ldpc_ms_x2.cuh"

namespace ldpc {

// Bytes of one (edge, tile) entry = FI * sizeof(Real), and bytes one thread moves per access (FPT frames).
// 128-byte entries (one full line: 32 fp32 / 16 fp64 frames per tile) moved as 16-byte vectors.  Measured on DVB-S2 fp32: 32-byte
// entries with 4-byte accesses 52 % of the HBM copy peak (all warps in long_scoreboard, 32 KB in flight per SM); 64-byte entries and
// 16-byte accesses 67 % (profiles/r1_summary.md); row weight / index words fetched one trip ahead 70 %; 128-byte entries 72 % (7.7 Gbit/s;
// needs the tile's decision words in the workspace instead of shared memory: FI x N bits; profiles/r2_summary.md).
#ifndef LDPC_TILE_BYTES
#define LDPC_TILE_BYTES 128
#endif
template <typename Real> struct TileFI { enum { value = LDPC_TILE_BYTES / sizeof(Real) }; };
template <typename Real, int W> struct alignas(sizeof(Real) * W) TilePack { Real x[W]; };

template <typename Real>
static inline size_t ms_tile_state_bytes(const CodeDev &c)
{
    const size_t FI = TileFI<Real>::value;
    return (((size_t)c.dvN + c.N) * sizeof(Real) * FI + 4 * FI * (size_t)((c.N + 31) / 32) + 255) & ~(size_t)255;   // messages, samples, decisions
}
template <typename Real>
static inline size_t ms_tile_smem_bytes(const CodeDev &c)
{
    const size_t FI = TileFI<Real>::value;
    return (16 * FI + 4 * (size_t)((c.N + 31) / 32) + 15) & ~(size_t)15;   // frame scratch + ONE frame's decision words (staging for finish_frame)
}

// DCMAX / DVMAX: compile-time bounds of the row / column weights (register arrays).
// VB: bytes one thread moves per access (VB / sizeof(Real) frame lanes per thread).
template <typename Real, typename IdxT, int DCMAX, int DVMAX, int NT_MAX, int VB>
__global__ void __launch_bounds__(NT_MAX, 1) ms_tile_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    constexpr int FI = TileFI<Real>::value, FPT = VB / (int)sizeof(Real), LPN = FI / FPT;     // LPN: lanes per node
    typedef TilePack<Real, FPT> PK;
    constexpr int VPL = IdxVec<IdxT>::VPL, NG = (DCMAX + VPL - 1) / VPL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);                       // [FI]
    const int N = c.N, M = c.M, nwords = (N + 31) >> 5, nblk = (N + 3) >> 2;
    Real *msg = reinterpret_cast<Real *>(io.workspace + (size_t)blockIdx.x * io.ws_stride);
    Real *yq = msg + (size_t)c.dvN * FI;
    // [FI][nwords] hard decisions: in the workspace too (written in the last iteration only; FI x N bits of shared memory would cap FI)
    uint32_t *dbits = reinterpret_cast<uint32_t *>(yq + (size_t)c.N * FI);
    uint32_t *dstage = reinterpret_cast<uint32_t *>(smem_raw + 16 * FI);                 // [nwords]
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint4 *cnv = reinterpret_cast<const uint4 *>(c.cn_pos);
    const Real INF = real_inf<Real>();
    const Real alpha = (Real)p.alpha, delta = (Real)p.delta;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    // redo launch of the exact-lattice kernel (ldpc_ms_tileh.cuh): frame q of the launch is frame frame_list[q] of the batch
    const long long n_launch = io.n_frames_dev ? min((long long)*io.n_frames_dev, io.n_frames) : io.n_frames;
    auto frame_of = [&](long long q) -> long long { return io.frame_list ? io.frame_list[q] : q; };
    const long long ntiles = (n_launch + FI - 1) / FI;
    CtaTotals tot; tot.clear();

    // SRC >= 0: the lean fp32 front ends of ldpc_ms_x2.cuh (same values as raw_samples4 + condition_ms_guarded, bit for bit; the
    // generic per-sample dispatch costs 215 lane-instructions per sample, twice the lean Philox body); SRC = -1: double conditioning
    const int src = ms_sample_source(io, p, N);
    auto front = [&](auto src_c, const long long f0) {
        constexpr int SRC = decltype(src_c)::value;
        for (int t = tid; t < nblk * FI; t += nt) {
            const int b = t / FI, fl = t % FI;
            const long long f = frame_of((f0 + fl < n_launch) ? f0 + fl : n_launch - 1);   // dead lanes replay the last frame, unreported
            const uint8_t *cw = codeword_row(io, c, f);
            double y4[4]; float vf4[4];
            if (SRC >= 0) {
                if (4 * b + 3 < N) ms_cond4_f32<(SRC >= 0 ? SRC : SRC_OTHER), true>(io, p, c, f, cw, b, qflags, fcond, vf4);
                else ms_cond4_f32<SRC_OTHER, true>(io, p, c, f, cw, b, qflags, fcond, vf4);   // ragged last block: the bounds-checked reader
            } else raw_samples4(io, p, c, f, cw, b, y4);
            uint32_t nib = 0; int unc = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                Real vr; bool rneg;
                if (SRC >= 0) {
                    vr = (Real)vf4[q]; rneg = !(vf4[q] > 0.0f);
                } else if (sizeof(Real) == 4 && fcond) {
                    const float vf = condition_ms_guarded(y4[q], p, qflags);
                    vr = (Real)vf; rneg = !(vf > 0.0f);
                } else {
                    double v = y4[q];
                    if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) v = quantize_ms(v, p);
                    if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) { if (v > p.Ymax) v = p.Ymax; if (v < -p.Ymax) v = -p.Ymax; }
                    rneg = !(v > 0); vr = (Real)v;
                }
                yq[(size_t)i * FI + fl] = vr;
                const int cb = cw ? cw[i] : 0;
                unc += (int)(rneg != (cb != 0));
                nib |= (uint32_t)rneg << q;
                const int deg = c.vn_deg[i];
                for (int s = 0; s < deg; s++) msg[((size_t)s * N + i) * FI + fl] = vr;
                if (io.out_soft && p.T == 0 && f0 + fl < n_launch) {
                    if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i] = (double)vr;
                    else ((float *)io.out_soft)[(size_t)f * N + i] = (float)vr;
                }
            }
            if (nib) atomicOr(&dbits[fl * nwords + ((4 * b) >> 5)], nib << ((4 * b) & 31));
            if (unc) atomicAdd(&fs[fl].uncoded, unc);
        }
    };

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long f0 = tile * FI;
        if (tid < FI) { fs[tid].uncoded = 0; fs[tid].errors = 0; fs[tid].flag = 0; }
        for (int w = tid; w < FI * nwords; w += nt) dbits[w] = 0u;
        __syncthreads();
        // ---- channel front end, (block of 4 samples, frame lane) per thread ---------------------
        if (sizeof(Real) == 4 && fcond && src == SRC_PHILOX) front(std::integral_constant<int, SRC_PHILOX>(), f0);   // the throughput entry's source
        else front(std::integral_constant<int, -1>(), f0);
        __syncthreads();

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            if (last) for (int w = tid; w < FI * nwords; w += nt) dbits[w] = 0u;
            // ---- check-node phase: (row, group of FPT frame lanes) per thread ----------------------------
            // fp32: the row's weight and index words are fetched one trip ahead: cn_deg -> index words -> gather is otherwise a chain
            // of two L2 latencies in front of every DRAM gather, with every warp waiting in long_scoreboard (r1j capture).  (The fp64
            // instantiation has no registers to spare for it: it spilled and lost 14 %.)
            constexpr bool AHEAD = sizeof(Real) == 4;
            int deg_n = 0; uint4 w_n[NG];
            if (AHEAD && tid < M * LPN) {
                const int j0 = tid / LPN;
                deg_n = c.cn_deg[j0];
#pragma unroll
                for (int g = 0; g < NG; g++) if (g * VPL < deg_n) w_n[g] = __ldg(&cnv[(size_t)g * M + j0]);
            }
            for (int t = tid; t < M * LPN; t += nt) {
                const int j = t / LPN, fl = (t % LPN) * FPT;
                const int deg = AHEAD ? deg_n : (int)c.cn_deg[j];
                uint4 w[NG];
#pragma unroll
                for (int g = 0; g < NG; g++) { if (AHEAD) w[g] = w_n[g]; else if (g * VPL < deg) w[g] = __ldg(&cnv[(size_t)g * M + j]); }
                if (AHEAD && t + nt < M * LPN) {
                    const int jn = (t + nt) / LPN;
                    deg_n = c.cn_deg[jn];
#pragma unroll
                    for (int g = 0; g < NG; g++) if (g * VPL < deg_n) w_n[g] = __ldg(&cnv[(size_t)g * M + jn]);
                }
                PK v[DCMAX];
#pragma unroll
                for (int k = 0; k < DCMAX; k++) if (k < deg) v[k] = *reinterpret_cast<const PK *>(&msg[(size_t)IdxVec<IdxT>::get(w[k / VPL], k % VPL) * FI + fl]);
                Real s1[FPT], s2[FPT], mm[FPT];
#pragma unroll
                for (int q = 0; q < FPT; q++) {
                    Real m1 = INF, m2 = INF;
                    typename SignOps<Real>::acc_t sg = SignOps<Real>::zero();
#pragma unroll
                    for (int k = 0; k < DCMAX; k++) if (k < deg) {
                        const Real a = absr(v[k].x[q]);
                        m2 = rmin(m2, rmax(m1, a)); m1 = rmin(m1, a);
                        SignOps<Real>::fold(sg, v[k].x[q]);
                    }
                    Real o1 = m1, o2 = m2;
                    if (normalized) {
                        o1 = o1 / alpha; o2 = o2 / alpha;                       // IEEE division in every precision, once per row (src/decodeMinSum.cpp:494-499)
                    }
                    if (offset) { o1 = o1 - delta; o1 = (o1 > 0) ? o1 : (Real)0; o2 = o2 - delta; o2 = (o2 > 0) ? o2 : (Real)0; }
                    s1[q] = SignOps<Real>::presign(o1, sg); s2[q] = SignOps<Real>::presign(o2, sg); mm[q] = m1;
                }
#pragma unroll
                for (int k = 0; k < DCMAX; k++) if (k < deg) {
                    PK o;
#pragma unroll
                    for (int q = 0; q < FPT; q++) {
                        const Real sel = (absr(v[k].x[q]) == mm[q]) ? s2[q] : s1[q];
                        o.x[q] = SignOps<Real>::apply(sel, v[k].x[q]);
                    }
                    *reinterpret_cast<PK *>(&msg[(size_t)IdxVec<IdxT>::get(w[k / VPL], k % VPL) * FI + fl]) = o;
                }
            }
            __syncthreads();
            // ---- variable-node phase: (variable, group of FPT frame lanes) per thread ---------------------
            // (two variables per trip was measured slower: 128 registers are not enough for both, 6.3 vs 7.1 Gbit/s)
            int vdeg_n = (AHEAD && tid < N * LPN) ? (int)c.vn_deg[tid / LPN] : 0;     // one trip ahead, as above
            for (int t = tid; t < N * LPN; t += nt) {
                const int i = t / LPN, fl = (t % LPN) * FPT;
                const int deg = AHEAD ? vdeg_n : (int)c.vn_deg[i];
                if (AHEAD && t + nt < N * LPN) vdeg_n = c.vn_deg[(t + nt) / LPN];
                PK cm[DVMAX];
                PK sum = *reinterpret_cast<const PK *>(&yq[(size_t)i * FI + fl]);
#pragma unroll
                for (int s = 0; s < DVMAX; s++) if (s < deg) cm[s] = *reinterpret_cast<const PK *>(&msg[((size_t)s * N + i) * FI + fl]);
#pragma unroll
                for (int s = 0; s < DVMAX; s++) if (s < deg) {                             // nlist order
#pragma unroll
                    for (int q = 0; q < FPT; q++) sum.x[q] += cm[s].x[q];
                }
#pragma unroll
                for (int s = 0; s < DVMAX; s++) if (s < deg) {
                    PK o;
#pragma unroll
                    for (int q = 0; q < FPT; q++) o.x[q] = sum.x[q] - cm[s].x[q];
                    *reinterpret_cast<PK *>(&msg[((size_t)s * N + i) * FI + fl]) = o;
                }
                if (last) {
#pragma unroll
                    for (int q = 0; q < FPT; q++) {
                        if (!(sum.x[q] > 0)) atomicOr(&dbits[(fl + q) * nwords + (i >> 5)], 1u << (i & 31));
                        if (io.out_soft && f0 + fl + q < n_launch) {
                            const size_t fo = (size_t)frame_of(f0 + fl + q);
                            if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[fo * N + i] = (double)sum.x[q];
                            else ((float *)io.out_soft)[fo * N + i] = (float)sum.x[q];
                        }
                    }
                }
            }
            __syncthreads();
        }
        for (int fl = 0; fl < FI; fl++) {
            if (f0 + fl >= n_launch) break;                          // uniform: dead lanes are not reported
            const long long f = frame_of(f0 + fl);
            const uint8_t *cw = codeword_row(io, c, f);
            for (int w = tid; w < nwords; w += nt) dstage[w] = __ldcg(&dbits[fl * nwords + w]);   // (set by L2 atomics: read at L2)
            __syncthreads();
            finish_frame(c, p, io, f, cw, dstage, &fs[fl], p.T, -1, 0, 0, 1, -1, tot);
        }
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
