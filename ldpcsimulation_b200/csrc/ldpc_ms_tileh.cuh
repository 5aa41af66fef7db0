// ldpc_ms_tileh.cuh -- the HBM-bound min-sum path (ldpc_ms_tile.cuh) on the exact lattice: plain / offset min-sum on quantised
// samples whose quantiser unit u = Ymax / (2^Q - 1) is a power of two (decodeOffsetMinSum's macro set with Ymax = 1.9375, Q = 5,
// delta = 0.125: every sample, message and sum is a multiple of 1/16).  While no check-to-variable magnitude exceeds the cap C with
// Ymax + dv_max C <= 2047 u, every value the decoder forms is an integer multiple of u below 2^11 u, binary16 holds it exactly, and
// every addition, subtraction, minimum and comparison below returns what the reference's doubles return (src/decodeMinSum.cpp:
// 417-524, 534-580): decisions, iteration counts and counters are those of the fp64 parity instantiation bit for bit.  A frame in
// which some |c2v| exceeded C is not reported by this kernel: its index goes to io.redo_list and the fp64 tile kernel decodes it
// (host: redo_after_sync / redo_pending, as for ldpc_ms_x2.cuh).  At the DVB-S2 operating points the largest |c2v| after ten
// iterations is about 6 against C = 15.75, so the list stays empty.
//
// What it buys: this path moves (4E + N) message words per frame-iteration through HBM and nothing else; 2-byte words halve the
// algorithmic bytes of the fp32 tile kernel (DVB-S2: 1.94 MB instead of 3.89 MB per frame-iteration).
//
// Layout: one CTA owns a tile of 64 frames; (edge, tile) and (variable, tile) entries are 64 binary16 = one 128-byte line,
//       msg[(s*N + i) * 64 + fl]     yq[i * 64 + fl]     dT[i * 8 + fl / 8]  (decision bits of the last iteration, frame-major bytes)
// in the CTA's slice of the workspace.  A thread is (node, 8 frame lanes) and moves 16 bytes per access as four half2 words;
// check rows keep (min1 with the running sign product riding on its sign bit, min2) per lane pair: min.xorsign.abs + two HMNMX2
// per edge and lane pair, HSET2.BF + HFMA2 + LOP3 for the second pass (the x2 kernel's arithmetic).
#pragma once
#include "ldpc_ms_tile.cuh"
#include "ldpc_ms_x2.cuh"

namespace ldpc {

enum { TILEH_FI = 64, TILEH_LPN = 8 };

static inline size_t ms_tileh_state_bytes(const CodeDev &c)
{
    return (((size_t)c.dvN + c.N) * 2 * TILEH_FI + 8 * (size_t)c.N + 4 * (size_t)TILEH_FI * ((c.N + 31) / 32) + 255) & ~(size_t)255;   // messages, samples, decision bytes, decision words
}
static inline size_t ms_tileh_smem_bytes(const CodeDev &c)
{
    return (16 * (size_t)TILEH_FI + 16 + 4 * (size_t)((c.N + 31) / 32) + 15) & ~(size_t)15;     // frame scratch, cap flags, ONE frame's decision words
}

// decision bits of eight frame lanes: bit e = 1 <-> !(sum_e > 0)   (d = -1, src/decodeMinSum.cpp:571-574)
LDPC_DEVINL uint32_t tileh_decision_byte(const uint32_t (&s)[4])
{
    uint32_t b = 0;
#pragma unroll
    for (int h = 0; h < 4; h++) {
        const uint32_t nb = ~__hgt2_mask(h2_from(s[h]), h2_from(0u));
        b |= ((nb & 1u) | ((nb >> 15) & 2u)) << (2 * h);
    }
    return b;
}

LDPC_DEVINL void tileh_cp_async16(void *smem, const void *g)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory");
}
LDPC_DEVINL void tileh_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND> LDPC_DEVINL void tileh_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(NPEND) : "memory"); }

// shared-memory bytes of the two staging buffers ((1 + DVMAX) 16-byte entries per thread and stage)
static inline size_t ms_tileh_pipe_bytes(int nt, int dvmax) { return (size_t)2 * (1 + dvmax) * nt * 16; }

// The NEXT trip's entries of a thread are copied asynchronously (cp.async, 16 bytes per thread and entry, L2 -> shared memory, no
// destination registers) into the thread's own slots of a two-stage shared-memory buffer while the current trip is computed from the
// other stage: twice the bytes in flight per SM, and the row is re-read from shared memory for the second pass instead of living in 32
// registers.  Every thread reads only the slots it filled itself, so cp.async.wait_group is the only synchronisation.  Measured against
// the register-staged form (one row of 8 x 16 bytes in registers, 124 registers, `prefetch.global.L2` of the next trip): 6.58 instead of
// 7.25 ms per iteration of 18 944 DVB-S2 frames (profiles/r2_summary.md section 7).
template <typename IdxT, int DCMAX, int DVMAX, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, 1) ms_tileh_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    constexpr int FI = TILEH_FI, LPN = TILEH_LPN;
    constexpr int VPL = IdxVec<IdxT>::VPL, NG = (DCMAX + VPL - 1) / VPL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);                       // [FI]
    uint32_t *capf = reinterpret_cast<uint32_t *>(smem_raw + 16 * FI);                   // [2]: bit fl = some |c2v| of frame lane fl exceeded the cap
    uint32_t *dstage = capf + 4;                                                         // [nwords]
    static_assert(DCMAX <= 1 + DVMAX, "the staging slots are sized for the variable phase");
    uint4 *sbuf = reinterpret_cast<uint4 *>(smem_raw + ((16 * (size_t)TILEH_FI + 16 + 4 * (size_t)((c.N + 31) / 32) + 15) & ~(size_t)15));   // [2][1 + DVMAX][nt] staging slots, behind ms_tileh_smem_bytes()
    const int N = c.N, M = c.M, nwords = (N + 31) >> 5, nblk = (N + 3) >> 2;
    __half *msg = reinterpret_cast<__half *>(io.workspace + (size_t)blockIdx.x * io.ws_stride);
    __half *yq = msg + (size_t)c.dvN * FI;
    uint8_t *dT = reinterpret_cast<uint8_t *>(yq + (size_t)N * FI);                      // [N][8]
    uint32_t *dbits = reinterpret_cast<uint32_t *>(dT + 8 * (size_t)N);                  // [FI][nwords]: dT transposed at the end of the tile
    uint4 *msgv = reinterpret_cast<uint4 *>(msg);                                        // entry e, lane group lg: msgv[e * 8 + lg]
    const uint4 *yqv = reinterpret_cast<const uint4 *>(yq);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, lgc = tid & (LPN - 1);   // nt is a multiple of 32: a thread keeps its lane group
    const uint4 *cnv = reinterpret_cast<const uint4 *>(c.cn_pos);
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    const long long ntiles = (io.n_frames + FI - 1) / FI;
    const uint32_t INF2 = 0x7bff7bffu;
    const __half2 zero2 = h2_from(0u), delta2 = h2_from(p.x2_delta2), cap2 = h2_from(p.x2_cap2);
    CtaTotals tot; tot.clear();

    // decisions (and the a-posteriori sums, when asked for) of (variable i, lane group lg)
    auto emit = [&](const int i, const int lg, const uint32_t (&s)[4], const long long f0) {
        dT[(size_t)i * 8 + lg] = (uint8_t)tileh_decision_byte(s);
        if (io.out_soft) {
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const long long f = f0 + lg * 8 + e;
                if (f >= io.n_frames) break;
                const __half2 w = h2_from(s[e >> 1]);
                const float x = (e & 1) ? __high2float(w) : __low2float(w);
                if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i] = (double)x;
                else ((float *)io.out_soft)[(size_t)f * N + i] = x;
            }
        }
    };

    const int src = ms_sample_source(io, p, N);
    auto front = [&](auto src_c, const long long f0) {
        constexpr int SRC = decltype(src_c)::value;
        constexpr bool HASCW = true;                                                     // (cw is NULL for the all-zero codeword)
        for (int t = tid; t < nblk * FI; t += nt) {
            const int b = t / FI, fl = t % FI;
            const long long f = (f0 + fl < io.n_frames) ? f0 + fl : io.n_frames - 1;     // dead lanes replay the last frame, unreported
            const uint8_t *cw = HASCW ? codeword_row(io, c, f) : nullptr;
            float vf[4];
            if (4 * b + 3 < N) ms_cond4_f32<SRC, HASCW>(io, p, c, f, cw, b, qflags, fcond, vf);
            else ms_cond4_f32<SRC_OTHER, HASCW>(io, p, c, f, cw, b, qflags, fcond, vf);      // ragged last block: the bounds-checked reader
            int unc = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                const __half vh = __float2half_rn(vf[q]);                                // exact: an odd multiple of u up to Ymax
                yq[(size_t)i * FI + fl] = vh;
                const int cb = (HASCW && cw) ? cw[i] : 0;
                unc += (int)(!(vf[q] > 0.0f) != (cb != 0));
                const int deg = c.vn_deg[i];
                for (int s = 0; s < deg; s++) msg[((size_t)s * N + i) * FI + fl] = vh;
            }
            if (unc) atomicAdd(&fs[fl].uncoded, unc);
        }
    };

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long f0 = tile * FI;
        if (tid < FI) { fs[tid].uncoded = 0; fs[tid].errors = 0; fs[tid].flag = 0; }
        if (tid < 2) capf[tid] = 0u;
        uint32_t cacc[4] = { 0u, 0u, 0u, 0u };
        __syncthreads();
        // ---- channel front end, (block of 4 samples, frame lane) per thread; one straight-line body per sample source (the
        // generic raw_samples4 dispatch costs 215 lane-instructions per sample here, 13 % of the kernel: r2bd capture) ----
        switch (src) {
        case SRC_PHILOX:      front(std::integral_constant<int, SRC_PHILOX>(), f0); break;
        case SRC_PHILOX_FAST: front(std::integral_constant<int, SRC_PHILOX_FAST>(), f0); break;
        case SRC_Q8:          front(std::integral_constant<int, SRC_Q8>(), f0); break;
        case SRC_QP:          front(std::integral_constant<int, SRC_QP>(), f0); break;
        default:              front(std::integral_constant<int, SRC_OTHER>(), f0); break;
        }
        __syncthreads();
        if (p.T == 0) {
            for (int t = tid; t < N * LPN; t += nt) {
                const int i = t / LPN, lg = t % LPN;
                const uint4 y = yqv[(size_t)i * 8 + lg];
                const uint32_t s[4] = { y.x, y.y, y.z, y.w };
                emit(i, lg, s, f0);
            }
            __syncthreads();
        }

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            // ---- check-node phase: (row, lane group) per thread; weight and index words one trip ahead (ldpc_ms_tile.cuh) ----
            {
            const int lim = M * LPN;
            auto slot = [&](const int st, const int k) -> uint4 * { return sbuf + ((size_t)(st * (1 + DVMAX) + k) * nt + tid); };
            auto load_idx = [&](const int t, int &deg, uint4 (&w)[NG]) {
                const int j = t / LPN;
                deg = c.cn_deg[j];
#pragma unroll
                for (int g = 0; g < NG; g++) if (g * VPL < deg) w[g] = __ldg(&cnv[(size_t)g * M + j]);
            };
            auto issue = [&](const int st, const int deg, const uint4 (&w)[NG]) {
#pragma unroll
                for (int k = 0; k < DCMAX; k++) if (k < deg) tileh_cp_async16(slot(st, k), &msgv[(size_t)IdxVec<IdxT>::get(w[k / VPL], k % VPL) * 8 + lgc]);
            };
            int deg_c = 0, deg_n = 0; uint4 w_c[NG], w_n[NG];
            if (tid < lim) { load_idx(tid, deg_c, w_c); issue(0, deg_c, w_c); }
            tileh_cp_async_commit();
            if (tid + nt < lim) load_idx(tid + nt, deg_n, w_n);
            int st = 0;
            for (int t = tid; t < lim; t += nt, st ^= 1) {
                if (t + nt < lim) issue(st ^ 1, deg_n, w_n);
                tileh_cp_async_commit();
                int deg_nn = 0; uint4 w_nn[NG];
                if (t + 2 * nt < lim) load_idx(t + 2 * nt, deg_nn, w_nn);
                tileh_cp_async_wait<1>();                                            // this trip's entries have landed
                const int deg = deg_c;
                uint32_t m1[4] = { INF2, INF2, INF2, INF2 }, m2[4] = { INF2, INF2, INF2, INF2 };
#pragma unroll
                for (int k = 0; k < DCMAX; k++) if (k < deg) {
                    const uint4 x = *slot(st, k);
                    const uint32_t xv[4] = { x.x, x.y, x.z, x.w };
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const __half2 hi = __hmax2(__habs2(h2_from(m1[h])), __habs2(h2_from(xv[h])));
                        m2[h] = h2_bits(__hmin2(h2_from(m2[h]), hi));
                        m1[h] = h2_min_xorsign_abs(m1[h], xv[h]);
                    }
                }
                uint32_t m1a[4], s1[4]; __half2 ds[4];
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const uint32_t sg = m1[h] & 0x80008000u;
                    m1a[h] = m1[h] & 0x7fff7fffu;
                    const __half2 t1 = __hmax2(__hsub2(h2_from(m1a[h]), delta2), zero2);
                    const __half2 t2 = __hmax2(__hsub2(h2_from(m2[h]), delta2), zero2);
                    cacc[h] |= __hgt2_mask(t2, cap2);
                    s1[h] = h2_bits(t1) ^ sg;
                    ds[h] = __hsub2(h2_from(h2_bits(t2) ^ sg), h2_from(s1[h]));
                }
#pragma unroll
                for (int k = 0; k < DCMAX; k++) if (k < deg) {
                    const uint4 x = *slot(st, k);
                    const uint32_t xv[4] = { x.x, x.y, x.z, x.w };
                    uint32_t oo[4];
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const __half2 eq = __heq2(__habs2(h2_from(xv[h])), h2_from(m1a[h]));
                        oo[h] = x2_and_xor(xv[h], 0x80008000u, h2_bits(__hfma2(eq, ds[h], h2_from(s1[h]))));
                    }
                    msgv[(size_t)IdxVec<IdxT>::get(w_c[k / VPL], k % VPL) * 8 + lgc] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
                }
                deg_c = deg_n; deg_n = deg_nn;
#pragma unroll
                for (int g = 0; g < NG; g++) { w_c[g] = w_n[g]; w_n[g] = w_nn[g]; }
            }
            tileh_cp_async_wait<0>();
            }
            __syncthreads();
            // ---- variable-node phase: (variable, lane group) per thread ------------------------------
            {
            const int lim = N * LPN;
            auto slot = [&](const int st, const int k) -> uint4 * { return sbuf + ((size_t)(st * (1 + DVMAX) + k) * nt + tid); };
            auto issue = [&](const int st, const int t, const int deg) {
                const int i = t / LPN;
                tileh_cp_async16(slot(st, 0), &yqv[(size_t)i * 8 + lgc]);
#pragma unroll
                for (int sl = 0; sl < DVMAX; sl++) if (sl < deg) tileh_cp_async16(slot(st, 1 + sl), &msgv[((size_t)sl * N + i) * 8 + lgc]);
            };
            int vdeg_c = (tid < lim) ? (int)c.vn_deg[tid / LPN] : 0;
            if (tid < lim) issue(0, tid, vdeg_c);
            tileh_cp_async_commit();
            int vdeg_n = (tid + nt < lim) ? (int)c.vn_deg[(tid + nt) / LPN] : 0;
            int st = 0;
            for (int t = tid; t < lim; t += nt, st ^= 1) {
                if (t + nt < lim) issue(st ^ 1, t + nt, vdeg_n);
                tileh_cp_async_commit();
                const int vdeg_nn = (t + 2 * nt < lim) ? (int)c.vn_deg[(t + 2 * nt) / LPN] : 0;
                tileh_cp_async_wait<1>();
                const int i = t / LPN, deg = vdeg_c;
                const uint4 y = *slot(st, 0);
                uint32_t sum[4] = { y.x, y.y, y.z, y.w };
#pragma unroll
                for (int sl = 0; sl < DVMAX; sl++) if (sl < deg) {                     // nlist order (exact here: any order gives the same sum)
                    const uint4 x = *slot(st, 1 + sl);
                    sum[0] = h2_bits(__hadd2(h2_from(sum[0]), h2_from(x.x))); sum[1] = h2_bits(__hadd2(h2_from(sum[1]), h2_from(x.y)));
                    sum[2] = h2_bits(__hadd2(h2_from(sum[2]), h2_from(x.z))); sum[3] = h2_bits(__hadd2(h2_from(sum[3]), h2_from(x.w)));
                }
#pragma unroll
                for (int sl = 0; sl < DVMAX; sl++) if (sl < deg) {
                    const uint4 x = *slot(st, 1 + sl);
                    uint4 o;
                    o.x = h2_bits(__hsub2(h2_from(sum[0]), h2_from(x.x))); o.y = h2_bits(__hsub2(h2_from(sum[1]), h2_from(x.y)));
                    o.z = h2_bits(__hsub2(h2_from(sum[2]), h2_from(x.z))); o.w = h2_bits(__hsub2(h2_from(sum[3]), h2_from(x.w)));
                    msgv[((size_t)sl * N + i) * 8 + lgc] = o;
                }
                if (last) emit(i, lgc, sum, f0);
                vdeg_c = vdeg_n; vdeg_n = vdeg_nn;
            }
            tileh_cp_async_wait<0>();
            }
            __syncthreads();
        }
        // ---- frames whose messages left the certified range go to the fp64 decoder -------------------
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const int fl = lgc * 8 + 2 * h;
            if (cacc[h] & 0x0000ffffu) atomicOr(&capf[fl >> 5], 1u << (fl & 31));
            if (cacc[h] & 0xffff0000u) atomicOr(&capf[(fl + 1) >> 5], 1u << ((fl + 1) & 31));
        }
        __syncthreads();
        // bit transpose, once per tile: the 64 frame bits of 32 variables (one coalesced 8-byte load per lane) -> word w of every
        // frame lane's decision vector (a per-frame byte gather was 6 % of the kernel, all of it L2 latency: r2bd capture)
        for (int w = tid >> 5; w < nwords; w += nt >> 5) {
            const int i = 32 * w + lane;
            uint2 x = make_uint2(0u, 0u);
            if (i < N) x = *reinterpret_cast<const uint2 *>(dT + (size_t)i * 8);
            uint32_t keep_lo = 0u, keep_hi = 0u;
#pragma unroll
            for (int fl = 0; fl < 32; fl++) {
                const uint32_t wl = __ballot_sync(0xffffffffu, (x.x >> fl) & 1u), wh = __ballot_sync(0xffffffffu, (x.y >> fl) & 1u);
                if (lane == fl) { keep_lo = wl; keep_hi = wh; }
            }
            dbits[(size_t)lane * nwords + w] = keep_lo; dbits[(size_t)(lane + 32) * nwords + w] = keep_hi;
        }
        __syncthreads();
        for (int fl = 0; fl < FI; fl++) {
            if (f0 + fl >= io.n_frames) break;                       // uniform: dead lanes are not reported
            if ((capf[fl >> 5] >> (fl & 31)) & 1u) {                 // uniform
                if (tid == 0) { const unsigned q = atomicAdd(io.redo_count, 1u); io.redo_list[q] = f0 + fl; atomicAdd(io.redo_total, 1ull); }
                continue;
            }
            const uint8_t *cw = codeword_row(io, c, f0 + fl);
            for (int w = tid; w < nwords; w += nt) dstage[w] = dbits[(size_t)fl * nwords + w];
            __syncthreads();
            finish_frame(c, p, io, f0 + fl, cw, dstage, &fs[fl], p.T, -1, 0, 0, 1, -1, tot);
        }
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
