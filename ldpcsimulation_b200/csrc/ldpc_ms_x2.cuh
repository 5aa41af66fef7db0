// ldpc_ms_x2.cuh -- min-sum / offset min-sum on an EXACT LATTICE, two frames per lane, decisions certified
// identical to the reference's (LDPC_GPU_PREC_F16X2 when the configuration admits it; else ms_h2rc_kernel).
//
// Lattice.  With `quantizeSamples` the channel values are k * step, step = 2 Ymax / (2^Q - 1), or +-Ymax
// (quantize(), src/decodeMinSum.cpp:480-489).  When u = step / 2 is a power of two (e.g. Ymax = 1.9375, Q = 5:
// u = 1/16) and delta is a multiple of u, every message and every sum of plain / offset min-sum
// (src/decodeMinSum.cpp:410-476, :503-515) is an integer multiple of u, in the reference's doubles as well.
// binary16 holds the integers up to 2048 exactly, so as long as every |value| <= 2047 u the packed-half
// arithmetic below (HADD2 / HMNMX2 on the same edge of two frames) IS the reference's arithmetic: no rounding
// happens anywhere, the operation order is irrelevant.
//
// Growth.  The reference's messages are unbounded: once a frame has converged they grow about (dv - 1)-fold per
// iteration (measured: 1e5 u and more at T = 10), while frames that have not converged stay small (max |c2v| about
// 22 u at 3.6 - 4.0 dB on the 802.3an code).  The kernel caps c2v magnitudes at C (`cap`), chosen so that
// |S| <= Ymax + dv C <= 2047 u, and PROVES per frame that the cap cannot have changed a decision:
//
//   Lemma (stable state).  Let x be a codeword (as +-1), dvm = min column weight >= 3, and suppose that before a
//   check phase every v2c on an edge of variable i has sign x_i and magnitude >= m, with
//   m >= m0 = (Ymax + (dvm - 1) delta) / (dvm - 2).  Then every c2v is x_i max(min_others - delta, 0) (min-sum with
//   all parities satisfied), of magnitude >= m - delta, so the next v2c = y_i + sum_{others} c2v has sign x_i and
//   magnitude >= (dvm - 1)(m - delta) - Ymax >= m, and S_i = y_i + sum c2v has sign x_i.  By induction every later
//   decision vector equals x.  The same induction holds when c2v magnitudes are capped at any C with
//   (dvm - 1) C - Ymax >= m0 (the bound becomes min(m, (dvm - 1) C - Ymax) >= m0).  So from a stable state on, the
//   capped and the uncapped decoder output the same decisions, x, at every iteration.
//
//   Certificate.  While the cap has never engaged the frame's arithmetic is exact (state EXACT).  The lemma's premise for
//   the inputs of check phase t+1 is verified in two cheap pieces: (a) the variable phase of iteration t (only when every row of
//   check phase t was satisfied with min |v2c| >= m0, i.e. when the frame looks converged; one extra pass over the variable's
//   c2v) checks that every v2c = S - c2v it implies has the sign of S and magnitude >= m0; (b) check phase t+1 reports, from
//   values it computes anyway, that every row has an even number of negative v2c (so sign(S) is a codeword) and min |v2c| >= m0.
//   (a) and (b) together are the premise, and the frame becomes CERTIFIED; from then on the cap is harmless.  If the cap engages
//   in a check phase whose inputs are not certified, the frame is UNCERTIFIED: it is not reported from here but appended to a
//   redo list, and the host library decodes the listed frames with the fp64 parity instantiation (ms_rc_kernel<double>) on the
//   same stream.  Either way every reported decision vector is the reference's, bit for bit; a-posteriori sums (out_soft) are
//   the reference's only while the cap has not engaged (they saturate near Ymax + dv C afterwards) -- callers that need the
//   sums of converged frames use fp32 / fp64.
//
// Kernel structure: ms_h2rc_kernel's (one CTA per frame pair, one thread per check row, the row's 32 c2v words
// resident in registers, a-posteriori sums published through shared memory), with the check row rewritten around the
// per-row constants: o1 / o2 (offset and cap applied once per row), second pass = HSET2 + 2 LOP3 per edge.
#pragma once
#include "ldpc_ms_h2rc.cuh"

namespace ldpc {

enum { X2_EXACT = 0, X2_CERT = 2, X2_UNCERT = 3 };

// lattice constants derived on the host (ldpc_gpu.cu: x2_lattice); all in LLR units (multiples of u)
// (binary16 pairs, both halves equal; kept in DecParams = the constant bank, so they cost no registers)

static inline size_t ms_x2_smem_bytes(const CodeDev &c)
{
    return (((size_t)64 + 4 * ((size_t)c.dvN + 2 * (size_t)c.N) + 8 * (size_t)((c.N + 31) / 32) + 15) & ~(size_t)15) + 2 * (size_t)c.N + 16;
}                                                                                 // ... + staging of two frames' packed levels (Q <= 8) + the mbarrier

// schedule word of step group g of row j.  The second pass re-reads the row's words through the `volatile` form: ptxas
// otherwise keeps the first pass's 32 offsets alive (merging the loads) and spills them around the 32 c2v registers.
LDPC_DEVINL uint4 x2_sched(const uint4 *p)
{
    uint4 r;
    asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
LDPC_DEVINL uint4 x2_sched_again(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// (a & b) ^ c as ONE LOP3 (NVVM re-associates the XOR chain into three otherwise)
LDPC_DEVINL uint32_t x2_and_xor(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One check row of a frame pair.  KEEP: the row's 32 schedule offsets stay in registers between the passes instead of being
// read again.  Returns the row's flag word: bit 0 / 16 = row unsatisfied or min |v2c| < m0 (frame a / b), bit 1 / 17 = cap engaged.
template <int DC, int DV, int NB, int M, bool KEEP>
LDPC_DEVINL uint32_t x2_check_row(unsigned char *msgb, const int slot, const uint4 *__restrict__ sched, const int j,
                                  uint32_t (&v)[DC], const DecParams &k)
{
    constexpr int NG = DC / 4;
    const uint32_t INF2 = 0x7bff7bffu;                                             // 65504, 65504
    uint32_t m1 = INF2, m2 = INF2;
    uint32_t offk[KEEP ? DC : 1];
    // the schedule words are read in order, one group ahead (`volatile` keeps ptxas from hoisting all eight loads -- 32 live
    // offsets next to the 32 c2v registers -- and from merging the two passes' loads)
    uint4 wn = KEEP ? x2_sched(&sched[j]) : x2_sched_again(&sched[j]);
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const uint4 w = wn;
        if (g + 1 < NG) wn = KEEP ? x2_sched(&sched[(size_t)(g + 1) * M + j]) : x2_sched_again(&sched[(size_t)(g + 1) * M + j]);
        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = g * 4 + q;
            if (KEEP) offk[e] = off[q];
            uint32_t so;
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(so) : "r"((uint32_t)slot), "r"((uint32_t)-NB), "r"(off[q]));
            const uint32_t s = *reinterpret_cast<const uint32_t *>(msgb + DV * NB + so);
            v[e] = h2_bits(__hsub2(h2_from(s), h2_from(v[e])));                    // v2c = sum - c2v (exact)
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
    const uint32_t sg = m1 & 0x80008000u;                                          // the rows' sign products (both frames)
    const uint32_t m1a = m1 & 0x7fff7fffu;
    const __half2 zero2 = h2_from(0u);
    // offset (:503-515; delta = 0 for plain min-sum) and cap, once per row
    const __half2 t1 = __hmax2(__hsub2(h2_from(m1a), h2_from(k.x2_delta2)), zero2);
    const __half2 t2 = __hmax2(__hsub2(h2_from(m2), h2_from(k.x2_delta2)), zero2);
    const __half2 o1 = __hmin2(t1, h2_from(k.x2_cap2)), o2 = __hmin2(t2, h2_from(k.x2_cap2));
    const uint32_t eng = __hgt2_mask(t2, h2_from(k.x2_cap2));                         // t2 >= t1
    const uint32_t bad = ~__hge2_mask(h2_from(m1a), h2_from(k.x2_m02)) | sg;             // bit 15 / 31 of each half
    const uint32_t flags = ((eng & 0x00010001u) << 1) | ((bad >> 15) & 0x00010001u);
    // second pass, per edge:  c2v = sign(v) * (row sign) * (|v| == min1 ? o2 : o1)  as  HSET2.BF (1.0 where the edge attains the row
    // minimum) -> HFMA2 eq * (s2 - s1) + s1 (exact on the lattice; the FMA pipe is the idle one in this phase, the ALU pipe --
    // HMNMX2 / LOP3 / HSET2 -- the busy one: r2e capture) -> one LOP3 for sign(v)
    const uint32_t s1 = h2_bits(o1) ^ sg, s2 = h2_bits(o2) ^ sg;
    const __half2 ds = __hsub2(h2_from(s2), h2_from(s1));
    const uint4 *sched2 = sched + k.zero;                                        // == sched, but ptxas cannot merge the second pass's loads with the first's
    if (!KEEP) wn = x2_sched_again(&sched2[j]);
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const uint4 w = wn;
        if (!KEEP && g + 1 < NG) wn = x2_sched_again(&sched2[(size_t)(g + 1) * M + j]);
        const uint32_t off[4] = { KEEP ? offk[4 * g] : w.x, KEEP ? offk[4 * g + 1] : w.y, KEEP ? offk[4 * g + 2] : w.z, KEEP ? offk[4 * g + 3] : w.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = g * 4 + q;
            const __half2 eq = __heq2(__habs2(h2_from(v[e])), h2_from(m1a));       // 1.0 where this edge attains the row minimum
            const uint32_t o = x2_and_xor(v[e], 0x80008000u, h2_bits(__hfma2(eq, ds, h2_from(s1))));   // sign(v) s1, or sign(v) s2 on the argmin edge
            v[e] = o;                                                              // c2v, kept for the next iteration
            *reinterpret_cast<uint32_t *>(msgb + off[q]) = o;
        }
    }
    return flags;
}

// Lean channel front ends of the min-sum family (fp32 conditioning): the sample source is a launch constant,
// so each source gets its own straight-line code instead of raw_samples4's per-sample dispatch through doubles.
enum { SRC_PHILOX = 0, SRC_PHILOX_FAST = 1, SRC_Q8 = 2, SRC_OTHER = 3, SRC_QP = 4, SRC_QP_STAGED = 5 };

// `staged`: SRC_QP_STAGED only, the frame's packed words in shared memory
template <int SRC, bool HASCW>
LDPC_DEVINL void ms_cond4_f32(const FrameIO &io, const DecParams &p, const CodeDev &c, const long long f, const uint8_t *cw, const int b,
                              const uint32_t qflags, const bool fcond, float (&vf)[4], const uint32_t *staged = nullptr)
{
    const int i0 = 4 * b;
    if (SRC == SRC_PHILOX) {                                   // same values as raw_samples4 + condition_ms_f32, bit for bit
        float n[4];
        normal4(io.seed, (unsigned long long)(io.frame_begin + f), (uint32_t)b, 0u, STREAM_CHANNEL, n);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double y = __dadd_rn(1.0, __dmul_rn(p.sigma, (double)n[q]));
            if (HASCW && cw && cw[i0 + q]) y = -y;                                // x (1 + sigma n), x = -1: an exact sign flip
            vf[q] = condition_ms_guarded(y, p, qflags);
        }
    } else if (SRC == SRC_PHILOX_FAST) {
        float n[4];
        normal4_fast(io.seed, (unsigned long long)(io.frame_begin + f), (uint32_t)b, 0u, STREAM_CHANNEL, n);
#pragma unroll
        for (int q = 0; q < 4; q++) vf[q] = condition_ms_guarded((double)fast_channel_sample(p, HASCW ? cw : nullptr, i0 + q, c.N, n[q]), p, qflags);
    } else if (SRC == SRC_Q8) {                                // four quantiser levels in one 32-bit load (N % 4 == 0, base 4-byte aligned)
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(io.y) + (((size_t)f * c.N) >> 2) + b);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = (int)(signed char)((w >> (8 * q)) & 0xffu);
            vf[q] = (k >= 32) ? p.Ymax_f : (k <= -32) ? -p.Ymax_f : (float)__dmul_rn((double)k, p.ms_step);
        }
    } else if (SRC == SRC_QP || SRC == SRC_QP_STAGED) {        // bit-packed quantiser levels
        double y4[4];
        if (SRC == SRC_QP_STAGED) packed_levels4_at<true>(staged, p, b, y4); else packed_levels4(io, p, c.N, f, b, y4);
#pragma unroll
        for (int q = 0; q < 4; q++) vf[q] = (float)y4[q];
    } else {
        double y4[4];
        raw_samples4(io, p, c, f, cw, b, y4);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (fcond) vf[q] = condition_ms_guarded(y4[q], p, qflags);
            else {
                double d = y4[q];
                if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) d = quantize_ms(d, p);
                if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) d = fmin(fmax(d, -p.Ymax), p.Ymax);
                vf[q] = (float)d;
            }
        }
    }
}

LDPC_DEVINL int ms_sample_source(const FrameIO &io, const DecParams &p, const int N)
{
    if (!io.y) return p.channel_mode == LDPC_GPU_CHANNEL_FAST ? SRC_PHILOX_FAST : SRC_PHILOX;
    if (io.y_dtype == LDPC_GPU_DT_Q8 && (N & 3) == 0 && ((size_t)io.y & 3) == 0) return SRC_Q8;
    if (io.y_dtype == LDPC_GPU_DT_QP) return SRC_QP;
    return SRC_OTHER;
}

// STAGED: the launch's samples are bit-packed levels whose frames are whole numbers of 16-byte units (checked by the host): they are
// staged one frame pair ahead by the TMA unit.  A separate instantiation, so that the staging state costs the other sources no registers
// (as a run-time branch of one kernel it pushed the register allocation into spills: 4.13 -> 4.26 ms per 65 536 frames on the Philox source).
template <int DC, int DV, int NFIX, int NT_MAX, int MINB, bool KEEP, bool STAGED = false>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_x2_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = NFIX, nwords = (N + 31) >> 5, nblk = (N + 3) >> 2, NB = N * 4;
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);                 // [2]
    int *st = reinterpret_cast<int *>(smem_raw + 32);                             // [0..1] certificate state, [2] flags of this check phase,
                                                                                  // (x2, by iteration parity), [4..5] failures of the variable-side check (x2)
    uint32_t *msg = reinterpret_cast<uint32_t *>(smem_raw + 64);                  // [DV*N] c2v pairs
    uint32_t *S = msg + DV * N;                                                   // [N] a-posteriori sums
    uint32_t *yq = S + N;                                                         // [N]
    uint32_t *dbits = yq + N;                                                     // [2][nwords]
    unsigned char *msgb = reinterpret_cast<unsigned char *>(msg);
    unsigned char *stage = smem_raw + (((size_t)64 + 4 * ((size_t)DV * N + 2 * (size_t)N) + 8 * (size_t)nwords + 15) & ~(size_t)15);   // [2][frame bytes <= N]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stage + 2 * N);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, M = c.M;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    const int src = STAGED ? (int)SRC_QP_STAGED : ms_sample_source(io, p, N);
    const uint32_t fbytes = (uint32_t)(((size_t)N * p.Q) >> 3);                    // STAGED: bytes of one frame's packed levels
    uint32_t stage_phase = 0;
    const bool has_row = tid < M;
    const int slot = has_row ? (int)__ldg(&c.row_slot[tid]) : 0;
    const long long npairs = (io.n_frames + 1) / 2;
    CtaTotals tot; tot.clear();
    uint32_t v[DC];                                                               // this thread's row: c2v of the previous iteration

    // channel front end of a frame pair, straight into yq / S (the iteration state of the previous pair is dead by then)
    auto gen_pair = [&](auto srct, auto cwt, const long long fa, const long long fb, const uint8_t *cwa, const uint8_t *cwb) {
        constexpr int SRC = decltype(srct)::value;
        constexpr bool HASCW = decltype(cwt)::value;
        int unca = 0, uncb = 0;
        for (int b = tid; b < nblk; b += nt) {
            float va[4], vb[4];
            ms_cond4_f32<SRC, HASCW>(io, p, c, fa, cwa, b, qflags, fcond, va, reinterpret_cast<const uint32_t *>(stage));
            ms_cond4_f32<SRC, HASCW>(io, p, c, fb, cwb, b, qflags, fcond, vb, reinterpret_cast<const uint32_t *>(stage + fbytes));
            const uint2 cc = __ldg(reinterpret_cast<const uint2 *>(c.col_of_var) + b);
            uint32_t niba = 0, nibb = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                niba |= (uint32_t)(!(va[q] > 0.0f)) << q; nibb |= (uint32_t)(!(vb[q] > 0.0f)) << q;
                const uint32_t w = h2_bits(__floats2half2_rn(va[q], vb[q]));      // exact: lattice points
                const int col = (int)(((q < 2 ? cc.x : cc.y) >> (16 * (q & 1))) & 0xffffu);
                yq[col] = w; S[col] = w;
                if (io.out_soft && p.T == 0) {                                    // T = 0: the conditioned samples themselves
                    if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)fa * N + i] = va[q]; if (fb != fa) ((double *)io.out_soft)[(size_t)fb * N + i] = vb[q]; }
                    else { ((float *)io.out_soft)[(size_t)fa * N + i] = va[q]; if (fb != fa) ((float *)io.out_soft)[(size_t)fb * N + i] = vb[q]; }
                }
            }
            if (p.T == 0) {                                                       // otherwise the decisions come from the last iteration
                if (niba) atomicOr(&dbits[(4 * b) >> 5], niba << ((4 * b) & 31));
                if (nibb) atomicOr(&dbits[nwords + ((4 * b) >> 5)], nibb << ((4 * b) & 31));
            }
            uint32_t ca = 0, cb = 0;                                              // the codeword bits of the block, as nibbles
            if (HASCW) {
                auto nibble = [&](const uint8_t *cw) -> uint32_t {
                    if (!cw) return 0u;
                    if (((size_t)cw & 3) == 0) { const uint32_t w = *reinterpret_cast<const uint32_t *>(cw + 4 * b); return (w & 1u) | ((w >> 7) & 2u) | ((w >> 14) & 4u) | ((w >> 21) & 8u); }
                    return (uint32_t)(cw[4 * b] != 0) | ((uint32_t)(cw[4 * b + 1] != 0) << 1) | ((uint32_t)(cw[4 * b + 2] != 0) << 2) | ((uint32_t)(cw[4 * b + 3] != 0) << 3);
                };
                ca = nibble(cwa); cb = nibble(cwb);
            }
            unca += __popc(niba ^ ca); uncb += __popc(nibb ^ cb);                 // uncodedErrors (:230-236)
        }
        unca = __reduce_add_sync(0xffffffffu, unca); uncb = __reduce_add_sync(0xffffffffu, uncb);
        if (lane == 0) { if (unca) atomicAdd(&fs[0].uncoded, unca); if (uncb) atomicAdd(&fs[1].uncoded, uncb); }
    };

    auto stage_pair = [&](const long long pr) {                                    // one thread: both frames of pair pr -> stage[]
        const long long fa = 2 * pr, fb = (2 * pr + 1 < io.n_frames) ? 2 * pr + 1 : 2 * pr;
        mbar_expect_tx(mbar, 2u * fbytes);
        tma_load_1d(stage, reinterpret_cast<const unsigned char *>(io.y) + (size_t)fa * fbytes, fbytes, mbar);
        tma_load_1d(stage + fbytes, reinterpret_cast<const unsigned char *>(io.y) + (size_t)fb * fbytes, fbytes, mbar);
    };
    if (STAGED) {
        if (tid == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (tid == 0 && (long long)blockIdx.x < npairs) stage_pair(blockIdx.x);
    }
    for (long long pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const long long fa = 2 * pr, fb = (2 * pr + 1 < io.n_frames) ? 2 * pr + 1 : 2 * pr;   // a dead lane replays frame fa, unreported
        const bool live_b = 2 * pr + 1 < io.n_frames;
        const uint8_t *cwa = codeword_row(io, c, fa), *cwb = codeword_row(io, c, fb);
        if (tid < 2) { fs[tid].uncoded = 0; fs[tid].errors = 0; fs[tid].flag = 0; }
        if (tid < 8) st[tid] = 0;
        if (p.T == 0) for (int w = tid; w < 2 * nwords; w += nt) dbits[w] = 0u;
        __syncthreads();
        if (STAGED) { mbar_wait(mbar, stage_phase); stage_phase ^= 1u; }          // this pair's levels have landed
        if (STAGED) {
            if (cwa || cwb) gen_pair(std::integral_constant<int, SRC_QP_STAGED>(), std::true_type(), fa, fb, cwa, cwb);
            else gen_pair(std::integral_constant<int, SRC_QP_STAGED>(), std::false_type(), fa, fb, cwa, cwb);
        } else if (cwa || cwb) {                                                  // launch constants: one straight-line front end per source
            switch (src) {
            case SRC_PHILOX:      gen_pair(std::integral_constant<int, SRC_PHILOX>(), std::true_type(), fa, fb, cwa, cwb); break;
            case SRC_PHILOX_FAST: gen_pair(std::integral_constant<int, SRC_PHILOX_FAST>(), std::true_type(), fa, fb, cwa, cwb); break;
            case SRC_Q8:          gen_pair(std::integral_constant<int, SRC_Q8>(), std::true_type(), fa, fb, cwa, cwb); break;
            case SRC_QP:          gen_pair(std::integral_constant<int, SRC_QP>(), std::true_type(), fa, fb, cwa, cwb); break;
            default:              gen_pair(std::integral_constant<int, SRC_OTHER>(), std::true_type(), fa, fb, cwa, cwb); break;
            }
        } else {
            switch (src) {
            case SRC_PHILOX:      gen_pair(std::integral_constant<int, SRC_PHILOX>(), std::false_type(), fa, fb, cwa, cwb); break;
            case SRC_PHILOX_FAST: gen_pair(std::integral_constant<int, SRC_PHILOX_FAST>(), std::false_type(), fa, fb, cwa, cwb); break;
            case SRC_Q8:          gen_pair(std::integral_constant<int, SRC_Q8>(), std::false_type(), fa, fb, cwa, cwb); break;
            case SRC_QP:          gen_pair(std::integral_constant<int, SRC_QP>(), std::false_type(), fa, fb, cwa, cwb); break;
            default:              gen_pair(std::integral_constant<int, SRC_OTHER>(), std::false_type(), fa, fb, cwa, cwb); break;
            }
        }
#pragma unroll
        for (int e = 0; e < DC; e++) v[e] = 0u;                                   // c2v = 0, S = y: the first v2c is the channel value (:364-370)
        __syncthreads();
        // everybody has read the staged levels: the next pair's can land while this one iterates
        if (STAGED && tid == 0 && pr + gridDim.x < npairs) stage_pair(pr + gridDim.x);

        // certificate state of the two frames: every thread runs the same few-instruction state machine on the same shared
        // flag words, so no thread has to publish a decision and no barrier is added.  Flag words are double-buffered by
        // iteration parity: fl[it & 1] collects check phase `it`, vb[it & 1] the variable-side check of iteration `it`.
        int sa = X2_EXACT, sb = X2_EXACT;
        uint32_t vn_ran = 0u;                                                     // bit h: the previous variable phase checked frame h
        unsigned int *fl = reinterpret_cast<unsigned int *>(&st[2]), *vb = reinterpret_cast<unsigned int *>(&st[4]);
        const bool cert_stop = (p.flags & LDPC_GPU_F_CERT_STOP) != 0;
        bool stopped = false;
        for (int it = 0; it < p.T; it++) {
            // LDPC_GPU_F_CERT_STOP: once neither frame is EXACT any more, no later iteration can change what is reported (a CERTIFIED
            // frame's decisions are fixed by the lemma, an UNCERTIFIED frame is redone anyway): leave the loop, decisions from S
            if (cert_stop && it > 0 && sa != X2_EXACT && sb != X2_EXACT) { stopped = true; break; }
            const bool last = (it == p.T - 1);
            if (last) for (int w = tid; w < 2 * nwords; w += nt) dbits[w] = 0u;
            // ---- check-node phase: one row per thread, both frames -----------------------------------
            uint32_t flags = 0u;
            if (has_row) flags = x2_check_row<DC, DV, NB, NT_MAX, KEEP>(msgb, slot, c.sched, tid, v, p);
            const bool open_cert = sa == X2_EXACT || sb == X2_EXACT;              // somebody still needs this phase's flags
            if (open_cert) {
                flags = __reduce_or_sync(0xffffffffu, flags);
                if (lane == 0 && flags) atomicOr(&fl[it & 1], flags);
            }
            __syncthreads();
            uint32_t vn_want = 0u;                                                // bit h: this variable phase checks frame h
            // (every thread runs this: keep the common case -- no frame of the pair looks converged yet, nothing engaged -- to a load,
            // a mask and a compare; the r2r capture had 6 % of all instructions in the general form below)
            const uint32_t all = open_cert ? fl[it & 1] : 0u;
            const uint32_t open_mask = ((sa == X2_EXACT) ? 1u : 0u) | ((sb == X2_EXACT) ? 0x00010000u : 0u);   // bit 0 of each open frame
            if (open_cert && !(vn_ran == 0u && (all & (open_mask * 3u)) == open_mask)) {     // open frames: rows not all satisfied, cap not engaged
                const uint32_t vbad = vb[(it + 1) & 1];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t f2 = all >> (16 * h);
                    int s = h ? sb : sa;
                    if (s == X2_EXACT) {
                        if (((vn_ran >> h) & 1u) && !((vbad >> h) & 1u) && !(f2 & 1u)) s = X2_CERT;   // (a) at it - 1 and (b) now
                        else if (f2 & 2u) s = X2_UNCERT;                          // the cap engaged on uncertified inputs
                        else if (!(f2 & 1u)) vn_want |= 1u << h;                  // looks converged: check the variable side now
                    }
                    if (h) sb = s; else sa = s;
                }
            }
            vn_ran = vn_want;
            // ---- variable-node phase: S = y + sum_s c2v, two adjacent storage columns per thread --------
            constexpr int VN_ROUNDS = (N / 2 + NT_MAX - 1) / NT_MAX;               // the host launches exactly NT_MAX threads
#pragma unroll
            for (int rr = 0; rr < VN_ROUNDS; rr++) {
                const int cp = tid + rr * NT_MAX;
                if (cp >= N / 2) break;
                const int col = 2 * cp;
                // S = y + sum_s c2v as a balanced tree: on the lattice the additions are exact, so their order is free, and the tree has
                // depth 3 instead of 7
                __half2 a0[DV + 1], a1[DV + 1];
                { const uint2 y2 = *reinterpret_cast<const uint2 *>(&yq[col]); a0[0] = h2_from(y2.x); a1[0] = h2_from(y2.y); }
#pragma unroll
                for (int s = 0; s < DV; s++) {
                    const uint2 cm = *reinterpret_cast<const uint2 *>(&msg[s * N + col]);
                    a0[s + 1] = h2_from(cm.x); a1[s + 1] = h2_from(cm.y);
                }
#pragma unroll
                for (int w = 1; w < DV + 1; w *= 2)
#pragma unroll
                    for (int s = 0; s + w < DV + 1; s += 2 * w) { a0[s] = __hadd2(a0[s], a0[s + w]); a1[s] = __hadd2(a1[s], a1[s + w]); }
                const __half2 s0 = a0[0], s1 = a1[0];
                *reinterpret_cast<uint2 *>(&S[col]) = make_uint2(h2_bits(s0), h2_bits(s1));
                if (last) {
                    const unsigned vv = __ldg(reinterpret_cast<const unsigned *>(c.var_of_col) + cp);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int i = h ? (int)(vv >> 16) : (int)(vv & 0xffffu);
                        const __half2 sum = h ? s1 : s0;
                        const float xa = __low2float(sum), xb = __high2float(sum);
                        if (!(xa > 0.0f)) atomicOr(&dbits[i >> 5], 1u << (i & 31));
                        if (!(xb > 0.0f)) atomicOr(&dbits[nwords + (i >> 5)], 1u << (i & 31));
                        if (io.out_soft) {
                            if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)fa * N + i] = xa; if (live_b) ((double *)io.out_soft)[(size_t)fb * N + i] = xb; }
                            else { ((float *)io.out_soft)[(size_t)fa * N + i] = xa; if (live_b) ((float *)io.out_soft)[(size_t)fb * N + i] = xb; }
                        }
                    }
                }
            }
            if (vn_want) {                                                        // rare (once or twice per frame): premise (a) as a second pass over
                uint32_t vfail = 0u;                                              // this thread's own columns: every next v2c = S - c2v keeps the
#pragma unroll 1                                                                  // sign of S and has magnitude >= m0
                for (int cp = tid; cp < N / 2; cp += NT_MAX) {
                    const int col = 2 * cp;
                    const uint2 s2 = *reinterpret_cast<const uint2 *>(&S[col]);
                    uint32_t w0 = 0x7bff7bffu, w1 = 0x7bff7bffu;                  // running minimum of  sign-agreement x min(|v2c|, |S|)
#pragma unroll
                    for (int s = 0; s < DV; s++) {
                        const uint2 cm = *reinterpret_cast<const uint2 *>(&msg[s * N + col]);
                        w0 = h2_bits(__hmin2(h2_from(w0), h2_from(h2_min_xorsign_abs(h2_bits(__hsub2(h2_from(s2.x), h2_from(cm.x))), s2.x))));
                        w1 = h2_bits(__hmin2(h2_from(w1), h2_from(h2_min_xorsign_abs(h2_bits(__hsub2(h2_from(s2.y), h2_from(cm.y))), s2.y))));
                    }
                    vfail |= ~(__hge2_mask(h2_from(w0), h2_from(p.x2_m02)) & __hge2_mask(h2_from(w1), h2_from(p.x2_m02)));
                }
                vfail = __reduce_or_sync(0xffffffffu, ((vfail >> 15) & 1u) | ((vfail >> 30) & 2u));
                if (lane == 0 && vfail) atomicOr(&vb[it & 1], vfail);
            }
            __syncthreads();
            if (tid == 0 && open_cert) { fl[it & 1] = 0u; vb[(it + 1) & 1] = 0u; }   // both were read by everybody before this barrier
        }
        if (stopped) {                                                            // decisions of the skipped iterations = sign(S) now
            for (int w = tid; w < 2 * nwords; w += nt) dbits[w] = 0u;
            __syncthreads();
            for (int col = tid; col < N; col += nt) {
                const int i = (int)__ldg(&c.var_of_col[col]);
                const __half2 sum = h2_from(S[col]);
                if (!(__low2float(sum) > 0.0f)) atomicOr(&dbits[i >> 5], 1u << (i & 31));
                if (!(__high2float(sum) > 0.0f)) atomicOr(&dbits[nwords + (i >> 5)], 1u << (i & 31));
            }
            __syncthreads();
        }
        // frames whose decisions are not certified go to the redo list instead of being reported
        const int enda = sa, endb = sb;
        if (tid == 0) {
            if (enda == X2_UNCERT) { const unsigned q = atomicAdd(io.redo_count, 1u); io.redo_list[q] = fa; atomicAdd(io.redo_total, 1ull); }
            if (live_b && endb == X2_UNCERT) { const unsigned q = atomicAdd(io.redo_count, 1u); io.redo_list[q] = fb; atomicAdd(io.redo_total, 1ull); }
        }
        if (enda != X2_UNCERT) finish_frame(c, p, io, fa, cwa, dbits, &fs[0], p.T, -1, 0, 0, 1, -1, tot);
        if (live_b && endb != X2_UNCERT) finish_frame(c, p, io, fb, cwb, dbits + nwords, &fs[1], p.T, -1, 0, 0, 1, -1, tot);
        __syncthreads();
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
