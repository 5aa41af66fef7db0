// ldpc_ms_rc.cuh -- min-sum family, scheduled kernel with the check-to-variable messages of a row
// kept in the REGISTERS of the row's thread from one iteration to the next.
//
// ms_sched_kernel (ldpc_ms_fast.cuh) moves every edge through shared memory four times per
// iteration: v2c read + c2v write in the check phase, c2v read + v2c write in the variable phase.
// Here the variable phase only publishes the a-posteriori sum  S_i = y_i + sum_s c2v_(s,i)  (one word
// per variable, the reference's `sum`, src/decodeMinSum.cpp:452-476), and the check thread rebuilds
//     v2c_k = S_i(k) - c2v_k(previous iteration)
// which is the reference's own expression (`sum - check_to_sym`, same operands, same order, so the
// result is bit-identical in fp64 and in fp32) from the c2v it produced itself one iteration earlier
// and still holds in registers.  Per edge and iteration: 3 shared-memory accesses instead of 4, and the
// variable phase shrinks from (dv loads, dv subtractions, dv stores) to (dv loads, dv additions, 1 store).
// The first iteration falls out of the same code: c2v = 0 and S = y give v2c = y
// (initializeSymMessages, :364-370), so the front end no longer writes dv copies of the channel value.
//
// Requirements (checked by the host, ldpc_gpu.cu): a conflict-free row schedule, one thread per check
// row (M <= blockDim), and every edge of a row sitting in the same slot of its variable's list
// (row_slot[j]; true for the 802.3an H, whose rows 64b..64b+63 touch every variable exactly once, and
// for any block-structured code with one block row per slot).  The slot makes the address of S_i a
// compile-time displacement from the address of the edge's message, so no per-edge address arithmetic
// is added to the ALU-bound check phase.
#pragma once
#include <type_traits>
#include "ldpc_ms_fast.cuh"

namespace ldpc {

template <typename Real> LDPC_DEVINL void fold2(typename SignOps<Real>::acc_t &a, Real x, Real y);
template <> LDPC_DEVINL void fold2<float>(uint32_t &a, float x, float y) { a = a ^ __float_as_uint(x) ^ __float_as_uint(y); }
template <> LDPC_DEVINL void fold2<double>(bool &a, double x, double y) { SignOps<double>::fold(a, x); SignOps<double>::fold(a, y); }

// One check row, any precision (the fp64 parity instantiation uses this one; fp32 uses rc_check_row_f32 below).
// The byte displacement from the edge's message word msg[slot*N + col] to S[col] is
// (DV - slot) * NB: the slot-dependent part is folded into each offset by one IMAD (the FMA pipe is idle in
// this phase, the ALU pipe is the busy one; written as PTX mad so that it is neither hoisted into an
// ALU-pipe IADD3 nor needs a replicated code path per slot, which thrashed the instruction cache:
// measured 10 % slower than ms_sched_kernel), the constant part is an immediate of the LDS.
template <typename Real, int DC, int DV, int NB, bool USLOT>
LDPC_DEVINL void rc_check_row(unsigned char *msgb, const int slot, const uint4 *__restrict__ sched, const int M, const int j, Real (&v)[DC],
                              const bool normalized, const bool offset, const Real alpha, const Real inv_alpha, const Real delta)
{
    constexpr int NG = DC / 4;
    Real m1 = real_inf<Real>(), m2 = real_inf<Real>();
    typename SignOps<Real>::acc_t sg = SignOps<Real>::zero();
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const uint4 w = __ldg(&sched[(size_t)g * M + j]);
        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = g * 4 + q;
            uint32_t so;
            if (USLOT) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(so) : "r"((uint32_t)slot), "r"((uint32_t)-NB), "r"(off[q]));
            else so = off[q] & (uint32_t)(NB - 1);                                         // rows that mix slots: NB is a power of two (checked by the host)
            v[k] = *reinterpret_cast<const Real *>(msgb + DV * NB + so) - v[k];            // v2c = sum - c2v
        }
        // two edges per update of (min1, min2): lo / hi of the pair, then three-input minima (FMNMX3 on
        // sm_100a): 5 min/max instructions per pair instead of 6.  min / max are exact, so any association
        // gives the reference's min1 and min2 bit for bit.
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
            fold2<Real>(sg, v[g * 4 + q], v[g * 4 + q + 1]);                               // one three-input LOP3 per pair
            const Real a = absr(v[g * 4 + q]), b = absr(v[g * 4 + q + 1]);
            const Real lo = rmin(a, b), hi = rmax(a, b);
            m2 = rmin(rmin(rmax(m1, lo), m2), hi);
            m1 = rmin(m1, lo);
        }
    }
    Real o1 = m1, o2 = m2;
    if (normalized) {
        o1 = o1 / alpha; o2 = o2 / alpha;
    }
    if (offset) { o1 = o1 - delta; o1 = (o1 > 0) ? o1 : (Real)0; o2 = o2 - delta; o2 = (o2 > 0) ? o2 : (Real)0; }
    const Real s1 = SignOps<Real>::presign(o1, sg), s2 = SignOps<Real>::presign(o2, sg);
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const uint4 w = __ldg(&sched[(size_t)g * M + j]);
        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = g * 4 + q;
            const Real sel = (absr(v[k]) == m1) ? s2 : s1;
            v[k] = SignOps<Real>::apply(sel, v[k]);                                      // c2v, kept for the next iteration
            *reinterpret_cast<Real *>(msgb + off[q]) = v[k];
        }
    }
}

#ifndef RC_NPRE
#define RC_NPRE 2          // schedule words (4 edges each) fetched ahead of the check phase's barrier
#endif

// d = min(|a|, |b|) carrying sign(a) XOR sign(b): FMNMX.XORSIGN, one ALU-pipe instruction (sm_86+).
LDPC_DEVINL float min_xorsign_abs(float a, float b)
{
    float d;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// fp32 check row.  Same results as rc_check_row<float> bit for bit, with the sign algebra folded into the
// min / max instructions:
//   pass 1: min1 is carried as a SIGNED value whose sign bit is the XOR of the signs seen so far
//           (min.xorsign.abs), so the row's sign product costs nothing; (min1, min2) advance two edges at a
//           time with a three-input minimum (FMNMX3): 5 ALU-pipe instructions per pair of edges;
//   pass 2: ts = min.xorsign.abs(v, min2) is min(|v|, min2) with the sign of v; min(|v|, min2) is min1 exactly
//           when this edge alone attains the row minimum and min2 otherwise, so the wanted magnitude
//           (|v| == min1 ? min2 : min1) has the bit pattern bits(min1) + bits(min2) - bits(min(|v|, min2)), and
//           subtracting bits(ts) instead leaves the sign of v in bit 31 (-2^31 = +2^31 mod 2^32).  One FMUL by
//           +-1/alpha (the row's sign product) finishes the message: the same rounding of the same operands
//           as scaling min1 / min2 once per row.  1 ALU-pipe + 2 FMA-pipe instructions per edge instead of
//           FSETP + FSEL + LOP3 on the ALU pipe, which is the pipe that bounds this phase.
template <int DC, int DV, int NB, bool USLOT>
LDPC_DEVINL void rc_check_row_f32(unsigned char *msgb, const int slot, const uint4 (&sw)[RC_NPRE], const uint4 *__restrict__ sched, const int M, const int j, float (&v)[DC],
                                  const bool normalized, const bool offset, const float inv_alpha, const float delta, const float alpha_div)
{
    constexpr int NG = DC / 4;
    float m1 = real_inf<float>(), m2 = real_inf<float>();
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const uint4 w = (g < RC_NPRE) ? sw[g] : __ldg(&sched[(size_t)g * M + j]);
        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = g * 4 + q;
            uint32_t so;
            if (USLOT) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(so) : "r"((uint32_t)slot), "r"((uint32_t)-NB), "r"(off[q]));
            else so = off[q] & (uint32_t)(NB - 1);
            v[k] = *reinterpret_cast<const float *>(msgb + DV * NB + so) - v[k];          // v2c = sum - c2v
        }
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
            const float a = v[g * 4 + q], b = v[g * 4 + q + 1];
            const float lo = min_xorsign_abs(a, b), hi = fmaxf(fabsf(a), fabsf(b));
            m2 = fminf(fminf(fmaxf(fabsf(m1), fabsf(lo)), m2), hi);
            m1 = min_xorsign_abs(m1, lo);
        }
    }
    const uint32_t sg = __float_as_uint(m1);                                             // bit 31 = the row's sign product
    m1 = fabsf(m1);
    const uint32_t K = __float_as_uint(m1) + __float_as_uint(m2);
    const float mult = SignOps<float>::presign(normalized ? inv_alpha : 1.0f, sg);
    // alpha_div != 0: alpha is not a power of two, so m * (1/alpha) is not always the correctly rounded m / alpha the
    // reference computes (src/decodeMinSum.cpp:494-499).  q' = fma(fma(-q, alpha, m), 1/alpha, q) with q = m * RN(1/alpha) is
    // (Markstein's correction step; no overflow / underflow at message magnitudes): two more FMA-pipe instructions per edge.
    const float nalpha = SignOps<float>::presign(-alpha_div, sg);
    // the offset variant is a separate copy of the loop: a per-edge `if (offset)` costs a branch per edge (0.35
    // branches per edge in the r1n capture) in the phase that is bound by issue slots
    auto second_pass = [&](auto with_offset, auto with_div) {
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const uint4 w = (g < RC_NPRE) ? sw[g] : __ldg(&sched[(size_t)g * M + j]);
            const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int k = g * 4 + q;
                const float ts = min_xorsign_abs(v[k], m2);
                uint32_t rb;
                asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rb) : "r"(__float_as_uint(ts)), "r"(K));   // K - bits(ts), FMA pipe
                float o = __fmul_rn(__uint_as_float(rb), mult);
                if (decltype(with_div)::value) o = __fmaf_rn(__fmaf_rn(o, nalpha, __uint_as_float(rb)), mult, o);
                // offset min-sum: sgn(o) max(|o| - delta, 0) = o - clamp(o, -delta, +delta), and the clamp is
                // min.xorsign.abs(o, delta): the same subtraction of the same operands as the reference's, sign-symmetric
                if (decltype(with_offset)::value) o = __fadd_rn(o, -min_xorsign_abs(o, delta));
                v[k] = o;                                                                // c2v, kept for the next iteration
                *reinterpret_cast<float *>(msgb + off[q]) = o;
            }
        }
    };
    if (alpha_div != 0.0f) { if (offset) second_pass(std::true_type(), std::true_type()); else second_pass(std::false_type(), std::true_type()); }
    else if (offset) second_pass(std::true_type(), std::false_type());
    else second_pass(std::false_type(), std::false_type());
}

// Sum-product check row on the register-resident structure (fp32 instantiation; src/decodeBP.cpp:353-377 in the phi domain, as in
// ms_sched_kernel<.., ALGO_BP>): v2c = clip(S - c2v, +-MAXLLR) (:399-402), |c2v_k| = phi(sum_{i != k} phi(|v_i|)).  The leave-one-out
// sum is total - own with the total carried as an unevaluated fp32 pair (hi, lo) -- Knuth's TwoSum per edge, on the FMA pipe -- which
// keeps the difference accurate when one weak message dominates the total.  (ms_sched_kernel carried it in fp64: a float->double and
// a double->float conversion per edge on the quarter-rate XU pipe, which phi's three MUFU per call already saturate.)
template <int DC, int DV, int NB, bool USLOT>
LDPC_DEVINL void rc_check_row_bp(unsigned char *msgb, const int slot, const uint4 *__restrict__ sched, const int M, const int j, float (&v)[DC], const float maxllr)
{
    constexpr int NG = DC / 4;
    float hi = 0.0f, lo = 0.0f;
    uint32_t sgb = 0u;
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const uint4 w = __ldg(&sched[(size_t)g * M + j]);
        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = g * 4 + q;
            uint32_t so;
            if (USLOT) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(so) : "r"((uint32_t)slot), "r"((uint32_t)-NB), "r"(off[q]));
            else so = off[q] & (uint32_t)(NB - 1);
            float x = *reinterpret_cast<const float *>(msgb + DV * NB + so) - v[k];      // v2c = sum - c2v
            x = fminf(fmaxf(x, -maxllr), maxllr);
            const float ph = bp_phi<float>(fabsf(x));
            const float s = __fadd_rn(hi, ph), bb = __fadd_rn(s, -hi);                   // TwoSum(hi, ph)
            lo = __fadd_rn(lo, __fadd_rn(__fadd_rn(hi, -__fadd_rn(s, -bb)), __fadd_rn(ph, -bb)));
            hi = s;
            sgb ^= __float_as_uint(x);
            v[k] = SignOps<float>::apply(ph, x);                                         // phi(|v|) carrying the sign of v
        }
    }
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const uint4 w = __ldg(&sched[(size_t)g * M + j]);
        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = g * 4 + q;
            const float own = v[k];
            const float excl = __fadd_rn(__fadd_rn(hi, -fabsf(own)), lo);
            const float mag = bp_phi<float>(fmaxf(excl, 0.0f));
            const float o = SignOps<float>::apply(SignOps<float>::presign(mag, sgb), own);
            v[k] = o;                                                                    // c2v, kept for the next iteration
            *reinterpret_cast<float *>(msgb + off[q]) = o;
        }
    }
}

// USLOT: every edge of a row sits in the same slot of its variables' lists (row_slot; the redundant-row 802.3an H).  Without it
// (the full-rank 802_3.alist: dv in {5, 6}, rows mix slots) the sum's address is the message offset modulo the plane size, and
// the planes of slots a variable does not have stay zero from the kernel's start.
template <typename Real, int DC, int DV, int NFIX, int NT_MAX, int MINB, bool USLOT = true, int ALGO = ALGO_MS>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_rc_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(DV <= 8, "slot dispatch covers dv <= 8");
    static_assert(ALGO == ALGO_MS || sizeof(Real) == 4, "the phi-domain sum-product row update is the fp32 path");
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);
    Real *msg = reinterpret_cast<Real *>(smem_raw + 16);                       // [DV*N] c2v
    constexpr int N = NFIX;
    Real *S = msg + DV * N;                                                    // [N] a-posteriori sums
    Real *yq = S + N;                                                          // [N] channel values
    uint32_t *dbits = reinterpret_cast<uint32_t *>(yq + N);
    unsigned char *msgb = reinterpret_cast<unsigned char *>(msg);
    constexpr int NB = N * (int)sizeof(Real);                                  // bytes per slot plane

    const int tid = threadIdx.x, nt = blockDim.x;
    const int M = c.M;
    constexpr int nwords = (N + 31) >> 5, nblk = (N + 3) >> 2;
    static_assert(DC % 4 == 0 && N % 32 == 0, "scheduled kernel needs dc % 4 == 0 and N % 32 == 0");
    const Real alpha = (Real)p.alpha, delta = (Real)p.delta, inv_alpha = (Real)p.inv_alpha_f;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    const bool has_row = tid < M;
    const int slot = (USLOT && has_row) ? (int)__ldg(&c.row_slot[tid]) : -1;
    if (!USLOT) { for (int q = tid; q < DV * N; q += nt) msg[q] = (Real)0; __syncthreads(); }

    CtaTotals tot; tot.clear();
    // channel front end, software-pipelined across frames exactly as in ms_sched_kernel
    Real *ybuf = reinterpret_cast<Real *>(dbits + nwords);
    uint32_t *rnext = reinterpret_cast<uint32_t *>(ybuf + N);
    int *unc_next = reinterpret_cast<int *>(rnext + nwords);
    auto gen = [&](long long fx, const uint8_t *cwx, int b) {
        double y4[4];
        raw_samples4(io, p, c, fx, cwx, b, y4);
        const uint2 cc = __ldg(reinterpret_cast<const uint2 *>(c.col_of_var) + b);
        uint32_t nib = 0; int unc = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = 4 * b + q;
            Real vr; bool rneg;
            if (ALGO == ALGO_BP) {                                // src/decodeBP.cpp:186-193
                double v = 4.0 * y4[q] / p.N0;
                if (fabs(v) > p.MAXLLR) v = (neg_ge(v) ? -1.0 : 1.0) * p.MAXLLR;
                rneg = neg_ge(v); vr = (Real)v;
            } else if (sizeof(Real) == 4 && fcond) {
                const float vf = condition_ms_guarded(y4[q], p, qflags);
                vr = (Real)vf; rneg = !(vf > 0.0f);
            } else {                                              // src/decodeMinSum.cpp:214-238
                double v = y4[q];
                if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) v = quantize_ms(v, p);
                if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) { if (v > p.Ymax) v = p.Ymax; if (v < -p.Ymax) v = -p.Ymax; }
                rneg = !(v > 0);
                vr = (Real)v;
            }
            const int col = (int)(((q < 2 ? cc.x : cc.y) >> (16 * (q & 1))) & 0xffffu);
            ybuf[col] = vr;
            unc += (int)(rneg != ((cwx ? cwx[i] : 0) != 0));
            nib |= (uint32_t)rneg << q;
        }
        if (nib) atomicOr(&rnext[(4 * b) >> 5], nib << ((4 * b) & 31));
        if (unc) atomicAdd(unc_next, unc);
    };

    // redo launches of the exact-lattice packed kernel (ldpc_ms_x2.cuh): the launch's frames are the batch frames
    // frame_list[0 .. *n_frames_dev)
    const long long n_launch = io.n_frames_dev ? min((long long)*io.n_frames_dev, io.n_frames) : io.n_frames;
    auto frame_of = [&](long long q) -> long long { return io.frame_list ? io.frame_list[q] : q; };
    if ((long long)blockIdx.x < n_launch) {
        for (int w = tid; w < nwords; w += nt) rnext[w] = 0u;
        if (tid == 0) *unc_next = 0;
        __syncthreads();
        const long long f0 = frame_of(blockIdx.x);
        const uint8_t *cw0 = codeword_row(io, c, f0);
        for (int b = tid; b < nblk; b += nt) gen(f0, cw0, b);
        __syncthreads();
    }

    Real v[DC];                                                                // this thread's row: c2v of the previous iteration
    // fp32: the row's schedule words are fetched BEFORE the barrier that opens a check phase (the registers
    // are free during the variable phase), so their L1 latency overlaps the barrier wait instead of stalling
    // all twelve warps at once right after it (long_scoreboard was 10 % of the check-phase samples)
    uint4 sw[RC_NPRE];
    auto fetch_schedule = [&]() {
        if constexpr (sizeof(Real) == 4 && ALGO == ALGO_MS) {
            if (has_row) {
#pragma unroll
                for (int g = 0; g < RC_NPRE; g++) sw[g] = __ldg(&c.sched[(size_t)g * M + tid]);
            }
        }
    };
    for (long long fq = blockIdx.x; fq < n_launch; fq += gridDim.x) {
        const long long f = frame_of(fq);
        const uint8_t *cw = codeword_row(io, c, f);
        // the staging words are cleared by the thread that consumes them, before the barrier that precedes the
        // next frame's generator (which accumulates into them with atomics)
        if (tid == 0) { fs->uncoded = *unc_next; *unc_next = 0; fs->errors = 0; fs->flag = 0; }
        for (int w = tid; w < nwords; w += nt) { dbits[w] = rnext[w]; rnext[w] = 0u; }
        for (int cp = tid; cp < N / 2; cp += nt) {
            typedef typename Vec2<Real>::type V2;
            const V2 vr = *reinterpret_cast<const V2 *>(&ybuf[2 * cp]);
            *reinterpret_cast<V2 *>(&yq[2 * cp]) = vr;
            *reinterpret_cast<V2 *>(&S[2 * cp]) = vr;
            if (io.out_soft && p.T == 0) {
                const unsigned vv = __ldg(reinterpret_cast<const unsigned *>(c.var_of_col) + cp);
                const int i0 = (int)(vv & 0xffffu), i1 = (int)(vv >> 16);
                if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)f * N + i0] = (double)vr.x; ((double *)io.out_soft)[(size_t)f * N + i1] = (double)vr.y; }
                else { ((float *)io.out_soft)[(size_t)f * N + i0] = (float)vr.x; ((float *)io.out_soft)[(size_t)f * N + i1] = (float)vr.y; }
            }
        }
#pragma unroll
        for (int k = 0; k < DC; k++) v[k] = (Real)0;
        fetch_schedule();
        __syncthreads();
        const bool have_next = fq + gridDim.x < n_launch;
        const long long fnext = have_next ? frame_of(fq + gridDim.x) : 0;
        const uint8_t *cwn = have_next ? codeword_row(io, c, fnext) : nullptr;
        int gen_done = 0;
        if (p.T == 0) __syncthreads();

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            if (last) for (int w = tid; w < nwords; w += nt) dbits[w] = 0u;
            // ---- check-node phase: one row per thread ------------------------------------------------
            if (has_row) {
                if constexpr (ALGO == ALGO_BP) rc_check_row_bp<DC, DV, NB, USLOT>(msgb, slot, c.sched, M, tid, v, (float)p.MAXLLR);
                else if constexpr (sizeof(Real) == 4) rc_check_row_f32<DC, DV, NB, USLOT>(msgb, slot, sw, c.sched, M, tid, v, normalized, offset, inv_alpha, delta, p.alpha_div_f);
                else rc_check_row<Real, DC, DV, NB, USLOT>(msgb, slot, c.sched, M, tid, v, normalized, offset, alpha, inv_alpha, delta);
            }
            // next frame's channel samples, one block per thread, right after the thread's row: the generator is a long
            // dependent chain on the FMA pipe and the check phase is issue / ALU bound, so the warps still in their rows
            // hide it (placed after the variable phase it cost 2.7 % more: 21.9 vs 22.5 Gbit/s)
            if (have_next && gen_done < nblk) { if (gen_done + tid < nblk) gen(fnext, cwn, gen_done + tid); gen_done += nt; }
            __syncthreads();
            // ---- variable-node phase: S = y + sum_s c2v, two adjacent storage columns per thread ------
            auto vn_pair = [&](const int cp) {
                typedef typename Vec2<Real>::type V2;
                const int col = 2 * cp;
                V2 sum = *reinterpret_cast<const V2 *>(&yq[col]);
#pragma unroll
                for (int s = 0; s < DV; s++) { const V2 cm = *reinterpret_cast<const V2 *>(&msg[s * N + col]); sum.x += cm.x; sum.y += cm.y; }
                *reinterpret_cast<V2 *>(&S[col]) = sum;
                if (last) {
                    const unsigned vv = __ldg(reinterpret_cast<const unsigned *>(c.var_of_col) + cp);
                    const int i0 = (int)(vv & 0xffffu), i1 = (int)(vv >> 16);
                    if (!(sum.x > 0)) atomicOr(&dbits[i0 >> 5], 1u << (i0 & 31));
                    if (!(sum.y > 0)) atomicOr(&dbits[i1 >> 5], 1u << (i1 & 31));
                    if (io.out_soft) {
                        if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)f * N + i0] = (double)sum.x; ((double *)io.out_soft)[(size_t)f * N + i1] = (double)sum.y; }
                        else { ((float *)io.out_soft)[(size_t)f * N + i0] = (float)sum.x; ((float *)io.out_soft)[(size_t)f * N + i1] = (float)sum.y; }
                    }
                }
            };
            // compile-time trip count (the host launches this kernel with exactly NT_MAX threads): no loop control, and the
            // loads of the rounds can overlap (22.4 -> 23.0 Gbit/s)
            constexpr int VN_ROUNDS = (N / 2 + NT_MAX - 1) / NT_MAX;
#pragma unroll
            for (int rr = 0; rr < VN_ROUNDS; rr++) { const int cp = tid + rr * NT_MAX; if (cp < N / 2) vn_pair(cp); }
            if (!last) fetch_schedule();
            __syncthreads();
        }
        finish_frame(c, p, io, f, cw, dbits, fs, p.T, -1, 0, 0, 1, -1, tot);      // ends with a barrier
        if (have_next && gen_done < nblk) {                          // what the iterations did not cover (T < 2); uniform
            for (int b = gen_done + tid; b < nblk; b += nt) gen(fnext, cwn, b);
            __syncthreads();
        }
    }
    if (tid == 0) tot.flush(io.counters);
}

template <typename Real>
static inline size_t ms_rc_smem_bytes(const CodeDev &c)
{
    const size_t nwords = (size_t)(c.N + 31) / 32;
    return (16 + sizeof(Real) * ((size_t)c.dvN + 3 * (size_t)c.N) + 8 * nwords + 16 + 15) & ~(size_t)15;
}

} // namespace ldpc
