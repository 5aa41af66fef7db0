// ldpc_ms_fast.cuh -- min-sum family, degree-specialised kernel (the headline path).
//
// Same data flow, same arithmetic order and the same on-chip layout as mp_kernel<.., ALGO_MS>
// (ldpc_mp_kernels.cuh), with everything the profiler showed to be wasted removed
// (profiles/r1a_*: 82 issued lane-instructions per edge-iteration in the generic loop):
//   * row / column weights are template parameters: loops fully unroll, a check thread keeps the
//     row's DC incoming messages in registers between its two passes;
//   * the argmin index is not tracked: c2v_k = (|v_k| == min1) ? min2 : min1 gives the reference's
//     result for every tie pattern (when the minimum is attained twice min2 == min1);
//   * normalisation / offset act on min1 and min2 once per row, not once per edge (the division is
//     sign-symmetric, so the value per edge is the reference's bit for bit);
//   * fp32: the sign product is an XOR of raw bit patterns and is applied with one LOP3 per edge.
//     (-0.0 then counts as negative where the reference's sgn() says +1; it can only occur when a
//     caller passes y = -0.0 exactly.  The fp64 parity instantiation keeps sgn()'s comparison.)
#pragma once
#include "ldpc_mp_kernels.cuh"

namespace ldpc {

template <typename Real> struct SignOps;
template <> struct SignOps<float> {
    typedef uint32_t acc_t;                     // XOR of the raw bit patterns: bit 31 = sign product
    static LDPC_DEVINL acc_t zero() { return 0u; }
    static LDPC_DEVINL void fold(acc_t &a, float v) { a ^= __float_as_uint(v); }
    // magnitude (>= 0) carrying the row's sign product
    static LDPC_DEVINL float presign(float mag, acc_t a) { return __uint_as_float(__float_as_uint(mag) ^ (a & 0x80000000u)); }
    // times sgn(v_k)
    static LDPC_DEVINL float apply(float presigned, float v) { return __uint_as_float(__float_as_uint(presigned) ^ (__float_as_uint(v) & 0x80000000u)); }
};
template <> struct SignOps<double> {
    typedef bool acc_t;
    static LDPC_DEVINL acc_t zero() { return false; }
    static LDPC_DEVINL void fold(acc_t &a, double v) { a ^= neg_ge(v); }
    static LDPC_DEVINL double presign(double mag, acc_t a) { return a ? -mag : mag; }
    static LDPC_DEVINL double apply(double presigned, double v) { return neg_ge(v) ? -presigned : presigned; }
};

template <typename Real> LDPC_DEVINL Real rmin(Real a, Real b);
template <> LDPC_DEVINL float rmin<float>(float a, float b) { return fminf(a, b); }
template <> LDPC_DEVINL double rmin<double>(double a, double b) { return fmin(a, b); }
template <typename Real> LDPC_DEVINL Real rmax(Real a, Real b);
template <> LDPC_DEVINL float rmax<float>(float a, float b) { return fmaxf(a, b); }
template <> LDPC_DEVINL double rmax<double>(double a, double b) { return fmax(a, b); }

// DC / DV: compile-time bounds of the row / column weight.  REGC / REGV: every row / column has
// exactly that weight (no per-slot predicate).
template <typename Real> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

// NT_MAX / MINB: launch bounds (threads per CTA, CTAs per SM the register allocation must allow).
template <typename Real, int DC, int DV, bool REGC, bool REGV, int NT_MAX, int MINB>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_fast_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);
    Real *msg = reinterpret_cast<Real *>(smem_raw + 16);
    Real *yq = msg + c.dvN;
    uint32_t *dbits = reinterpret_cast<uint32_t *>(yq + c.N);

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int N = c.N, M = c.M;
    const int nwords = (N + 31) >> 5, npad = nwords << 5, nblk = (N + 3) >> 2;
    constexpr int VPL = 8, NG = (DC + VPL - 1) / VPL;
    const uint4 *cnv = reinterpret_cast<const uint4 *>(c.cn_pos);
    const Real INF = real_inf<Real>();
    const Real alpha = (Real)p.alpha, delta = (Real)p.delta;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;      // fp32 instantiation: fp32 front end on fp32 samples

    CtaTotals tot; tot.clear();

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        if (tid == 0) { fs->uncoded = 0; fs->errors = 0; fs->flag = 0; }
        for (int w = tid; w < nwords; w += nt) dbits[w] = 0u;
        __syncthreads();

        // ---- channel front end (src/decodeMinSum.cpp:214-240) --------------------------------
        int unc = 0;
        for (int b = tid; b < nblk; b += nt) {
            double y4[4];
            raw_samples4(io, p, c, f, cw, b, y4);
            uint32_t nib = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                Real vr; bool rneg;
                if (sizeof(Real) == 4 && fcond) {
                    const float vf = condition_ms_guarded(y4[q], p, qflags);
                    vr = (Real)vf; rneg = !(vf > 0.0f);
                } else {
                    double v = y4[q];
                    if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) v = quantize_ms(v, p);
                    if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) { if (v > p.Ymax) v = p.Ymax; if (v < -p.Ymax) v = -p.Ymax; }
                    rneg = !(v > 0);
                    vr = (Real)v;
                }
                yq[i] = vr;
                const int cb = cw ? cw[i] : 0;
                unc += (int)(rneg != (cb != 0));
                nib |= (uint32_t)rneg << q;
                const int deg = REGV ? DV : (int)c.vn_deg[i];
#pragma unroll
                for (int s = 0; s < DV; s++) if (REGV || s < deg) msg[s * N + i] = vr;
                if (io.out_soft && p.T == 0) {
                    if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i] = (double)vr;
                    else ((float *)io.out_soft)[(size_t)f * N + i] = (float)vr;
                }
            }
            if (nib) atomicOr(&dbits[(4 * b) >> 5], nib << ((4 * b) & 31));
        }
        for (int o = 16; o; o >>= 1) unc += __shfl_xor_sync(0xffffffffu, unc, o);
        if (lane == 0 && unc) atomicAdd(&fs->uncoded, unc);
        __syncthreads();

        for (int it = 0; it < p.T; it++) {
            // ---- check-node phase: src/decodeMinSum.cpp:410-450 (+ :494-515) --------------------
            for (int j = tid; j < M; j += nt) {
                const int deg = REGC ? DC : (int)c.cn_deg[j];
                Real v[DC];
                Real m1 = INF, m2 = INF;
                typename SignOps<Real>::acc_t sg = SignOps<Real>::zero();
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const uint4 w = __ldg(&cnv[(size_t)g * M + j]);
#pragma unroll
                    for (int q = 0; q < VPL; q++) {
                        const int k = g * VPL + q;
                        if (k < DC && (REGC || k < deg)) {
                            v[k] = msg[IdxVec<uint16_t>::get(w, q)];
                            const Real a = absr(v[k]);
                            m2 = rmin(m2, rmax(m1, a));
                            m1 = rmin(m1, a);
                            SignOps<Real>::fold(sg, v[k]);
                        }
                    }
                }
                Real o1 = m1, o2 = m2;
                if (normalized) {
                    o1 = o1 / alpha; o2 = o2 / alpha;                       // IEEE division in every precision, once per row (src/decodeMinSum.cpp:494-499)
                }
                if (offset) { o1 = o1 - delta; o1 = (o1 > 0) ? o1 : (Real)0; o2 = o2 - delta; o2 = (o2 > 0) ? o2 : (Real)0; }
                const Real s1 = SignOps<Real>::presign(o1, sg), s2 = SignOps<Real>::presign(o2, sg);
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const uint4 w = __ldg(&cnv[(size_t)g * M + j]);
#pragma unroll
                    for (int q = 0; q < VPL; q++) {
                        const int k = g * VPL + q;
                        if (k < DC && (REGC || k < deg)) {
                            const Real sel = (absr(v[k]) == m1) ? s2 : s1;
                            msg[IdxVec<uint16_t>::get(w, q)] = SignOps<Real>::apply(sel, v[k]);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- variable-node phase: src/decodeMinSum.cpp:452-476 ----------------------------
            const bool last = (it == p.T - 1);
            for (int i0 = tid; i0 < npad; i0 += nt) {
                bool dneg = false;
                if (i0 < N) {
                    const int deg = REGV ? DV : (int)c.vn_deg[i0];
                    Real cm[DV];
                    Real sum = yq[i0];
#pragma unroll
                    for (int s = 0; s < DV; s++) if (REGV || s < deg) { cm[s] = msg[s * N + i0]; sum += cm[s]; }
#pragma unroll
                    for (int s = 0; s < DV; s++) if (REGV || s < deg) msg[s * N + i0] = sum - cm[s];
                    dneg = !(sum > 0);
                    if (last && io.out_soft) {
                        if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i0] = (double)sum;
                        else ((float *)io.out_soft)[(size_t)f * N + i0] = (float)sum;
                    }
                }
                if (last) { const unsigned bal = __ballot_sync(0xffffffffu, dneg); if (lane == 0) dbits[i0 >> 5] = bal; }
            }
            __syncthreads();
        }
        finish_frame(c, p, io, f, cw, dbits, fs, p.T, /*satisfied: evaluated on demand*/ -1, 0, 0, 1, -1, tot);
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc

namespace ldpc {

// ---------------------------------------------------------------------------------------------
// Scheduled variant for regular codes whose size is known at compile time (the 802.3an H):
// identical arithmetic, but
//   * the check phase walks each row in the bank-conflict-free order of ldpc_schedule.h, reading
//     pre-multiplied BYTE offsets four at a time (no index unpacking, no address arithmetic);
//   * variables live at storage column col(i); the variable phase walks columns at stride 1 with
//     compile-time slot strides (immediate offsets);
//   * decisions are scattered back to true variable order once, in the last iteration.
// ---------------------------------------------------------------------------------------------
// ALGO = ALGO_BP (fp32 only): sum-product on the same skeleton.  The row update is done in the
// phi domain, phi(x) = -ln tanh(x/2) = log1p(2/expm1(x)):  |c2v_k| = phi(sum_{i != k} phi(|v_i|)),
// because in fp32 tanhf saturates to 1 and the reference's product form (src/decodeBP.cpp:353-377)
// yields inf.  The leave-one-out sum is total - own with the total kept in fp64, which keeps full fp32
// accuracy even when one weak message dominates the total; cost O(dc) per row instead of the
// reference's O(dc^2).  The fp64 parity instantiation stays on mp_kernel (reference order, O(dc^2)).
template <typename Real, int DC, int DV, int NFIX, int NT_MAX, int MINB, int ALGO = ALGO_MS>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_sched_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    static_assert(ALGO == ALGO_MS || sizeof(Real) == 4, "the phi-domain sum-product row update is the fp32 path");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);
    Real *msg = reinterpret_cast<Real *>(smem_raw + 16);
    constexpr int N = NFIX;
    Real *yq = msg + DV * N;
    uint32_t *dbits = reinterpret_cast<uint32_t *>(yq + N);
    unsigned char *msgb = reinterpret_cast<unsigned char *>(msg);

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int M = c.M;
    constexpr int nwords = (N + 31) >> 5, nblk = (N + 3) >> 2;
    constexpr int NG = DC / 4;
    static_assert(DC % 4 == 0 && N % 32 == 0, "scheduled kernel needs dc % 4 == 0 and N % 32 == 0");
    const Real INF = real_inf<Real>();
    const Real alpha = (Real)p.alpha, delta = (Real)p.delta;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;      // fp32 instantiation: fp32 front end on fp32 samples

    CtaTotals tot; tot.clear();
    // ---- channel front end, software-pipelined across frames ----------------------------------------
    // gen(fx, b) produces the four conditioned samples of block b of frame fx into the staging buffer
    // ybuf (storage-column order), the raw hard decisions into rnext and the uncodedErrors increments
    // into *unc_next.  The first frame of a CTA is generated up front by all threads; every later frame
    // is generated WHILE the previous one iterates, by the threads that have no column left in the last
    // round of the variable phase (N = 2048 columns over 384 threads: 256 threads idle at the barrier),
    // so the Philox / Box-Muller / quantiser work (11 % of the kernel's instructions, profiles/
    // r1_summary.md) disappears from the critical path.
    Real *ybuf = reinterpret_cast<Real *>(dbits + nwords);
    uint32_t *rnext = reinterpret_cast<uint32_t *>(ybuf + N);
    int *unc_next = reinterpret_cast<int *>(rnext + nwords);
    auto gen = [&](long long fx, const uint8_t *cwx, int b) {
        double y4[4];
        raw_samples4(io, p, c, fx, cwx, b, y4);
        const uint2 cc = __ldg(reinterpret_cast<const uint2 *>(c.col_of_var) + b);      // four uint16 columns
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
    const int vn_rem = (N / 2) % nt;                              // threads >= vn_rem idle in the last variable round
    const int gen_threads = vn_rem ? nt - vn_rem : nt;
    const int gen_id = vn_rem ? tid - vn_rem : tid;

    if ((long long)blockIdx.x < io.n_frames) {                        // the CTA's first frame
        for (int w = tid; w < nwords; w += nt) rnext[w] = 0u;
        if (tid == 0) *unc_next = 0;
        __syncthreads();
        const uint8_t *cw0 = codeword_row(io, c, blockIdx.x);
        for (int b = tid; b < nblk; b += nt) gen(blockIdx.x, cw0, b);
        __syncthreads();
    }

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        // install the staged frame: messages start as the channel value (initializeSymMessages, :364-370)
        if (tid == 0) { fs->uncoded = *unc_next; fs->errors = 0; fs->flag = 0; }
        for (int w = tid; w < nwords; w += nt) dbits[w] = rnext[w];
        for (int col = tid; col < N; col += nt) {
            const Real vr = ybuf[col];
            yq[col] = vr;
#pragma unroll
            for (int s = 0; s < DV; s++) msg[s * N + col] = vr;
            if (io.out_soft && p.T == 0) {
                const int i = (int)__ldg(&c.var_of_col[col]);
                if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i] = (double)vr;
                else ((float *)io.out_soft)[(size_t)f * N + i] = (float)vr;
            }
        }
        __syncthreads();
        for (int w = tid; w < nwords; w += nt) rnext[w] = 0u;
        if (tid == 0) *unc_next = 0;
        const long long fnext = f + gridDim.x;
        const bool have_next = fnext < io.n_frames;
        const uint8_t *cwn = have_next ? codeword_row(io, c, fnext) : nullptr;
        int gen_done = 0;                                             // blocks of the next frame staged so far (uniform)
        if (p.T == 0) __syncthreads();

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            if (last) for (int w = tid; w < nwords; w += nt) dbits[w] = 0u;          // re-filled below, in true variable order
            // ---- check-node phase ----------------------------------------------------------------
            for (int j = tid; j < M; j += nt) {
                if (ALGO == ALGO_BP) {
                    Real t[DC];                                       // phi(|v_k|) carrying the sign of v_k
                    double tot_phi = 0.0;
                    typename SignOps<Real>::acc_t sgb = SignOps<Real>::zero();
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        const uint4 w = __ldg(&c.sched[(size_t)g * M + j]);
                        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const Real x = *reinterpret_cast<const Real *>(msgb + off[q]);
                            const Real ph = bp_phi<Real>(absr(x));
                            tot_phi += (double)ph;
                            SignOps<Real>::fold(sgb, x);
                            t[g * 4 + q] = SignOps<Real>::apply(ph, x);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        const uint4 w = __ldg(&c.sched[(size_t)g * M + j]);
                        const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const Real own = t[g * 4 + q];
                            const Real excl = (Real)(tot_phi - (double)absr(own));
                            const Real mag = bp_phi<Real>(excl > (Real)0 ? excl : (Real)0);
                            *reinterpret_cast<Real *>(msgb + off[q]) = SignOps<Real>::apply(SignOps<Real>::presign(mag, sgb), own);
                        }
                    }
                    continue;
                }
                Real v[DC];
                Real m1 = INF, m2 = INF;
                typename SignOps<Real>::acc_t sg = SignOps<Real>::zero();
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const uint4 w = __ldg(&c.sched[(size_t)g * M + j]);
                    const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int k = g * 4 + q;
                        v[k] = *reinterpret_cast<const Real *>(msgb + off[q]);
                        const Real a = absr(v[k]);
                        m2 = rmin(m2, rmax(m1, a));
                        m1 = rmin(m1, a);
                        SignOps<Real>::fold(sg, v[k]);
                    }
                }
                Real o1 = m1, o2 = m2;
                if (normalized) {
                    o1 = o1 / alpha; o2 = o2 / alpha;                       // IEEE division in every precision, once per row (src/decodeMinSum.cpp:494-499)
                }
                if (offset) { o1 = o1 - delta; o1 = (o1 > 0) ? o1 : (Real)0; o2 = o2 - delta; o2 = (o2 > 0) ? o2 : (Real)0; }
                const Real s1 = SignOps<Real>::presign(o1, sg), s2 = SignOps<Real>::presign(o2, sg);
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const uint4 w = __ldg(&c.sched[(size_t)g * M + j]);
                    const uint32_t off[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int k = g * 4 + q;
                        const Real sel = (absr(v[k]) == m1) ? s2 : s1;
                        *reinterpret_cast<Real *>(msgb + off[q]) = SignOps<Real>::apply(sel, v[k]);
                    }
                }
            }
            __syncthreads();
            // ---- variable-node phase (storage columns) --------------------------------------------
            // Two adjacent columns per thread, moved as 8/16-byte vectors: same bytes and layout, half the
            // LDS/STS instructions.  The kernel is bound by the SM-wide issue rate of memory instructions
            // (profiles/r1_summary.md), not by shared-memory bytes.
            for (int cp = tid; cp < N / 2; cp += nt) {
                typedef typename Vec2<Real>::type V2;
                const int col = 2 * cp;
                V2 cm[DV];
                V2 sum = *reinterpret_cast<const V2 *>(&yq[col]);
#pragma unroll
                for (int s = 0; s < DV; s++) { cm[s] = *reinterpret_cast<const V2 *>(&msg[s * N + col]); sum.x += cm[s].x; sum.y += cm[s].y; }
#pragma unroll
                for (int s = 0; s < DV; s++) {
                    V2 o; o.x = sum.x - cm[s].x; o.y = sum.y - cm[s].y;
                    if (ALGO == ALGO_BP) {                            // src/decodeBP.cpp:399-402
                        o.x = rmin(rmax(o.x, -(Real)p.MAXLLR), (Real)p.MAXLLR); o.y = rmin(rmax(o.y, -(Real)p.MAXLLR), (Real)p.MAXLLR);
                    }
                    *reinterpret_cast<V2 *>(&msg[s * N + col]) = o;
                }
                if (last) {
                    const unsigned vv = __ldg(reinterpret_cast<const unsigned *>(c.var_of_col) + cp);   // two uint16 variable ids
                    const int i0 = (int)(vv & 0xffffu), i1 = (int)(vv >> 16);
                    if (!(sum.x > 0)) atomicOr(&dbits[i0 >> 5], 1u << (i0 & 31));
                    if (!(sum.y > 0)) atomicOr(&dbits[i1 >> 5], 1u << (i1 & 31));
                    if (io.out_soft) {
                        if (io.y_dtype == LDPC_GPU_DT_F64) { ((double *)io.out_soft)[(size_t)f * N + i0] = (double)sum.x; ((double *)io.out_soft)[(size_t)f * N + i1] = (double)sum.y; }
                        else { ((float *)io.out_soft)[(size_t)f * N + i0] = (float)sum.x; ((float *)io.out_soft)[(size_t)f * N + i1] = (float)sum.y; }
                    }
                }
            }
            if (have_next && gen_done < nblk) {                       // idle threads stage the next frame
                if (gen_id >= 0 && gen_done + gen_id < nblk) gen(fnext, cwn, gen_done + gen_id);
                gen_done += gen_threads;
            }
            __syncthreads();
        }
        finish_frame(c, p, io, f, cw, dbits, fs, p.T, /*satisfied: evaluated on demand*/ -1, 0, 0, 1, -1, tot);
        if (have_next) for (int b = gen_done + tid; b < nblk; b += nt) gen(fnext, cwn, b);      // what T iterations did not cover
        __syncthreads();
    }
    if (tid == 0) tot.flush(io.counters);
}

// Dynamic shared memory of ms_sched_kernel: scratch, messages, channel values, decisions, and the
// staging buffer / raw decisions / uncoded count of the next frame.
template <typename Real>
static inline size_t ms_sched_smem_bytes(const CodeDev &c)
{
    const size_t nwords = (size_t)(c.N + 31) / 32;
    return (16 + sizeof(Real) * ((size_t)c.dvN + 2 * (size_t)c.N) + 8 * nwords + 16 + 15) & ~(size_t)15;
}

} // namespace ldpc
