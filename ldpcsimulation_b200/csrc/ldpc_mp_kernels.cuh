// ldpc_mp_kernels.cuh -- flooding message passing: min-sum family, sum-product, DD-BMP.
//
// One CTA owns one frame at a time (persistent over the launch's frames); several CTAs share an
// SM.  All per-frame state is on chip:
//     msg[dv_max*N]   edge messages, variable-major / slot-major.  After a variable phase an entry
//                     holds v2c, after a check phase the same entry holds c2v: every edge belongs to
//                     exactly one check thread in the check phase and one variable thread in the
//                     variable phase, so both update in place.
//     yq[N]           conditioned channel values
//     mem[dv_max*N]   DD-BMP accumulator memories only
//     dbits[N/32]     hard decisions, bit-packed through warp ballots
// The edge map (cn_pos) is read-only, identical for every frame and every CTA, and is streamed
// through L1 with 16-byte vector loads (8 edges per load for uint16 positions).
//
// Arithmetic order is the reference's: checks walk their mlist row front to back, variables add
// c2v in nlist order starting from the channel value, v2c = sum - c2v (src/decodeMinSum.cpp:410-476).
// The library is compiled with -fmad=false so no multiply-add is ever fused.
#pragma once
#include "ldpc_common.cuh"

namespace ldpc {

enum { ALGO_MS = 0, ALGO_BP = 1, ALGO_DDBMP = 2 };

template <typename IdxT> struct IdxVec;
template <> struct IdxVec<uint16_t> { enum { VPL = 8 }; static LDPC_DEVINL uint32_t get(const uint4 &w, int q) {
    const uint32_t x = (q < 2) ? w.x : (q < 4) ? w.y : (q < 6) ? w.z : w.w; return (q & 1) ? (x >> 16) : (x & 0xffffu); } };
template <> struct IdxVec<uint32_t> { enum { VPL = 4 }; static LDPC_DEVINL uint32_t get(const uint4 &w, int q) {
    return q == 0 ? w.x : q == 1 ? w.y : q == 2 ? w.z : w.w; } };

template <typename Real> LDPC_DEVINL Real bp_phi(Real x);
// phi(x) = -ln tanh(x/2) = log1p(2/expm1(x)); phi(0)=inf, phi(inf)=0
// fp32: branch-free, on the hardware exp2 / log2 / reciprocal (MUFU), ~20 instructions.  log1pf(2/expm1f(x)) cost
// ~95 instructions per call with its IEEE division and special-case paths, and a two-branch version diverged on
// almost every warp (both are in the history of this file; profiles/r1_summary.md).
//   e = exp(-x) = ex2(-x log2 e)                       (relative error ~2^-22, growing to ~1e-6 only where phi < 1e-6)
//   e <  0.1 :  phi = 2 atanh(e) = 2e (1 + z/3 + z^2/5), z = e^2                       (next term 1.4e-7)
//   e >= 0.1 :  phi = ln2 * lg2((2 - u)/u),  u = 1 - e^-x;  phi >= 0.2 there, so the 2^-22 ABSOLUTE error of
//               MUFU.LG2 is a relative error below 1e-6;  for x < 1/8, where 1 - e cancels, u comes from the series
//               x (1 - x/2 + x^2/6 - x^3/24 + x^4/120)                                    (next term 4e-8)
// phi(0) = +inf, phi(inf) = 0.  Accuracy is what tests/test_gpu_parity.py::test_f32_within_tolerance checks: sums
// within 2e-5 of the frame's largest |LLR| against the double oracle.
LDPC_DEVINL float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
LDPC_DEVINL float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
LDPC_DEVINL float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <> LDPC_DEVINL float bp_phi<float>(float x)
{
    const float e = mufu_ex2(__fmul_rn(x, -1.4426950408889634f));
    const float z = __fmul_rn(e, e);
    float a = __fmaf_rn(z, 0.2f, 0.3333333432674408f);
    a = __fmul_rn(__fadd_rn(e, e), __fmaf_rn(a, z, 1.0f));
    float pu = __fmaf_rn(x, 8.3333337679505348e-3f, -4.1666667908430099e-2f);
    pu = __fmaf_rn(pu, x, 1.6666667163372040e-1f);
    pu = __fmaf_rn(pu, x, -0.5f);
    pu = __fmul_rn(__fmaf_rn(pu, x, 1.0f), x);
    const float u = (x < 0.125f) ? pu : __fadd_rn(1.0f, -e);
    const float b = __fmul_rn(0.693147182464599609375f, mufu_lg2(__fmul_rn(__fadd_rn(2.0f, -u), mufu_rcp(u))));
    return (e < 0.1f) ? a : b;
}
template <> LDPC_DEVINL double bp_phi<double>(double x) { return log1p(2.0 / expm1(x)); }

// GSTATE: the message / sample arrays live in a per-CTA slice of an HBM workspace instead of shared
// memory (same code, same arithmetic): the path for codes such as DVB-S2 (N=64800, E=226799) whose
// messages exceed one SM.  Decisions and scratch stay in shared memory.
template <typename Real, typename IdxT, int ALGO, bool GSTATE>
__global__ void mp_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);
    unsigned char *state = GSTATE ? io.workspace + (size_t)blockIdx.x * io.ws_stride : smem_raw + 16;
    Real *msg = reinterpret_cast<Real *>(state);
    Real *yq = msg + c.dvN;
    Real *mem = yq + c.N;                                             // DD-BMP only
    uint32_t *dbits = GSTATE ? reinterpret_cast<uint32_t *>(smem_raw + 16)
                             : reinterpret_cast<uint32_t *>(ALGO == ALGO_DDBMP ? mem + c.dvN : mem);

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int N = c.N, M = c.M;
    const int nwords = (N + 31) >> 5, npad = nwords << 5, nblk = (N + 3) >> 2;
    constexpr int VPL = IdxVec<IdxT>::VPL;
    const uint4 *cnv = reinterpret_cast<const uint4 *>(c.cn_pos);
    const Real INF = real_inf<Real>();
    const Real alpha = (Real)p.alpha, delta = (Real)p.delta, MAXLLR = (Real)p.MAXLLR;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;

    CtaTotals tot; tot.clear();

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        if (tid == 0) { fs->uncoded = 0; fs->errors = 0; fs->flag = 0; }
        for (int w = tid; w < nwords; w += nt) dbits[w] = 0u;
        __syncthreads();

        // ---- channel front end: src/decodeMinSum.cpp:214-240, decodeBP.cpp:174-201, decodeDDBMP.cpp:173-187
        int unc = 0;
        for (int b = tid; b < nblk; b += nt) {
            double y4[4];
            raw_samples4(io, p, c, f, cw, b, y4);
            uint32_t nib = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                double v = y4[q];
                bool rneg;
                if (ALGO == ALGO_MS) {
                    if (p.flags & LDPC_GPU_F_QUANTIZE_SAMPLES) v = quantize_ms(v, p);
                    if (p.flags & LDPC_GPU_F_SATURATE_SAMPLES) { if (v > p.Ymax) v = p.Ymax; if (v < -p.Ymax) v = -p.Ymax; }
                    rneg = !(v > 0);
                } else if (ALGO == ALGO_BP) {
                    v = 4.0 * v / p.N0;
                    if (fabs(v) > p.MAXLLR) v = (neg_ge(v) ? -1.0 : 1.0) * p.MAXLLR;
                    rneg = neg_ge(v);
                } else {
                    v = quantize_ms(v, p);
                    rneg = !(v > 0);
                }
                const Real vr = (Real)v;
                yq[i] = vr;
                const int cb = cw ? cw[i] : 0;
                unc += (int)(rneg != (cb != 0));                      // r*c < 0
                nib |= (uint32_t)rneg << q;
                const int deg = c.vn_deg[i];
                for (int s = 0; s < deg; s++) {                       // initializeSymMessages
                    if (ALGO == ALGO_DDBMP) { msg[s * N + i] = neg_ge(vr) ? (Real)-1 : (Real)1; mem[s * N + i] = vr; }
                    else msg[s * N + i] = vr;
                }
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

        int it = 0, satisfied = -1;
        for (it = 0; it < p.T; it++) {
            // ---- check-node phase ------------------------------------------------------------
            for (int j = tid; j < M; j += nt) {
                const int deg = c.cn_deg[j];
                if (ALGO == ALGO_MS) {                               // src/decodeMinSum.cpp:410-450
                    Real m1 = INF, m2 = INF; int idx = -1; unsigned long long signs = 0ull; bool pneg = false;
                    for (int k0 = 0; k0 < deg; k0 += VPL) {
                        const uint4 w = __ldg(&cnv[(size_t)(k0 / VPL) * M + j]);
#pragma unroll
                        for (int q = 0; q < VPL; q++) {
                            const int k = k0 + q;
                            if (k < deg) {
                                const Real v = msg[IdxVec<IdxT>::get(w, q)];
                                const Real a = absr(v);
                                const bool ng = neg_ge(v);
                                signs |= (unsigned long long)ng << k; pneg ^= ng;
                                if (a <= m1) { m2 = m1; m1 = a; idx = k; } else if (a < m2) m2 = a;
                            }
                        }
                    }
                    for (int k0 = 0; k0 < deg; k0 += VPL) {
                        const uint4 w = __ldg(&cnv[(size_t)(k0 / VPL) * M + j]);
#pragma unroll
                        for (int q = 0; q < VPL; q++) {
                            const int k = k0 + q;
                            if (k < deg) {
                                const Real mag = (k == idx) ? m2 : m1;
                                const bool ng = pneg ^ (bool)((signs >> k) & 1ull);
                                Real out = ng ? -mag : mag;           // prod*minMag*sgn(msg)
                                if (normalized) out = out / alpha;                                // :494-499
                                if (offset) {                         // :503-515
                                    const Real mg = absr(out) - delta;
                                    out = (mg > 0) ? (neg_ge(out) ? -mg : mg) : (Real)0;
                                }
                                msg[IdxVec<IdxT>::get(w, q)] = out;
                            }
                        }
                    }
                } else if (ALGO == ALGO_DDBMP) {                      // src/decodeDDBMP.cpp:350-372
                    unsigned long long signs = 0ull; bool pneg = false;
                    for (int k0 = 0; k0 < deg; k0 += VPL) {
                        const uint4 w = __ldg(&cnv[(size_t)(k0 / VPL) * M + j]);
#pragma unroll
                        for (int q = 0; q < VPL; q++) {
                            const int k = k0 + q;
                            if (k < deg) { const bool ng = neg_ge(msg[IdxVec<IdxT>::get(w, q)]); signs |= (unsigned long long)ng << k; pneg ^= ng; }
                        }
                    }
                    for (int k0 = 0; k0 < deg; k0 += VPL) {
                        const uint4 w = __ldg(&cnv[(size_t)(k0 / VPL) * M + j]);
#pragma unroll
                        for (int q = 0; q < VPL; q++) {
                            const int k = k0 + q;
                            if (k < deg) msg[IdxVec<IdxT>::get(w, q)] = (pneg ^ (bool)((signs >> k) & 1ull)) ? (Real)-1 : (Real)1;
                        }
                    }
                } else {                                              // BP, src/decodeBP.cpp:353-377
                    Real t[64];
                    for (int k0 = 0; k0 < deg; k0 += VPL) {
                        const uint4 w = __ldg(&cnv[(size_t)(k0 / VPL) * M + j]);
#pragma unroll
                        for (int q = 0; q < VPL; q++) {
                            const int k = k0 + q;
                            if (k < deg) {
                                const Real v = msg[IdxVec<IdxT>::get(w, q)];
                                if (sizeof(Real) == 8) t[k] = (Real)tanh((double)v / 2.0);
                                else t[k] = (v < 0) ? -bp_phi<Real>(-v) : bp_phi<Real>(v);   // sign-carrying phi(|v|); +-0 -> +inf
                            }
                        }
                    }
                    for (int k0 = 0; k0 < deg; k0 += VPL) {
                        const uint4 w = __ldg(&cnv[(size_t)(k0 / VPL) * M + j]);
#pragma unroll
                        for (int q = 0; q < VPL; q++) {
                            const int k = k0 + q;
                            if (k < deg) {
                                Real out;
                                if (sizeof(Real) == 8) {              // the reference's O(dc^2) product, in its order
                                    double prod = 1.0;
                                    for (int k2 = 0; k2 < deg; k2++) if (k2 != k) prod *= (double)t[k2];
                                    out = (Real)log((1.0 + prod) / (1.0 - prod));
                                } else {                              // fp32: same leave-one-out, in the phi domain
                                    Real s = 0; bool ng = false;      // (tanhf saturates to 1 -> inf in the product form)
                                    for (int k2 = 0; k2 < deg; k2++) if (k2 != k) { s += absr(t[k2]); ng ^= (t[k2] < 0); }
                                    const Real mag = bp_phi<Real>(s);
                                    out = ng ? -mag : mag;
                                }
                                msg[IdxVec<IdxT>::get(w, q)] = out;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- variable-node phase -----------------------------------------------------------
            const bool last = (it == p.T - 1);
            for (int i0 = tid; i0 < npad; i0 += nt) {
                const bool valid = i0 < N;
                bool dneg = false;
                if (valid) {
                    const int deg = c.vn_deg[i0];
                    const Real ych = yq[i0];
                    Real sum = ych;
                    for (int s = 0; s < deg; s++) sum += msg[s * N + i0];          // nlist order
                    if (ALGO == ALGO_MS) {                                         // src/decodeMinSum.cpp:452-476
                        for (int s = 0; s < deg; s++) msg[s * N + i0] = sum - msg[s * N + i0];
                        dneg = !(sum > 0);
                    } else if (ALGO == ALGO_BP) {                                  // src/decodeBP.cpp:379-409
                        for (int s = 0; s < deg; s++) {
                            Real o = sum - msg[s * N + i0];
                            if (absr(o) > MAXLLR) o = neg_ge(o) ? -MAXLLR : MAXLLR;
                            msg[s * N + i0] = o;
                        }
                        dneg = !(sum > 0);
                    } else {                                                       // src/decodeDDBMP.cpp:396-423
                        Real dsum = neg_ge(ych) ? (Real)-1 : (Real)1;
                        for (int s = 0; s < deg; s++) {
                            const Real mm = mem[s * N + i0] + (sum - msg[s * N + i0]);
                            mem[s * N + i0] = mm;
                            const Real o = neg_ge(mm) ? (Real)-1 : (Real)1;
                            msg[s * N + i0] = o;
                            dsum += o;
                        }
                        dneg = !(dsum > 0);
                    }
                    if (io.out_soft && (last || ALGO == ALGO_DDBMP)) {
                        if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i0] = (double)sum;
                        else ((float *)io.out_soft)[(size_t)f * N + i0] = (float)sum;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, dneg);
                if (lane == 0) dbits[i0 >> 5] = bal;
            }
            __syncthreads();
            if (ALGO == ALGO_DDBMP) {                                 // checkStoppingCondition, :375-393
                satisfied = syndrome_ok(c, dbits);
                if (satisfied) break;                                 // `it` is not incremented on this exit (:203-204)
            }
        }
        if (ALGO == ALGO_DDBMP && p.T == 0) satisfied = syndrome_ok(c, dbits);
        // MS / BP: the reference keeps no syndrome; finish_frame evaluates it on demand (satisfied == -1)
        finish_frame(c, p, io, f, cw, dbits, fs, it, satisfied, 0, 0, 1, -1, tot);
    }
    if (tid == 0) tot.flush(io.counters);
}

// Bytes of per-frame message / sample state, and the dynamic shared memory of mp_kernel.
template <typename Real>
static inline size_t mp_state_bytes(const CodeDev &c, int algo)
{
    size_t n = sizeof(Real) * ((size_t)c.dvN + c.N);
    if (algo == ALGO_DDBMP) n += sizeof(Real) * (size_t)c.dvN;
    return (n + 15) & ~(size_t)15;
}
template <typename Real>
static inline size_t mp_smem_bytes(const CodeDev &c, int algo, bool gstate = false)
{
    size_t n = 16 + 4 * (size_t)((c.N + 31) / 32) + (gstate ? 0 : mp_state_bytes<Real>(c, algo));
    return (n + 15) & ~(size_t)15;
}

} // namespace ldpc
