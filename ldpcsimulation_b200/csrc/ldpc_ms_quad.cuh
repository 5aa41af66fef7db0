// ldpc_ms_quad.cuh -- fp32 min-sum family for SMALL codes (M <= 512 rows, row weight <= 8: the (3,6) PEG code of BASELINE
// configs[0]), four frames per thread.
//
// A small code leaves a thread-per-row kernel with six edges of work between two CTA barriers, and every edge pays its own
// index load, address computation and 4-byte shared-memory access (ms_fast_kernel: 3.4 Gbit/s at T = 50 on PEGReg504x1008).
// Here a CTA owns a TILE of four frames and every shared-memory word is a float4 holding the same value of the four frames:
//
//   S4[i]   a-posteriori sum of variable i      msg4[s*N + i]   c2v on slot s of variable i      y4[i]   channel value
//
//   * check thread = one row for all four frames: the row's c2v stay in registers between iterations (ms_rc_kernel's structure),
//     v2c = S - c2v is rebuilt from the published sums -- the reference's own expression (src/decodeMinSum.cpp:452-476), same
//     operands, same order, so the results are bit-identical to ms_fast_kernel<float> -- and the edges' byte offsets are computed
//     once per kernel, not once per iteration;
//   * one 16-byte access moves four frames: a quarter of the LSU instructions and address arithmetic per edge-frame;
//   * the row update runs frame by frame over the register-resident v2c (min1 / min2 / sign of ONE frame live at a time), which
//     is what lets 32 c2v registers + 8 packed offsets fit the 64-register budget of two 512-thread CTAs per SM;
//   * a 16-byte access is served in four phases of eight lanes.  The host sorts the rows by weight, relabels the variables
//     (storage columns) and orders each row's edges so that the eight lanes of a phase fall into eight distinct 16-byte bank groups
//     (ldpc_schedule.h, build_group_schedule: min / second min / sign parity do not depend on the visiting order), for the gather
//     of S and the scatter of c2v alike: the r2v capture had 39 % of all shared-memory wavefronts in conflict replays and the
//     shared-memory pipe busy for 7.5 of the kernel's 10.1 ms;
//   * a step in which a row has no edge (rows lighter than their group's heaviest: PEGReg504x1008 has weights 5 .. 8) is a padding
//     edge to a virtual variable whose sum is +inf: its v2c = inf - c2v = +inf changes neither the minima nor the sign product, so
//     every lane of a warp runs the same straight-line code.  A warp runs 6 or 8 steps, whichever its heaviest row needs (the
//     sort puts the 28 heavy rows of the PEG code into one warp);
//   * variable thread = storage column for all four frames: S = y + sum_s c2v in nlist order, one float4 store.
//
// Arithmetic contract: that of the fp32 instantiation (ldpc_ms_fast.cuh); tests/test_gpu_parity.py compares the two kernels
// bit for bit (LDPC_GPU_NO_QUAD selects ms_fast_kernel).
#pragma once
#include "ldpc_ms_rc.cuh"

namespace ldpc {

static inline size_t ms_quad_smem_bytes(const CodeDev &c)
{
    return (64 + 16 * ((size_t)c.dvN + 1 + 2 * (size_t)c.N + 1) + 16 * (size_t)((c.N + 31) / 32) + 15) & ~(size_t)15;
}

LDPC_DEVINL float &q4_at(float4 &a, const int q) { return q == 0 ? a.x : q == 1 ? a.y : q == 2 ? a.z : a.w; }
LDPC_DEVINL float q4_get(const float4 &a, const int q) { return q == 0 ? a.x : q == 1 ? a.y : q == 2 ? a.z : a.w; }

// DC: row weight bound (lighter rows are padded with virtual edges).  DV: column weight bound; REGV: every column has weight DV.
template <int DC, int DV, bool REGV, int NT_MAX, int MINB>
__global__ void __launch_bounds__(NT_MAX, MINB) ms_quad_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    constexpr int FI = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameScratch *fs = reinterpret_cast<FrameScratch *>(smem_raw);                  // [FI]
    unsigned char *msgb = smem_raw + 16 * FI;                                       // msg4 [dvN + 1 (trash)], S4 [N + 1 (+inf)], y4 [N]
    float4 *msg4 = reinterpret_cast<float4 *>(msgb);
    const int N = c.N, M = c.M, nwords = (N + 31) >> 5, nblk = (N + 3) >> 2;
    const uint32_t S_OFF = 16u * (uint32_t)(c.dvN + 1), Y_OFF = S_OFF + 16u * (uint32_t)(N + 1);
    float4 *S4 = reinterpret_cast<float4 *>(msgb + S_OFF);
    float4 *y4 = reinterpret_cast<float4 *>(msgb + Y_OFF);
    uint32_t *dbits = reinterpret_cast<uint32_t *>(msgb + Y_OFF + 16u * (uint32_t)N);  // [FI][nwords]
    float *yf = reinterpret_cast<float *>(y4), *Sf = reinterpret_cast<float *>(S4), *msgf = reinterpret_cast<float *>(msg4);
    const int tid = threadIdx.x, nt = blockDim.x;
    const float INF = real_inf<float>();
    const float alpha = (float)p.alpha, delta = (float)p.delta;
    const bool normalized = (p.flags & LDPC_GPU_F_NORMALIZED_MS) != 0, offset = (p.flags & LDPC_GPU_F_OFFSET_MS) != 0;
    const uint32_t qflags = p.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);
    const bool fcond = !io.y || io.y_dtype != LDPC_GPU_DT_F64;
    const long long ntiles = (io.n_frames + FI - 1) / FI;
    CtaTotals tot; tot.clear();

    // this thread's row (c.quad_edge is indexed by thread): (storage column << 16) | message position of every step, as 16-byte
    // indices; an idle step reads the +inf sum S4[N] and writes the trash word msg4[dvN]
    const bool has_row = tid < M;
    uint32_t edge[DC];
#pragma unroll
    for (int k = 0; k < DC; k++) edge[k] = has_row ? __ldg(&c.quad_edge[(size_t)k * M + tid]) : 0u;
    const int wsteps = __reduce_max_sync(0xffffffffu, has_row ? (int)c.quad_steps[tid] : 0);   // warp-uniform
    // unpacked on the FMA pipe (IMAD.HI / IMAD): the ALU pipe -- FMNMX, LOP3 -- is the one that bounds the check phase (r2v capture)
    auto s_off = [&](const uint32_t e) -> uint32_t {
        uint32_t hi, r;
        asm("mul.hi.u32 %0, %1, 0x10000;" : "=r"(hi) : "r"(e));
        asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(r) : "r"(hi), "r"(S_OFF));
        return r;
    };
    auto m_off = [&](const uint32_t e) -> uint32_t {
        uint32_t hi, t, r;
        asm("mul.hi.u32 %0, %1, 0x10000;" : "=r"(hi) : "r"(e));
        asm("mul.lo.u32 %0, %1, 16;" : "=r"(t) : "r"(e));
        asm("mad.lo.u32 %0, %1, 0xfff00000, %2;" : "=r"(r) : "r"(hi), "r"(t));       // 16 e - (hi << 20) = 16 (e & 0xffff)
        return r;
    };
    const float inv_alpha = p.inv_alpha_f, alpha_div = normalized ? p.alpha_div_f : 0.0f;
    if (tid == 0) S4[N] = make_float4(INF, INF, INF, INF);
    float4 v[DC];                                                                    // c2v of the previous iteration (then v2c, then the new c2v)

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long f0 = tile * FI;
        if (tid < FI) { fs[tid].uncoded = 0; fs[tid].errors = 0; fs[tid].flag = 0; }
        for (int w = tid; w < FI * nwords; w += nt) dbits[w] = 0u;
        __syncthreads();
        // ---- channel front end, (block of 4 samples, frame lane) per thread; S = y, c2v = 0 (src/decodeMinSum.cpp:214-240, :364-370)
        for (int t = tid; t < nblk * FI; t += nt) {
            const int b = t / FI, fl = t % FI;
            const long long f = (f0 + fl < io.n_frames) ? f0 + fl : io.n_frames - 1;     // dead lanes replay the last frame, unreported
            const uint8_t *cw = codeword_row(io, c, f);
            double s4[4];
            raw_samples4(io, p, c, f, cw, b, s4);
            uint32_t nib = 0; int unc = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = 4 * b + q;
                if (i >= N) break;
                float vr;
                if (fcond) vr = condition_ms_guarded(s4[q], p, qflags);
                else {
                    double d = s4[q];
                    if (qflags & LDPC_GPU_F_QUANTIZE_SAMPLES) d = quantize_ms(d, p);
                    if (qflags & LDPC_GPU_F_SATURATE_SAMPLES) { if (d > p.Ymax) d = p.Ymax; if (d < -p.Ymax) d = -p.Ymax; }
                    vr = (float)d;
                }
                const bool rneg = !(vr > 0.0f);
                const int col = (int)__ldg(&c.quad_col_of_var[i]);
                yf[col * FI + fl] = vr; Sf[col * FI + fl] = vr;
                const int cb = cw ? cw[i] : 0;
                unc += (int)(rneg != (cb != 0));
                nib |= (uint32_t)rneg << q;
                if (io.out_soft && p.T == 0 && f0 + fl < io.n_frames) {
                    if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)f * N + i] = (double)vr;
                    else ((float *)io.out_soft)[(size_t)f * N + i] = vr;
                }
            }
            if (nib) atomicOr(&dbits[fl * nwords + ((4 * b) >> 5)], nib << ((4 * b) & 31));
            if (unc) atomicAdd(&fs[fl].uncoded, unc);
        }
#pragma unroll
        for (int k = 0; k < DC; k++) v[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        __syncthreads();

        for (int it = 0; it < p.T; it++) {
            const bool last = (it == p.T - 1);
            // ---- check-node phase: src/decodeMinSum.cpp:410-450 (+ :494-515), one row x four frames per thread ----
            auto check_phase = [&](auto dcw) {                                      // DCW steps: 6 or 8, whichever this warp's heaviest row needs
                constexpr int DCW = decltype(dcw)::value;
#pragma unroll
                for (int k = 0; k < DCW; k++) {                                      // v2c = sum - c2v
                    const float4 s = *reinterpret_cast<const float4 *>(msgb + s_off(edge[k]));
                    v[k] = make_float4(s.x - v[k].x, s.y - v[k].y, s.z - v[k].z, s.w - v[k].w);
                }
                // frame by frame (one frame's row state live at a time), with rc_check_row_f32's arithmetic (ldpc_ms_rc.cuh): min1 carried
                // signed through min.xorsign.abs, select as K - bits(min.xorsign.abs(v, min2)), sign product in the multiplier
                auto row_update = [&](auto with_offset, auto with_div) {
#pragma unroll
                    for (int q = 0; q < FI; q++) {
                        float m1 = INF, m2 = INF;
#pragma unroll
                        for (int k = 0; k < DCW; k += 2) {
                            const float a = q4_get(v[k], q), b = q4_get(v[k + 1], q);
                            const float lo = min_xorsign_abs(a, b), hi = fmaxf(fabsf(a), fabsf(b));
                            m2 = fminf(fminf(fmaxf(fabsf(m1), fabsf(lo)), m2), hi);
                            m1 = min_xorsign_abs(m1, lo);
                        }
                        const uint32_t sg = __float_as_uint(m1);                     // bit 31 = the row's sign product
                        m1 = fabsf(m1);
                        const uint32_t K = __float_as_uint(m1) + __float_as_uint(m2);
                        const float mult = SignOps<float>::presign(normalized ? inv_alpha : 1.0f, sg);
                        const float nalpha = SignOps<float>::presign(-alpha_div, sg);
#pragma unroll
                        for (int k = 0; k < DCW; k++) {
                            const float ts = min_xorsign_abs(q4_get(v[k], q), m2);
                            uint32_t rb;
                            asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rb) : "r"(__float_as_uint(ts)), "r"(K));   // K - bits(ts), FMA pipe
                            float o = __fmul_rn(__uint_as_float(rb), mult);
                            if (decltype(with_div)::value) o = __fmaf_rn(__fmaf_rn(o, nalpha, __uint_as_float(rb)), mult, o);
                            if (decltype(with_offset)::value) o = __fadd_rn(o, -min_xorsign_abs(o, delta));
                            q4_at(v[k], q) = o;
                        }
                    }
                };
                if (alpha_div != 0.0f) { if (offset) row_update(std::true_type(), std::true_type()); else row_update(std::false_type(), std::true_type()); }
                else { if (offset) row_update(std::true_type(), std::false_type()); else row_update(std::false_type(), std::false_type()); }
#pragma unroll
                for (int k = 0; k < DCW; k++) *reinterpret_cast<float4 *>(msgb + m_off(edge[k])) = v[k];
            };
            if (last) for (int w = tid; w < FI * nwords; w += nt) dbits[w] = 0u;        // refilled by the variable phase, behind the barrier below
            if (has_row) {
                if (DC > 6 && wsteps > 6) check_phase(std::integral_constant<int, DC>());
                else check_phase(std::integral_constant<int, (DC < 6 ? DC : 6)>());
            }
            __syncthreads();
            // ---- variable-node phase: src/decodeMinSum.cpp:452-476, one storage column x four frames per thread ----
            // two columns per trip with all loads issued before the first add: the phase is bound by the latency of its dependent
            // LDS.128 (r2 capture: 42 % short_scoreboard), not by their number
            auto vn_finish = [&](const int col, const float4 sum) {
                S4[col] = sum;
                if (last) {
                    const int i = (int)__ldg(&c.quad_var_of_col[col]);
#pragma unroll
                    for (int q = 0; q < FI; q++) {
                        const float x = q4_get(sum, q);
                        if (!(x > 0.0f)) atomicOr(&dbits[q * nwords + (i >> 5)], 1u << (i & 31));
                        if (io.out_soft && f0 + q < io.n_frames) {
                            if (io.y_dtype == LDPC_GPU_DT_F64) ((double *)io.out_soft)[(size_t)(f0 + q) * N + i] = (double)x;
                            else ((float *)io.out_soft)[(size_t)(f0 + q) * N + i] = x;
                        }
                    }
                }
            };
            for (int col = tid; col < N; col += 2 * nt) {
                const int col1 = col + nt;
                const bool two = col1 < N;
                const int cb = two ? col1 : col;                                     // (a lone last column is simply read twice)
                int dv0 = DV, dv1 = DV;
                if (!REGV) { dv0 = (int)c.vn_deg[__ldg(&c.quad_var_of_col[col])]; dv1 = (int)c.vn_deg[__ldg(&c.quad_var_of_col[cb])]; }
                float4 m0[DV], m1v[DV];
                float4 s0 = y4[col], s1 = y4[cb];
#pragma unroll
                for (int s = 0; s < DV; s++) {
                    if (REGV || s < dv0) m0[s] = msg4[s * N + col];
                    if (REGV || s < dv1) m1v[s] = msg4[s * N + cb];
                }
#pragma unroll
                for (int s = 0; s < DV; s++) {                                       // nlist order
                    if (REGV || s < dv0) s0 = make_float4(s0.x + m0[s].x, s0.y + m0[s].y, s0.z + m0[s].z, s0.w + m0[s].w);
                    if (REGV || s < dv1) s1 = make_float4(s1.x + m1v[s].x, s1.y + m1v[s].y, s1.z + m1v[s].z, s1.w + m1v[s].w);
                }
                vn_finish(col, s0);
                if (two) vn_finish(col1, s1);
            }
            __syncthreads();
        }
        for (int fl = 0; fl < FI; fl++) {
            if (f0 + fl >= io.n_frames) break;                                       // uniform: dead lanes are not reported
            const uint8_t *cw = codeword_row(io, c, f0 + fl);
            finish_frame(c, p, io, f0 + fl, cw, dbits + fl * nwords, &fs[fl], p.T, -1, 0, 0, 1, -1, tot);
        }
    }
    if (tid == 0) tot.flush(io.counters);
}

} // namespace ldpc
