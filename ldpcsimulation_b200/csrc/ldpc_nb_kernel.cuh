// ldpc_nb_kernel.cuh -- non-binary GF(q) LDPC, min-max decoding (BASELINE.json configs[4]).
//
// The reference's SystemC/NB-LDPC tree is unfinished (it does not compile, its check node enumerates all q^dc
// combinations in the probability domain and its symbol node has no edge-value permutation: inc/nodes.h:137,195-293;
// min_max.py:74-76 is an empty stub), so there is nothing to be bit-compatible WITH: parity is UNPINNED.  What is kept from the
// reference: the code format (src/alist.cpp:23-56,97-124: header `N M q`, (index, value) pairs) and the symbol <-> bit
// mapping of its symbol node (bit b of the integer value, least significant first, inc/nodes.h:104-108).  The decoder is the
// textbook min-max algorithm (Savin, "Min-Max decoding for non binary LDPC codes", ISIT 2008) in its forward / backward form:
//   channel   gamma_i[a] = sum over the bits b of a that disagree with the hard decision of sample y_(i,b) of |y_(i,b)|
//   check j   A_k[x] = alpha_k[h_k^-1 x];  F, B = running (min, max)-convolutions over GF(q) addition,
//             (U * V)[z] = min over x of max(U[x], V[z + x]);  beta_k[a] = (F_(k-1) * B_(k+1))[h_k a]
//   variable  alpha_k[a] = gamma[a] + sum_(k' != k) beta_k'[a]  minus its minimum;  decision = first argmin of gamma + sum beta
//   stop      when every check sum_k h_k d_k vanishes, else after T iterations
// GF(2^m) elements are the integers 0 .. q-1 in the polynomial basis (addition = XOR), multiplication modulo the primitive
// polynomials x^2+x+1, x^3+x+1, x^4+x+1, x^5+x^2+1, x^6+x+1.  All arithmetic is double, in the order of the C restatement in
// the test oracle, which this kernel equals bit for bit.
//
// One CTA per frame; messages live in an HBM / L2 workspace slice per CTA (E q doubles each way), one thread per check in the
// check phase, q lanes per variable in the variable phase.  The check update keeps its three q-vectors (the two operands and the result
// of one (min, max)-convolution) in registers with compile-time indices -- the first cut indexed A / F / B[8][q] dynamically, i.e.
// in local memory: 0.27 Gbit/s on the GF(16) code -- and stores only the forward / backward vectors the outputs need
// (F_1 .. F_(dc-2), B_(dc-2) .. B_1: 3 (dc - 2) convolutions per check instead of 3 dc - 4; the skipped ones feed no output).
#pragma once
#include "ldpc_common.cuh"

namespace ldpc {

struct NbCodeDev {
    int N, M, q, m, E, dv_max, dc_max;
    const int32_t *cn_var;    // [M][dc_max] variable of slot k of check j, -1 padded
    const uint8_t *cn_val;    // [M][dc_max] h_jk
    const uint8_t *cn_deg;    // [M]
    const int32_t *vn_edge;   // [N][dv_max] edge id (j * dc_max + k) of slot s of variable i, -1 padded
    const uint8_t *vn_deg;    // [N]
    const uint8_t *mul;       // [q][q] product table
    const uint8_t *inv;       // [q] inverses (inv[0] = 0)
};

struct NbIO {
    long long n_frames, frame_begin;
    const double *y;          // device [F][N*m] bit samples, or NULL -> Philox channel
    double sigma;
    unsigned long long seed;
    int T;
    uint8_t *out_symbols;     // device [F][N] or NULL
    int *out_iters;           // device [F] or NULL
    unsigned long long *counters;   // CNT_* layout; errors = BIT errors
    double *workspace; size_t ws_stride;   // per CTA: alpha[E*q], beta[E*q], fwd[E*q], bwd[E*q], gamma[N*q]; uint8 d[N] behind them
};

// W[z] = min over x of max(U[x], V[z ^ x]), in the restatement's order (x = 0 first)
template <int Q>
LDPC_DEVINL void nb_conv(const double (&U)[Q], const double (&V)[Q], double (&W)[Q])
{
#pragma unroll
    for (int z = 0; z < Q; z++) {
        double best = fmax(U[0], V[z]);
#pragma unroll
        for (int x = 1; x < Q; x++) best = fmin(best, fmax(U[x], V[z ^ x]));
        W[z] = best;
    }
}

template <int Q>
__global__ void __launch_bounds__(256, 2) nb_minmax_kernel(const NbCodeDev c, const NbIO io)
{
    __shared__ uint8_t s_mul[Q * Q], s_inv[Q];
    __shared__ int s_flag, s_biterr, s_symerr;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = c.N, M = c.M, m = c.m, E = M * c.dc_max;
    for (int i = tid; i < Q * Q; i += nt) s_mul[i] = c.mul[i];
    for (int i = tid; i < Q; i += nt) s_inv[i] = c.inv[i];
    double *alpha = io.workspace + (size_t)blockIdx.x * io.ws_stride, *beta = alpha + (size_t)E * Q, *fwd = beta + (size_t)E * Q, *bwd = fwd + (size_t)E * Q,
           *gamma = bwd + (size_t)E * Q;
    uint8_t *dsym = reinterpret_cast<uint8_t *>(gamma + (size_t)N * Q);
    unsigned long long tot[CNT_N];
    for (int k = 0; k < CNT_N; k++) tot[k] = 0ull;
    __syncthreads();

    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        // ---- channel: bit samples -> symbol cost vectors, v->c messages, hard decisions --------------------
        for (int i = tid; i < N; i += nt) {
            double ab[6]; int hd = 0;
            for (int b = 0; b < m; b++) {
                double yb;
                if (io.y) yb = io.y[((size_t)f * N + i) * m + b];
                else {                                                 // all-zero codeword: x = +1, y = 1 + sigma n
                    const unsigned long long idx = (unsigned long long)i * m + b;
                    float n4[4]; normal4(io.seed, (unsigned long long)(io.frame_begin + f), (uint32_t)(idx >> 2), 0u, STREAM_CHANNEL, n4);
                    yb = __dadd_rn(1.0, __dmul_rn(io.sigma, (double)n4[idx & 3]));
                }
                ab[b] = fabs(yb);
                if (!(yb > 0)) hd |= 1 << b;
            }
            for (int a = 0; a < Q; a++) {
                double g = 0.0;
                for (int b = 0; b < m; b++) if (((a ^ hd) >> b) & 1) g += ab[b];
                gamma[(size_t)i * Q + a] = g;
            }
            dsym[i] = (uint8_t)hd;
            const int deg = c.vn_deg[i];
            for (int s = 0; s < deg; s++) {
                const int e = c.vn_edge[(size_t)i * c.dv_max + s];
                for (int a = 0; a < Q; a++) alpha[(size_t)e * Q + a] = gamma[(size_t)i * Q + a];
            }
        }
        __syncthreads();
        int it = 0, ok = 0;
        for (;;) {
            // ---- syndrome of the current decisions ----------------------------------------------------------
            if (tid == 0) s_flag = 0;
            __syncthreads();
            int bad = 0;
            for (int j = tid; j < M; j += nt) {
                int syn = 0;
                const int deg = c.cn_deg[j];
                for (int k = 0; k < deg; k++) syn ^= s_mul[c.cn_val[(size_t)j * c.dc_max + k] * Q + dsym[c.cn_var[(size_t)j * c.dc_max + k]]];
                bad |= syn;
            }
            if (bad) s_flag = 1;
            __syncthreads();
            ok = !s_flag;
            if (ok || it >= io.T) break;
            // ---- check phase --------------------------------------------------------------------------------
            for (int j = tid; j < M; j += nt) {
                const int deg = c.cn_deg[j];
                const size_t e0 = (size_t)j * c.dc_max;
                auto load_A = [&](const int k, double (&A)[Q]) {                   // A_k[x] = alpha_k[h_k^-1 x]
                    const int hinv = s_inv[c.cn_val[e0 + k]];
#pragma unroll
                    for (int x = 0; x < Q; x++) A[x] = alpha[(e0 + k) * Q + s_mul[hinv * Q + x]];
                };
                auto load_V = [&](const double *src, const int k, double (&A)[Q]) {
#pragma unroll
                    for (int x = 0; x < Q; x++) A[x] = src[(e0 + k) * Q + x];
                };
                auto store_V = [&](double *dst, const int k, const double (&A)[Q]) {
#pragma unroll
                    for (int x = 0; x < Q; x++) dst[(e0 + k) * Q + x] = A[x];
                };
                auto emit = [&](const int k, const double (&out)[Q]) {              // beta_k[a] = out[h_k a]
                    const int hinv = s_inv[c.cn_val[e0 + k]];
#pragma unroll
                    for (int z = 0; z < Q; z++) beta[(e0 + k) * Q + s_mul[hinv * Q + z]] = out[z];
                };
                double U[Q], V[Q], W[Q];
                load_A(0, U);                                                       // F_0 = A_0
                for (int k = 1; k <= deg - 2; k++) { load_A(k, V); nb_conv<Q>(U, V, W); store_V(fwd, k, W);
#pragma unroll
                    for (int x = 0; x < Q; x++) U[x] = W[x]; }
                emit(deg - 1, U);                                                   // beta_(dc-1) = F_(dc-2)
                load_A(deg - 1, U);                                                 // B_(dc-1) = A_(dc-1)
                for (int k = deg - 2; k >= 1; k--) { load_A(k, V); nb_conv<Q>(U, V, W); store_V(bwd, k, W);
#pragma unroll
                    for (int x = 0; x < Q; x++) U[x] = W[x]; }
                emit(0, U);                                                         // beta_0 = B_1
                for (int k = 1; k <= deg - 2; k++) {                                // beta_k = F_(k-1) * B_(k+1)
                    if (k - 1 == 0) load_A(0, U); else load_V(fwd, k - 1, U);
                    if (k + 1 == deg - 1) load_A(deg - 1, V); else load_V(bwd, k + 1, V);
                    nb_conv<Q>(U, V, W);
                    emit(k, W);
                }
                // (a single call site for the convolution -- one loop over the three kinds of step -- cuts the 14.5 k instructions of this
                // phase to a third, but ptxas then keeps the q-vectors in local memory: 1.8 x slower, measured)
            }
            __syncthreads();
            // ---- variable phase -----------------------------------------------------------------------------
            // Q lanes per variable (lane a owns symbol value a): the q-vectors of an edge are read and written as one coalesced line,
            // the minimum / first argmin over a are shuffle reductions inside the group of Q lanes (exact: min, and ties -> lowest a)
            for (int base = 0; base < N * Q; base += nt) {
                const int t = base + tid;
                const bool valid = t < N * Q;
                const int i = valid ? t / Q : 0, a = t % Q;
                const int deg = valid ? (int)c.vn_deg[i] : 0;
                const double g = gamma[(size_t)i * Q + a];
                double post = g;
                for (int s = 0; s < deg; s++) post += beta[(size_t)c.vn_edge[(size_t)i * c.dv_max + s] * Q + a];
                double bv = post; int bi = a;
#pragma unroll
                for (int o = Q / 2; o; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (valid && a == 0) dsym[i] = (uint8_t)bi;
                for (int s = 0; s < c.dv_max; s++) {                                 // (uniform trip count: the shuffles below need every lane)
                    const bool on = s < deg;
                    double v = g;
                    for (int s2 = 0; s2 < deg; s2++) if (s2 != s && on) v += beta[(size_t)c.vn_edge[(size_t)i * c.dv_max + s2] * Q + a];
                    double mn = v;
#pragma unroll
                    for (int o = Q / 2; o; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    if (on) alpha[(size_t)c.vn_edge[(size_t)i * c.dv_max + s] * Q + a] = v - mn;
                }
            }
            it++;
            __syncthreads();
        }
        // ---- accounting (all-zero codeword): bit and frame errors ----------------------------------------------
        if (tid == 0) { s_biterr = 0; s_symerr = 0; }
        __syncthreads();
        int be = 0, se = 0;
        for (int i = tid; i < N; i += nt) {
            const int d = dsym[i];
            be += __popc((unsigned)d); se += d != 0;
            if (io.out_symbols) io.out_symbols[(size_t)f * N + i] = (uint8_t)d;
        }
        if (be) { atomicAdd(&s_biterr, be); atomicAdd(&s_symerr, se); }
        __syncthreads();
        if (tid == 0) {
            if (io.out_iters) io.out_iters[f] = it;
            tot[CNT_BITS] += (unsigned long long)N * m; tot[CNT_WORDS] += 1ull; tot[CNT_ITERS] += (unsigned long long)it;
            if (s_biterr > 0) { tot[CNT_ERRORS] += (unsigned long long)s_biterr; tot[CNT_WORDERRS] += 1ull; if (ok) tot[CNT_UNDETECTED] += 1ull; }
            tot[CNT_SMOOTH] += (unsigned long long)s_symerr;                 // (symbol errors ride in the otherwise unused smoothing slot)
        }
        __syncthreads();
    }
    if (tid == 0 && io.counters) for (int k = 0; k < CNT_N; k++) if (tot[k]) atomicAdd(&io.counters[k], tot[k]);
}

} // namespace ldpc
