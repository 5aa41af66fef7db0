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
// check phase, one per variable in the variable phase.  Built for correctness and for frame-batched scaling, not tuned.
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
    double *workspace; size_t ws_stride;   // per CTA: alpha[E*q], beta[E*q], gamma[N*q]; uint8 d[N] behind them
};

template <int Q>
__global__ void nb_minmax_kernel(const NbCodeDev c, const NbIO io)
{
    constexpr int DCM = 8;
    __shared__ uint8_t s_mul[Q * Q], s_inv[Q];
    __shared__ int s_flag, s_biterr, s_symerr;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = c.N, M = c.M, m = c.m, E = M * c.dc_max;
    for (int i = tid; i < Q * Q; i += nt) s_mul[i] = c.mul[i];
    for (int i = tid; i < Q; i += nt) s_inv[i] = c.inv[i];
    double *alpha = io.workspace + (size_t)blockIdx.x * io.ws_stride, *beta = alpha + (size_t)E * Q, *gamma = beta + (size_t)E * Q;
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
                double A[DCM][Q], F[DCM][Q], B[DCM][Q];
                for (int k = 0; k < deg; k++) {
                    const int e = j * c.dc_max + k, hinv = s_inv[c.cn_val[e]];
                    for (int x = 0; x < Q; x++) A[k][x] = alpha[(size_t)e * Q + s_mul[hinv * Q + x]];
                }
                for (int x = 0; x < Q; x++) { F[0][x] = A[0][x]; B[deg - 1][x] = A[deg - 1][x]; }
                for (int k = 1; k < deg; k++)
                    for (int z = 0; z < Q; z++) {
                        double best = fmax(F[k - 1][0], A[k][z]);
                        for (int x = 1; x < Q; x++) best = fmin(best, fmax(F[k - 1][x], A[k][z ^ x]));
                        F[k][z] = best;
                    }
                for (int k = deg - 2; k >= 0; k--)
                    for (int z = 0; z < Q; z++) {
                        double best = fmax(B[k + 1][0], A[k][z]);
                        for (int x = 1; x < Q; x++) best = fmin(best, fmax(B[k + 1][x], A[k][z ^ x]));
                        B[k][z] = best;
                    }
                for (int k = 0; k < deg; k++) {
                    const int e = j * c.dc_max + k, h = c.cn_val[e];
                    for (int a = 0; a < Q; a++) {
                        const int z = s_mul[h * Q + a];
                        double out;
                        if (k == 0) out = B[1][z];
                        else if (k == deg - 1) out = F[deg - 2][z];
                        else {
                            out = fmax(F[k - 1][0], B[k + 1][z]);
                            for (int x = 1; x < Q; x++) out = fmin(out, fmax(F[k - 1][x], B[k + 1][z ^ x]));
                        }
                        beta[(size_t)e * Q + a] = out;
                    }
                }
            }
            __syncthreads();
            // ---- variable phase -----------------------------------------------------------------------------
            for (int i = tid; i < N; i += nt) {
                const int deg = c.vn_deg[i];
                double post[Q];
                for (int a = 0; a < Q; a++) post[a] = gamma[(size_t)i * Q + a];
                for (int s = 0; s < deg; s++) {
                    const int e = c.vn_edge[(size_t)i * c.dv_max + s];
                    for (int a = 0; a < Q; a++) post[a] += beta[(size_t)e * Q + a];
                }
                int best = 0;
                for (int a = 1; a < Q; a++) if (post[a] < post[best]) best = a;
                dsym[i] = (uint8_t)best;
                for (int s = 0; s < deg; s++) {
                    const int e = c.vn_edge[(size_t)i * c.dv_max + s];
                    double mn = 0.0;
                    for (int a = 0; a < Q; a++) {
                        double v = gamma[(size_t)i * Q + a];
                        for (int s2 = 0; s2 < deg; s2++) if (s2 != s) v += beta[(size_t)c.vn_edge[(size_t)i * c.dv_max + s2] * Q + a];
                        alpha[(size_t)e * Q + a] = v;
                        mn = (a == 0) ? v : fmin(mn, v);
                    }
                    for (int a = 0; a < Q; a++) alpha[(size_t)e * Q + a] -= mn;
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
