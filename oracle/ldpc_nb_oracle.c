/* ldpc_nb_oracle.c -- TEST INFRASTRUCTURE ONLY: CPU restatement of min-max decoding of non-binary GF(q) LDPC codes
 * (SURVEY.md 8(f) N5).  PARITY UNPINNED: the reference's SystemC/NB-LDPC tree does not compile and contains no min-max / EMS
 * decoder (inc/nodes.h:137 "TODO: Figure out permutation", :195-293 brute-force probability-domain check node; min_max.py:74-76
 * is an empty stub), so this file restates the published algorithm (V. Savin, "Min-Max decoding for non binary LDPC codes",
 * ISIT 2008, forward / backward form) and keeps from the reference only the code format (src/alist.cpp:23-56,97-124) and the
 * symbol <-> bit mapping (inc/nodes.h:104-108).  It is the checker of csrc/ldpc_nb_kernel.cuh and is written independently of
 * it (row-at-a-time, plain arrays); both use double arithmetic, so equality is bit for bit. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ldpc_oracle.h"

typedef struct oracle_nb_code {
    int N, M, q, m, dvm, dcm;
    int *col_deg, *row_deg, *nlist, *mlist, *mvals;      /* 0-based indices, -1 padded */
    unsigned char mul[64 * 64], inv[64];
} oracle_nb_code;

static int gf_mul_slow(int a, int b, int m, int prim)
{
    int r = 0;
    for (int k = 0; k < m; k++) { if ((b >> k) & 1) r ^= a; a <<= 1; if (a & (1 << m)) a ^= prim; }
    return r;
}

oracle_nb_code *oracle_nb_code_load_alist(const char *path)
{   /* SystemC/NB-LDPC/src/alist.cpp:23-56: `N M q`, `dv dc`, weights, N rows then M rows of (index value) pairs */
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    oracle_nb_code *c = (oracle_nb_code *)calloc(1, sizeof *c);
    if (fscanf(f, "%d %d %d %d %d", &c->N, &c->M, &c->q, &c->dvm, &c->dcm) != 5) { fclose(f); free(c); return NULL; }
    static const int prim[7] = { 0, 0x3, 0x7, 0xB, 0x13, 0x25, 0x43 };
    while ((1 << c->m) < c->q) c->m++;
    for (int a = 0; a < c->q; a++) for (int b = 0; b < c->q; b++) c->mul[a * c->q + b] = (unsigned char)gf_mul_slow(a, b, c->m, prim[c->m]);
    for (int a = 1; a < c->q; a++) for (int b = 1; b < c->q; b++) if (c->mul[a * c->q + b] == 1) c->inv[a] = (unsigned char)b;
    c->col_deg = (int *)calloc((size_t)c->N, sizeof(int)); c->row_deg = (int *)calloc((size_t)c->M, sizeof(int));
    c->nlist = (int *)malloc(sizeof(int) * (size_t)c->N * c->dvm); c->mlist = (int *)malloc(sizeof(int) * (size_t)c->M * c->dcm);
    c->mvals = (int *)malloc(sizeof(int) * (size_t)c->M * c->dcm);
    int ok = 1, v, h;
    for (int i = 0; i < c->N; i++) ok &= fscanf(f, "%d", &c->col_deg[i]) == 1;
    for (int j = 0; j < c->M; j++) ok &= fscanf(f, "%d", &c->row_deg[j]) == 1;
    for (int i = 0; i < c->N * c->dvm; i++) { ok &= fscanf(f, "%d %d", &v, &h) == 2; c->nlist[i] = v - 1; }
    for (int i = 0; i < c->M * c->dcm; i++) { ok &= fscanf(f, "%d %d", &v, &h) == 2; c->mlist[i] = v - 1; c->mvals[i] = h; }
    fclose(f);
    if (!ok) { free(c); return NULL; }
    return c;
}

void oracle_nb_code_free(oracle_nb_code *c)
{
    if (!c) return;
    free(c->col_deg); free(c->row_deg); free(c->nlist); free(c->mlist); free(c->mvals); free(c);
}

/* one frame; y: N*m bit samples (bit b of symbol i at i*m + b, bit 0 <-> +1).  Returns iterations executed. */
static int nb_frame(const oracle_nb_code *c, int T, const double *y, int *d, int *satisfied)
{
    const int N = c->N, M = c->M, q = c->q, m = c->m, dcm = c->dcm, dvm = c->dvm;
    double *gamma = (double *)malloc(sizeof(double) * (size_t)N * q);
    double *alpha = (double *)malloc(sizeof(double) * (size_t)M * dcm * q), *beta = (double *)malloc(sizeof(double) * (size_t)M * dcm * q);
    for (int i = 0; i < N; i++) {                                  /* channel costs: disagreement with the hard decisions */
        int hd = 0;
        for (int b = 0; b < m; b++) if (!(y[i * m + b] > 0)) hd |= 1 << b;
        for (int a = 0; a < q; a++) {
            double g = 0.0;
            for (int b = 0; b < m; b++) if (((a ^ hd) >> b) & 1) g += fabs(y[i * m + b]);
            gamma[i * q + a] = g;
        }
        d[i] = hd;
    }
    for (int j = 0; j < M; j++) for (int k = 0; k < c->row_deg[j]; k++)
        for (int a = 0; a < q; a++) alpha[((size_t)j * dcm + k) * q + a] = gamma[c->mlist[j * dcm + k] * q + a];
    int it = 0;
    for (;;) {
        int ok = 1;
        for (int j = 0; j < M; j++) {
            int syn = 0;
            for (int k = 0; k < c->row_deg[j]; k++) syn ^= c->mul[c->mvals[j * dcm + k] * q + d[c->mlist[j * dcm + k]]];
            if (syn) ok = 0;
        }
        *satisfied = ok;
        if (ok || it >= T) break;
        for (int j = 0; j < M; j++) {                              /* check node: forward / backward (min, max)-convolutions */
            const int deg = c->row_deg[j];
            double A[8][64], F[8][64], B[8][64];
            for (int k = 0; k < deg; k++) {
                const int hinv = c->inv[c->mvals[j * dcm + k]];
                for (int x = 0; x < q; x++) A[k][x] = alpha[((size_t)j * dcm + k) * q + c->mul[hinv * q + x]];
            }
            memcpy(F[0], A[0], sizeof(double) * q); memcpy(B[deg - 1], A[deg - 1], sizeof(double) * q);
            for (int k = 1; k < deg; k++) for (int z = 0; z < q; z++) {
                double best = INFINITY;
                for (int x = 0; x < q; x++) { double v = fmax(F[k - 1][x], A[k][z ^ x]); if (v < best) best = v; }
                F[k][z] = best;
            }
            for (int k = deg - 2; k >= 0; k--) for (int z = 0; z < q; z++) {
                double best = INFINITY;
                for (int x = 0; x < q; x++) { double v = fmax(B[k + 1][x], A[k][z ^ x]); if (v < best) best = v; }
                B[k][z] = best;
            }
            for (int k = 0; k < deg; k++) {
                const int h = c->mvals[j * dcm + k];
                for (int a = 0; a < q; a++) {
                    const int z = c->mul[h * q + a];
                    double out;
                    if (k == 0) out = B[1][z];
                    else if (k == deg - 1) out = F[deg - 2][z];
                    else { out = INFINITY; for (int x = 0; x < q; x++) { double v = fmax(F[k - 1][x], B[k + 1][z ^ x]); if (v < out) out = v; } }
                    beta[((size_t)j * dcm + k) * q + a] = out;
                }
            }
        }
        for (int i = 0; i < N; i++) {                              /* variable node: sums in nlist order */
            const int deg = c->col_deg[i];
            int ed[16];
            for (int s = 0; s < deg; s++) {
                const int j = c->nlist[i * dvm + s];
                ed[s] = -1;
                for (int k = 0; k < c->row_deg[j]; k++) if (c->mlist[j * dcm + k] == i) ed[s] = j * dcm + k;
            }
            int best = 0; double pbest = 0.0;
            for (int a = 0; a < q; a++) {
                double pa = gamma[i * q + a];
                for (int s = 0; s < deg; s++) pa += beta[(size_t)ed[s] * q + a];
                if (a == 0 || pa < pbest) { pbest = pa; best = a; }
            }
            d[i] = best;
            for (int s = 0; s < deg; s++) {
                double mn = 0.0;
                for (int a = 0; a < q; a++) {
                    double v = gamma[i * q + a];
                    for (int s2 = 0; s2 < deg; s2++) if (s2 != s) v += beta[(size_t)ed[s2] * q + a];
                    alpha[(size_t)ed[s] * q + a] = v;
                    if (a == 0 || v < mn) mn = v;
                }
                for (int a = 0; a < q; a++) alpha[(size_t)ed[s] * q + a] -= mn;
            }
        }
        it++;
    }
    free(gamma); free(alpha); free(beta);
    return it;
}

/* counters: [0] bit errors, [1] total bits, [2] words, [3] word errors, [4] iterations, [5] undetected, [6] symbol errors */
int oracle_nb_decode(const oracle_nb_code *c, int T, int64_t n_frames, const double *y, uint8_t *out_symbols, int32_t *out_iters, int64_t *cnt)
{
    int *d = (int *)malloc(sizeof(int) * (size_t)c->N);
    for (int64_t f = 0; f < n_frames; f++) {
        int sat = 0;
        const int it = nb_frame(c, T, y + (size_t)f * c->N * c->m, d, &sat);
        int be = 0, se = 0;
        for (int i = 0; i < c->N; i++) { be += __builtin_popcount((unsigned)d[i]); se += d[i] != 0; if (out_symbols) out_symbols[(size_t)f * c->N + i] = (uint8_t)d[i]; }
        if (out_iters) out_iters[f] = it;
        if (cnt) { cnt[1] += (int64_t)c->N * c->m; cnt[2]++; cnt[4] += it; if (be) { cnt[0] += be; cnt[3]++; cnt[6] += se; if (sat) cnt[5]++; } }
    }
    free(d);
    return 0;
}

int oracle_nb_simulate(const oracle_nb_code *c, int T, double snr_db, double R, uint64_t seed, int64_t frame_begin, int64_t n_frames, int64_t *cnt)
{
    const double sigma = sqrt(pow(10.0, -snr_db / 10.0) / R / 2.0);
    const int L = c->N * c->m;
    double *y = (double *)malloc(sizeof(double) * (size_t)L);
    for (int64_t f = 0; f < n_frames; f++) {
        for (int blk = 0; blk * 4 < L; blk++) {
            float n4[4]; oracle_normal4(seed, (uint64_t)(frame_begin + f), (uint32_t)blk, 0, 0, n4);
            for (int qx = 0; qx < 4 && blk * 4 + qx < L; qx++) y[blk * 4 + qx] = 1.0 + sigma * (double)n4[qx];
        }
        oracle_nb_decode(c, T, 1, y, NULL, NULL, cnt);
    }
    free(y);
    return 0;
}

int oracle_nb_dims(const oracle_nb_code *c, int *N, int *M, int *q, int *m) { *N = c->N; *M = c->M; *q = c->q; *m = c->m; return 0; }
