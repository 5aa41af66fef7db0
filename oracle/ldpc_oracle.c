/* ldpc_oracle.c -- CPU restatement of the reference decode loop.  TEST INFRASTRUCTURE ONLY.
 *
 * See ldpc_oracle.h for who may use this and how it is pinned to the reference.
 * Reference paths are relative to /root/reference/C_implementations/.
 *
 * Layout differs from the reference on purpose: flat arrays and an edge permutation
 * computed once (the reference recomputes it with find(), src/decodeMinSum.cpp:527-536,
 * on every access).  The arithmetic, its order, and every tie rule follow the reference.
 * Compile with -O2 -ffp-contract=off (no fused multiply-add: the reference is built for
 * baseline x86-64, which has none).
 */
#include "ldpc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

static char g_err[256];
const char *oracle_last_error(void) { return g_err; }
static int fail(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return LDPC_GPU_ERR_INVALID_ARG; }

/* ------------------------------------------------------------------------------------ */
/* Code structure (a1: inc/alist.h:21-36, src/alist.cpp:22-95)                           */
/* ------------------------------------------------------------------------------------ */
void oracle_code_free(oracle_code *c)
{
    if (!c) return;
    free(c->col_deg); free(c->row_deg); free(c->nlist); free(c->mlist); free(c->vn_slot); free(c->cn_slot);
    free(c);
}

/* Same inputs as loadFile()'s outputs: 1-based indices, rows padded with zeros to the
 * biggest weight.  Builds the reverse slot maps the reference derives with find();
 * like find() (last match wins) a duplicated entry resolves to its last occurrence. */
oracle_code *oracle_code_create(int N, int M, int dvm, int dcm,
                                const int *num_nlist, const int *num_mlist,
                                const int *nlist_flat, const int *mlist_flat)
{
    if (N <= 0 || M <= 0 || dvm <= 0 || dcm <= 0) { fail("bad dimensions"); return NULL; }
    oracle_code *c = (oracle_code *)calloc(1, sizeof *c);
    c->N = N; c->M = M; c->dv_max = dvm; c->dc_max = dcm;
    c->col_deg = (int *)malloc(sizeof(int) * N);
    c->row_deg = (int *)malloc(sizeof(int) * M);
    c->nlist = (int *)malloc(sizeof(int) * (size_t)N * dvm);
    c->mlist = (int *)malloc(sizeof(int) * (size_t)M * dcm);
    c->cn_slot = (int *)malloc(sizeof(int) * (size_t)N * dvm);
    c->vn_slot = (int *)malloc(sizeof(int) * (size_t)M * dcm);
    int E = 0;
    for (int i = 0; i < N; i++) {
        c->col_deg[i] = num_nlist[i];
        if (num_nlist[i] < 0 || num_nlist[i] > dvm) { fail("column weight out of range"); oracle_code_free(c); return NULL; }
        for (int s = 0; s < dvm; s++) {
            int v = (s < num_nlist[i]) ? nlist_flat[(size_t)i * dvm + s] - 1 : -1;
            if (s < num_nlist[i] && (v < 0 || v >= M)) { fail("nlist index out of range"); oracle_code_free(c); return NULL; }
            c->nlist[(size_t)i * dvm + s] = v;
            c->cn_slot[(size_t)i * dvm + s] = -1;
        }
    }
    for (int j = 0; j < M; j++) {
        c->row_deg[j] = num_mlist[j];
        if (num_mlist[j] < 0 || num_mlist[j] > dcm) { fail("row weight out of range"); oracle_code_free(c); return NULL; }
        E += num_mlist[j];
        for (int k = 0; k < dcm; k++) {
            int v = (k < num_mlist[j]) ? mlist_flat[(size_t)j * dcm + k] - 1 : -1;
            if (k < num_mlist[j] && (v < 0 || v >= N)) { fail("mlist index out of range"); oracle_code_free(c); return NULL; }
            c->mlist[(size_t)j * dcm + k] = v;
            c->vn_slot[(size_t)j * dcm + k] = -1;
        }
    }
    c->E = E;
    /* reverse maps: find(H.nlist[snode], ..., j) and find(H.mlist[cnode], ..., i) */
    for (int j = 0; j < M; j++)
        for (int k = 0; k < c->row_deg[j]; k++) {
            int i = c->mlist[(size_t)j * dcm + k], slot = -1;
            for (int s = 0; s < c->col_deg[i]; s++) if (c->nlist[(size_t)i * dvm + s] == j) slot = s;
            if (slot < 0) { fail("mlist entry has no matching nlist entry"); oracle_code_free(c); return NULL; }
            c->vn_slot[(size_t)j * dcm + k] = slot;
        }
    for (int i = 0; i < N; i++)
        for (int s = 0; s < c->col_deg[i]; s++) {
            int j = c->nlist[(size_t)i * dvm + s], slot = -1;
            for (int k = 0; k < c->row_deg[j]; k++) if (c->mlist[(size_t)j * dcm + k] == i) slot = k;
            if (slot < 0) { fail("nlist entry has no matching mlist entry"); oracle_code_free(c); return NULL; }
            c->cn_slot[(size_t)i * dvm + s] = slot;
        }
    return c;
}

/* alist text -> arrays.  Token semantics of src/alist.cpp:71-91 + src/r.cpp:277-302,448-464
 * (the default, zero-padded branch), extended to accept the unpadded layout of the
 * -DCPPSTYLE branch (src/alist.cpp:26-62): rows are read line by line. */
oracle_code *oracle_code_load_alist(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) { fail("cannot open alist file"); return NULL; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 2);
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); fail("short read"); return NULL; }
    fclose(f); buf[sz] = '\n'; buf[sz + 1] = 0;
    /* split into non-empty lines of integers */
    size_t nlines = 0, cap = 1024;
    char **lines = (char **)malloc(cap * sizeof *lines);
    for (char *p = buf; *p;) {
        char *e = strchr(p, '\n'); *e = 0;
        char *q = p; while (*q && isspace((unsigned char)*q)) q++;
        if (*q) { if (nlines == cap) { cap *= 2; lines = (char **)realloc(lines, cap * sizeof *lines); } lines[nlines++] = q; }
        p = e + 1;
    }
    oracle_code *code = NULL;
    int *num_n = NULL, *num_m = NULL, *nl = NULL, *ml = NULL;
    int N = 0, M = 0, dvm = 0, dcm = 0;
    if (nlines < 4 || sscanf(lines[0], "%d %d", &N, &M) != 2 || sscanf(lines[1], "%d %d", &dvm, &dcm) != 2
        || N <= 0 || M <= 0 || dvm <= 0 || dcm <= 0) { fail("bad alist header"); goto done; }
    if (nlines != (size_t)(4 + N + M)) { fail("alist line count does not match header (transposed or truncated file?)"); goto done; }
    num_n = (int *)calloc((size_t)N, sizeof(int)); num_m = (int *)calloc((size_t)M, sizeof(int));
    nl = (int *)calloc((size_t)N * dvm, sizeof(int)); ml = (int *)calloc((size_t)M * dcm, sizeof(int));
    {
        char *p = lines[2]; for (int i = 0; i < N; i++) num_n[i] = (int)strtol(p, &p, 10);
        p = lines[3];       for (int j = 0; j < M; j++) num_m[j] = (int)strtol(p, &p, 10);
    }
    for (int i = 0; i < N; i++) {
        char *p = lines[4 + i]; int cnt = 0;
        for (;;) { char *e; long v = strtol(p, &e, 10); if (e == p) break; p = e;
                   if (v == 0) continue; if (cnt >= dvm) { fail("column has more entries than biggest_num_n"); goto done; }
                   nl[(size_t)i * dvm + cnt++] = (int)v; }
        if (cnt != num_n[i]) { fail("column weight does not match its entry count"); goto done; }
    }
    for (int j = 0; j < M; j++) {
        char *p = lines[4 + N + j]; int cnt = 0;
        for (;;) { char *e; long v = strtol(p, &e, 10); if (e == p) break; p = e;
                   if (v == 0) continue; if (cnt >= dcm) { fail("row has more entries than biggest_num_m"); goto done; }
                   ml[(size_t)j * dcm + cnt++] = (int)v; }
        if (cnt != num_m[j]) { fail("row weight does not match its entry count"); goto done; }
    }
    code = oracle_code_create(N, M, dvm, dcm, num_n, num_m, nl, ml);
done:
    free(num_n); free(num_m); free(nl); free(ml); free(lines); free(buf);
    return code;
}

/* ------------------------------------------------------------------------------------ */
/* Scalar helpers                                                                        */
/* ------------------------------------------------------------------------------------ */
/* src/decodeMinSum.cpp:518-523 (also BP :412-417, DDBMP :426-431): sgn(0) = +1 */
static double sgn_ge(double x) { return (x >= 0.0) ? 1.0 : -1.0; }
/* src/decodeGDBF.cpp:495-501 (also NGDBFhw.cpp): sgn(0) = -1 */
static double sgn_gt(double x) { return (x > 0) ? 1.0 : -1.0; }

/* a3: MS / DD-BMP sample quantiser, src/decodeMinSum.cpp:480-489, src/decodeDDBMP.cpp:434-443 */
double oracle_quantize_ms(double x, double Ymax, double Nq)
{
    if (fabs(x) > Ymax) return sgn_ge(x) * Ymax;
    double qval = sgn_ge(x) * (floor(fabs(x) * (Nq - 1) / (2.0 * Ymax)) + 0.0) * (2 * Ymax / (Nq - 1));
    if (qval == 0.0) qval = sgn_ge(x) * 2.0 * Ymax / (Nq - 1);
    return qval;
}

/* a3: GDBF sample quantiser, src/decodeGDBF.cpp:488-493 */
double oracle_quantize_gdbf(double x, double Ymax, int NQ)
{
    double qmax = pow(2, (NQ - 1));
    double lmax = Ymax / 2.0;
    return sgn_gt(x) * floor((fabs(x) * qmax) / (2 * lmax) + 0.5) * (2.0 * lmax / qmax);
}

/* a3/a17: NGDBFhw 5-bit sign-magnitude code of a pre-scaled sample: quantize(double) + pack(),
 * src/NGDBFhw.cpp:639-663.  NQ = 5 is a compile-time constant there (:56). */
int oracle_hw_pack(double ymod, double Ymax, double w)
{
    const int NQ = 5;
    double qmax = pow(2, (NQ));
    double lmax = Ymax / (2.0 * w);
    double NL = qmax - 1;
    int yq = (int)(sgn_gt(ymod) * round(floor(fabs(ymod) * NL / (2 * lmax))));
    int sign = (int)sgn_gt(ymod);
    unsigned mag = (unsigned)abs(yq) & ((1u << NQ) - 1u);   /* bitset<NQ>(abs(sample)) keeps NQ bits */
    if (sign < 0) mag |= 1u << (NQ - 1);
    return (int)mag;
}

/* a17: unpack(), src/NGDBFhw.cpp:665-677: odd-integer reconstruction +-(2k+1) */
int oracle_hw_unpack(int code)
{
    const int NQ = 5;
    unsigned b = ((unsigned)code << 1) & ((1u << (NQ + 1)) - 1u);   /* bitset<NQ+1> b(sample); b <<= 1 */
    b |= 1u;
    if (b & (1u << NQ)) { b &= ~(1u << NQ); return -(int)b; }
    return (int)b;
}

static int count_errors(const int *d, const int *c, int N)
{   /* a19: countDecisionErrors, src/decodeMinSum.cpp:382-393 */
    int e = 0;
    for (int i = 0; i < N; i++) if (d[i] != c[i]) e++;
    return e;
}

/* ------------------------------------------------------------------------------------ */
/* Per-frame workspaces and result                                                       */
/* ------------------------------------------------------------------------------------ */
typedef struct frame_result {
    int it;            /* `it` as accounted by the reference (summed over phases for RNGDBF)   */
    int satisfied;     /* the reference's `satisfied` flag where it has one, else H.d == 0     */
    int smoothed;      /* smoothing vote replaced d                                            */
    int smoothing_used;/* increments of smoothingUsed (one per phase at most)                  */
    int phases;        /* phases used (RNGDBF), else 1                                         */
    int uncoded;       /* increments of uncodedErrors                                          */
    int errors;        /* newErrors / leastErrors                                              */
} frame_result;

typedef struct workspace {
    double *yq, *sum, *v2c, *c2v, *mem, *theta, *pert, *shape, *Emet;
    int *d, *r, *c, *c2s, *dsum;
} workspace;

static void ws_alloc(workspace *w, const oracle_code *H)
{
    size_t N = (size_t)H->N, M = (size_t)H->M;
    w->yq = (double *)calloc(N, sizeof(double));     w->sum = (double *)calloc(N, sizeof(double));
    w->v2c = (double *)calloc(N * H->dv_max, sizeof(double)); w->c2v = (double *)calloc(M * H->dc_max, sizeof(double));
    w->mem = (double *)calloc(N * H->dv_max, sizeof(double));
    w->theta = (double *)calloc(N, sizeof(double));  w->pert = (double *)calloc(N, sizeof(double));
    w->shape = (double *)calloc(N, sizeof(double));  w->Emet = (double *)calloc(N, sizeof(double));
    w->d = (int *)calloc(N, sizeof(int)); w->r = (int *)calloc(N, sizeof(int)); w->c = (int *)calloc(N, sizeof(int));
    w->c2s = (int *)calloc(M, sizeof(int)); w->dsum = (int *)calloc(N, sizeof(int));
}
static void ws_free(workspace *w)
{
    free(w->yq); free(w->sum); free(w->v2c); free(w->c2v); free(w->mem); free(w->theta); free(w->pert);
    free(w->shape); free(w->Emet); free(w->d); free(w->r); free(w->c); free(w->c2s); free(w->dsum);
}

/* syndrome of bipolar decisions: 1 if every parity product is +1 */
static int all_checks_ok_bipolar(const oracle_code *H, const int *d)
{
    for (int j = 0; j < H->M; j++) {
        int prod = 1;
        for (int k = 0; k < H->row_deg[j]; k++) prod *= d[H->mlist[(size_t)j * H->dc_max + k]];
        if (prod < 0) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------ */
/* A.2 min-sum family: src/decodeMinSum.cpp                                              */
/* ------------------------------------------------------------------------------------ */
static void ms_frame(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, const double *y, workspace *w, frame_result *res)
{
    const int N = H->N, M = H->M, dvm = H->dv_max, dcm = H->dc_max;
    const double Nq = pow(2.0, cfg->Q);                                   /* :125 */
    res->uncoded = 0;
    for (int i = 0; i < N; i++) {                                         /* channel conditioning :214-238 */
        double yq = (cfg->flags & LDPC_GPU_F_QUANTIZE_SAMPLES) ? oracle_quantize_ms(y[i], cfg->Ymax, Nq) : y[i];
        if (cfg->flags & LDPC_GPU_F_SATURATE_SAMPLES) { if (yq > cfg->Ymax) yq = cfg->Ymax; if (yq < -cfg->Ymax) yq = -cfg->Ymax; }
        w->yq[i] = yq;
        w->r[i] = (yq > 0) ? 1 : -1;
        w->d[i] = w->r[i];
        if (w->r[i] * w->c[i] < 0) res->uncoded++;
        w->sum[i] = yq;
        for (int s = 0; s < H->col_deg[i]; s++) w->v2c[(size_t)i * dvm + s] = yq;   /* initializeSymMessages :364-370 */
    }
    int it;
    for (it = 0; it < cfg->num_iterations; it++) {                        /* :247-263 */
        for (int j = 0; j < M; j++) {                                     /* checkNodeUpdates :410-450 */
            double minMag = INFINITY, minMag2 = INFINITY, prod = 1.0; int minIdx = -1;
            const int deg = H->row_deg[j];
            for (int k = 0; k < deg; k++) {
                int i = H->mlist[(size_t)j * dcm + k];
                double msg = w->v2c[(size_t)i * dvm + H->vn_slot[(size_t)j * dcm + k]];
                prod *= sgn_ge(msg);
                if (fabs(msg) <= minMag) { minMag2 = minMag; minMag = fabs(msg); minIdx = k; }
                else if (fabs(msg) < minMag2) minMag2 = fabs(msg);
            }
            for (int k = 0; k < deg; k++) {
                int i = H->mlist[(size_t)j * dcm + k];
                double msg = w->v2c[(size_t)i * dvm + H->vn_slot[(size_t)j * dcm + k]];
                double out = (k == minIdx) ? prod * minMag2 * sgn_ge(msg) : prod * minMag * sgn_ge(msg);
                if (cfg->flags & LDPC_GPU_F_NORMALIZED_MS) out /= cfg->alpha;                  /* applyNormalization :494-499 */
                if (cfg->flags & LDPC_GPU_F_OFFSET_MS) {                                       /* applyOffset :503-515 */
                    double mag = fabs(out) - cfg->delta;
                    out = (mag > 0) ? sgn_ge(out) * mag : 0;
                }
                w->c2v[(size_t)j * dcm + k] = out;
            }
        }
        for (int i = 0; i < N; i++) {                                     /* symNodeUpdates :452-476 */
            double sum = w->yq[i];
            for (int s = 0; s < H->col_deg[i]; s++) {
                int j = H->nlist[(size_t)i * dvm + s];
                sum += w->c2v[(size_t)j * dcm + H->cn_slot[(size_t)i * dvm + s]];
            }
            for (int s = 0; s < H->col_deg[i]; s++) {
                int j = H->nlist[(size_t)i * dvm + s];
                w->v2c[(size_t)i * dvm + s] = sum - w->c2v[(size_t)j * dcm + H->cn_slot[(size_t)i * dvm + s]];
            }
            w->sum[i] = sum;
            w->d[i] = (sum > 0) ? 1 : -1;
        }
    }
    res->it = it;
    res->satisfied = all_checks_ok_bipolar(H, w->d);   /* extension: the reference MS keeps no syndrome */
    res->smoothed = 0; res->smoothing_used = 0; res->phases = 1;
}

/* ------------------------------------------------------------------------------------ */
/* A.3 sum-product: src/decodeBP.cpp                                                     */
/* ------------------------------------------------------------------------------------ */
static void bp_frame(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, double N0, const double *y, workspace *w, frame_result *res)
{
    const int N = H->N, M = H->M, dvm = H->dv_max, dcm = H->dc_max;
    const double MAXLLR = cfg->MAXLLR;
    res->uncoded = 0;
    for (int i = 0; i < N; i++) {                                         /* :174-199 */
        double yq = 4.0 * y[i] / N0;
        if (fabs(yq) > MAXLLR) yq = sgn_ge(yq) * MAXLLR;
        w->yq[i] = yq;
        w->r[i] = (int)sgn_ge(yq);
        w->d[i] = w->r[i];
        if (w->r[i] * w->c[i] < 0) res->uncoded++;
        w->sum[i] = yq;
        for (int s = 0; s < H->col_deg[i]; s++) w->v2c[(size_t)i * dvm + s] = yq;
    }
    int it;
    for (it = 0; it < cfg->num_iterations; it++) {                        /* :206-213 */
        for (int j = 0; j < M; j++) {                                     /* checkNodeUpdates :353-377, O(dc^2) as written */
            const int deg = H->row_deg[j];
            for (int k = 0; k < deg; k++) {
                double prod = 1.0;
                for (int k2 = 0; k2 < deg; k2++) if (k2 != k) {
                    int i = H->mlist[(size_t)j * dcm + k2];
                    double msg = w->v2c[(size_t)i * dvm + H->vn_slot[(size_t)j * dcm + k2]];
                    prod *= tanh(msg / 2.0);
                }
                w->c2v[(size_t)j * dcm + k] = log((1.0 + prod) / (1.0 - prod));
            }
        }
        for (int i = 0; i < N; i++) {                                     /* symNodeUpdates :379-409 */
            double sum = w->yq[i];
            for (int s = 0; s < H->col_deg[i]; s++) {
                int j = H->nlist[(size_t)i * dvm + s];
                sum += w->c2v[(size_t)j * dcm + H->cn_slot[(size_t)i * dvm + s]];
            }
            for (int s = 0; s < H->col_deg[i]; s++) {
                int j = H->nlist[(size_t)i * dvm + s];
                double out = sum - w->c2v[(size_t)j * dcm + H->cn_slot[(size_t)i * dvm + s]];
                if (fabs(out) > MAXLLR) out = MAXLLR * sgn_ge(out);
                w->v2c[(size_t)i * dvm + s] = out;
            }
            w->sum[i] = sum;
            w->d[i] = (sum > 0) ? 1 : -1;
        }
    }
    res->it = it;
    res->satisfied = all_checks_ok_bipolar(H, w->d);
    res->smoothed = 0; res->smoothing_used = 0; res->phases = 1;
}

/* ------------------------------------------------------------------------------------ */
/* A.6 DD-BMP: src/decodeDDBMP.cpp                                                       */
/* ------------------------------------------------------------------------------------ */
static void ddbmp_frame(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, const double *y, workspace *w, frame_result *res)
{
    const int N = H->N, M = H->M, dvm = H->dv_max, dcm = H->dc_max;
    const double Nq = pow(2.0, cfg->Q);                                   /* :106 */
    res->uncoded = 0;
    for (int i = 0; i < N; i++) {                                         /* :173-185 */
        double yq = oracle_quantize_ms(y[i], cfg->Ymax, Nq);
        w->yq[i] = yq;
        w->r[i] = (yq > 0) ? 1 : -1;
        w->d[i] = w->r[i];
        if (w->r[i] * w->c[i] < 0) res->uncoded++;
        w->sum[i] = yq;
        for (int s = 0; s < H->col_deg[i]; s++) {                         /* initializeSymMessages :301-310 */
            w->v2c[(size_t)i * dvm + s] = sgn_ge(yq);
            w->mem[(size_t)i * dvm + s] = yq;
        }
    }
    int it, satisfied = all_checks_ok_bipolar(H, w->d);
    for (it = 0; it < cfg->num_iterations; it++) {                        /* :191-205 */
        for (int j = 0; j < M; j++) {                                     /* checkNodeUpdates :350-372 */
            double prod = 1.0; const int deg = H->row_deg[j];
            for (int k = 0; k < deg; k++) {
                int i = H->mlist[(size_t)j * dcm + k];
                prod *= sgn_ge(w->v2c[(size_t)i * dvm + H->vn_slot[(size_t)j * dcm + k]]);
            }
            for (int k = 0; k < deg; k++) {
                int i = H->mlist[(size_t)j * dcm + k];
                w->c2v[(size_t)j * dcm + k] = prod * sgn_ge(w->v2c[(size_t)i * dvm + H->vn_slot[(size_t)j * dcm + k]]);
            }
        }
        for (int i = 0; i < N; i++) {                                     /* symNodeUpdates :396-423 */
            double sum = w->yq[i], dsum = sgn_ge(w->yq[i]);
            for (int s = 0; s < H->col_deg[i]; s++) {
                int j = H->nlist[(size_t)i * dvm + s];
                sum += w->c2v[(size_t)j * dcm + H->cn_slot[(size_t)i * dvm + s]];
            }
            for (int s = 0; s < H->col_deg[i]; s++) {
                int j = H->nlist[(size_t)i * dvm + s];
                w->mem[(size_t)i * dvm + s] += sum - w->c2v[(size_t)j * dcm + H->cn_slot[(size_t)i * dvm + s]];
                w->v2c[(size_t)i * dvm + s] = sgn_ge(w->mem[(size_t)i * dvm + s]);
                dsum += w->v2c[(size_t)i * dvm + s];
            }
            w->sum[i] = sum;
            w->d[i] = (dsum > 0) ? 1 : -1;
        }
        satisfied = all_checks_ok_bipolar(H, w->d);                       /* checkStoppingCondition :375-393; d is +-1 so sgn(d)=d */
        if (satisfied) break;                                             /* `it` is NOT incremented on this exit */
    }
    res->it = it;
    res->satisfied = satisfied;
    res->smoothed = 0; res->smoothing_used = 0; res->phases = 1;
}

/* ------------------------------------------------------------------------------------ */
/* A.4 GDBF family: src/decodeGDBF.cpp, and src/RNGDBF.cpp when LDPC_GPU_F_REDECODE       */
/* ------------------------------------------------------------------------------------ */
static double normalCDF(double value) { return 0.5 * erfc(-value * M_SQRT1_2); }   /* decodeGDBF.cpp:66-69 */

static double gdbf_objective(const oracle_code *H, const int *d, const double *y, const int *c2s)
{   /* evaluateObjectiveFunction :624-633 */
    double f = 0;
    for (int i = 0; i < H->N; i++) f += d[i] * y[i];
    for (int j = 0; j < H->M; j++) f += c2s[j];
    return f;
}

/* rows of `noise` one flip step consumes */
static int gdbf_rows_per_step(uint32_t flags)
{
    return ((flags & LDPC_GPU_F_ADDNOISE) ? 1 : 0) + ((flags & LDPC_GPU_F_QUANTIZEPROBABILITIES) ? 1 : 0);
}

static int gdbf_frame(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, double sigma, const double *y,
                      const double *noise, int64_t noise_rows, workspace *w, frame_result *res)
{
    const int N = H->N, M = H->M, dvm = H->dv_max, dcm = H->dc_max;
    const uint32_t fl = cfg->flags;
    const int T = cfg->num_iterations, W = cfg->windowsize;
    const int redecode = (fl & LDPC_GPU_F_REDECODE) != 0;
    const int maxphase = redecode ? cfg->maxphase : 1;
    int64_t row = 0;                                                      /* next unread row of `noise` */
    res->uncoded = 0;
    for (int i = 0; i < N; i++) {                                         /* decodeGDBF.cpp:251-274 / RNGDBF.cpp:251-275 */
        double yq = y[i];
        if (fl & LDPC_GPU_F_SATURATE_SAMPLES) if (fabs(yq) > cfg->Ymax) yq *= cfg->Ymax / fabs(yq);
        w->r[i] = (yq > 0) ? 1 : -1;
        if (fl & LDPC_GPU_F_QUANTIZE_SAMPLES) yq = oracle_quantize_gdbf(yq, cfg->Ymax, cfg->NQ);
        w->yq[i] = yq;
        if (w->r[i] * w->c[i] < 0) res->uncoded++;
        w->d[i] = w->r[i];
        w->dsum[i] = 0;
        w->shape[i] = 0.0;    /* noiseSamples: the reference never resets it between frames (decodeGDBF.cpp:201);
                                 this framework defines it per frame so frames stay independent */
        w->theta[i] = cfg->theta;
    }
    const double noiseSigma = sigma * cfg->noiseScale;                    /* :296 */
    int it = 0, satisfied = 1, phase = 0, total_it = 0;
    res->smoothed = 0; res->smoothing_used = 0;
    while (phase < maxphase) {                                            /* RNGDBF.cpp:280-400; runs once otherwise */
        if (redecode) for (int i = 0; i < N; i++) { w->d[i] = w->r[i]; w->dsum[i] = 0; }
        int mu = (fl & LDPC_GPU_F_SEQUENTIALMODE) ? 0 : 1;                /* :284-289 */
        if (fl & LDPC_GPU_F_THRESHOLDADAPTATION) for (int i = 0; i < N; i++) w->theta[i] = cfg->theta;   /* :291-294 */
        double f1 = 0, f2 = 0;
        for (it = 0; it < T; it++) {                                      /* :298 */
            satisfied = 1;
            for (int j = 0; j < M; j++) {                                 /* checkNodeUpdates :517-534 */
                int prod = 1;
                for (int k = 0; k < H->row_deg[j]; k++) prod *= w->d[H->mlist[(size_t)j * dcm + k]];
                if (prod < 0) satisfied = 0;
                w->c2s[j] = prod;
            }
            if (satisfied) break;                                         /* :305-306 */
            if ((fl & LDPC_GPU_F_MODESWITCHING) && it > cfg->Tswitch) f1 = gdbf_objective(H, w->d, w->yq, w->c2s);
            if (fl & LDPC_GPU_F_ADDNOISE) {                               /* :318-333 */
                if (!noise || row >= noise_rows) return fail("gdbf: noise array exhausted");
                const double *nr = noise + (size_t)row * N; row++;
                for (int i = 0; i < N; i++) {
                    double newSample = (fl & LDPC_GPU_F_UNIFORMNOISE) ? sqrt(3) * noiseSigma * 2.0 * (nr[i] - 0.5)
                                                                      : noiseSigma * nr[i];
                    if (fl & LDPC_GPU_F_NOISESHAPING) { w->pert[i] = newSample - w->shape[i]; w->shape[i] = newSample; }
                    else w->pert[i] = newSample;
                }
            }
            const double *ur = NULL;
            if (fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) {
                if (!noise || row >= noise_rows) return fail("gdbf: noise array exhausted");
                ur = noise + (size_t)row * N; row++;
            }
            /* symNodeUpdates :536-621 */
            double Emin = INFINITY; int mindx = -1; double wgt = 1;
            for (int i = 0; i < N; i++) {
                int flip = 0;
                double E = w->d[i] * w->yq[i];
                if (fl & LDPC_GPU_F_WEIGHTSYNDROMES)
                    wgt = redecode ? cfg->alpha * cfg->Ymax / H->col_deg[i]   /* RNGDBF.cpp:566 */
                                   : cfg->alpha;                              /* decodeGDBF.cpp:550 */
                for (int s = 0; s < H->col_deg[i]; s++) E += wgt * w->c2s[H->nlist[(size_t)i * dvm + s]];
                if (fl & LDPC_GPU_F_ADDNOISE) E += w->pert[i];
                if (fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) {              /* :561-597 */
                    double pcdf = normalCDF((-E + w->theta[i]) / noiseSigma);
                    static const double pr_levels[8] = { 0, 0.0625, 0.125, 0.25, 0.34375, 0.4106, 0.68359, 1 };
                    double min_dist = 1; int min_idx = 0;
                    for (int q = 0; q < 8; q++) { double t = pr_levels[q] - pcdf; t = t * t; if (t < min_dist) { min_dist = t; min_idx = q; } }
                    if (ur[i] < pr_levels[min_idx]) { flip = 1; w->d[i] = -w->d[i]; }
                } else {
                    if (mu == 1 && E < w->theta[i]) { flip = 1; w->d[i] = -w->d[i]; }
                    if (mu == 0) if (E < Emin) { flip = 1; Emin = E; mindx = i; }
                }
                if (fl & LDPC_GPU_F_THRESHOLDADAPTATION) if (!flip) w->theta[i] *= cfg->lambda;   /* :612-617 */
            }
            if (mu == 0 && mindx >= 0) w->d[mindx] = -w->d[mindx];        /* :619-620 */
            if ((fl & LDPC_GPU_F_MODESWITCHING) && it > cfg->Tswitch) {   /* :338-346 */
                f2 = gdbf_objective(H, w->d, w->yq, w->c2s);
                if (f1 >= f2) mu = 0;
            }
            if (fl & LDPC_GPU_F_OUTPUTSMOOTHING) if (it > T - W) for (int i = 0; i < N; i++) w->dsum[i] += w->d[i];   /* :348-354 */
        }
        if (fl & LDPC_GPU_F_OUTPUTSMOOTHING) {
            if (!satisfied) { for (int i = 0; i < N; i++) w->d[i] = (w->dsum[i] > 0) ? 1 : -1; res->smoothed = 1; }   /* :358-367 */
            else res->smoothed = 0;
            if (it > T - W) res->smoothing_used++;                        /* :371-374 */
        }
        total_it += it;                                                   /* RNGDBF.cpp:394 / decodeGDBF.cpp:399 */
        phase++;
        if (!redecode || satisfied) break;                                /* RNGDBF.cpp:398-399 */
    }
    res->it = total_it; res->satisfied = satisfied; res->phases = phase;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* A.5 NGDBFhw: src/NGDBFhw.cpp (d in {0,1}; integer flip metric)                        */
/* ------------------------------------------------------------------------------------ */
static int hw_frame(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, double sigma, const double *y,
                    const double *noise, int qpointer, workspace *w, frame_result *res, int *d01, const int *c01)
{
    const int N = H->N, M = H->M, dvm = H->dv_max, dcm = H->dc_max;
    const int QB = LDPC_GPU_HW_QBUF;
    if (N >= QB) return fail("NGDBFhw needs N < 2648");
    if (!noise) return fail("NGDBFhw needs the per-frame noise buffer");
    const double wgt = cfg->w, Ymax = cfg->Ymax;
    const double noiseSigma = sigma * cfg->noiseScale;                    /* :129 */
    const double qmax = pow(2, 5), lmax = Ymax / (2.0 * wgt), NL = qmax - 1;   /* :171-173 */
    const int theta = oracle_hw_unpack(oracle_hw_pack(2.0, Ymax, wgt) & 0xF); /* :175: unpack(pack(quantize(2),1)) */
    const int Smult = (int)round(NL / lmax);                              /* :176 */
    int *yprime = w->dsum;            /* reuse int scratch */
    int *qprime = (int *)malloc(sizeof(int) * QB);
    res->uncoded = 0;
    for (int i = 0; i < N; i++) {                                         /* :218-235 */
        double yi = y[i];
        if (fabs(yi) > Ymax) yi *= Ymax / fabs(yi);
        w->r[i] = (yi > 0) ? 1 : -1;
        if (w->r[i] * c01[i] < 0) res->uncoded++;                         /* c is 0/1 here, :141,230 */
        yprime[i] = oracle_hw_pack(yi / (2.0 * wgt), Ymax, wgt);         /* :234,237 */
    }
    for (int i = 0; i < QB; i++) {                                        /* :239-252 */
        double q = noiseSigma * noise[i];
        double qm = ((q - cfg->theta0) / (2.0 * wgt) - 1.0);
        if (qm > lmax) qm = lmax; else if (qm < -lmax) qm = -lmax;
        qprime[i] = oracle_hw_pack(qm, Ymax, wgt);
    }
    const int T = cfg->num_iterations, maxPhases = cfg->maxphase > 0 ? cfg->maxphase : 1;
    int leastIterations = T, leastErrors = N, satisfied = 1, it = 0;
    for (int phase = 0; phase < maxPhases; phase++) {                     /* :280-373 */
        for (int i = 0; i < N; i++) d01[i] = (1 - w->r[i]) / 2;
        for (it = 0; it < T; it++) {
            satisfied = 1;
            for (int j = 0; j < M; j++) {                                 /* checkNodeUpdates :546-563 */
                int prod = 1;
                for (int k = 0; k < H->row_deg[j]; k++) prod *= (1 - 2 * d01[H->mlist[(size_t)j * dcm + k]]);
                if (prod < 0) satisfied = 0;
                w->c2s[j] = (1 - prod) / 2;
            }
            if (satisfied) break;
            for (int i = 0; i < N; i++) {                                 /* symNodeUpdates :565-593 */
                int E = (1 - 2 * d01[i]) * oracle_hw_unpack(yprime[i]);
                int SSum = 0;
                for (int s = 0; s < H->col_deg[i]; s++) SSum += 1 - w->c2s[H->nlist[(size_t)i * dvm + s]];
                E += SSum * Smult + oracle_hw_unpack(qprime[i + qpointer]);
                if (E <= theta) d01[i] = 1 - d01[i];
            }
            qpointer++;                                                   /* :356-358 */
            if (qpointer >= QB - N) qpointer = 0;
        }
        int newErrors = count_errors(d01, c01, N);                        /* :362-372 */
        if (newErrors < leastErrors) leastErrors = newErrors;
        if (it < leastIterations) leastIterations = it;
    }
    free(qprime);
    res->it = leastIterations; res->errors = leastErrors; res->satisfied = satisfied;
    res->smoothed = 0; res->smoothing_used = 0; res->phases = maxPhases;
    return qpointer;   /* >= 0: the window position the next frame would inherit */
}

/* ------------------------------------------------------------------------------------ */
/* LDPC_GPU_KIND_NGDBF_SC: restatement of the reference's SystemC NGDBF model              */
/* (/root/reference/SystemC/NGDBF/inc/nodes.h:76-138,167-201; inc/decoder.h:183-254;       */
/* inc/ldpcsim.h:85-119).  PARITY UNPINNED: SystemC is absent, the model cannot be run    */
/* here; this restatement is the only checker of the GPU kernel.         */
/* ------------------------------------------------------------------------------------ */
static double sc_quantize(const double *thr, const double *val, int Nq, double Y)
{   /* inc/ldpcsim.h:98-119 */
    int k = 0;
    for (int i = 0; i < Nq - 1; i++) if (Y > thr[i]) k = i + 1;
    return val[k];
}

static int sc_frame(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, double sigma, const double *y,
                    const double *noise, int64_t nlen, workspace *w, frame_result *res)
{
    const int N = H->N, M = H->M, dvm = H->dv_max, dcm = H->dc_max, T = cfg->num_iterations, W = cfg->windowsize;
    if (cfg->Q < 1 || cfg->Q > 8) return fail("NGDBF_SC: 1 <= Q <= 8");
    if (!noise || nlen < (int64_t)N + T + 1) return fail("NGDBF_SC needs N + T + 1 noise values per frame");
    const int Nq = 1 << cfg->Q;                                           /* src/ldpcsim.cpp:114 */
    double thr[256], val[256];
    for (int i = 0; i < Nq - 1; i++) {                                    /* inc/ldpcsim.h:85-96 */
        thr[i] = -cfg->Ymax * (Nq - 2.0) / (Nq - 1.0) + i * (2.0 * cfg->Ymax / (Nq - 1.0));
        val[i] = -cfg->Ymax + i * (2.0 * cfg->Ymax / (Nq - 1.0));
    }
    val[Nq - 1] = cfg->Ymax;
    double *r = w->yq, *thl = w->theta;
    double *q = (double *)malloc(sizeof(double) * (size_t)(N + T + 1));
    int *updown = w->dsum, *x = w->d;
    const double noiseSigma = sigma * cfg->noiseScale;
    res->uncoded = 0;
    for (int i = 0; i < N; i++) {                                         /* decoder.h:212, nodes.h:80-97 */
        r[i] = sc_quantize(thr, val, Nq, y[i]);
        x[i] = (r[i] > 0) ? 1 : -1;
        thl[i] = cfg->theta; updown[i] = 0;
        if (x[i] * w->c[i] < 0) res->uncoded++;
    }
    for (int k = 0; k < N + T + 1; k++) q[k] = sc_quantize(thr, val, Nq, noiseSigma * noise[k]);   /* decoder.h:187 */
    int t, all_stop = 0;
    for (t = 1; ; t++) {
        all_stop = 1;                                                     /* nodes.h:167-201: products of the previous clock's messages */
        for (int j = 0; j < M; j++) {
            int prod = 1;
            for (int k = 0; k < H->row_deg[j]; k++) prod *= x[H->mlist[(size_t)j * dcm + k]];
            w->c2s[j] = prod;
            if (prod != 1) all_stop = 0;
        }
        if (t > T - W) for (int i = 0; i < N; i++) updown[i] += x[i];    /* decoder.h:221-227 (1 - 2 Sd = x) */
        if (all_stop || t > T) break;                                     /* :248-249 */
        for (int i = 0; i < N; i++) {                                     /* nodes.h:104-137 */
            const double wgt = cfg->alpha * cfg->Ymax / H->col_deg[i];    /* :59 */
            double E = x[i] * r[i] + q[(N - 1 - i) + (t - 1)];            /* :107; the chain: node i reads what node i-1 read a clock earlier */
            for (int s2 = 0; s2 < H->col_deg[i]; s2++) E += wgt * w->c2s[H->nlist[(size_t)i * dvm + s2]];   /* :113-114 */
            if (E < sc_quantize(thr, val, Nq, thl[i])) { thl[i] = thl[i] / cfg->lambda; x[i] = -x[i]; }      /* :117-121 */
            else thl[i] = thl[i] * cfg->lambda;                           /* :123-126 */
        }
    }
    res->smoothed = (t >= T);                                             /* decoder.h:236-246 */
    if (res->smoothed) for (int i = 0; i < N; i++) x[i] = (updown[i] > 0) ? 1 : -1;
    res->it = t - 1; res->satisfied = all_stop; res->smoothing_used = res->smoothed; res->phases = 1;
    free(q);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Batch driver + a19 accounting (src/decodeMinSum.cpp:270-288 and siblings)             */
/* ------------------------------------------------------------------------------------ */
static int iter_hist_len(const ldpc_gpu_decoder_cfg *cfg)
{
    int ph = ((cfg->flags & LDPC_GPU_F_REDECODE) && cfg->kind == LDPC_GPU_KIND_GDBF && cfg->maxphase > 1) ? cfg->maxphase : 1;
    return cfg->num_iterations * ph + 1;
}

int oracle_decode_batch(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, const ldpc_gpu_channel *ch,
                        const ldpc_gpu_batch *b, ldpc_gpu_counters *cnt)
{
    if (!H || !cfg || !ch || !b) return fail("null argument");
    if (b->y_dtype != LDPC_GPU_DT_F64 || b->mem != LDPC_GPU_MEM_HOST) return fail("oracle takes host f64 samples");
    if (cfg->num_iterations < 0) return fail("negative T");
    const int N = H->N;
    const double N0 = pow(10.0, -ch->snr_db / 10.0) / ch->R;              /* decodeMinSum.cpp:146 */
    const double sigma = sqrt(N0 / 2.0);                                  /* :147 */
    const int hw = cfg->kind == LDPC_GPU_KIND_NGDBF_HW;
    workspace w; ws_alloc(&w, H);
    int *d01 = (int *)calloc((size_t)N, sizeof(int)), *c01 = (int *)calloc((size_t)N, sizeof(int));
    const size_t bpf = (size_t)(N + 7) / 8;
    int rc = 0;
    for (int64_t f = 0; f < b->n_frames && rc >= 0; f++) {
        const double *y = (const double *)b->y + (size_t)f * N;
        for (int i = 0; i < N; i++) {                                     /* codeword: bit 0 -> +1 (:204-210) */
            int bit = b->codeword ? b->codeword[(size_t)f * N + i] : 0;
            w.c[i] = bit ? -1 : 1; c01[i] = bit;
        }
        frame_result res; memset(&res, 0, sizeof res);
        switch (cfg->kind) {
        case LDPC_GPU_KIND_MINSUM: ms_frame(H, cfg, y, &w, &res); break;
        case LDPC_GPU_KIND_BP:     bp_frame(H, cfg, N0, y, &w, &res); break;
        case LDPC_GPU_KIND_DDBMP:  ddbmp_frame(H, cfg, y, &w, &res); break;
        case LDPC_GPU_KIND_GDBF:
            rc = gdbf_frame(H, cfg, sigma, y, b->noise ? b->noise + (size_t)f * b->noise_rows * N : NULL, b->noise_rows, &w, &res);
            break;
        case LDPC_GPU_KIND_NGDBF_HW:
            rc = hw_frame(H, cfg, sigma, y, b->noise ? b->noise + (size_t)f * LDPC_GPU_HW_QBUF : NULL,
                          b->qpointer0 ? b->qpointer0[f] : 0, &w, &res, d01, c01);
            break;
        case LDPC_GPU_KIND_NGDBF_SC:
            rc = sc_frame(H, cfg, sigma, y, b->noise ? b->noise + (size_t)f * b->noise_rows : NULL, b->noise_rows, &w, &res);
            break;
        default: rc = fail("unknown decoder kind");
        }
        if (rc < 0) break;
        if (!hw) res.errors = count_errors(w.d, w.c, N);
        if (b->out_bits) {
            uint8_t *ob = b->out_bits + f * bpf; memset(ob, 0, bpf);
            for (int i = 0; i < N; i++) { int one = hw ? d01[i] : (w.d[i] < 0); if (one) ob[i >> 3] |= (uint8_t)(1u << (i & 7)); }
        }
        if (b->out_iters) b->out_iters[f] = res.it;
        if (b->out_errors) b->out_errors[f] = res.errors;
        if (b->out_flags) b->out_flags[f] = (uint8_t)((res.satisfied ? 1 : 0) | (res.smoothed ? 2 : 0) | ((res.phases & 15) << 4));
        if (b->out_soft && !hw && cfg->kind != LDPC_GPU_KIND_GDBF)
            memcpy((double *)b->out_soft + (size_t)f * N, w.sum, sizeof(double) * N);
        if (cnt) {
            if (res.errors > 0) {
                cnt->errors += res.errors; cnt->wordErrors++;
                if (cnt->error_weight_hist) cnt->error_weight_hist[res.errors - 1]++;
                if (res.satisfied) cnt->undetectedWords++;
            }
            cnt->uncodedErrors += res.uncoded;
            cnt->totalWords++; cnt->totalBits += N; cnt->totalIterations += res.it;
            cnt->smoothingUsed += res.smoothing_used;
            if (cnt->iter_hist && res.it < iter_hist_len(cfg)) cnt->iter_hist[res.it]++;
            if (cnt->phase_hist && (cfg->flags & LDPC_GPU_F_REDECODE) && cfg->kind == LDPC_GPU_KIND_GDBF) cnt->phase_hist[res.phases - 1]++;
        }
        rc = 0;
    }
    ws_free(&w); free(d01); free(c01);
    return rc < 0 ? rc : 0;
}

/* ------------------------------------------------------------------------------------ */
/* The framework's counter-based channel (no reference counterpart; replaces inc/rand.h). */
/* Independent restatement of ldpcsimulation_b200/csrc/ldpc_rng.cuh -- see DESIGN.md.     */
/* ------------------------------------------------------------------------------------ */
void oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{   /* Philox4x32-10 (Salmon et al., SC'11): 10 rounds, multipliers D2511F53 / CD9E8D57, Weyl 9E3779B9 / BB67AE85 */
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* One Box-Muller pair from two 32-bit words, in fp32 with every rounding spelled out
 * (fmaf / single operations only) so that the CUDA kernel reproduces it bit for bit.
 *   angle  = 2*pi * (a>>8)/2^24, quadrant-reduced, cephes sinf/cosf minimax polynomials
 *   radius = sqrt(-2 ln((b+1)/2^32)), cephes logf polynomial on the top 24 bits of b+1 */
static void box_muller_pair(uint32_t a, uint32_t b, float *n0, float *n1)
{
    int32_t a24 = (int32_t)(a >> 8);
    int32_t k = (a24 + (1 << 21)) >> 22;                   /* nearest quarter turn, 0..4 */
    int32_t rem = a24 - (k << 22);                         /* [-2^21, 2^21] */
    float phi = (float)rem * 5.9604644775390625e-08f;      /* rem * 2^-24, exact */
    phi = phi * 6.2831855f;
    float z = phi * phi;
    float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    float s = fmaf(ps * z, phi, phi);
    float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    float cs, sn;
    switch (k & 3) { case 0: cs = c; sn = s; break; case 1: cs = -s; sn = c; break;
                     case 2: cs = -c; sn = -s; break; default: cs = s; sn = -c; }
    uint64_t v = (uint64_t)b + 1u;                         /* [1, 2^32] */
    int e = 63 - __builtin_clzll(v);
    uint32_t top = (uint32_t)((v << (63 - e)) >> 40);      /* 24 bits, leading one set */
    float m = (float)top * 1.1920928955078125e-07f;        /* top * 2^-23 in [1,2), exact */
    e -= 32;
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float x = m - 1.0f;
    float zz = x * x;
    float p = 7.0376836292e-2f;
    p = fmaf(p, x, -1.1514610310e-1f); p = fmaf(p, x, 1.1676998740e-1f); p = fmaf(p, x, -1.2420140846e-1f);
    p = fmaf(p, x, 1.4249322787e-1f);  p = fmaf(p, x, -1.6668057665e-1f); p = fmaf(p, x, 2.0000714765e-1f);
    p = fmaf(p, x, -2.4999993993e-1f); p = fmaf(p, x, 3.3333331174e-1f);
    float fe = (float)e;
    float yv = (x * zz) * p;
    yv = fmaf(-2.12194440e-4f, fe, yv);
    yv = fmaf(-0.5f, zz, yv);
    float ln = x + yv;
    ln = fmaf(0.693359375f, fe, ln);
    float t = -2.0f * ln;
    if (t < 0.0f) t = 0.0f;
    float rad = sqrtf(t);
    *n0 = rad * cs; *n1 = rad * sn;
}

/* Four standard normals for (seed, frame, block-of-4 index, row, stream).
 * stream 0 = channel noise, 1 = decoder perturbation noise. */
void oracle_normal4(uint64_t seed, uint64_t frame, uint32_t block, uint32_t row, uint32_t stream, float out[4])
{
    uint32_t ctr[4] = { block, (row << 2) | (stream & 3u), (uint32_t)frame, (uint32_t)(frame >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, r[4];
    oracle_philox4x32(ctr, key, r);
    box_muller_pair(r[0], r[1], &out[0], &out[1]);
    box_muller_pair(r[2], r[3], &out[2], &out[3]);
}

static void uniform4(uint64_t seed, uint64_t frame, uint32_t block, uint32_t row, uint32_t stream, double out[4])
{   /* ranu() shape, inc/rand.h:12-13, on 31 bits of each Philox word */
    uint32_t ctr[4] = { block, (row << 2) | (stream & 3u), (uint32_t)frame, (uint32_t)(frame >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, r[4];
    oracle_philox4x32(ctr, key, r);
    for (int q = 0; q < 4; q++) out[q] = (1.0 + (double)(r[q] >> 1)) / (2.0 + (double)0x7fffffff);
}

static int64_t noise_rows_needed(const ldpc_gpu_decoder_cfg *cfg)
{
    if (cfg->kind != LDPC_GPU_KIND_GDBF) return 0;
    int ph = (cfg->flags & LDPC_GPU_F_REDECODE) ? (cfg->maxphase > 0 ? cfg->maxphase : 1) : 1;
    return (int64_t)cfg->num_iterations * ph * gdbf_rows_per_step(cfg->flags);
}

int oracle_channel_dump(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, const ldpc_gpu_channel *ch,
                        uint64_t seed, int64_t frame_begin, int64_t n_frames,
                        const uint8_t *codewords, int64_t n_codewords,
                        double *y, double *noise, int64_t noise_rows)
{
    const int N = H->N;
    const double N0 = pow(10.0, -ch->snr_db / 10.0) / ch->R;
    const double sigma = sqrt(N0 / 2.0);
    for (int64_t f = 0; f < n_frames; f++) {
        uint64_t fid = (uint64_t)(frame_begin + f);
        const uint8_t *cw = (codewords && n_codewords > 0) ? codewords + (size_t)(fid % (uint64_t)n_codewords) * N : NULL;
        for (int blk = 0; blk * 4 < N; blk++) {
            float n4[4]; oracle_normal4(seed, fid, (uint32_t)blk, 0, 0, n4);
            for (int q = 0; q < 4 && blk * 4 + q < N; q++) {
                int i = blk * 4 + q;
                double x = (cw && cw[i]) ? -1.0 : 1.0;
                y[(size_t)f * N + i] = x * (1.0 + sigma * (double)n4[q]);     /* decodeMinSum.cpp:216 */
            }
        }
        if (!noise) continue;
        if (cfg->kind == LDPC_GPU_KIND_NGDBF_HW) {
            double *nf = noise + (size_t)f * LDPC_GPU_HW_QBUF;
            for (int blk = 0; blk * 4 < LDPC_GPU_HW_QBUF; blk++) {
                float n4[4]; oracle_normal4(seed, fid, (uint32_t)blk, 0, 1, n4);
                for (int q = 0; q < 4 && blk * 4 + q < LDPC_GPU_HW_QBUF; q++) nf[blk * 4 + q] = (double)n4[q];
            }
        } else if (cfg->kind == LDPC_GPU_KIND_NGDBF_SC) {
            double *nf = noise + (size_t)f * noise_rows;
            for (int blk = 0; blk * 4 < noise_rows; blk++) {
                float n4[4]; oracle_normal4(seed, fid, (uint32_t)blk, 0, 1, n4);
                for (int q = 0; q < 4 && blk * 4 + q < noise_rows; q++) nf[blk * 4 + q] = (double)n4[q];
            }
        } else if (cfg->kind == LDPC_GPU_KIND_GDBF) {
            const int rps = gdbf_rows_per_step(cfg->flags);
            for (int64_t row = 0; row < noise_rows; row++) {
                /* row kinds inside one step: [perturbation][flip uniforms] */
                int which = rps ? (int)(row % rps) : 0;
                int is_uniform = (which == 0 && (cfg->flags & LDPC_GPU_F_ADDNOISE)) ? ((cfg->flags & LDPC_GPU_F_UNIFORMNOISE) != 0) : 1;
                double *nr = noise + ((size_t)f * noise_rows + row) * N;
                for (int blk = 0; blk * 4 < N; blk++) {
                    if (is_uniform) { double u4[4]; uniform4(seed, fid, (uint32_t)blk, (uint32_t)row, 1, u4);
                                      for (int q = 0; q < 4 && blk * 4 + q < N; q++) nr[blk * 4 + q] = u4[q]; }
                    else { float n4[4]; oracle_normal4(seed, fid, (uint32_t)blk, (uint32_t)row, 1, n4);
                           for (int q = 0; q < 4 && blk * 4 + q < N; q++) nr[blk * 4 + q] = (double)n4[q]; }
                }
            }
        }
    }
    return 0;
}

int oracle_simulate(const oracle_code *H, const ldpc_gpu_decoder_cfg *cfg, const ldpc_gpu_channel *ch,
                    const ldpc_gpu_sim_args *a, const uint8_t *codewords, int64_t n_codewords,
                    ldpc_gpu_counters *cnt)
{
    const int N = H->N;
    const int64_t rows = cfg->kind == LDPC_GPU_KIND_NGDBF_SC ? (int64_t)N + cfg->num_iterations + 1 : noise_rows_needed(cfg);
    double *y = (double *)malloc(sizeof(double) * N);
    size_t nsz = cfg->kind == LDPC_GPU_KIND_NGDBF_HW ? LDPC_GPU_HW_QBUF : cfg->kind == LDPC_GPU_KIND_NGDBF_SC ? (size_t)rows : (size_t)rows * N;
    double *noise = nsz ? (double *)malloc(sizeof(double) * nsz) : NULL;
    uint8_t *cw = (uint8_t *)calloc((size_t)N, 1);
    int rc = 0;
    for (int64_t f = 0; f < a->n_frames; f++) {
        if ((a->stop_errors > 0 || a->stop_word_errors > 0) &&
            !(cnt->errors < a->stop_errors || cnt->wordErrors < a->stop_word_errors)) break;   /* decodeMinSum.cpp:189 */
        int64_t fid = a->frame_begin + f;
        oracle_channel_dump(H, cfg, ch, a->seed, fid, 1, codewords, n_codewords, y, noise, rows);
        if (codewords && n_codewords > 0) memcpy(cw, codewords + (size_t)((uint64_t)fid % (uint64_t)n_codewords) * N, (size_t)N);
        ldpc_gpu_batch b; memset(&b, 0, sizeof b);
        b.n_frames = 1; b.mem = LDPC_GPU_MEM_HOST; b.y_dtype = LDPC_GPU_DT_F64; b.y = y; b.noise = noise; b.noise_rows = rows;
        b.codeword = (codewords && n_codewords > 0) ? cw : NULL;
        rc = oracle_decode_batch(H, cfg, ch, &b, cnt);
        if (rc) break;
    }
    free(y); free(noise); free(cw);
    return rc;
}
