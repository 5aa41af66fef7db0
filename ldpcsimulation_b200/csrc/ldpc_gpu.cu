// ldpc_gpu.cu -- the C ABI of include/ldpc_gpu.h over the sm_100a kernels.
//
// Host side of the drop-in boundary: alist parsing and graph compilation (a1), decoder handles
// (the reference's -D macros + parameter globals), the parity entry (caller samples), the
// throughput entry (Philox channel), counters (a19) and the one collective (e).
// There is no CPU implementation of any decoder in this library.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "ldpc_mp_kernels.cuh"
#include "ldpc_ms_fast.cuh"
#include "ldpc_ms_rc.cuh"
#include "ldpc_schedule.h"
#include "ldpc_ms_tile.cuh"
#include "ldpc_ms_quad.cuh"
#include "ldpc_ms_h2.cuh"
#include "ldpc_ms_h2rc.cuh"
#include "ldpc_ms_x2.cuh"
#include "ldpc_ms_tileh.cuh"
#include "ldpc_bf_kernels.cuh"
#include "ldpc_sc_kernel.cuh"
#include "ldpc_nb_kernel.cuh"

using namespace ldpc;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int set_err(int code, const std::string &msg) { g_err = msg; return code; }
#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return set_err(e_ == cudaErrorMemoryAllocation ? LDPC_GPU_ERR_NOMEM : LDPC_GPU_ERR_CUDA,   \
                           std::string(#expr) + ": " + cudaGetErrorString(e_));                        \
    } while (0)

extern "C" const char *ldpc_gpu_last_error(void) { return g_err.c_str(); }
extern "C" int ldpc_gpu_version(void) { return 100; }

static std::vector<int> g_devices;
extern "C" int ldpc_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int ldpc_gpu_init(const int *ords, int n)
{
    int cnt = ldpc_gpu_device_count();
    if (cnt <= 0) return set_err(LDPC_GPU_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    g_devices.clear();
    if (n <= 0 || !ords) g_devices.push_back(0);
    else for (int q = 0; q < n; q++) {
        if (ords[q] < 0 || ords[q] >= cnt) return set_err(LDPC_GPU_ERR_INVALID_ARG, "device ordinal out of range");
        g_devices.push_back(ords[q]);
    }
    for (int d : g_devices) { CU_TRY(cudaSetDevice(d)); CU_TRY(cudaFree(0)); }
    CU_TRY(cudaSetDevice(g_devices[0]));
    return LDPC_GPU_OK;
}
extern "C" int ldpc_gpu_shutdown(void) { g_devices.clear(); return LDPC_GPU_OK; }

// ------------------------------------------------------------------------------------------------
// a1: parity-check matrix
// ------------------------------------------------------------------------------------------------
struct ldpc_gpu_code {
    int N = 0, M = 0, E = 0, dv_max = 0, dc_max = 0;
    std::vector<int> col_deg, row_deg;
    std::vector<int> nlist;   // [N*dv_max] 0-based check of slot s of variable i, -1 padded
    std::vector<int> mlist;   // [M*dc_max] 0-based variable of slot k of check j, -1 padded
    std::vector<int> vn_slot; // [M*dc_max] slot of check j inside variable mlist[j][k]'s nlist row
    // encoder state (ldpc_gpu_code_random_codewords), built on first use
    std::mutex enc_mutex;
    bool enc_ready = false;
    int rank = 0;
    std::vector<int> pivot_col;                 // [rank] pivot column of RREF row r
    std::vector<int> info_col;                  // [N - rank] free columns
    std::vector<std::vector<int>> parity_of;    // [rank] free-column indices (into info_col) that RREF row r sums
};

extern "C" int ldpc_gpu_code_create(int N, int M, int dvm, int dcm, const int *num_nlist, const int *num_mlist,
                                    const int *nlist_flat, const int *mlist_flat, ldpc_gpu_code **out)
{
    if (!out) return set_err(LDPC_GPU_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (N <= 0 || M <= 0 || dvm <= 0 || dcm <= 0 || !num_nlist || !num_mlist || !nlist_flat || !mlist_flat)
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "bad dimensions or NULL array");
    if (dvm > 255 || dcm > 255) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "node degree above 255");
    std::unique_ptr<ldpc_gpu_code> c(new ldpc_gpu_code);
    c->N = N; c->M = M; c->dv_max = dvm; c->dc_max = dcm;
    c->col_deg.assign(num_nlist, num_nlist + N); c->row_deg.assign(num_mlist, num_mlist + M);
    c->nlist.assign((size_t)N * dvm, -1); c->mlist.assign((size_t)M * dcm, -1); c->vn_slot.assign((size_t)M * dcm, -1);
    long long En = 0, Em = 0;
    for (int i = 0; i < N; i++) {
        if (num_nlist[i] < 0 || num_nlist[i] > dvm) return set_err(LDPC_GPU_ERR_BAD_CODE, "column weight exceeds biggest_num_n");
        En += num_nlist[i];
        for (int s = 0; s < num_nlist[i]; s++) {
            int v = nlist_flat[(size_t)i * dvm + s] - 1;          // indices stay 1-based in the file (inc/alist.h)
            if (v < 0 || v >= M) return set_err(LDPC_GPU_ERR_BAD_CODE, "nlist entry out of range (zero inside a row's weight?)");
            c->nlist[(size_t)i * dvm + s] = v;
        }
    }
    for (int j = 0; j < M; j++) {
        if (num_mlist[j] < 0 || num_mlist[j] > dcm) return set_err(LDPC_GPU_ERR_BAD_CODE, "row weight exceeds biggest_num_m");
        Em += num_mlist[j];
        for (int k = 0; k < num_mlist[j]; k++) {
            int v = mlist_flat[(size_t)j * dcm + k] - 1;
            if (v < 0 || v >= N) return set_err(LDPC_GPU_ERR_BAD_CODE, "mlist entry out of range");
            c->mlist[(size_t)j * dcm + k] = v;
        }
    }
    if (En != Em) return set_err(LDPC_GPU_ERR_BAD_CODE, "nlist and mlist hold different numbers of edges");
    c->E = (int)Em;
    // the edge permutation the reference recomputes with find() (last match wins, decodeMinSum.cpp:527-536)
    for (int j = 0; j < M; j++)
        for (int k = 0; k < c->row_deg[j]; k++) {
            const int i = c->mlist[(size_t)j * dcm + k];
            int slot = -1;
            for (int s = 0; s < c->col_deg[i]; s++) if (c->nlist[(size_t)i * dvm + s] == j) slot = s;
            if (slot < 0) return set_err(LDPC_GPU_ERR_BAD_CODE, "nlist is not the transpose of mlist");
            c->vn_slot[(size_t)j * dcm + k] = slot;
        }
    // and the other direction must be consistent too (every nlist entry owned by exactly one mlist entry)
    std::vector<unsigned char> seen((size_t)N * dvm, 0);
    for (int j = 0; j < M; j++)
        for (int k = 0; k < c->row_deg[j]; k++) {
            const size_t pos = (size_t)c->mlist[(size_t)j * dcm + k] * dvm + c->vn_slot[(size_t)j * dcm + k];
            if (seen[pos]) return set_err(LDPC_GPU_ERR_BAD_CODE, "duplicate entry in a row of mlist");
            seen[pos] = 1;
        }
    *out = c.release();
    return LDPC_GPU_OK;
}

// alist text: header, weights, N column rows, M row rows (src/alist.cpp:71-91).  Rows are read line
// by line, so both the zero-padded layout the reference's default loader needs and the unpadded
// layout of its -DCPPSTYLE branch (src/alist.cpp:26-62) are accepted.
static int load_alist(const char *path, bool transposed, ldpc_gpu_code **out)
{
    if (!out) return set_err(LDPC_GPU_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!path) return set_err(LDPC_GPU_ERR_INVALID_ARG, "path is NULL");
    std::ifstream f(path);
    if (!f) return set_err(LDPC_GPU_ERR_IO, std::string("cannot open ") + path);
    std::vector<std::vector<long>> rows;
    std::string line;
    while (std::getline(f, line)) {
        std::vector<long> v; const char *p = line.c_str(); char *e;
        for (;;) { long x = strtol(p, &e, 10); if (e == p) break; v.push_back(x); p = e; }
        while (*p == ' ' || *p == '\t' || *p == '\r') p++;
        if (*p) return set_err(LDPC_GPU_ERR_BAD_CODE, "non-numeric token in alist");
        if (!v.empty()) rows.push_back(std::move(v));
    }
    if (rows.size() < 4 || rows[0].size() != 2 || rows[1].size() != 2) return set_err(LDPC_GPU_ERR_BAD_CODE, "bad alist header");
    const long N = rows[0][0], M = rows[0][1], dvm = rows[1][0], dcm = rows[1][1];
    if (N <= 0 || M <= 0 || dvm <= 0 || dcm <= 0 || N > (1 << 24) || M > (1 << 24)) return set_err(LDPC_GPU_ERR_BAD_CODE, "bad alist dimensions");
    if ((long)rows.size() != 4 + N + M) return set_err(LDPC_GPU_ERR_BAD_CODE, "alist line count does not match its header");
    if ((long)rows[2].size() != N || (long)rows[3].size() != M) {
        if ((long)rows[2].size() == M && (long)rows[3].size() == N)
            return set_err(LDPC_GPU_ERR_BAD_CODE, "alist header is transposed (M N): SystemC-tree convention, not C_implementations'");
        return set_err(LDPC_GPU_ERR_BAD_CODE, "weight vectors do not match the header");
    }
    std::vector<int> num_n(N), num_m(M), nl((size_t)N * dvm, 0), ml((size_t)M * dcm, 0);
    for (long i = 0; i < N; i++) num_n[i] = (int)rows[2][i];
    for (long j = 0; j < M; j++) num_m[j] = (int)rows[3][j];
    for (long i = 0; i < N; i++) {
        int cnt = 0;
        for (long v : rows[4 + i]) { if (v == 0) continue; if (cnt >= dvm) return set_err(LDPC_GPU_ERR_BAD_CODE, "column longer than biggest_num_n"); nl[(size_t)i * dvm + cnt++] = (int)v; }
        if (cnt != num_n[i]) return set_err(LDPC_GPU_ERR_BAD_CODE, "column weight does not match its entries");
    }
    for (long j = 0; j < M; j++) {
        int cnt = 0;
        for (long v : rows[4 + N + j]) { if (v == 0) continue; if (cnt >= dcm) return set_err(LDPC_GPU_ERR_BAD_CODE, "row longer than biggest_num_m"); ml[(size_t)j * dcm + cnt++] = (int)v; }
        if (cnt != num_m[j]) return set_err(LDPC_GPU_ERR_BAD_CODE, "row weight does not match its entries");
    }
    if (transposed)      // the file describes H^T: its "columns" are the checks (SystemC/NGDBF/src/ldpcsim.cpp:107-110)
        return ldpc_gpu_code_create((int)M, (int)N, (int)dcm, (int)dvm, num_m.data(), num_n.data(), ml.data(), nl.data(), out);
    return ldpc_gpu_code_create((int)N, (int)M, (int)dvm, (int)dcm, num_n.data(), num_m.data(), nl.data(), ml.data(), out);
}
extern "C" int ldpc_gpu_code_load_alist(const char *path, ldpc_gpu_code **out) { return load_alist(path, false, out); }
extern "C" int ldpc_gpu_code_load_alist_transposed(const char *path, ldpc_gpu_code **out) { return load_alist(path, true, out); }

// GF(2) reduced row-echelon form of H, dense bit rows.
static int build_encoder(ldpc_gpu_code *c)
{
    const int N = c->N, M = c->M;
    if ((long long)N * M > (1ll << 28)) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "encoder: dense GF(2) elimination is limited to M*N <= 2^28 bits");
    const int W = (N + 63) / 64;
    std::vector<uint64_t> A((size_t)M * W, 0);
    for (int j = 0; j < M; j++) for (int k = 0; k < c->row_deg[j]; k++) { const int i = c->mlist[(size_t)j * c->dc_max + k]; A[(size_t)j * W + (i >> 6)] ^= 1ull << (i & 63); }
    c->pivot_col.clear();
    int r = 0;
    for (int col = 0; col < N && r < M; col++) {
        int piv = -1;
        for (int j = r; j < M; j++) if (A[(size_t)j * W + (col >> 6)] >> (col & 63) & 1) { piv = j; break; }
        if (piv < 0) continue;
        if (piv != r) for (int w = 0; w < W; w++) std::swap(A[(size_t)piv * W + w], A[(size_t)r * W + w]);
        for (int j = 0; j < M; j++) if (j != r && (A[(size_t)j * W + (col >> 6)] >> (col & 63) & 1))
            for (int w = 0; w < W; w++) A[(size_t)j * W + w] ^= A[(size_t)r * W + w];
        c->pivot_col.push_back(col); r++;
    }
    c->rank = r;
    std::vector<int> free_index(N, -1);
    c->info_col.clear();
    { size_t pi = 0; for (int col = 0; col < N; col++) { if (pi < c->pivot_col.size() && c->pivot_col[pi] == col) { pi++; continue; } free_index[col] = (int)c->info_col.size(); c->info_col.push_back(col); } }
    c->parity_of.assign(r, std::vector<int>());
    for (int q = 0; q < r; q++) for (int col = 0; col < N; col++) if (free_index[col] >= 0 && (A[(size_t)q * W + (col >> 6)] >> (col & 63) & 1)) c->parity_of[q].push_back(free_index[col]);
    c->enc_ready = true;
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_code_random_codewords(ldpc_gpu_code *c, uint64_t seed, int64_t n, uint8_t *bits01, int *rank)
{
    if (!c || (n > 0 && !bits01) || n < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "bad argument");
    {
        std::lock_guard<std::mutex> lock(c->enc_mutex);
        if (!c->enc_ready) { int rc = build_encoder(c); if (rc) return rc; }
    }
    if (rank) *rank = c->rank;
    const int N = c->N, K = (int)c->info_col.size();
    std::vector<uint8_t> u(K);
    for (int64_t f = 0; f < n; f++) {
        uint8_t *cw = bits01 + (size_t)f * N;
        for (int b = 0; b * 128 < K; b++) {                        // 128 information bits per Philox block
            uint32_t r4[4];
            philox4x32_10((uint32_t)b, 3u, (uint32_t)f, (uint32_t)((uint64_t)f >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r4);
            for (int q = 0; q < 128 && b * 128 + q < K; q++) u[b * 128 + q] = (uint8_t)((r4[q >> 5] >> (q & 31)) & 1u);
        }
        for (int k = 0; k < K; k++) cw[c->info_col[k]] = u[k];
        for (int q = 0; q < c->rank; q++) { uint8_t pbit = 0; for (int k : c->parity_of[q]) pbit ^= u[k]; cw[c->pivot_col[q]] = pbit; }
    }
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_code_dims(const ldpc_gpu_code *c, int *N, int *M, int *E, int *dv, int *dc)
{
    if (!c) return set_err(LDPC_GPU_ERR_INVALID_ARG, "code is NULL");
    if (N) *N = c->N; if (M) *M = c->M; if (E) *E = c->E; if (dv) *dv = c->dv_max; if (dc) *dc = c->dc_max;
    return LDPC_GPU_OK;
}
extern "C" int ldpc_gpu_code_destroy(ldpc_gpu_code *c) { delete c; return LDPC_GPU_OK; }

// ------------------------------------------------------------------------------------------------
// decoder handle
// ------------------------------------------------------------------------------------------------
typedef void (*KernelFn)(const CodeDev, const DecParams, const FrameIO);

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return LDPC_GPU_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) { cudaGetLastError(); return set_err(LDPC_GPU_ERR_NOMEM, "cudaMalloc failed for a staging buffer"); }
        cap = n; return LDPC_GPU_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Slot {                    // one in-flight chunk of a host-memory batch
    cudaStream_t st = nullptr; cudaEvent_t k0 = nullptr, k1 = nullptr;
    DevBuf y, noise, cw, qp, bits, iters, soft, errs, flags;
    bool used = false;
    ldpc::FrameIO io;            // the chunk in flight (for the redo launch of the exact-lattice kernel)
    long long f0 = 0, nf = 0;
};

struct ldpc_gpu_decoder {
    int device = 0, n_sm = 0;
    ldpc_gpu_decoder_cfg cfg;
    CodeDev dev;                 // device pointers owned below
    std::vector<void *> owned;
    KernelFn fn = nullptr;
    KernelFn fn_staged = nullptr;   // exact-lattice kernel, bit-packed levels staged by the TMA unit (ldpc_ms_x2.cuh); same geometry as fn
    int block = 0, smem = 0, ctas_per_sm = 0, grid_full = 0;
    bool gstate = false; size_t ws_stride = 0; unsigned char *d_ws = nullptr;   // HBM-resident frame state
    int frames_per_cta = 1;                                                      // > 1: frame-interleaved tile kernel
    DecParams base;
    Slot slot[2];
    unsigned long long *d_counters = nullptr, *d_ew = nullptr, *d_it = nullptr, *d_ph = nullptr;
    uint8_t *d_cwtab = nullptr; long long n_cw = 0;
    double last_kernel_ms = 0; long long last_launches = 0;
    int N = 0, M = 0;
    // exact-lattice packed kernel (ldpc_ms_x2.cuh): frames it cannot certify are decoded by an fp64 decoder on the same stream
    bool x2 = false; float x2_cap = 0, x2_m0 = 0;
    bool x2_tile = false;        // ... its HBM-bound form (ldpc_ms_tileh.cuh): same redo protocol, no certificate
    ldpc_gpu_decoder *redo = nullptr;
    long long *d_redo_list = nullptr; size_t redo_cap = 0;
    unsigned int *d_redo_count = nullptr; unsigned long long *d_redo_total = nullptr;
    unsigned int *h_redo = nullptr;          // pinned [2]: the two slots' redo counts, copied back behind every launch
    std::vector<int> h_mlist, h_rowdeg; int h_dcm = 0;   // host copy of the rows of H (syndromes of ldpc_gpu_replay_frame)
};

// Does the configuration put plain / offset min-sum on a lattice binary16 holds exactly (ldpc_ms_x2.cuh)?
// u = Ymax / (2^Q - 1) = step / 2 must be a power of two, delta a multiple of u, and the c2v cap C must leave room on both
// sides: |S| <= Ymax + dv_max C <= 2047 u, and (dv_min - 1) C - Ymax >= m0, the stable-state bound of the lemma.
static bool x2_lattice(const ldpc_gpu_decoder_cfg &c, int dv_min, int dv_max, float *cap, float *m0)
{
    if (c.kind != LDPC_GPU_KIND_MINSUM || !(c.flags & LDPC_GPU_F_QUANTIZE_SAMPLES) || (c.flags & LDPC_GPU_F_NORMALIZED_MS)) return false;
    if (c.Q < 2 || c.Q > 8 || !(c.Ymax > 0) || dv_min < 3) return false;
    const double Yu = pow(2.0, c.Q) - 1.0, u = c.Ymax / Yu;
    int ex;
    if (frexp(u, &ex) != 0.5 || ex - 1 < -14 || ex - 1 > 4) return false;
    const double du = (c.flags & LDPC_GPU_F_OFFSET_MS) ? c.delta / u : 0.0;
    if (du < 0 || du != floor(du) || du > 64) return false;
    double cap_u = floor((2047.0 - Yu) / dv_max);
    if (const char *e = getenv("LDPC_GPU_X2_CAP_UNITS")) cap_u = std::min(cap_u, (double)atoi(e));   // test hook: a low cap exercises the redo path
    const double m0_u = ceil((Yu + (dv_min - 1) * du) / (dv_min - 2));
    if ((dv_min - 1) * cap_u - Yu < m0_u || cap_u < m0_u) return false;
    *cap = (float)(cap_u * u); *m0 = (float)(m0_u * u);
    return true;
}

// The HBM-bound form (ldpc_ms_tileh.cuh) runs T fixed iterations and needs the lattice and the cap only: every value stays an
// integer multiple of u below 2^11 u while |c2v| <= C, Ymax + dv_max C <= 2047 u.
static bool tile_lattice(const ldpc_gpu_decoder_cfg &c, int dv_max, float *cap)
{
    if (c.kind != LDPC_GPU_KIND_MINSUM || !(c.flags & LDPC_GPU_F_QUANTIZE_SAMPLES) || (c.flags & LDPC_GPU_F_NORMALIZED_MS)) return false;
    if (c.Q < 2 || c.Q > 8 || !(c.Ymax > 0) || dv_max < 1) return false;
    const double Yu = pow(2.0, c.Q) - 1.0, u = c.Ymax / Yu;
    int ex;
    if (frexp(u, &ex) != 0.5 || ex - 1 < -14 || ex - 1 > 4) return false;
    const double du = (c.flags & LDPC_GPU_F_OFFSET_MS) ? c.delta / u : 0.0;
    if (du < 0 || du != floor(du) || du > 64) return false;
    double cap_u = floor((2047.0 - Yu) / dv_max);
    if (const char *e = getenv("LDPC_GPU_X2_CAP_UNITS")) cap_u = std::min(cap_u, (double)atoi(e));   // test hook: a low cap exercises the redo path
    if (cap_u < 1) return false;
    *cap = (float)(cap_u * u);
    return true;
}

static int rows_per_step(uint32_t fl) { return ((fl & LDPC_GPU_F_ADDNOISE) ? 1 : 0) + ((fl & LDPC_GPU_F_QUANTIZEPROBABILITIES) ? 1 : 0); }
static int iter_hist_len(const ldpc_gpu_decoder_cfg &c)
{
    int ph = ((c.flags & LDPC_GPU_F_REDECODE) && c.kind == LDPC_GPU_KIND_GDBF && c.maxphase > 1) ? c.maxphase : 1;
    return c.num_iterations * ph + 1;
}

extern "C" int ldpc_gpu_iter_hist_len(const ldpc_gpu_decoder_cfg *c) { return c ? iter_hist_len(*c) : 0; }

extern "C" int ldpc_gpu_decoder_cfg_default(int kind, ldpc_gpu_decoder_cfg *c)
{
    if (!c) return set_err(LDPC_GPU_ERR_INVALID_ARG, "cfg is NULL");
    memset(c, 0, sizeof *c);
    c->kind = kind; c->precision = LDPC_GPU_PREC_F64; c->num_iterations = 10;
    c->Ymax = 2.25; c->Q = 5; c->NQ = 16;                         // decodeGDBF.cpp:48-56
    c->alpha = 2.25; c->delta = 0; c->theta = -0.6; c->lambda = 0.991; c->noiseScale = 1.0;
    c->windowsize = 64; c->maxphase = 7; c->Tswitch = 0;
    c->w = 0.185; c->theta0 = -0.525; c->MAXLLR = 20;             // NGDBFhw.cpp:48-57, decodeBP.cpp:58
    if (kind == LDPC_GPU_KIND_NGDBF_HW) { c->num_iterations = 600; c->Ymax = 1.625; c->noiseScale = 0.95; c->maxphase = 1; c->NQ = 5; }
    if (kind == LDPC_GPU_KIND_MINSUM) c->alpha = 1.25;
    if (kind == LDPC_GPU_KIND_NGDBF_SC) {                         // SystemC/NGDBF/example.sh
        c->num_iterations = 100; c->theta = -0.5; c->lambda = 0.975; c->Q = 4; c->Ymax = 3.0; c->alpha = 0.95; c->windowsize = 32; c->noiseScale = 1.0;
    }
    return LDPC_GPU_OK;
}

template <typename T> static int upload(ldpc_gpu_decoder *d, const std::vector<T> &h, const T **out)
{
    void *p = nullptr;
    CU_TRY(cudaMalloc(&p, std::max<size_t>(16, h.size() * sizeof(T))));
    d->owned.push_back(p);
    CU_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)p;
    return LDPC_GPU_OK;
}

// Regular codes: conflict-free check schedule + variable relabelling (ldpc_schedule.h).
static int build_schedule(ldpc_gpu_decoder *d, const ldpc_gpu_code *c, size_t real_bytes)
{
    CodeDev &v = d->dev;
    v.sched = nullptr; v.col_of_var = nullptr; v.var_of_col = nullptr; v.row_slot = nullptr;
    if (v.regular_dc <= 0 || v.regular_dc % 4 || v.N > 65535 || getenv("LDPC_GPU_NO_SCHED")) return LDPC_GPU_OK;
    const int N = v.N, M = v.M, dc = v.regular_dc, dvm = c->dv_max, dcm = c->dc_max;
    std::vector<int> ml((size_t)M * dc);
    for (int j = 0; j < M; j++) for (int k = 0; k < dc; k++) ml[(size_t)j * dc + k] = c->mlist[(size_t)j * dcm + k];
    RowSchedule rs = build_row_schedule(N, M, dc, ml);
    if (!rs.ok) return LDPC_GPU_OK;                               // no perfect schedule: the unscheduled kernel is used
    std::vector<uint32_t> tab((size_t)(dc / 4) * M * 4, 0);
    for (int j = 0; j < M; j++) for (int t = 0; t < dc; t++) {
        const int k = rs.order[(size_t)j * dc + t];
        const int i = c->mlist[(size_t)j * dcm + k], s = c->vn_slot[(size_t)j * dcm + k];
        tab[((size_t)(t / 4) * M + j) * 4 + (t % 4)] = (uint32_t)(((size_t)s * N + rs.col[i]) * real_bytes);
    }
    (void)dvm;
    std::vector<uint16_t> cov(N), voc(N);
    for (int i = 0; i < N; i++) { cov[i] = (uint16_t)rs.col[i]; voc[i] = (uint16_t)rs.var_of_col[i]; }
    const uint32_t *pt; const uint16_t *p1, *p2; int rc;
    if ((rc = upload(d, tab, &pt)) || (rc = upload(d, cov, &p1)) || (rc = upload(d, voc, &p2))) return rc;
    v.sched = reinterpret_cast<const uint4 *>(pt); v.col_of_var = p1; v.var_of_col = p2;
    // rows whose edges all sit in one slot of their variables' lists (block-structured codes): ldpc_ms_rc.cuh
    std::vector<uint8_t> rslot(M);
    bool uniform = true;
    for (int j = 0; j < M && uniform; j++) {
        rslot[j] = (uint8_t)c->vn_slot[(size_t)j * dcm];
        for (int k = 1; k < dc; k++) if (c->vn_slot[(size_t)j * dcm + k] != rslot[j]) { uniform = false; break; }
    }
    if (uniform) { const uint8_t *p3; if ((rc = upload(d, rslot, &p3))) return rc; v.row_slot = p3; }
    return LDPC_GPU_OK;
}

// Small codes (ldpc_ms_quad.cuh): rows sorted by weight, one per thread; per-thread edge words in an order in which the eight lanes of
// a 16-byte access phase fall into distinct 16-byte bank groups (ldpc_schedule.h, build_group_schedule), natural order otherwise.
static int build_quad_tables(ldpc_gpu_decoder *d, const ldpc_gpu_code *c)
{
    CodeDev &v = d->dev;
    v.quad_edge = nullptr; v.quad_steps = nullptr; v.quad_col_of_var = nullptr; v.quad_var_of_col = nullptr;
    const int N = v.N, M = v.M, dvm = c->dv_max, dcm = c->dc_max;
    if (M > 512 || dcm > 8 || dvm > 4 || v.dvN > 65535 || N > 65534) return LDPC_GPU_OK;
    for (int j = 0; j < M; j++) if (c->row_deg[j] < 2) return LDPC_GPU_OK;            // the padded row update needs two finite minima
    std::vector<int> rdeg(M), ml((size_t)M * dcm, 0);
    for (int j = 0; j < M; j++) { rdeg[j] = c->row_deg[j]; for (int k = 0; k < rdeg[j]; k++) ml[(size_t)j * dcm + k] = c->mlist[(size_t)j * dcm + k]; }
    GroupSchedule gs;
    if (!getenv("LDPC_GPU_NO_SCHED")) gs = build_group_schedule(N, M, dcm, rdeg, ml, 8);
    if (!gs.ok) {                                                                      // natural order, rows still sorted by weight
        gs.col.resize(N); gs.var_of_col.resize(N); gs.row_of_thread.resize(M); gs.step.assign((size_t)M * dcm, -1);
        std::iota(gs.col.begin(), gs.col.end(), 0); std::iota(gs.var_of_col.begin(), gs.var_of_col.end(), 0);
        std::iota(gs.row_of_thread.begin(), gs.row_of_thread.end(), 0);
        std::stable_sort(gs.row_of_thread.begin(), gs.row_of_thread.end(), [&](int a, int b) { return rdeg[a] > rdeg[b]; });
        for (int t = 0; t < M; t++) for (int k = 0; k < rdeg[gs.row_of_thread[t]]; k++) gs.step[(size_t)t * dcm + k] = k;
    }
    std::vector<uint32_t> tab((size_t)8 * M, ((uint32_t)N << 16) | (uint32_t)v.dvN);
    std::vector<uint8_t> steps(M, 0);
    for (int t = 0; t < M; t++) {
        const int j = gs.row_of_thread[t];
        for (int s = 0; s < dcm; s++) {
            const int k = gs.step[(size_t)t * dcm + s];
            if (k < 0) continue;
            const int i = c->mlist[(size_t)j * dcm + k], sl = c->vn_slot[(size_t)j * dcm + k];
            tab[(size_t)s * M + t] = ((uint32_t)gs.col[i] << 16) | (uint32_t)(sl * N + gs.col[i]);
            steps[t] = (uint8_t)(s + 1);
        }
    }
    std::vector<uint16_t> cov(N), voc(N);
    for (int i = 0; i < N; i++) { cov[i] = (uint16_t)gs.col[i]; voc[i] = (uint16_t)gs.var_of_col[i]; }
    int rc;
    if ((rc = upload(d, tab, &v.quad_edge)) || (rc = upload(d, steps, &v.quad_steps)) || (rc = upload(d, cov, &v.quad_col_of_var)) ||
        (rc = upload(d, voc, &v.quad_var_of_col))) return rc;
    return LDPC_GPU_OK;
}

static int build_device_code(ldpc_gpu_decoder *d, const ldpc_gpu_code *c)
{
    CodeDev &v = d->dev;
    const int N = c->N, M = c->M, dvm = c->dv_max, dcm = c->dc_max;
    v.N = N; v.M = M; v.E = c->E; v.dv_max = dvm; v.dc_max = dcm; v.dvN = dvm * N;
    v.idx16 = (v.dvN <= 65535) ? 1 : 0;
    v.regular_dc = c->row_deg[0]; for (int j = 0; j < M; j++) if (c->row_deg[j] != v.regular_dc) v.regular_dc = -1;
    v.regular_dv = c->col_deg[0]; for (int i = 0; i < N; i++) if (c->col_deg[i] != v.regular_dv) v.regular_dv = -1;
    std::vector<uint8_t> cdeg(M), vdeg(N);
    for (int j = 0; j < M; j++) cdeg[j] = (uint8_t)c->row_deg[j];
    for (int i = 0; i < N; i++) vdeg[i] = (uint8_t)c->col_deg[i];
    std::vector<uint32_t> cn_var((size_t)dcm * M, 0), vn_chk((size_t)dvm * N, 0);
    for (int j = 0; j < M; j++) for (int k = 0; k < c->row_deg[j]; k++) cn_var[(size_t)k * M + j] = (uint32_t)c->mlist[(size_t)j * dcm + k];
    for (int i = 0; i < N; i++) for (int s = 0; s < c->col_deg[i]; s++) vn_chk[(size_t)s * N + i] = (uint32_t)c->nlist[(size_t)i * dvm + s];
    const int VPL = v.idx16 ? 8 : 4;
    int groups = (dcm + VPL - 1) / VPL;
    if (v.idx16 && dcm > 8 && dcm <= 32) groups = 4;             // ms_fast_kernel<.., DC=32, ..> loads four vectors per row
    int rc;
    if (v.idx16) {
        std::vector<uint16_t> pos((size_t)groups * M * VPL, 0);
        for (int j = 0; j < M; j++) for (int k = 0; k < c->row_deg[j]; k++)
            pos[((size_t)(k / VPL) * M + j) * VPL + (k % VPL)] = (uint16_t)(c->vn_slot[(size_t)j * dcm + k] * N + c->mlist[(size_t)j * dcm + k]);
        const uint16_t *p; if ((rc = upload(d, pos, &p))) return rc; v.cn_pos = p;
    } else {
        std::vector<uint32_t> pos((size_t)groups * M * VPL, 0);
        for (int j = 0; j < M; j++) for (int k = 0; k < c->row_deg[j]; k++)
            pos[((size_t)(k / VPL) * M + j) * VPL + (k % VPL)] = (uint32_t)(c->vn_slot[(size_t)j * dcm + k] * N + c->mlist[(size_t)j * dcm + k]);
        const uint32_t *p; if ((rc = upload(d, pos, &p))) return rc; v.cn_pos = p;
    }
    if ((rc = upload(d, cdeg, &v.cn_deg))) return rc;
    if ((rc = upload(d, vdeg, &v.vn_deg))) return rc;
    if ((rc = upload(d, cn_var, &v.cn_var))) return rc;
    if ((rc = upload(d, vn_chk, &v.vn_chk))) return rc;
    return LDPC_GPU_OK;
}

static int round32(int x) { return (x + 31) & ~31; }

static int pick_kernel(ldpc_gpu_decoder *d)
{
    const CodeDev &v = d->dev; const int kind = d->cfg.kind; const bool f64 = d->cfg.precision == LDPC_GPU_PREC_F64;
    size_t smem = 0; int block = 128;
    int max_optin = 0;
    CU_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d->device));
    if (kind == LDPC_GPU_KIND_MINSUM || kind == LDPC_GPU_KIND_BP || kind == LDPC_GPU_KIND_DDBMP) {
        if (v.dc_max > 64) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "message-passing kernels hold a row's signs in 64 bits: dc_max > 64");
        const int algo = kind == LDPC_GPU_KIND_MINSUM ? ALGO_MS : kind == LDPC_GPU_KIND_BP ? ALGO_BP : ALGO_DDBMP;
        if (d->cfg.precision == LDPC_GPU_PREC_F16X2 && algo == ALGO_MS && v.dv_max <= 8 && v.dc_max <= 8 && tile_lattice(d->cfg, v.dv_max, &d->x2_cap) &&
            (mp_smem_bytes<float>(v, algo) > (size_t)max_optin || getenv("LDPC_GPU_FORCE_HBM_STATE"))) {
            // exact lattice, messages in HBM (DVB-S2-class codes): binary16 tile kernel, fp64 redo of the frames it cannot vouch for
            d->x2 = true; d->x2_tile = true; d->gstate = true;
            // 640 threads: the pipelined kernel needs 80 registers, and the compute-bound front end gains from 20 warps (measured 512: 16.0, 640: 16.4, 704: 16.1 Gbit/s)
            d->fn = v.idx16 ? (KernelFn)ms_tileh_kernel<uint16_t, 8, 8, 640> : (KernelFn)ms_tileh_kernel<uint32_t, 8, 8, 640>;
            block = 640; d->frames_per_cta = TILEH_FI;
            d->ws_stride = ms_tileh_state_bytes(v);
            smem = ms_tileh_smem_bytes(v) + ms_tileh_pipe_bytes(block, 8);
            goto geometry;
        }
        if (d->cfg.precision == LDPC_GPU_PREC_F16X2) {
            if (!(v.sched && v.regular_dc == 32 && v.regular_dv == 6 && v.N == 2048 && v.M <= 384))
                return set_err(LDPC_GPU_ERR_UNSUPPORTED, "LDPC_GPU_PREC_F16X2 is built for regular (6,32) codes of length 2048 with a conflict-free schedule (the 802.3an H)");
            d->fn = (KernelFn)ms_h2_kernel<32, 6, 2048, 384, 2>;
            d->frames_per_cta = 2;
            smem = ms_h2_smem_bytes(v); block = 384;
            if (v.row_slot && !getenv("LDPC_GPU_NO_RC")) {                 // c2v pairs resident in registers (ldpc_ms_h2rc.cuh)
                d->fn = (KernelFn)ms_h2rc_kernel<32, 6, 2048, 384, 2>;
                smem = ms_h2rc_smem_bytes(v);
                // exact lattice: decisions certified identical to the reference's, fp64 redo of the rest (ldpc_ms_x2.cuh)
                if (!getenv("LDPC_GPU_NO_X2") && v.M == 384 && x2_lattice(d->cfg, 6, 6, &d->x2_cap, &d->x2_m0)) {
                    d->x2 = true;
                    const char *xv = getenv("LDPC_GPU_X2_VARIANT");   // A/B switch: offsets kept in registers | re-read, 2 or 3 CTAs per SM
                    const int variant = xv ? atoi(xv) : 1;
                    d->fn = variant == 0 ? (KernelFn)ms_x2_kernel<32, 6, 2048, 384, 2, true>
                          : variant == 2 ? (KernelFn)ms_x2_kernel<32, 6, 2048, 384, 3, false> : (KernelFn)ms_x2_kernel<32, 6, 2048, 384, 2, false>;
                    if (variant == 1 && !getenv("LDPC_GPU_NO_TMA")) d->fn_staged = (KernelFn)ms_x2_kernel<32, 6, 2048, 384, 2, false, true>;
                    smem = ms_x2_smem_bytes(v);
                }
            }
            goto geometry;
        }
        smem = f64 ? mp_smem_bytes<double>(v, algo) : mp_smem_bytes<float>(v, algo);
        if (smem > (size_t)max_optin || getenv("LDPC_GPU_FORCE_HBM_STATE")) {
            d->gstate = true;
            d->ws_stride = f64 ? mp_state_bytes<double>(v, algo) : mp_state_bytes<float>(v, algo);
            smem = f64 ? mp_smem_bytes<double>(v, algo, true) : mp_smem_bytes<float>(v, algo, true);
        }
#define MP_PICK2(A, G)                                                                                                     \
    d->fn = f64 ? (v.idx16 ? (KernelFn)mp_kernel<double, uint16_t, A, G> : (KernelFn)mp_kernel<double, uint32_t, A, G>)   \
                : (v.idx16 ? (KernelFn)mp_kernel<float, uint16_t, A, G> : (KernelFn)mp_kernel<float, uint32_t, A, G>)
#define MP_PICK(A) do { if (d->gstate) MP_PICK2(A, true); else MP_PICK2(A, false); } while (0)
        if (algo == ALGO_MS) MP_PICK(ALGO_MS); else if (algo == ALGO_BP) MP_PICK(ALGO_BP); else MP_PICK(ALGO_DDBMP);
#undef MP_PICK
#undef MP_PICK2
        block = std::min(1024, std::max(128, round32(v.M)));
        if (algo == ALGO_MS && d->gstate && v.dv_max <= 8 && v.dc_max <= 64 && !getenv("LDPC_GPU_NO_TILE")) {
            // HBM-bound codes: frame-interleaved tile kernel (ldpc_ms_tile.cuh)
#define TILE_PICK(DC, NT, VB) (f64 ? (v.idx16 ? (KernelFn)ms_tile_kernel<double, uint16_t, DC, 8, NT, (VB < 8 ? 8 : VB)> : (KernelFn)ms_tile_kernel<double, uint32_t, DC, 8, NT, (VB < 8 ? 8 : VB)>) \
                                   : (v.idx16 ? (KernelFn)ms_tile_kernel<float, uint16_t, DC, 8, NT, VB> : (KernelFn)ms_tile_kernel<float, uint32_t, DC, 8, NT, VB>))
            // 16-byte accesses where the row fits the register file (dc <= 8: 8 x 16 B per thread in flight)
            if (v.dc_max <= 8) { d->fn = TILE_PICK(8, 512, 16); block = 512; }
            else if (v.dc_max <= 32) { d->fn = TILE_PICK(32, 512, 4); block = 512; }
            else { d->fn = TILE_PICK(64, 256, 4); block = 256; }
#undef TILE_PICK
            d->frames_per_cta = f64 ? TileFI<double>::value : TileFI<float>::value;
            d->ws_stride = f64 ? ms_tile_state_bytes<double>(v) : ms_tile_state_bytes<float>(v);
            smem = f64 ? ms_tile_smem_bytes<double>(v) : ms_tile_smem_bytes<float>(v);
        }
        if (algo == ALGO_BP && !f64 && !d->gstate && v.sched && v.regular_dc == 32 && v.regular_dv == 6 && v.N == 2048 && v.M <= 384 &&
            !getenv("LDPC_GPU_GENERIC_BP")) {
            d->fn = (KernelFn)ms_sched_kernel<float, 32, 6, 2048, 384, 2, ALGO_BP>;    // O(dc) phi-domain sum-product
            block = 384; smem = ms_sched_smem_bytes<float>(v);
            if (v.row_slot && !getenv("LDPC_GPU_NO_RC")) {                              // c2v resident in the row thread's registers (ldpc_ms_rc.cuh)
                d->fn = (KernelFn)ms_rc_kernel<float, 32, 6, 2048, 384, 2, true, ALGO_BP>;   // (one CTA per SM at 160 registers, no spills: 6.8 vs 7.0 Gbit/s)
                smem = ms_rc_smem_bytes<float>(v);
            }
        }
        if (algo == ALGO_MS && v.idx16 && !d->gstate && !getenv("LDPC_GPU_GENERIC_MS")) {
            // degree-specialised min-sum kernel where an instantiation covers the code
            const bool rc = v.regular_dc > 0, rv = v.regular_dv > 0;
            KernelFn fast = nullptr;
#define MS_FAST(DC, DV, RC, RV) (f64 ? (KernelFn)ms_fast_kernel<double, DC, DV, RC, RV, 1024, 1> : (KernelFn)ms_fast_kernel<float, DC, DV, RC, RV, 1024, 1>)
            if (v.sched && v.regular_dc == 32 && v.regular_dv == 6 && v.N == 2048 && v.M <= 384) {   // the 802.3an H, scheduled
                smem = f64 ? ms_sched_smem_bytes<double>(v) : ms_sched_smem_bytes<float>(v);
                // 2 CTAs/SM at 80 registers is the measured optimum for fp32 (1 CTA: -15 %, 3 CTAs: spills, -30 %)
                fast = f64 ? (KernelFn)ms_sched_kernel<double, 32, 6, 2048, 384, 1> : (KernelFn)ms_sched_kernel<float, 32, 6, 2048, 384, 2>;
                if (v.row_slot && !getenv("LDPC_GPU_NO_RC")) {        // c2v resident in the row thread's registers (ldpc_ms_rc.cuh)
                    smem = f64 ? ms_rc_smem_bytes<double>(v) : ms_rc_smem_bytes<float>(v);
                    fast = f64 ? (KernelFn)ms_rc_kernel<double, 32, 6, 2048, 384, 1> : (KernelFn)ms_rc_kernel<float, 32, 6, 2048, 384, 2>;
                    block = 384;
                }
            }
            else if (v.sched && v.regular_dc == 32 && v.dv_max == 6 && v.N == 2048 && v.M <= 384 && !getenv("LDPC_GPU_NO_RC")) {
                // the full-rank 802.3an H (802_3.alist: M = 325, dv in {5, 6}): the same register-resident kernel, rows that mix slots
                smem = f64 ? ms_rc_smem_bytes<double>(v) : ms_rc_smem_bytes<float>(v);
                fast = f64 ? (KernelFn)ms_rc_kernel<double, 32, 6, 2048, 384, 1, false> : (KernelFn)ms_rc_kernel<float, 32, 6, 2048, 384, 2, false>;
                block = 384;
            }
            else if (v.regular_dc == 32 && v.regular_dv == 6) fast = MS_FAST(32, 6, true, true);
            else if (v.regular_dc == 32 && v.dv_max <= 6) fast = MS_FAST(32, 6, true, false);
            else if (!f64 && v.quad_edge && !getenv("LDPC_GPU_NO_QUAD") && ms_quad_smem_bytes(v) * 2 <= (size_t)max_optin) {
                // small codes (BASELINE configs[0], the (3,6) PEG code): four frames per thread, c2v resident in registers (ldpc_ms_quad.cuh)
                fast = (v.regular_dv == 3) ? (KernelFn)ms_quad_kernel<8, 3, true, 512, 2> : (KernelFn)ms_quad_kernel<8, 4, false, 512, 2>;
                d->frames_per_cta = 4;
                smem = ms_quad_smem_bytes(v);
                block = std::max(128, round32(v.M));
            }
            else if (v.dc_max <= 8 && v.regular_dv == 3 && v.M <= 512 && !f64 && !getenv("LDPC_GPU_NO_SMALL"))
                fast = (KernelFn)ms_fast_kernel<float, 8, 3, false, true, 512, 4>;     // small (3,6)-class codes: 4 frames per SM at 32 registers (3.17 -> 3.46 Gbit/s on PEG, T = 50; 3 frames at 40 registers: 3.39)
            else if (v.dc_max <= 8 && v.regular_dv == 3) fast = MS_FAST(8, 3, false, true);
            else if (v.regular_dc == 8 && v.regular_dv == 4) fast = MS_FAST(8, 4, true, true);
            else if (v.dc_max <= 8 && v.dv_max <= 8) fast = MS_FAST(8, 8, false, false);
            else if (v.dc_max <= 32 && v.dv_max <= 8) fast = MS_FAST(32, 8, false, false);
#undef MS_FAST
            (void)rc; (void)rv;
            if (fast) d->fn = fast;
        }
    } else if (kind == LDPC_GPU_KIND_GDBF) {
        smem = f64 ? gdbf_smem_bytes<double>(v) : gdbf_smem_bytes<float>(v);
        if (smem > (size_t)max_optin || getenv("LDPC_GPU_FORCE_HBM_STATE")) {
            d->gstate = true;
            d->ws_stride = f64 ? gdbf_state_bytes<double>(v) : gdbf_state_bytes<float>(v);
            smem = f64 ? gdbf_smem_bytes<double>(v, true) : gdbf_smem_bytes<float>(v, true);
        }
        d->fn = d->gstate ? (f64 ? (KernelFn)gdbf_kernel<double, true> : (KernelFn)gdbf_kernel<float, true>)
                          : (f64 ? (KernelFn)gdbf_kernel<double, false> : (KernelFn)gdbf_kernel<float, false>);
        // parallel-flipping variants on codes that fit an SM: bit-packed kernel with incremental syndromes
        const size_t par_smem = f64 ? gdbf_par_smem_bytes<double>(v) : gdbf_par_smem_bytes<float>(v);
        if (!d->gstate && !(d->cfg.flags & (LDPC_GPU_F_SEQUENTIALMODE | LDPC_GPU_F_MODESWITCHING)) && v.M <= 65535 &&
            par_smem <= (size_t)max_optin && !getenv("LDPC_GPU_GENERIC_GDBF")) {
            d->fn = f64 ? (KernelFn)gdbf_par_kernel<double> : (KernelFn)gdbf_par_kernel<float>;
            smem = par_smem;
        }
        block = std::min(1024, std::max(128, round32(std::max(v.M, (v.N + 3) / 4))));
        if (v.M <= 512 && v.N <= 4096) block = 256;                           // measured on the 802.3an H: 256 > 384 > 512
    } else if (kind == LDPC_GPU_KIND_NGDBF_HW) {
        if (v.N >= LDPC_GPU_HW_QBUF) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "NGDBFhw's 2648-entry noise window needs N < 2648 (src/NGDBFhw.cpp:151)");
        d->fn = v.regular_dv == 6 ? (KernelFn)hw_kernel<6> : (KernelFn)hw_kernel<0>;
        smem = hw_smem_bytes(v);
        // measured on the 802.3an H (524 288 frames at 4.5 dB): 384 threads x 3 CTAs/SM 18.3 Gbit/s, 256 x 4 19.2, 192 x 6 19.3, 128 x 7 19.6: the two
        // barriers of an iteration stall a whole CTA, so many small CTAs (independent frames) hide them better than few large ones
        block = 128;
    } else if (kind == LDPC_GPU_KIND_NGDBF_SC) {
        d->fn = (KernelFn)sc_kernel;
        smem = sc_smem_bytes(v, d->cfg.num_iterations, d->cfg.Q);
        block = std::min(1024, std::max(128, round32(std::max(v.M, (v.N + 1) / 2))));
    } else return set_err(LDPC_GPU_ERR_INVALID_ARG, "unknown decoder kind");
geometry:
    if (smem > (size_t)max_optin)
        return set_err(LDPC_GPU_ERR_UNSUPPORTED, "shared-memory scratch (" + std::to_string(smem) + " B) exceeds one SM");
    if ((d->cfg.flags & LDPC_GPU_F_CERT_STOP) && (!d->x2 || d->x2_tile))
        return set_err(LDPC_GPU_ERR_UNSUPPORTED, "LDPC_GPU_F_CERT_STOP needs LDPC_GPU_PREC_F16X2 on an exact lattice (plain / offset min-sum, dyadic quantiser step, the 802.3an H)");
    CU_TRY(cudaFuncSetAttribute((const void *)d->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (d->fn_staged) CU_TRY(cudaFuncSetAttribute((const void *)d->fn_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    CU_TRY(cudaFuncGetAttributes(&fa, (const void *)d->fn));
    const int by_regs = (65536 / std::max(1, fa.numRegs)) & ~31;      // one CTA must fit the register file
    const int block_wanted = block;
    block = std::max(32, std::min(block, std::min(by_regs, fa.maxThreadsPerBlock & ~31)));
    if (block != block_wanted && (d->fn == (KernelFn)ms_rc_kernel<float, 32, 6, 2048, 384, 2> || d->fn == (KernelFn)ms_rc_kernel<float, 32, 6, 2048, 384, 2, true, ALGO_BP> || d->fn == (KernelFn)ms_rc_kernel<double, 32, 6, 2048, 384, 1> ||
                                  d->fn == (KernelFn)ms_rc_kernel<float, 32, 6, 2048, 384, 2, false> || d->fn == (KernelFn)ms_rc_kernel<double, 32, 6, 2048, 384, 1, false>))
        return set_err(LDPC_GPU_ERR_UNSUPPORTED, "ms_rc_kernel is written for blocks of exactly 384 threads");
    int nb = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)d->fn, block, smem));
    if (nb < 1) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "kernel does not fit on an SM");
    d->block = block; d->smem = (int)smem; d->ctas_per_sm = nb; d->grid_full = nb * d->n_sm;
    if (d->gstate) {
        // HBM-resident state: enough CTAs to fill the SMs, bounded so the workspace stays modest
        d->ws_stride = (d->ws_stride + 255) & ~(size_t)255;
        const size_t budget = (size_t)12 << 30;
        while (d->grid_full > d->n_sm && (size_t)d->grid_full * d->ws_stride > budget) d->grid_full -= d->n_sm;
        // two workspaces: the host pipeline keeps one launch per slot in flight, and CTA b of the second may start while CTA b of
        // the first is still running
        CU_TRY(cudaMalloc(&d->d_ws, 2 * (size_t)d->grid_full * d->ws_stride));
    }
    return LDPC_GPU_OK;
}

static int validate_cfg(const ldpc_gpu_decoder_cfg &c)
{
    if (c.num_iterations < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "num_iterations < 0");
    if (c.precision != LDPC_GPU_PREC_F64 && c.precision != LDPC_GPU_PREC_F32 && c.precision != LDPC_GPU_PREC_F16X2)
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "unknown precision");
    if (c.channel_mode != LDPC_GPU_CHANNEL_EXACT && c.channel_mode != LDPC_GPU_CHANNEL_FAST) return set_err(LDPC_GPU_ERR_INVALID_ARG, "unknown channel_mode");
    if (c.channel_mode == LDPC_GPU_CHANNEL_FAST && c.kind != LDPC_GPU_KIND_MINSUM && c.kind != LDPC_GPU_KIND_BP && c.kind != LDPC_GPU_KIND_DDBMP)
        return set_err(LDPC_GPU_ERR_UNSUPPORTED, "LDPC_GPU_CHANNEL_FAST exists for the message-passing decoders (min-sum family, sum-product, DD-BMP)");
    if (c.precision == LDPC_GPU_PREC_F16X2 && c.kind != LDPC_GPU_KIND_MINSUM)
        return set_err(LDPC_GPU_ERR_UNSUPPORTED, "LDPC_GPU_PREC_F16X2 exists for the min-sum family only");
    if (c.kind == LDPC_GPU_KIND_GDBF) {
        if ((c.flags & LDPC_GPU_F_REDECODE) && (c.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_QUANTIZEPROBABILITIES)))
            return set_err(LDPC_GPU_ERR_UNSUPPORTED, "RNGDBF.cpp has no quantizeSamples / quantizeProbabilities code");
        if ((c.flags & LDPC_GPU_F_REDECODE) && (c.maxphase < 1 || c.maxphase > 15)) return set_err(LDPC_GPU_ERR_INVALID_ARG, "maxphase must be 1..15");
        if (c.num_iterations < 1) return set_err(LDPC_GPU_ERR_INVALID_ARG, "GDBF needs T >= 1 (the reference reads `satisfied` uninitialised at T = 0)");
    }
    if (c.kind == LDPC_GPU_KIND_NGDBF_HW) {
        if (c.num_iterations < 1 || c.w <= 0 || c.Ymax <= 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "NGDBFhw needs T >= 1, w > 0, Ymax > 0");
        if (c.maxphase > 15) return set_err(LDPC_GPU_ERR_INVALID_ARG, "maxphase must be <= 15");
    }
    if (c.kind == LDPC_GPU_KIND_NGDBF_SC) {
        if (c.num_iterations < 1 || c.Q < 1 || c.Q > 8 || !(c.Ymax > 0) || !(c.lambda > 0) || c.windowsize < 0)
            return set_err(LDPC_GPU_ERR_INVALID_ARG, "NGDBF_SC needs T >= 1, 1 <= Q <= 8, Ymax > 0, lambda > 0, windowsize >= 0");
        if (c.precision != LDPC_GPU_PREC_F64) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "NGDBF_SC exists in the double arithmetic of the SystemC model only (LDPC_GPU_PREC_F64)");
    }
    if ((c.kind == LDPC_GPU_KIND_MINSUM && (c.flags & (LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES))) || c.kind == LDPC_GPU_KIND_DDBMP)
        if (!(c.Ymax > 0) || (((c.flags & LDPC_GPU_F_QUANTIZE_SAMPLES) || c.kind == LDPC_GPU_KIND_DDBMP) && (c.Q < 1 || c.Q > 30)))
            return set_err(LDPC_GPU_ERR_INVALID_ARG, "quantiser needs Ymax > 0 and 1 <= Q <= 30");
    if (c.kind == LDPC_GPU_KIND_MINSUM && (c.flags & LDPC_GPU_F_NORMALIZED_MS) && !(c.alpha > 0)) return set_err(LDPC_GPU_ERR_INVALID_ARG, "normalised min-sum needs alpha > 0");
    if (c.kind == LDPC_GPU_KIND_MINSUM && (c.flags & LDPC_GPU_F_OFFSET_MS) && !(c.delta >= 0))
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "offset min-sum needs delta >= 0 (the kernels implement sgn(m) max(|m| - delta, 0), src/decodeMinSum.cpp:503-515, in forms that assume it)");
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_decoder_destroy(ldpc_gpu_decoder *d)
{
    if (!d) return LDPC_GPU_OK;
    cudaSetDevice(d->device);
    for (void *p : d->owned) cudaFree(p);
    for (Slot &s : d->slot) {
        for (DevBuf *b : { &s.y, &s.noise, &s.cw, &s.qp, &s.bits, &s.iters, &s.soft, &s.errs, &s.flags }) b->release();
        if (s.k0) cudaEventDestroy(s.k0); if (s.k1) cudaEventDestroy(s.k1);
        if (s.st) cudaStreamDestroy(s.st);
    }
    if (d->redo) ldpc_gpu_decoder_destroy(d->redo);
    if (d->d_redo_list) cudaFree(d->d_redo_list);
    if (d->d_redo_count) cudaFree(d->d_redo_count);
    if (d->d_redo_total) cudaFree(d->d_redo_total);
    if (d->h_redo) cudaFreeHost(d->h_redo);
    if (d->d_counters) cudaFree(d->d_counters);
    if (d->d_ws) cudaFree(d->d_ws);
    if (d->d_cwtab) cudaFree(d->d_cwtab);
    delete d;
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_decoder_create(const ldpc_gpu_code *code, const ldpc_gpu_decoder_cfg *cfg, int device, ldpc_gpu_decoder **out)
{
    if (!out) return set_err(LDPC_GPU_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!code || !cfg) return set_err(LDPC_GPU_ERR_INVALID_ARG, "code or cfg is NULL");
    int rc = validate_cfg(*cfg); if (rc) return rc;
    if (ldpc_gpu_device_count() <= 0) return set_err(LDPC_GPU_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CU_TRY(cudaSetDevice(device));
    ldpc_gpu_decoder *d = new ldpc_gpu_decoder;
    d->device = device; d->cfg = *cfg; d->N = code->N; d->M = code->M;
    d->h_mlist = code->mlist; d->h_rowdeg = code->row_deg; d->h_dcm = code->dc_max;
    cudaDeviceGetAttribute(&d->n_sm, cudaDevAttrMultiProcessorCount, device);
    if ((rc = build_device_code(d, code)) ||
        ((cfg->kind == LDPC_GPU_KIND_MINSUM || (cfg->kind == LDPC_GPU_KIND_BP && cfg->precision != LDPC_GPU_PREC_F64)) &&
         (rc = build_schedule(d, code, cfg->precision == LDPC_GPU_PREC_F64 ? 8 : 4))) ||
        (cfg->kind == LDPC_GPU_KIND_MINSUM && cfg->precision == LDPC_GPU_PREC_F32 && (rc = build_quad_tables(d, code))) ||
        (rc = pick_kernel(d))) { ldpc_gpu_decoder_destroy(d); return rc; }
    if (d->x2) {
        ldpc_gpu_decoder_cfg c64 = *cfg; c64.precision = LDPC_GPU_PREC_F64; c64.flags &= ~(uint32_t)LDPC_GPU_F_CERT_STOP;
        if ((rc = ldpc_gpu_decoder_create(code, &c64, device, &d->redo))) { ldpc_gpu_decoder_destroy(d); return rc; }
        if (cudaMalloc(&d->d_redo_count, 2 * sizeof(unsigned int)) != cudaSuccess || cudaMalloc(&d->d_redo_total, sizeof(unsigned long long)) != cudaSuccess ||
            cudaMemset(d->d_redo_total, 0, sizeof(unsigned long long)) != cudaSuccess ||
            cudaHostAlloc((void **)&d->h_redo, 2 * sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess) { ldpc_gpu_decoder_destroy(d); return set_err(LDPC_GPU_ERR_NOMEM, "redo-list allocation failed"); }
    }
    for (Slot &s : d->slot) {
        if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&s.k0) != cudaSuccess ||
            cudaEventCreate(&s.k1) != cudaSuccess) { ldpc_gpu_decoder_destroy(d); return set_err(LDPC_GPU_ERR_CUDA, "stream/event creation failed"); }
    }
    const size_t nhist = (size_t)d->N + iter_hist_len(*cfg) + 16;
    if (cudaMalloc(&d->d_counters, sizeof(unsigned long long) * (CNT_N + nhist)) != cudaSuccess) { ldpc_gpu_decoder_destroy(d); return set_err(LDPC_GPU_ERR_NOMEM, "counter allocation failed"); }
    d->d_ew = d->d_counters + CNT_N; d->d_it = d->d_ew + d->N; d->d_ph = d->d_it + iter_hist_len(*cfg);
    // configuration-only part of DecParams
    DecParams &p = d->base; memset(&p, 0, sizeof p);
    p.kind = cfg->kind; p.flags = cfg->flags; p.T = cfg->num_iterations; p.Q = cfg->Q; p.NQ = cfg->NQ;
    p.windowsize = cfg->windowsize; p.maxphase = cfg->maxphase; p.Tswitch = cfg->Tswitch;
    p.Ymax = cfg->Ymax; p.alpha = cfg->alpha; p.delta = cfg->delta; p.theta = cfg->theta; p.lambda = cfg->lambda;
    p.noiseScale = cfg->noiseScale; p.w = cfg->w; p.theta0 = cfg->theta0; p.MAXLLR = cfg->MAXLLR;
    const double Nq = pow(2.0, cfg->Q);                            // decodeMinSum.cpp:125
    p.ms_Nq1 = Nq - 1; p.ms_twoY = 2.0 * cfg->Ymax; p.ms_step = 2 * cfg->Ymax / (Nq - 1);
    { int ex; p.ms_inv_twoY = (p.ms_twoY > 0 && frexp(p.ms_twoY, &ex) == 0.5) ? 1.0 / p.ms_twoY : 0.0; }
    const double gq = pow(2, (cfg->NQ - 1)), gl = cfg->Ymax / 2.0; // decodeGDBF.cpp:490-491
    p.g_qmax = gq; p.g_twol = 2 * gl; p.g_step = 2.0 * gl / gq;
    if (cfg->kind == LDPC_GPU_KIND_NGDBF_HW) {                    // NGDBFhw.cpp:171-176
        const double qmax = pow(2, 5), lmax = cfg->Ymax / (2.0 * cfg->w), NL = qmax - 1;
        p.hw_lmax = lmax; p.hw_NL = NL; p.hw_two_lmax = 2 * lmax; p.hw_two_w = 2.0 * cfg->w;
        const int q2 = (int)(1.0 * round(floor(fabs(2.0) * NL / (2 * lmax))));      // quantize(2)
        const int code = q2 & 31;                                                    // pack(q2, +1)
        p.hw_theta = ((code << 1) & 31) | 1;                                         // unpack()
        if (code & 16) p.hw_theta = -p.hw_theta;
        p.hw_Smult = (int)round(NL / lmax);
    }
    p.inv_alpha_f = cfg->alpha != 0 ? (float)(1.0 / (double)(float)cfg->alpha) : 0.0f;
    { int ex; const bool pow2 = cfg->alpha > 0 && frexp(cfg->alpha, &ex) == 0.5;
      p.alpha_div_f = ((cfg->flags & LDPC_GPU_F_NORMALIZED_MS) && !pow2 && !getenv("LDPC_GPU_NO_FDIV")) ? (float)cfg->alpha : 0.0f; }
    { int ex; p.ms_step_dyadic = (p.ms_step > 0 && frexp(p.ms_step, &ex) == 0.5 && cfg->Q <= 16) ? 1 : 0; }
    p.ms_scale_f = (float)((Nq - 1) / (2.0 * cfg->Ymax)); p.ms_step_f = (float)p.ms_step; p.Ymax_f = (float)cfg->Ymax;
    p.iter_hist_len = iter_hist_len(*cfg);
    p.rows_per_step = rows_per_step(cfg->flags);
    p.channel_mode = cfg->channel_mode;
    {
        auto h2 = [](float x) { const __half h = __float2half_rn(x); unsigned short b; memcpy(&b, &h, 2); return (uint32_t)b | ((uint32_t)b << 16); };
        p.x2_delta2 = h2((cfg->flags & LDPC_GPU_F_OFFSET_MS) ? (float)cfg->delta : 0.0f);
        p.x2_cap2 = h2(d->x2_cap); p.x2_m02 = h2(d->x2_m0);
    }
    *out = d;
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_decoder_set_codewords(ldpc_gpu_decoder *d, const uint8_t *bits01, int64_t n)
{
    if (!d) return set_err(LDPC_GPU_ERR_INVALID_ARG, "decoder is NULL");
    CU_TRY(cudaSetDevice(d->device));
    if (d->d_cwtab) { cudaFree(d->d_cwtab); d->d_cwtab = nullptr; }
    d->n_cw = 0;
    if (n <= 0 || !bits01) return LDPC_GPU_OK;
    CU_TRY(cudaMalloc(&d->d_cwtab, (size_t)n * d->N));
    CU_TRY(cudaMemcpy(d->d_cwtab, bits01, (size_t)n * d->N, cudaMemcpyHostToDevice));
    d->n_cw = n;
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_decoder_geometry(const ldpc_gpu_decoder *d, int *grid, int *block, int *smem, int *fpc)
{
    if (!d) return set_err(LDPC_GPU_ERR_INVALID_ARG, "decoder is NULL");
    if (grid) *grid = d->grid_full; if (block) *block = d->block; if (smem) *smem = d->smem; if (fpc) *fpc = d->ctas_per_sm * d->frames_per_cta;
    return LDPC_GPU_OK;
}
extern "C" int ldpc_gpu_last_timing(const ldpc_gpu_decoder *d, double *ms, int64_t *launches)
{
    if (!d) return set_err(LDPC_GPU_ERR_INVALID_ARG, "decoder is NULL");
    if (ms) *ms = d->last_kernel_ms; if (launches) *launches = d->last_launches;
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_decoder_stats(const ldpc_gpu_decoder *d, int64_t *redo_frames, int32_t *exact_lattice)
{
    if (!d) return set_err(LDPC_GPU_ERR_INVALID_ARG, "decoder is NULL");
    if (exact_lattice) *exact_lattice = d->x2 ? 1 : 0;
    if (redo_frames) {
        *redo_frames = 0;
        if (d->x2) {
            unsigned long long h = 0;
            CU_TRY(cudaSetDevice(d->device));
            CU_TRY(cudaDeviceSynchronize());
            CU_TRY(cudaMemcpy(&h, d->d_redo_total, sizeof h, cudaMemcpyDeviceToHost));
            *redo_frames = (int64_t)h;
        }
    }
    return LDPC_GPU_OK;
}

// per-call channel constants, with the reference's expressions (decodeMinSum.cpp:146-147)
static int channel_params(const ldpc_gpu_decoder *d, const ldpc_gpu_channel *ch, DecParams &p)
{
    p = d->base;
    if (!ch || !(ch->R > 0)) return set_err(LDPC_GPU_ERR_INVALID_ARG, "channel needs R > 0");
    p.N0 = pow(10.0, -ch->snr_db / 10.0) / ch->R;
    p.sigma = sqrt(p.N0 / 2.0);
    p.noiseSigma = p.sigma * p.noiseScale;                        // decodeGDBF.cpp:296
    p.uni_scale = sqrt(3) * p.noiseSigma * 2.0;                   // decodeGDBF.cpp:322
    p.sigma_f = (float)p.sigma;
    return LDPC_GPU_OK;
}

static int zero_counters(ldpc_gpu_decoder *d, cudaStream_t st)
{
    const size_t n = CNT_N + (size_t)d->N + d->base.iter_hist_len + 16;
    CU_TRY(cudaMemsetAsync(d->d_counters, 0, sizeof(unsigned long long) * n, st));
    return LDPC_GPU_OK;
}

static int fetch_counters(ldpc_gpu_decoder *d, ldpc_gpu_counters *out, cudaStream_t st)
{
    const size_t n = CNT_N + (size_t)d->N + d->base.iter_hist_len + 16;
    std::vector<unsigned long long> h(n);
    CU_TRY(cudaMemcpyAsync(h.data(), d->d_counters, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    out->errors += (int64_t)h[CNT_ERRORS]; out->uncodedErrors += (int64_t)h[CNT_UNCODED]; out->totalBits += (int64_t)h[CNT_BITS];
    out->totalWords += (int64_t)h[CNT_WORDS]; out->wordErrors += (int64_t)h[CNT_WORDERRS]; out->totalIterations += (int64_t)h[CNT_ITERS];
    out->smoothingUsed += (int64_t)h[CNT_SMOOTH]; out->undetectedWords += (int64_t)h[CNT_UNDETECTED];
    const unsigned long long *ew = h.data() + CNT_N, *it = ew + d->N, *ph = it + d->base.iter_hist_len;
    if (out->error_weight_hist) for (int i = 0; i < d->N; i++) out->error_weight_hist[i] += (int64_t)ew[i];
    if (out->iter_hist) for (int i = 0; i < d->base.iter_hist_len; i++) out->iter_hist[i] += (int64_t)it[i];
    if (out->phase_hist && (d->cfg.flags & LDPC_GPU_F_REDECODE)) for (int i = 0; i < std::min(15, std::max(1, d->cfg.maxphase)); i++) out->phase_hist[i] += (int64_t)ph[i];
    return LDPC_GPU_OK;
}

static int launch(ldpc_gpu_decoder *d, const DecParams &p, const FrameIO &io, cudaStream_t st)
{
    const long long want = std::min<long long>((io.n_frames + d->frames_per_cta - 1) / d->frames_per_cta, d->grid_full);
    if (want <= 0) return LDPC_GPU_OK;
    FrameIO io2 = io; io2.ws_stride = d->ws_stride;
    io2.workspace = d->d_ws ? d->d_ws + ((st == d->slot[1].st) ? (size_t)d->grid_full * d->ws_stride : 0) : nullptr;
    if (d->x2) {
        if ((size_t)io.n_frames > d->redo_cap) {          // grows only (streams of this decoder may still read the old list: drain them)
            for (Slot &s : d->slot) if (s.st) CU_TRY(cudaStreamSynchronize(s.st));
            if (d->d_redo_list) cudaFree(d->d_redo_list);
            d->d_redo_list = nullptr; d->redo_cap = 0;
            const size_t cap = std::max<size_t>((size_t)io.n_frames, 1 << 16);
            CU_TRY(cudaMalloc(&d->d_redo_list, 2 * cap * sizeof(long long)));      // one half per pipeline slot
            d->redo_cap = cap;
        }
        const int half = (st == d->slot[1].st) ? 1 : 0;
        io2.redo_list = d->d_redo_list + (size_t)half * d->redo_cap;
        io2.redo_count = d->d_redo_count + half;
        io2.redo_total = d->d_redo_total;
        CU_TRY(cudaMemsetAsync(io2.redo_count, 0, sizeof(unsigned int), st));
    }
    KernelFn fn = d->fn;
    if (d->fn_staged && io2.y && io2.y_dtype == LDPC_GPU_DT_QP && ((size_t)io2.y & 15) == 0 && ((((size_t)d->N * d->cfg.Q) >> 3) & 15) == 0 &&
        (((size_t)d->N * d->cfg.Q) >> 3) <= (size_t)d->N) fn = d->fn_staged;
    fn<<<(unsigned)want, d->block, d->smem, st>>>(d->dev, p, io2);
    CU_TRY(cudaGetLastError());
    d->last_launches++;
    if (d->x2) {                                          // how many frames the packed kernel could not certify: read by redo_after_sync
        const int half = (st == d->slot[1].st) ? 1 : 0;
        CU_TRY(cudaMemcpyAsync(&d->h_redo[half], d->d_redo_count + half, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    }
    return LDPC_GPU_OK;
}

// launch of the fp64 redo decoder over n frames (its own workspace when its state lives in HBM)
static int launch_redo(ldpc_gpu_decoder *d, const DecParams &p, FrameIO io, long long n, cudaStream_t st)
{
    ldpc_gpu_decoder *r = d->redo;
    io.workspace = r->d_ws; io.ws_stride = r->ws_stride;
    const long long want = std::min<long long>((n + r->frames_per_cta - 1) / r->frames_per_cta, r->grid_full);
    r->fn<<<(unsigned)want, r->block, r->smem, st>>>(r->dev, p, io);
    CU_TRY(cudaGetLastError());
    d->last_launches++;
    return LDPC_GPU_OK;
}

// Exact-lattice kernel: after the stream of a launch has been synchronised, decode the frames it could not certify (about 5 in
// 10^6 at the operating point) with the fp64 parity instantiation, same FrameIO, same stream.  *ran tells the caller whether
// outputs changed.  Nothing is launched in the common case.
static int redo_after_sync(ldpc_gpu_decoder *d, const DecParams &p, const FrameIO &io, cudaStream_t st, bool *ran)
{
    if (ran) *ran = false;
    if (!d->x2) return LDPC_GPU_OK;
    const int half = (st == d->slot[1].st) ? 1 : 0;
    const unsigned int n = d->h_redo[half];
    if (n == 0) return LDPC_GPU_OK;
    FrameIO io3 = io; io3.workspace = nullptr; io3.ws_stride = 0;
    io3.frame_list = d->d_redo_list + (size_t)half * d->redo_cap; io3.n_frames_dev = d->d_redo_count + half;
    { const int rc = launch_redo(d, p, io3, (long long)n, st); if (rc) return rc; }
    CU_TRY(cudaStreamSynchronize(st));
    if (ran) *ran = true;
    return LDPC_GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// parity entry
// ------------------------------------------------------------------------------------------------
extern "C" int ldpc_gpu_decode_batch(ldpc_gpu_decoder *d, const ldpc_gpu_channel *ch, const ldpc_gpu_batch *b, ldpc_gpu_counters *cnt)
{
    if (!d || !b) return set_err(LDPC_GPU_ERR_INVALID_ARG, "decoder or batch is NULL");
    if (b->n_frames < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "n_frames < 0");
    if (b->n_frames > 0 && !b->y) return set_err(LDPC_GPU_ERR_INVALID_ARG, "batch.y is NULL");
    if (b->y_dtype != LDPC_GPU_DT_F64 && b->y_dtype != LDPC_GPU_DT_F32 && b->y_dtype != LDPC_GPU_DT_F16 && b->y_dtype != LDPC_GPU_DT_Q8 && b->y_dtype != LDPC_GPU_DT_QP)
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "unknown y_dtype");
    const int N = d->N, kind = d->cfg.kind;
    if ((d->cfg.flags & LDPC_GPU_F_CERT_STOP) && b->out_soft)
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "LDPC_GPU_F_CERT_STOP reports decisions, not the a-posteriori sums of iteration T: out_soft must be NULL");
    if (b->y_dtype == LDPC_GPU_DT_Q8 && !(kind == LDPC_GPU_KIND_MINSUM && (d->cfg.flags & LDPC_GPU_F_QUANTIZE_SAMPLES) && d->cfg.Q >= 2 && d->cfg.Q <= 6))
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "LDPC_GPU_DT_Q8 carries quantiser levels: it needs a min-sum decoder with LDPC_GPU_F_QUANTIZE_SAMPLES and 2 <= Q <= 6");
    if (b->y_dtype == LDPC_GPU_DT_QP && !(kind == LDPC_GPU_KIND_MINSUM && (d->cfg.flags & LDPC_GPU_F_QUANTIZE_SAMPLES) && d->cfg.Q >= 2 && d->cfg.Q <= 8 &&
                                          ((long long)N * d->cfg.Q) % 32 == 0 && ((size_t)b->y & 3) == 0))
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "LDPC_GPU_DT_QP carries bit-packed quantiser levels: it needs a min-sum decoder with LDPC_GPU_F_QUANTIZE_SAMPLES, 2 <= Q <= 8, N*Q a multiple of 32 and a 4-byte aligned buffer");
    const int rps = rows_per_step(d->cfg.flags);
    if (kind == LDPC_GPU_KIND_GDBF && rps > 0) {
        const int ph = (d->cfg.flags & LDPC_GPU_F_REDECODE) ? std::max(1, d->cfg.maxphase) : 1;
        if (!b->noise) return set_err(LDPC_GPU_ERR_INVALID_ARG, "this GDBF variant draws random numbers: batch.noise is required (use ldpc_gpu_simulate for on-device noise)");
        if (b->noise_rows < (int64_t)d->cfg.num_iterations * ph * rps) return set_err(LDPC_GPU_ERR_INVALID_ARG, "batch.noise_rows is smaller than T * phases * rows-per-step");
    }
    if (kind == LDPC_GPU_KIND_NGDBF_HW && !b->noise) return set_err(LDPC_GPU_ERR_INVALID_ARG, "NGDBFhw needs its per-frame noise buffer in batch.noise");
    if (kind == LDPC_GPU_KIND_NGDBF_SC && (!b->noise || b->noise_rows < (int64_t)N + d->cfg.num_iterations + 1))
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "NGDBF_SC needs batch.noise = [F][noise_rows] normals with noise_rows >= N + T + 1 (the node-to-node noise chain)");
    if (kind == LDPC_GPU_KIND_NGDBF_HW && b->qpointer0 && b->mem == LDPC_GPU_MEM_HOST)
        for (int64_t f = 0; f < b->n_frames; f++)
            if (b->qpointer0[f] < 0 || b->qpointer0[f] >= LDPC_GPU_HW_QBUF - N)
                return set_err(LDPC_GPU_ERR_INVALID_ARG, "batch.qpointer0 outside [0, LDPC_GPU_HW_QBUF - N): the noise window wraps there (src/NGDBFhw.cpp:356-358)");
    DecParams p; int rc = channel_params(d, ch, p); if (rc) return rc;
    if (b->y_dtype == LDPC_GPU_DT_Q8 || b->y_dtype == LDPC_GPU_DT_QP) p.flags &= ~(uint32_t)(LDPC_GPU_F_QUANTIZE_SAMPLES | LDPC_GPU_F_SATURATE_SAMPLES);   // the levels ARE the quantiser's output
    CU_TRY(cudaSetDevice(d->device));
    d->last_kernel_ms = 0; d->last_launches = 0;
    cudaStream_t st0 = d->slot[0].st;
    if (cnt) { if ((rc = zero_counters(d, st0))) return rc; CU_TRY(cudaStreamSynchronize(st0)); }

    const size_t esz = b->y_dtype == LDPC_GPU_DT_F64 ? 8 : b->y_dtype == LDPC_GPU_DT_F32 ? 4 : b->y_dtype == LDPC_GPU_DT_F16 ? 2 : 1, bpf = (size_t)(N + 7) / 8;
    const size_t ybytes = b->y_dtype == LDPC_GPU_DT_QP ? ((size_t)N * d->cfg.Q) / 8 : esz * N;   // sample bytes per frame
    const size_t ssz = b->y_dtype == LDPC_GPU_DT_F64 ? 8 : 4;     // out_soft element size
    const size_t noise_pf = kind == LDPC_GPU_KIND_NGDBF_HW ? (size_t)LDPC_GPU_HW_QBUF : kind == LDPC_GPU_KIND_NGDBF_SC ? (size_t)b->noise_rows
                          : (kind == LDPC_GPU_KIND_GDBF && b->noise ? (size_t)b->noise_rows * N : 0);
    FrameIO io; memset(&io, 0, sizeof io);
    io.y_dtype = b->y_dtype; io.noise_rows = b->noise_rows;
    if (cnt) { io.counters = d->d_counters; io.ew_hist = d->d_ew; io.it_hist = d->d_it; io.ph_hist = d->d_ph; }

    if (b->mem == LDPC_GPU_MEM_DEVICE) {
        io.n_frames = b->n_frames; io.y = b->y; io.noise = b->noise; io.codeword = b->codeword; io.qpointer0 = b->qpointer0;
        io.out_bits = b->out_bits; io.out_iters = b->out_iters; io.out_soft = b->out_soft; io.out_errors = b->out_errors; io.out_flags = b->out_flags;
        CU_TRY(cudaEventRecord(d->slot[0].k0, st0));
        if ((rc = launch(d, p, io, st0))) return rc;
        CU_TRY(cudaEventRecord(d->slot[0].k1, st0));
        CU_TRY(cudaStreamSynchronize(st0));
        float ms = 0; cudaEventElapsedTime(&ms, d->slot[0].k0, d->slot[0].k1); d->last_kernel_ms = ms;
        if ((rc = redo_after_sync(d, p, io, st0, nullptr))) return rc;
    } else if (b->mem == LDPC_GPU_MEM_HOST) {
        // two-slot pipeline: H2D / kernel / D2H of consecutive chunks overlap on two streams
        const size_t per_frame = ybytes + (b->out_soft ? ssz * N : 0) + 8 * noise_pf + (b->codeword ? N : 0) + bpf + 16;
        // chunks small enough that the first H2D and the last D2H (the parts no kernel hides) are a few percent
        // of the batch, large enough to give every CTA tens of frames
        const char *cb = getenv("LDPC_GPU_CHUNK_MB");
        const size_t chunk_bytes = (size_t)(cb ? atoi(cb) : 16) << 20;
        long long chunk = (long long)std::max<size_t>(1, chunk_bytes / per_frame);
        chunk = std::max<long long>(chunk, 16ll * d->grid_full * d->frames_per_cta);
        chunk = std::min<long long>(chunk, std::max<long long>(1, (b->n_frames + 1) / 2));
        if (b->n_frames <= 2 * (long long)d->grid_full) chunk = std::max<long long>(1, b->n_frames);
        auto copy_out = [&](Slot &s) -> int {             // results of the slot's chunk, device -> caller
            const long long f0 = s.f0, nf = s.nf;
            if (b->out_bits)   CU_TRY(cudaMemcpyAsync(b->out_bits + (size_t)f0 * bpf, s.bits.p, bpf * nf, cudaMemcpyDeviceToHost, s.st));
            if (b->out_iters)  CU_TRY(cudaMemcpyAsync(b->out_iters + f0, s.iters.p, 4 * (size_t)nf, cudaMemcpyDeviceToHost, s.st));
            if (b->out_soft)   CU_TRY(cudaMemcpyAsync((char *)b->out_soft + (size_t)f0 * N * ssz, s.soft.p, ssz * N * nf, cudaMemcpyDeviceToHost, s.st));
            if (b->out_errors) CU_TRY(cudaMemcpyAsync(b->out_errors + f0, s.errs.p, 4 * (size_t)nf, cudaMemcpyDeviceToHost, s.st));
            if (b->out_flags)  CU_TRY(cudaMemcpyAsync(b->out_flags + f0, s.flags.p, (size_t)nf, cudaMemcpyDeviceToHost, s.st));
            return LDPC_GPU_OK;
        };
        // Frames the exact-lattice kernel could not certify (about 5 in 10^6) are collected and decoded by the fp64 instantiation
        // AFTER the last chunk: a redo launch in the middle of the pipeline needs a whole SM's shared memory, so it waits for the
        // other slot's kernel to drain while the host, blocked on it, cannot queue the next chunk (3.5 % of the batch time, measured).
        std::vector<long long> pending;
        auto drain = [&](Slot &s) -> int {                // wait for the slot's chunk; note what the exact-lattice kernel left open
            CU_TRY(cudaStreamSynchronize(s.st));
            float ms = 0; cudaEventElapsedTime(&ms, s.k0, s.k1); d->last_kernel_ms += ms;
            if (d->x2) {
                const int half = (s.st == d->slot[1].st) ? 1 : 0;
                const unsigned int n = d->h_redo[half];
                if (n) {
                    std::vector<long long> ids(n);
                    CU_TRY(cudaMemcpy(ids.data(), d->d_redo_list + (size_t)half * d->redo_cap, n * sizeof(long long), cudaMemcpyDeviceToHost));
                    for (long long q : ids) pending.push_back(s.f0 + q);
                }
            }
            return LDPC_GPU_OK;
        };
        auto redo_pending = [&]() -> int {                // the collected frames, compacted, through the fp64 decoder; results scattered back
            if (pending.empty()) return LDPC_GPU_OK;
            Slot &s = d->slot[0];
            const size_t n = pending.size();
            int r2;
            if ((r2 = s.y.reserve(ybytes * n))) return r2;
            if (b->codeword && (r2 = s.cw.reserve((size_t)N * n))) return r2;
            for (size_t k = 0; k < n; k++) {
                CU_TRY(cudaMemcpyAsync((char *)s.y.p + k * ybytes, (const char *)b->y + (size_t)pending[k] * ybytes, ybytes, cudaMemcpyHostToDevice, s.st));
                if (b->codeword) CU_TRY(cudaMemcpyAsync((char *)s.cw.p + k * N, b->codeword + (size_t)pending[k] * N, (size_t)N, cudaMemcpyHostToDevice, s.st));
            }
            FrameIO ior = io;
            ior.y = s.y.p; ior.n_frames = (long long)n; ior.frame_begin = 0; ior.codeword = b->codeword ? (const uint8_t *)s.cw.p : nullptr;
            ior.noise = nullptr; ior.qpointer0 = nullptr; ior.workspace = nullptr; ior.ws_stride = 0;
            ior.frame_list = nullptr; ior.n_frames_dev = nullptr; ior.redo_list = nullptr; ior.redo_count = nullptr; ior.redo_total = nullptr;
            ior.out_bits = nullptr; ior.out_iters = nullptr; ior.out_soft = nullptr; ior.out_errors = nullptr; ior.out_flags = nullptr;
            if (b->out_bits)   { if ((r2 = s.bits.reserve(bpf * n))) return r2;  ior.out_bits = (uint8_t *)s.bits.p; }
            if (b->out_iters)  { if ((r2 = s.iters.reserve(4 * n))) return r2; ior.out_iters = (int *)s.iters.p; }
            if (b->out_soft)   { if ((r2 = s.soft.reserve(ssz * N * n))) return r2; ior.out_soft = s.soft.p; }
            if (b->out_errors) { if ((r2 = s.errs.reserve(4 * n))) return r2; ior.out_errors = (int *)s.errs.p; }
            if (b->out_flags)  { if ((r2 = s.flags.reserve(n))) return r2; ior.out_flags = (uint8_t *)s.flags.p; }
            if ((r2 = launch_redo(d, p, ior, (long long)n, s.st))) return r2;
            for (size_t k = 0; k < n; k++) {
                const size_t f = (size_t)pending[k];
                if (b->out_bits)   CU_TRY(cudaMemcpyAsync(b->out_bits + f * bpf, (char *)s.bits.p + k * bpf, bpf, cudaMemcpyDeviceToHost, s.st));
                if (b->out_iters)  CU_TRY(cudaMemcpyAsync(b->out_iters + f, (int *)s.iters.p + k, 4, cudaMemcpyDeviceToHost, s.st));
                if (b->out_soft)   CU_TRY(cudaMemcpyAsync((char *)b->out_soft + f * N * ssz, (char *)s.soft.p + k * N * ssz, ssz * N, cudaMemcpyDeviceToHost, s.st));
                if (b->out_errors) CU_TRY(cudaMemcpyAsync(b->out_errors + f, (int *)s.errs.p + k, 4, cudaMemcpyDeviceToHost, s.st));
                if (b->out_flags)  CU_TRY(cudaMemcpyAsync(b->out_flags + f, (uint8_t *)s.flags.p + k, 1, cudaMemcpyDeviceToHost, s.st));
            }
            CU_TRY(cudaStreamSynchronize(s.st));
            return LDPC_GPU_OK;
        };
        // any failure inside the loop must not leave async copies in flight on the caller's buffers
        auto run_chunks = [&]() -> int {
        int c = 0;
        for (long long f0 = 0; f0 < b->n_frames; f0 += chunk, c++) {
            const long long nf = std::min<long long>(chunk, b->n_frames - f0);
            Slot &s = d->slot[c & 1];
            if (s.used && (rc = drain(s))) return rc;     // this slot's previous chunk
            s.used = true; s.f0 = f0; s.nf = nf;
            if ((rc = s.y.reserve(ybytes * nf))) return rc;
            CU_TRY(cudaMemcpyAsync(s.y.p, (const char *)b->y + (size_t)f0 * ybytes, ybytes * nf, cudaMemcpyHostToDevice, s.st));
            io.y = s.y.p; io.n_frames = nf; io.frame_begin = f0;
            io.noise = nullptr; io.codeword = nullptr; io.qpointer0 = nullptr;
            if (b->noise && noise_pf) {
                if ((rc = s.noise.reserve(8 * noise_pf * nf))) return rc;
                CU_TRY(cudaMemcpyAsync(s.noise.p, b->noise + (size_t)f0 * noise_pf, 8 * noise_pf * nf, cudaMemcpyHostToDevice, s.st));
                io.noise = (const double *)s.noise.p;
            }
            if (b->codeword) {
                if ((rc = s.cw.reserve((size_t)N * nf))) return rc;
                CU_TRY(cudaMemcpyAsync(s.cw.p, b->codeword + (size_t)f0 * N, (size_t)N * nf, cudaMemcpyHostToDevice, s.st));
                io.codeword = (const uint8_t *)s.cw.p;
            }
            if (b->qpointer0) {
                if ((rc = s.qp.reserve(4 * (size_t)nf))) return rc;
                CU_TRY(cudaMemcpyAsync(s.qp.p, b->qpointer0 + f0, 4 * (size_t)nf, cudaMemcpyHostToDevice, s.st));
                io.qpointer0 = (const int *)s.qp.p;
            }
            io.out_bits = nullptr; io.out_iters = nullptr; io.out_soft = nullptr; io.out_errors = nullptr; io.out_flags = nullptr;
            if (b->out_bits)   { if ((rc = s.bits.reserve(bpf * nf))) return rc;  io.out_bits = (uint8_t *)s.bits.p; }
            if (b->out_iters)  { if ((rc = s.iters.reserve(4 * (size_t)nf))) return rc; io.out_iters = (int *)s.iters.p; }
            if (b->out_soft)   { if ((rc = s.soft.reserve(ssz * N * nf))) return rc; io.out_soft = s.soft.p; }
            if (b->out_errors) { if ((rc = s.errs.reserve(4 * (size_t)nf))) return rc; io.out_errors = (int *)s.errs.p; }
            if (b->out_flags)  { if ((rc = s.flags.reserve((size_t)nf))) return rc; io.out_flags = (uint8_t *)s.flags.p; }
            CU_TRY(cudaEventRecord(s.k0, s.st));
            s.io = io;
            if ((rc = launch(d, p, io, s.st))) return rc;
            CU_TRY(cudaEventRecord(s.k1, s.st));
            if ((rc = copy_out(s))) return rc;
        }
        return LDPC_GPU_OK;
        };
        rc = run_chunks();
        if (rc) {                                         // drain both slots, keep the first error's message
            const std::string msg = g_err;
            for (Slot &s : d->slot) { cudaStreamSynchronize(s.st); s.used = false; }
            cudaGetLastError();
            return set_err(rc, msg);
        }
        for (Slot &s : d->slot) if (s.used) {
            s.used = false;
            if ((rc = drain(s))) { const std::string msg = g_err; for (Slot &t : d->slot) { cudaStreamSynchronize(t.st); t.used = false; } return set_err(rc, msg); }
        }
        if ((rc = redo_pending())) { const std::string msg = g_err; for (Slot &t : d->slot) cudaStreamSynchronize(t.st); cudaGetLastError(); return set_err(rc, msg); }
    } else return set_err(LDPC_GPU_ERR_INVALID_ARG, "unknown batch.mem");
    if (cnt) return fetch_counters(d, cnt, st0);
    return LDPC_GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// throughput entry
// ------------------------------------------------------------------------------------------------
extern "C" int ldpc_gpu_simulate(ldpc_gpu_decoder *d, const ldpc_gpu_channel *ch, const ldpc_gpu_sim_args *a, ldpc_gpu_counters *cnt)
{
    if (!d || !a || !cnt) return set_err(LDPC_GPU_ERR_INVALID_ARG, "NULL argument");
    if (a->n_frames < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "n_frames < 0");
    DecParams p; int rc = channel_params(d, ch, p); if (rc) return rc;
    CU_TRY(cudaSetDevice(d->device));
    cudaStream_t st = d->slot[0].st;
    d->last_kernel_ms = 0; d->last_launches = 0;
    if ((rc = zero_counters(d, st))) return rc;
    FrameIO io; memset(&io, 0, sizeof io);
    io.seed = a->seed; io.cw_table = d->d_cwtab; io.n_cw = d->n_cw;
    io.counters = d->d_counters; io.ew_hist = d->d_ew; io.it_hist = d->d_it; io.ph_hist = d->d_ph;
    const bool stop_rule = a->stop_errors > 0 || a->stop_word_errors > 0;
    long long per_launch = a->poll_frames > 0 ? a->poll_frames : (stop_rule ? (long long)d->grid_full * 4 : (1ll << 24));
    long long done = 0;
    ldpc_gpu_counters snap; memset(&snap, 0, sizeof snap);
    CU_TRY(cudaEventRecord(d->slot[0].k0, st));
    while (done < a->n_frames) {
        const long long nf = std::min<long long>(per_launch, a->n_frames - done);
        io.n_frames = nf; io.frame_begin = a->frame_begin + done;
        if ((rc = launch(d, p, io, st))) return rc;
        if (d->x2) { CU_TRY(cudaStreamSynchronize(st)); if ((rc = redo_after_sync(d, p, io, st, nullptr))) return rc; }
        done += nf;
        if (stop_rule && done < a->n_frames) {            // poll the reference's loop condition (decodeMinSum.cpp:189)
            unsigned long long h[CNT_N];
            CU_TRY(cudaMemcpyAsync(h, d->d_counters, sizeof h, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            const int64_t e = cnt->errors + (int64_t)h[CNT_ERRORS], we = cnt->wordErrors + (int64_t)h[CNT_WORDERRS];
            if (!(e < a->stop_errors || we < a->stop_word_errors)) break;
        }
    }
    CU_TRY(cudaEventRecord(d->slot[0].k1, st));
    rc = fetch_counters(d, cnt, st);
    float ms = 0; cudaEventElapsedTime(&ms, d->slot[0].k0, d->slot[0].k1); d->last_kernel_ms = ms;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// re-decode statistics (src/redecodeStatistics.cpp): every frame of the range is decoded
// n_redecodes times from the SAME channel samples, each time with fresh decoder noise, and the error weight of
// every outcome is recorded.  The channel is a function of (seed, frame id) and re-decode r draws the rows
// r * rows_per_decode ... of the frame's decoder-noise stream, so the matrix is reproducible entry by entry.
// ------------------------------------------------------------------------------------------------
extern "C" int ldpc_gpu_redecode_stats(ldpc_gpu_decoder *d, const ldpc_gpu_channel *ch, const ldpc_gpu_sim_args *a, int32_t n_redecodes,
                                       int32_t *outcomes, ldpc_gpu_counters *cnt)
{
    if (!d || !a || !outcomes) return set_err(LDPC_GPU_ERR_INVALID_ARG, "NULL argument");
    if (a->n_frames < 0 || n_redecodes < 1) return set_err(LDPC_GPU_ERR_INVALID_ARG, "n_frames < 0 or n_redecodes < 1");
    if (d->cfg.kind != LDPC_GPU_KIND_GDBF && d->cfg.kind != LDPC_GPU_KIND_NGDBF_HW && d->cfg.kind != LDPC_GPU_KIND_NGDBF_SC)
        return set_err(LDPC_GPU_ERR_UNSUPPORTED, "re-decode statistics need a decoder with its own noise (GDBF family, NGDBFhw, NGDBF_SC)");
    DecParams p; int rc = channel_params(d, ch, p); if (rc) return rc;
    CU_TRY(cudaSetDevice(d->device));
    cudaStream_t st = d->slot[0].st;
    d->last_kernel_ms = 0; d->last_launches = 0;
    if ((rc = zero_counters(d, st))) return rc;
    long long rows_per_decode = 1;                                    // NGDBFhw: one noise buffer per decode
    if (d->cfg.kind == LDPC_GPU_KIND_GDBF) {
        const uint32_t f = d->cfg.flags;
        const long long per_step = ((f & LDPC_GPU_F_ADDNOISE) ? 1 : 0) + ((f & LDPC_GPU_F_QUANTIZEPROBABILITIES) ? 1 : 0);
        const long long ph = ((f & LDPC_GPU_F_REDECODE) && d->cfg.maxphase > 1) ? d->cfg.maxphase : 1;
        rows_per_decode = std::max<long long>(1, per_step * (long long)d->cfg.num_iterations * ph);
    }
    const long long per_launch = std::min<long long>(std::max<long long>(a->n_frames, 1), 1ll << 20);
    int *d_err = nullptr;
    CU_TRY(cudaMalloc(&d_err, sizeof(int) * (size_t)per_launch));
    std::vector<int> h_err((size_t)per_launch);
    FrameIO io; memset(&io, 0, sizeof io);
    io.seed = a->seed; io.cw_table = d->d_cwtab; io.n_cw = d->n_cw;
    io.counters = d->d_counters; io.ew_hist = d->d_ew; io.it_hist = d->d_it; io.ph_hist = d->d_ph;
    io.out_errors = d_err;
    if (cudaEventRecord(d->slot[0].k0, st) != cudaSuccess) { cudaFree(d_err); return set_err(LDPC_GPU_ERR_CUDA, "cudaEventRecord failed"); }
    for (long long done = 0; done < a->n_frames && !rc; done += per_launch) {
        const long long nf = std::min<long long>(per_launch, a->n_frames - done);
        io.n_frames = nf; io.frame_begin = a->frame_begin + done;
        for (int r = 0; r < n_redecodes && !rc; r++) {
            io.noise_row_base = (long long)r * rows_per_decode;
            if ((rc = launch(d, p, io, st))) break;
            if (cudaMemcpyAsync(h_err.data(), d_err, sizeof(int) * (size_t)nf, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess) { rc = set_err(LDPC_GPU_ERR_CUDA, "copying the outcomes failed"); break; }
            for (long long f = 0; f < nf; f++) outcomes[(size_t)(done + f) * n_redecodes + r] = h_err[(size_t)f];
        }
    }
    cudaEventRecord(d->slot[0].k1, st);
    cudaFree(d_err);
    if (rc) return rc;
    if (cnt) rc = fetch_counters(d, cnt, st); else CU_TRY(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, d->slot[0].k0, d->slot[0].k1); d->last_kernel_ms = ms;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// replay of one seed-addressed frame with a per-iteration trace (second half; src/replayGDBF.cpp:306-373
// writes, per executed flip step, the decisions after the step and the syndromes the step started from; NGDBFhw's
// LOG_PROCESSING dump, src/NGDBFhw.cpp:304-335, carries the same two vectors).  The frame is a function of (seed, frame id), so
// the trace is produced kernel-agnostically: the frame is decoded with T = 0, 1, 2, ... iterations (same channel samples, same
// decoder-noise rows) and row t pairs the decisions after t + 1 iterations with the syndrome of the decisions after t.
// A debugging tool: O(T) single-frame launches.
// ------------------------------------------------------------------------------------------------
extern "C" int ldpc_gpu_replay_frame(ldpc_gpu_decoder *d, const ldpc_gpu_channel *ch, uint64_t seed, int64_t frame_id, int32_t max_rows,
                                     uint8_t *trace_d, uint8_t *trace_syn, int32_t *n_rows, int32_t *final_errors)
{
    if (!d || !trace_d || !n_rows || max_rows < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "bad argument");
    if (d->cfg.kind == LDPC_GPU_KIND_NGDBF_SC)
        return set_err(LDPC_GPU_ERR_UNSUPPORTED, "replay: the SystemC-model decoder's smoothing window is tied to T, a run with fewer iterations is not a prefix of the full run");
    DecParams p; int rc = channel_params(d, ch, p); if (rc) return rc;
    CU_TRY(cudaSetDevice(d->device));
    const int N = d->N, M = d->M, T = d->cfg.num_iterations;
    const size_t bpf = (size_t)(N + 7) / 8, spf = (size_t)(M + 7) / 8;
    // the trace is the raw decision trajectory: output smoothing only post-processes it (decodeGDBF.cpp:358-367), re-decoding
    // phases restart it (RNGDBF.cpp:280): phase 1 is traced
    p.flags &= ~(uint32_t)LDPC_GPU_F_OUTPUTSMOOTHING;
    if (p.maxphase > 1) p.maxphase = 1;
    cudaStream_t st = d->slot[0].st;
    DevBuf bits, iters, errs;
    if ((rc = bits.reserve(bpf)) || (rc = iters.reserve(4)) || (rc = errs.reserve(4))) { bits.release(); iters.release(); errs.release(); return rc; }
    std::vector<uint8_t> prev(bpf), cur(bpf);
    FrameIO io; memset(&io, 0, sizeof io);
    io.n_frames = 1; io.frame_begin = frame_id; io.seed = seed; io.cw_table = d->d_cwtab; io.n_cw = d->n_cw;
    io.out_bits = (uint8_t *)bits.p; io.out_iters = (int *)iters.p; io.out_errors = (int *)errs.p;
    auto run = [&](int t, std::vector<uint8_t> &out, int *it, int *err) -> int {
        DecParams q = p; q.T = t;
        int r = launch(d, q, io, st); if (r) return r;
        if (cudaStreamSynchronize(st) != cudaSuccess) return set_err(LDPC_GPU_ERR_CUDA, "replay: kernel failed");
        if ((r = redo_after_sync(d, q, io, st, nullptr))) return r;
        if (cudaMemcpy(out.data(), bits.p, bpf, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(it, iters.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(err, errs.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return set_err(LDPC_GPU_ERR_CUDA, "replay: copy failed");
        return LDPC_GPU_OK;
    };
    int it = 0, err = 0, rows = 0;
    rc = run(0, prev, &it, &err);
    for (int t = 1; t <= T && !rc; t++) {
        rc = run(t, cur, &it, &err);
        if (rc) break;
        // bit-flipping decoders and DD-BMP stop early: a run that executed fewer than t steps adds no row
        const bool early = (d->cfg.kind == LDPC_GPU_KIND_GDBF || d->cfg.kind == LDPC_GPU_KIND_NGDBF_HW || d->cfg.kind == LDPC_GPU_KIND_NGDBF_SC) ? it < t
                         : (d->cfg.kind == LDPC_GPU_KIND_DDBMP ? it < t - 1 : false);
        if (early) break;
        if (rows < max_rows) {
            memcpy(trace_d + (size_t)rows * bpf, cur.data(), bpf);
            if (trace_syn) {
                uint8_t *sy = trace_syn + (size_t)rows * spf;
                memset(sy, 0, spf);
                for (int j = 0; j < M; j++) {
                    int par = 0;
                    for (int k = 0; k < d->h_rowdeg[j]; k++) { const int i = d->h_mlist[(size_t)j * d->h_dcm + k]; par ^= (prev[i >> 3] >> (i & 7)) & 1; }
                    if (par) sy[j >> 3] |= (uint8_t)(1u << (j & 7));
                }
            }
        }
        rows++;
        prev.swap(cur);
    }
    bits.release(); iters.release(); errs.release();
    if (rc) return rc;
    *n_rows = rows;
    if (final_errors) *final_errors = err;
    return LDPC_GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// channel dump: the samples the throughput entry feeds its decoder
// ------------------------------------------------------------------------------------------------
__global__ void dump_kernel(const CodeDev c, const DecParams p, const FrameIO io)
{
    const int N = c.N, nblk = (N + 3) >> 2;
    for (long long f = blockIdx.x; f < io.n_frames; f += gridDim.x) {
        const uint8_t *cw = codeword_row(io, c, f);
        const unsigned long long fid = (unsigned long long)(io.frame_begin + f);
        for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
            double y4[4];
            raw_samples4(io, p, c, f, cw, b, y4);
            for (int q = 0; q < 4; q++) if (4 * b + q < N) io.dump_y[(size_t)f * N + 4 * b + q] = y4[q];
        }
        if (!io.dump_noise) continue;
        if (p.kind == LDPC_GPU_KIND_NGDBF_HW) {
            for (int b = threadIdx.x; b < (LDPC_GPU_HW_QBUF + 3) / 4; b += blockDim.x) {
                float n4[4]; normal4(io.seed, fid, (uint32_t)b, (uint32_t)io.noise_row_base, STREAM_DECODER, n4);
                for (int q = 0; q < 4; q++) if (4 * b + q < LDPC_GPU_HW_QBUF) io.dump_noise[(size_t)f * LDPC_GPU_HW_QBUF + 4 * b + q] = (double)n4[q];
            }
        } else if (p.kind == LDPC_GPU_KIND_NGDBF_SC) {
            for (int b = threadIdx.x; b < (int)((io.noise_rows + 3) / 4); b += blockDim.x) {
                float n4[4]; normal4(io.seed, fid, (uint32_t)b, (uint32_t)io.noise_row_base, STREAM_DECODER, n4);
                for (int q = 0; q < 4; q++) if (4 * b + q < io.noise_rows) io.dump_noise[(size_t)f * io.noise_rows + 4 * b + q] = (double)n4[q];
            }
        } else if (p.kind == LDPC_GPU_KIND_GDBF && p.rows_per_step > 0) {
            for (long long row = 0; row < io.noise_rows; row++) {
                const int which = (int)(row % p.rows_per_step);
                const bool uniform = !(which == 0 && (p.flags & LDPC_GPU_F_ADDNOISE)) || (p.flags & LDPC_GPU_F_UNIFORMNOISE);
                for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
                    double v[4];
                    if (uniform) uniform4(io.seed, fid, (uint32_t)b, (uint32_t)(row + io.noise_row_base), STREAM_DECODER, v);
                    else { float n4[4]; normal4(io.seed, fid, (uint32_t)b, (uint32_t)(row + io.noise_row_base), STREAM_DECODER, n4); for (int q = 0; q < 4; q++) v[q] = (double)n4[q]; }
                    for (int q = 0; q < 4; q++) if (4 * b + q < N) io.dump_noise[((size_t)f * io.noise_rows + row) * N + 4 * b + q] = v[q];
                }
            }
        }
    }
}

extern "C" int ldpc_gpu_channel_dump(ldpc_gpu_decoder *d, const ldpc_gpu_channel *ch, uint64_t seed, int64_t frame_begin,
                                     int64_t n_frames, double *y, double *noise, int64_t noise_rows)
{
    if (!d || !y || n_frames < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "bad argument");
    DecParams p; int rc = channel_params(d, ch, p); if (rc) return rc;
    CU_TRY(cudaSetDevice(d->device));
    const int N = d->N;
    const size_t npf = d->cfg.kind == LDPC_GPU_KIND_NGDBF_HW ? (size_t)LDPC_GPU_HW_QBUF : d->cfg.kind == LDPC_GPU_KIND_NGDBF_SC ? (size_t)noise_rows : (size_t)noise_rows * N;
    DevBuf by, bn;
    if ((rc = by.reserve(8 * (size_t)N * n_frames))) return rc;
    if (noise && npf) if ((rc = bn.reserve(8 * npf * n_frames))) { by.release(); return rc; }
    FrameIO io; memset(&io, 0, sizeof io);
    io.n_frames = n_frames; io.frame_begin = frame_begin; io.seed = seed; io.cw_table = d->d_cwtab; io.n_cw = d->n_cw;
    io.noise_rows = noise_rows; io.dump_y = (double *)by.p; io.dump_noise = (noise && npf) ? (double *)bn.p : nullptr;
    cudaStream_t st = d->slot[0].st;
    dump_kernel<<<(unsigned)std::min<long long>(std::max<long long>(n_frames, 1), 4 * d->n_sm), 256, 0, st>>>(d->dev, p, io);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(y, by.p, 8 * (size_t)N * n_frames, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && io.dump_noise) e = cudaMemcpyAsync(noise, bn.p, 8 * npf * n_frames, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    by.release(); bn.release();
    if (e != cudaSuccess) return set_err(LDPC_GPU_ERR_CUDA, std::string("channel dump: ") + cudaGetErrorString(e));
    return LDPC_GPU_OK;
}

__global__ void philox_kernel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out)
{
    uint32_t r[4]; philox4x32_10(c0, c1, c2, c3, k0, k1, r);
    for (int q = 0; q < 4; q++) out[q] = r[q];
}
extern "C" int ldpc_gpu_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    if (ldpc_gpu_device_count() <= 0) return set_err(LDPC_GPU_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    uint32_t *dv = nullptr;
    CU_TRY(cudaMalloc(&dv, 16));
    philox_kernel<<<1, 1>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], dv);
    cudaError_t e = cudaMemcpy(out, dv, 16, cudaMemcpyDeviceToHost);
    cudaFree(dv);
    if (e != cudaSuccess) return set_err(LDPC_GPU_ERR_CUDA, cudaGetErrorString(e));
    return LDPC_GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// (e) the one collective: sum of the counter struct over ranks, NCCL over NVLink.
// NCCL is bound at run time (dlopen) so that the library itself links only the CUDA runtime.
// ------------------------------------------------------------------------------------------------
namespace {
struct Id128 { char b[128]; };                 // ncclUniqueId, passed by value
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl; void *g_comm = nullptr; int g_comm_dev = 0; cudaStream_t g_comm_stream = nullptr;

int load_nccl()
{
    if (g_nccl.lib) return LDPC_GPU_OK;
    const char *names[] = { getenv("LDPC_GPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    for (const char *n : names) { if (!n || !*n) continue; g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
    if (!g_nccl.lib) return set_err(LDPC_GPU_ERR_COMM, "cannot dlopen NCCL (set LDPC_GPU_NCCL_LIB to libnccl.so.2)");
    *(void **)&g_nccl.GetUniqueId = dlsym(g_nccl.lib, "ncclGetUniqueId");
    *(void **)&g_nccl.CommInitRank = dlsym(g_nccl.lib, "ncclCommInitRank");
    *(void **)&g_nccl.AllReduce = dlsym(g_nccl.lib, "ncclAllReduce");
    *(void **)&g_nccl.CommDestroy = dlsym(g_nccl.lib, "ncclCommDestroy");
    *(void **)&g_nccl.GetErrorString = dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return set_err(LDPC_GPU_ERR_COMM, "NCCL library lacks a required symbol");
    return LDPC_GPU_OK;
}
int nccl_err(int r, const char *what)
{
    return set_err(LDPC_GPU_ERR_COMM, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
} // namespace

extern "C" int ldpc_gpu_comm_unique_id(uint8_t id[128])
{
    int rc = load_nccl(); if (rc) return rc;
    int r = g_nccl.GetUniqueId(id); if (r) return nccl_err(r, "ncclGetUniqueId");
    return LDPC_GPU_OK;
}
extern "C" int ldpc_gpu_comm_init(const uint8_t id[128], int rank, int nranks, int device)
{
    int rc = load_nccl(); if (rc) return rc;
    if (g_comm) return set_err(LDPC_GPU_ERR_COMM, "communicator already initialised");
    CU_TRY(cudaSetDevice(device));
    Id128 idv; memcpy(idv.b, id, 128);
    int r = g_nccl.CommInitRank(&g_comm, nranks, idv, rank);
    if (r) { g_comm = nullptr; return nccl_err(r, "ncclCommInitRank"); }
    g_comm_dev = device;
    CU_TRY(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
    return LDPC_GPU_OK;
}
extern "C" int ldpc_gpu_comm_destroy(void)
{
    if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
    if (g_comm_stream) { cudaStreamDestroy(g_comm_stream); g_comm_stream = nullptr; }
    return LDPC_GPU_OK;
}
extern "C" int ldpc_gpu_allreduce_counters(ldpc_gpu_counters *c, int N, const ldpc_gpu_decoder_cfg *cfg)
{
    if (!c || !cfg) return set_err(LDPC_GPU_ERR_INVALID_ARG, "counters or cfg is NULL");
    const int maxphase = cfg->maxphase;
    if (!g_comm) return set_err(LDPC_GPU_ERR_COMM, "communicator not initialised");
    CU_TRY(cudaSetDevice(g_comm_dev));
    const size_t n_ew = c->error_weight_hist ? (size_t)N : 0, n_it = c->iter_hist ? (size_t)iter_hist_len(*cfg) : 0, n_ph = c->phase_hist ? (size_t)std::max(1, maxphase) : 0;
    std::vector<int64_t> h(8 + n_ew + n_it + n_ph);
    int64_t *q = h.data();
    q[0] = c->errors; q[1] = c->uncodedErrors; q[2] = c->totalBits; q[3] = c->totalWords; q[4] = c->wordErrors;
    q[5] = c->totalIterations; q[6] = c->smoothingUsed; q[7] = c->undetectedWords; q += 8;
    if (n_ew) { memcpy(q, c->error_weight_hist, 8 * n_ew); q += n_ew; }
    if (n_it) { memcpy(q, c->iter_hist, 8 * n_it); q += n_it; }
    if (n_ph) { memcpy(q, c->phase_hist, 8 * n_ph); q += n_ph; }
    void *dv = nullptr;
    CU_TRY(cudaMalloc(&dv, 8 * h.size()));
    cudaError_t e = cudaMemcpyAsync(dv, h.data(), 8 * h.size(), cudaMemcpyHostToDevice, g_comm_stream);
    int r = 0;
    if (e == cudaSuccess) r = g_nccl.AllReduce(dv, dv, h.size(), /*ncclInt64*/ 4, /*ncclSum*/ 0, g_comm, g_comm_stream);
    if (e == cudaSuccess && !r) e = cudaMemcpyAsync(h.data(), dv, 8 * h.size(), cudaMemcpyDeviceToHost, g_comm_stream);
    if (e == cudaSuccess && !r) e = cudaStreamSynchronize(g_comm_stream);
    cudaFree(dv);
    if (r) return nccl_err(r, "ncclAllReduce");
    if (e != cudaSuccess) return set_err(LDPC_GPU_ERR_CUDA, cudaGetErrorString(e));
    q = h.data();
    c->errors = q[0]; c->uncodedErrors = q[1]; c->totalBits = q[2]; c->totalWords = q[3]; c->wordErrors = q[4];
    c->totalIterations = q[5]; c->smoothingUsed = q[6]; c->undetectedWords = q[7]; q += 8;
    if (n_ew) { memcpy(c->error_weight_hist, q, 8 * n_ew); q += n_ew; }
    if (n_it) { memcpy(c->iter_hist, q, 8 * n_it); q += n_it; }
    if (n_ph) { memcpy(c->phase_hist, q, 8 * n_ph); q += n_ph; }
    return LDPC_GPU_OK;
}


// ------------------------------------------------------------------------------------------------
// non-binary GF(q) codes, min-max decoding (csrc/ldpc_nb_kernel.cuh).  Parity unpinned.
// ------------------------------------------------------------------------------------------------
struct ldpc_gpu_nb_code {
    int N = 0, M = 0, q = 0, m = 0, E = 0, dv_max = 0, dc_max = 0;
    std::vector<int> col_deg, row_deg, nlist, mlist;      // 0-based, -1 padded
    std::vector<uint8_t> mvals;                           // [M*dc_max] h_jk
    std::vector<uint8_t> mul, inv;
};

static int gf_tables(int q, std::vector<uint8_t> &mul, std::vector<uint8_t> &inv)
{
    int m = 0; while ((1 << m) < q) m++;
    if ((1 << m) != q || m < 1 || m > 6) return -1;
    static const int prim[7] = { 0, 0x3, 0x7, 0xB, 0x13, 0x25, 0x43 };        // x+1, x^2+x+1, x^3+x+1, x^4+x+1, x^5+x^2+1, x^6+x+1
    mul.assign((size_t)q * q, 0); inv.assign(q, 0);
    for (int a = 0; a < q; a++) for (int b = 0; b < q; b++) {
        int r = 0, aa = a;
        for (int k = 0; k < m; k++) { if ((b >> k) & 1) r ^= aa; aa <<= 1; if (aa & q) aa ^= prim[m]; }
        mul[(size_t)a * q + b] = (uint8_t)r;
    }
    for (int a = 1; a < q; a++) for (int b = 1; b < q; b++) if (mul[(size_t)a * q + b] == 1) inv[a] = (uint8_t)b;
    return m;
}

extern "C" int ldpc_gpu_nb_code_create(int N, int M, int q, int dvm, int dcm, const int *num_nlist, const int *num_mlist,
                                       const int *nlist_flat, const int *nvals_flat, const int *mlist_flat, const int *mvals_flat,
                                       ldpc_gpu_nb_code **out)
{
    if (!out) return set_err(LDPC_GPU_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (N <= 0 || M <= 0 || dvm <= 0 || dcm <= 0 || !num_nlist || !num_mlist || !nlist_flat || !mlist_flat || !mvals_flat)
        return set_err(LDPC_GPU_ERR_INVALID_ARG, "bad dimensions or NULL array");
    if (dcm > 8 || dvm > 16) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "non-binary decoder: dc_max <= 8, dv_max <= 16");
    std::unique_ptr<ldpc_gpu_nb_code> c(new ldpc_gpu_nb_code);
    c->m = gf_tables(q, c->mul, c->inv);
    if (c->m < 0) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "non-binary decoder: q must be 2, 4, 8, 16, 32 or 64");
    c->N = N; c->M = M; c->q = q; c->dv_max = dvm; c->dc_max = dcm;
    c->col_deg.assign(num_nlist, num_nlist + N); c->row_deg.assign(num_mlist, num_mlist + M);
    c->nlist.assign((size_t)N * dvm, -1); c->mlist.assign((size_t)M * dcm, -1); c->mvals.assign((size_t)M * dcm, 0);
    long long En = 0, Em = 0;
    for (int j = 0; j < M; j++) {
        if (num_mlist[j] < 2 || num_mlist[j] > dcm) return set_err(LDPC_GPU_ERR_BAD_CODE, "row weight outside [2, biggest_num_m]");
        Em += num_mlist[j];
        for (int k = 0; k < num_mlist[j]; k++) {
            const int v = mlist_flat[(size_t)j * dcm + k] - 1, h = mvals_flat[(size_t)j * dcm + k];
            if (v < 0 || v >= N || h <= 0 || h >= q) return set_err(LDPC_GPU_ERR_BAD_CODE, "mlist entry or GF value out of range");
            c->mlist[(size_t)j * dcm + k] = v; c->mvals[(size_t)j * dcm + k] = (uint8_t)h;
        }
    }
    for (int i = 0; i < N; i++) {
        if (num_nlist[i] < 0 || num_nlist[i] > dvm) return set_err(LDPC_GPU_ERR_BAD_CODE, "column weight exceeds biggest_num_n");
        En += num_nlist[i];
        for (int s2 = 0; s2 < num_nlist[i]; s2++) {
            const int j = nlist_flat[(size_t)i * dvm + s2] - 1;
            if (j < 0 || j >= M) return set_err(LDPC_GPU_ERR_BAD_CODE, "nlist entry out of range");
            c->nlist[(size_t)i * dvm + s2] = j;
            int found = -1;                                   // the two lists must describe the same matrix, values included
            for (int k = 0; k < c->row_deg[j]; k++) if (c->mlist[(size_t)j * dcm + k] == i) found = k;
            if (found < 0) return set_err(LDPC_GPU_ERR_BAD_CODE, "nlist is not the transpose of mlist");
            if (nvals_flat && nvals_flat[(size_t)i * dvm + s2] != c->mvals[(size_t)j * dcm + found]) return set_err(LDPC_GPU_ERR_BAD_CODE, "nlist and mlist disagree on a GF value");
        }
    }
    if (En != Em) return set_err(LDPC_GPU_ERR_BAD_CODE, "nlist and mlist hold different numbers of edges");
    c->E = (int)Em;
    *out = c.release();
    return LDPC_GPU_OK;
}

// header `N M q`, then the binary layout with every index followed by its GF value, `0 0` padding
// (/root/reference/SystemC/NB-LDPC/src/alist.cpp:23-56,97-124)
extern "C" int ldpc_gpu_nb_code_load_alist(const char *path, ldpc_gpu_nb_code **out)
{
    if (!out || !path) return set_err(LDPC_GPU_ERR_INVALID_ARG, "NULL argument");
    *out = nullptr;
    std::ifstream f(path);
    if (!f) return set_err(LDPC_GPU_ERR_IO, std::string("cannot open ") + path);
    std::vector<long> t; long x;
    while (f >> x) t.push_back(x);
    if (t.size() < 5) return set_err(LDPC_GPU_ERR_BAD_CODE, "bad non-binary alist header");
    const long N = t[0], M = t[1], q = t[2], dvm = t[3], dcm = t[4];
    if (N <= 0 || M <= 0 || dvm <= 0 || dcm <= 0 || N > (1 << 24) || M > (1 << 24)) return set_err(LDPC_GPU_ERR_BAD_CODE, "bad non-binary alist dimensions");
    const size_t need = 5 + (size_t)N + M + 2 * ((size_t)N * dvm + (size_t)M * dcm);
    if (t.size() != need) return set_err(LDPC_GPU_ERR_BAD_CODE, "non-binary alist: token count does not match its header (rows must be padded with `0 0`)");
    std::vector<int> nn(N), nm(M), nl((size_t)N * dvm), nv((size_t)N * dvm), ml((size_t)M * dcm), mv((size_t)M * dcm);
    size_t p = 5;
    for (long i = 0; i < N; i++) nn[i] = (int)t[p++];
    for (long j = 0; j < M; j++) nm[j] = (int)t[p++];
    for (size_t e = 0; e < (size_t)N * dvm; e++) { nl[e] = (int)t[p++]; nv[e] = (int)t[p++]; }
    for (size_t e = 0; e < (size_t)M * dcm; e++) { ml[e] = (int)t[p++]; mv[e] = (int)t[p++]; }
    return ldpc_gpu_nb_code_create((int)N, (int)M, (int)q, (int)dvm, (int)dcm, nn.data(), nm.data(), nl.data(), nv.data(), ml.data(), mv.data(), out);
}
extern "C" int ldpc_gpu_nb_code_dims(const ldpc_gpu_nb_code *c, int *N, int *M, int *q, int *E)
{
    if (!c) return set_err(LDPC_GPU_ERR_INVALID_ARG, "code is NULL");
    if (N) *N = c->N; if (M) *M = c->M; if (q) *q = c->q; if (E) *E = c->E;
    return LDPC_GPU_OK;
}
extern "C" int ldpc_gpu_nb_code_destroy(ldpc_gpu_nb_code *c) { delete c; return LDPC_GPU_OK; }

struct ldpc_gpu_nb_decoder {
    int device = 0, T = 0, grid = 0, block = 256;
    NbCodeDev dev; std::vector<void *> owned;
    double *d_ws = nullptr; size_t ws_stride = 0;
    unsigned long long *d_counters = nullptr;
    cudaStream_t st = nullptr; cudaEvent_t k0 = nullptr, k1 = nullptr;
    double last_ms = 0;
};
typedef void (*NbKernelFn)(const NbCodeDev, const NbIO);

extern "C" int ldpc_gpu_nb_decoder_destroy(ldpc_gpu_nb_decoder *d)
{
    if (!d) return LDPC_GPU_OK;
    cudaSetDevice(d->device);
    for (void *p : d->owned) cudaFree(p);
    if (d->d_ws) cudaFree(d->d_ws);
    if (d->d_counters) cudaFree(d->d_counters);
    if (d->k0) cudaEventDestroy(d->k0); if (d->k1) cudaEventDestroy(d->k1);
    if (d->st) cudaStreamDestroy(d->st);
    delete d;
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_nb_decoder_create(const ldpc_gpu_nb_code *c, int num_iterations, int device, ldpc_gpu_nb_decoder **out)
{
    if (!out) return set_err(LDPC_GPU_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!c || num_iterations < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "code is NULL or num_iterations < 0");
    if (c->q > 16) return set_err(LDPC_GPU_ERR_UNSUPPORTED, "non-binary kernel instantiations: q = 2, 4, 8, 16");
    if (ldpc_gpu_device_count() <= 0) return set_err(LDPC_GPU_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CU_TRY(cudaSetDevice(device));
    ldpc_gpu_nb_decoder *d = new ldpc_gpu_nb_decoder;
    d->device = device; d->T = num_iterations;
    auto up = [&](const void *h, size_t n, const void **p) -> int {
        void *q2 = nullptr;
        if (cudaMalloc(&q2, std::max<size_t>(16, n)) != cudaSuccess) return set_err(LDPC_GPU_ERR_NOMEM, "cudaMalloc failed");
        d->owned.push_back(q2);
        if (cudaMemcpy(q2, h, n, cudaMemcpyHostToDevice) != cudaSuccess) return set_err(LDPC_GPU_ERR_CUDA, "cudaMemcpy failed");
        *p = q2; return LDPC_GPU_OK;
    };
    const int N = c->N, M = c->M, dvm = c->dv_max, dcm = c->dc_max;
    std::vector<int32_t> vn_edge((size_t)N * dvm, -1);
    std::vector<uint8_t> cdeg(M), vdeg(N);
    for (int j = 0; j < M; j++) cdeg[j] = (uint8_t)c->row_deg[j];
    for (int i = 0; i < N; i++) {
        vdeg[i] = (uint8_t)c->col_deg[i];
        for (int s2 = 0; s2 < c->col_deg[i]; s2++) {
            const int j = c->nlist[(size_t)i * dvm + s2];
            for (int k = 0; k < c->row_deg[j]; k++) if (c->mlist[(size_t)j * dcm + k] == i) vn_edge[(size_t)i * dvm + s2] = j * dcm + k;
        }
    }
    NbCodeDev &v = d->dev;
    v.N = N; v.M = M; v.q = c->q; v.m = c->m; v.E = c->E; v.dv_max = dvm; v.dc_max = dcm;
    int rc;
    if ((rc = up(c->mlist.data(), c->mlist.size() * 4, (const void **)&v.cn_var)) || (rc = up(c->mvals.data(), c->mvals.size(), (const void **)&v.cn_val)) ||
        (rc = up(cdeg.data(), cdeg.size(), (const void **)&v.cn_deg)) || (rc = up(vn_edge.data(), vn_edge.size() * 4, (const void **)&v.vn_edge)) ||
        (rc = up(vdeg.data(), vdeg.size(), (const void **)&v.vn_deg)) || (rc = up(c->mul.data(), c->mul.size(), (const void **)&v.mul)) ||
        (rc = up(c->inv.data(), c->inv.size(), (const void **)&v.inv))) { ldpc_gpu_nb_decoder_destroy(d); return rc; }
    int n_sm = 0; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    d->grid = 2 * n_sm;
    d->ws_stride = ((4 * (size_t)M * dcm + (size_t)N) * c->q + ((size_t)N + 7) / 8 + 31) & ~(size_t)31;        // doubles
    if (cudaMalloc(&d->d_ws, d->ws_stride * sizeof(double) * d->grid) != cudaSuccess || cudaMalloc(&d->d_counters, sizeof(unsigned long long) * CNT_N) != cudaSuccess ||
        cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&d->k0) != cudaSuccess || cudaEventCreate(&d->k1) != cudaSuccess) {
        ldpc_gpu_nb_decoder_destroy(d); return set_err(LDPC_GPU_ERR_NOMEM, "non-binary decoder: allocation failed");
    }
    *out = d;
    return LDPC_GPU_OK;
}

static int nb_run(ldpc_gpu_nb_decoder *d, const ldpc_gpu_channel *ch, NbIO io, ldpc_gpu_counters *cnt)
{
    if (!ch || !(ch->R > 0)) return set_err(LDPC_GPU_ERR_INVALID_ARG, "channel needs R > 0");
    io.sigma = sqrt(pow(10.0, -ch->snr_db / 10.0) / ch->R / 2.0);
    io.T = d->T; io.workspace = d->d_ws; io.ws_stride = d->ws_stride; io.counters = cnt ? d->d_counters : nullptr;
    if (cnt) CU_TRY(cudaMemsetAsync(d->d_counters, 0, sizeof(unsigned long long) * CNT_N, d->st));
    const NbKernelFn fn = d->dev.q == 2 ? (NbKernelFn)nb_minmax_kernel<2> : d->dev.q == 4 ? (NbKernelFn)nb_minmax_kernel<4>
                        : d->dev.q == 8 ? (NbKernelFn)nb_minmax_kernel<8> : (NbKernelFn)nb_minmax_kernel<16>;
    const long long want = std::min<long long>(io.n_frames, d->grid);
    CU_TRY(cudaEventRecord(d->k0, d->st));
    if (want > 0) { fn<<<(unsigned)want, d->block, 0, d->st>>>(d->dev, io); CU_TRY(cudaGetLastError()); }
    CU_TRY(cudaEventRecord(d->k1, d->st));
    if (cnt) {
        unsigned long long h[CNT_N];
        CU_TRY(cudaMemcpyAsync(h, d->d_counters, sizeof h, cudaMemcpyDeviceToHost, d->st));
        CU_TRY(cudaStreamSynchronize(d->st));
        cnt->errors += (int64_t)h[CNT_ERRORS]; cnt->totalBits += (int64_t)h[CNT_BITS]; cnt->totalWords += (int64_t)h[CNT_WORDS];
        cnt->wordErrors += (int64_t)h[CNT_WORDERRS]; cnt->totalIterations += (int64_t)h[CNT_ITERS]; cnt->undetectedWords += (int64_t)h[CNT_UNDETECTED];
        cnt->smoothingUsed += (int64_t)h[CNT_SMOOTH];
    } else CU_TRY(cudaStreamSynchronize(d->st));
    float ms = 0; cudaEventElapsedTime(&ms, d->k0, d->k1); d->last_ms = ms;
    return LDPC_GPU_OK;
}

extern "C" int ldpc_gpu_nb_decode_batch(ldpc_gpu_nb_decoder *d, const ldpc_gpu_channel *ch, int64_t n_frames, const double *y,
                                        uint8_t *out_symbols, int32_t *out_iters, ldpc_gpu_counters *cnt)
{
    if (!d || n_frames < 0 || (n_frames > 0 && !y)) return set_err(LDPC_GPU_ERR_INVALID_ARG, "bad argument");
    CU_TRY(cudaSetDevice(d->device));
    const size_t ny = (size_t)n_frames * d->dev.N * d->dev.m;
    DevBuf by, bs, bi;
    int rc;
    if ((rc = by.reserve(8 * ny)) || (out_symbols && (rc = bs.reserve((size_t)n_frames * d->dev.N))) || (out_iters && (rc = bi.reserve(4 * (size_t)n_frames)))) { by.release(); bs.release(); bi.release(); return rc; }
    NbIO io; memset(&io, 0, sizeof io);
    io.n_frames = n_frames; io.y = (const double *)by.p; io.out_symbols = out_symbols ? (uint8_t *)bs.p : nullptr; io.out_iters = out_iters ? (int *)bi.p : nullptr;
    cudaError_t e = ny ? cudaMemcpy(by.p, y, 8 * ny, cudaMemcpyHostToDevice) : cudaSuccess;
    rc = e == cudaSuccess ? nb_run(d, ch, io, cnt) : set_err(LDPC_GPU_ERR_CUDA, "copy failed");
    if (!rc && out_symbols && n_frames) e = cudaMemcpy(out_symbols, bs.p, (size_t)n_frames * d->dev.N, cudaMemcpyDeviceToHost);
    if (!rc && e == cudaSuccess && out_iters && n_frames) e = cudaMemcpy(out_iters, bi.p, 4 * (size_t)n_frames, cudaMemcpyDeviceToHost);
    by.release(); bs.release(); bi.release();
    if (!rc && e != cudaSuccess) return set_err(LDPC_GPU_ERR_CUDA, "copy failed");
    return rc;
}

extern "C" int ldpc_gpu_nb_simulate(ldpc_gpu_nb_decoder *d, const ldpc_gpu_channel *ch, uint64_t seed, int64_t frame_begin, int64_t n_frames,
                                    ldpc_gpu_counters *cnt, double *kernel_ms)
{
    if (!d || !cnt || n_frames < 0) return set_err(LDPC_GPU_ERR_INVALID_ARG, "bad argument");
    CU_TRY(cudaSetDevice(d->device));
    NbIO io; memset(&io, 0, sizeof io);
    io.n_frames = n_frames; io.frame_begin = frame_begin; io.seed = seed;
    int rc = nb_run(d, ch, io, cnt);
    if (kernel_ms) *kernel_ms = d->last_ms;
    return rc;
}
