// ldpc_types.cuh -- device-side views shared by the kernels and the C-ABI host code.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/ldpc_gpu.h"

namespace ldpc {

// Parity-check matrix as the kernels see it.  Built once by ldpc_gpu_code_create from the
// arrays loadFile() yields (inc/alist.h:21-36).  The reverse-slot lookup the reference redoes
// with find() on every message access (src/decodeMinSum.cpp:527-536, 74 % of its run time) is
// folded into `cn_pos` here.
//
// Edge messages live variable-major, slot-major:  msg[s*N + i]  = the message on the edge
// between variable i and the s-th check of its nlist row.  A variable thread therefore walks
// its edges at stride N (conflict-free across a warp of consecutive i), and a check thread
// reaches its k-th edge through cn_pos.
struct CodeDev {
    int N, M, E, dv_max, dc_max;
    int dvN;                 // dv_max * N : extent of one message array
    int idx16;               // cn_pos holds uint16_t (dvN <= 65535) else uint32_t
    int regular_dc;          // common row weight, or -1
    int regular_dv;          // common column weight, or -1
    const void    *cn_pos;   // [(k / VPL)][j][VPL]  message position s*N+i of slot k of check j, VPL = 16/sizeof(idx);
                             //                      one 16-byte vector per thread per load; padded with 0
    const uint8_t *cn_deg;   // [M]
    const uint8_t *vn_deg;   // [N]
    const uint32_t*cn_var;   // [k*M + j]  variable of slot k of check j (bit-flipping kernels), padded with 0
    const uint32_t*vn_chk;   // [s*N + i]  check of slot s of variable i, padded with 0
    // bank-conflict-free check schedule of a regular code (ldpc_schedule.h), or NULL
    const uint4   *sched;    // [(t/4)][j]  four BYTE offsets (s*N + col(i))*sizeof(Real) of steps 4(t/4)..+3 of row j
    const uint16_t*col_of_var; // [N] storage column of variable i
    const uint16_t*var_of_col; // [N] inverse
    const uint8_t *row_slot; // [M] slot shared by every edge of row j (ldpc_ms_rc.cuh), or NULL when rows mix slots
    // small codes, four frames per 16-byte word (ldpc_ms_quad.cuh; ldpc_schedule.h: build_group_schedule), or NULL
    const uint32_t*quad_edge;  // [s*M + t]  step s of thread t: (storage column << 16) | (slot*N + storage column); idle steps: (N << 16) | dvN
    const uint8_t *quad_steps; // [M] steps thread t needs (its last busy step + 1)
    const uint16_t*quad_col_of_var; // [N] storage column of variable i
    const uint16_t*quad_var_of_col; // [N] inverse
};

// Decoder configuration + per-call channel constants, all derived on the host in double with the
// reference's own expressions so that the device never re-derives a constant differently.
struct DecParams {
    int      kind; uint32_t flags; int T; int Q, NQ; int windowsize, maxphase, Tswitch;
    double   Ymax, alpha, delta, theta, lambda, noiseScale, w, theta0, MAXLLR;
    double   ms_Nq1;        // Nq - 1,             Nq = pow(2,Q)          decodeMinSum.cpp:125
    double   ms_twoY;       // 2.0 * Ymax
    double   ms_inv_twoY;   // 1/(2*Ymax) when 2*Ymax is a power of two (the division is then exact), else 0
    double   ms_step;       // 2*Ymax/(Nq-1)                              :485
    double   g_qmax;        // pow(2, NQ-1)                               decodeGDBF.cpp:490
    double   g_twol;        // 2 * (Ymax/2.0)
    double   g_step;        // 2.0*lmax/qmax
    double   hw_lmax, hw_NL, hw_two_lmax, hw_two_w;   // NGDBFhw.cpp:171-173
    int      hw_theta, hw_Smult;                      // :175-176
    double   sigma, N0, noiseSigma;                   // decodeMinSum.cpp:146-147, decodeGDBF.cpp:296
    double   uni_scale;     // (sqrt(3)*noiseSigma)*2.0                   decodeGDBF.cpp:322
    float    inv_alpha_f;   // fp32 instantiation: RN(1/alpha)
    float    alpha_div_f;   // (float)alpha when the fp32 normalisation needs the division-correction step (alpha not a power of two), else 0
    int      ms_step_dyadic; // 2*Ymax/(Nq-1) is a power of two times a small integer: level * step is exact in fp32
    float    ms_scale_f, ms_step_f, Ymax_f;   // fp32 front end (fp32 instantiation fed by the Philox channel or fp32 samples)
    uint32_t x2_delta2, x2_cap2, x2_m02;   // exact-lattice packed kernel (ldpc_ms_x2.cuh), binary16 pairs: offset (0 for plain min-sum), c2v cap, stable-state bound
    int      channel_mode;  // LDPC_GPU_CHANNEL_*
    int      zero;          // always 0: an offset the compiler cannot see through (ldpc_ms_x2.cuh)
    float    sigma_f;       // fast channel: (float)sigma
    int      iter_hist_len;
    int      rows_per_step; // GDBF noise rows consumed per flip step
};

// One launch worth of frames.
struct FrameIO {
    long long n_frames;
    long long frame_begin;        // global id of frame 0 of this launch (Philox counter)
    const void   *y;              // device [F][N] raw samples, or NULL -> Philox channel
    int           y_dtype;
    const double *noise;          // device, ldpc_gpu_batch layout, or NULL -> Philox
    long long     noise_rows;
    const uint8_t*codeword;       // device [F][N] 0/1, or NULL
    const uint8_t*cw_table;       // device [n_cw][N] 0/1 (simulate), or NULL
    long long     n_cw;
    const int    *qpointer0;      // device [F] or NULL
    uint8_t      *out_bits;       // device [F][ceil(N/8)] or NULL
    int          *out_iters;
    void         *out_soft;
    int          *out_errors;
    uint8_t      *out_flags;
    unsigned long long *counters; // device [8] or NULL: order of ldpc_gpu_counters scalars
    unsigned long long *ew_hist;  // device [N] or NULL
    unsigned long long *it_hist;  // device [iter_hist_len] or NULL
    unsigned long long *ph_hist;  // device [maxphase] or NULL
    unsigned long long seed;
    long long     noise_row_base;  // added to the row index of the decoder-noise stream (Philox path): re-decode r of a frame
                                  // draws rows r * rows_per_decode ... (ldpc_gpu_redecode_stats)
    unsigned char *workspace;     // device, gridDim.x * ws_stride bytes: per-CTA frame state of the HBM-resident
    size_t         ws_stride;     //   instantiations (codes whose state exceeds one SM's shared memory)
    // exact-lattice packed kernel: frames whose decisions it cannot certify (batch-relative indices), decoded by the fp64 kernel
    long long     *redo_list;
    unsigned int  *redo_count;
    unsigned long long *redo_total; // running total over the decoder's life (statistics)
    // frame indirection of the redo launch: frame f of the launch is frame frame_list[f] of the batch; the launch size is read
    // on the device so that no host synchronisation sits between the two kernels
    const long long *frame_list;
    const unsigned int *n_frames_dev;
    // channel_dump outputs
    double       *dump_y;
    double       *dump_noise;
};

enum { CNT_ERRORS = 0, CNT_UNCODED, CNT_BITS, CNT_WORDS, CNT_WORDERRS, CNT_ITERS, CNT_SMOOTH, CNT_UNDETECTED, CNT_N };

} // namespace ldpc
