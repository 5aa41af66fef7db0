/* ldpc_gpu.h -- C ABI of the B200-native Monte-Carlo LDPC decode path.
 *
 * Drop-in boundary for the hot path of ereiss123/LDPCsimulation's
 * C_implementations/ (reference paths below are relative to that tree).
 * The reference has no library interface: every decoder is a main() that calls
 * file-local node-update functions once per frame / iteration, with the
 * algorithm variant picked by -D macros and its parameters held in mutable
 * globals.  This header is the batch-granular replacement of that seam:
 *
 *   reference seam                                         replaced by
 *   -----------------------------------------------------  -------------------------------
 *   loadFile()                  inc/alist.h:39              ldpc_gpu_code_create / _load_alist
 *   -D macros  Makefile:24-71, globals decodeGDBF.cpp:48-56 ldpc_gpu_decoder_cfg
 *   channel block          src/decodeMinSum.cpp:214-238     fused into decode/simulate kernels
 *   initializeSymMessages  src/decodeMinSum.cpp:364-370   \
 *   checkNodeUpdates       src/decodeMinSum.cpp:410-450    |
 *   applyNormalization/Offset          :494-515            |
 *   symNodeUpdates         src/decodeMinSum.cpp:452-476    |  ldpc_gpu_decode_batch  (caller's samples)
 *   BP nodes               src/decodeBP.cpp:353-409        |  ldpc_gpu_simulate      (on-device channel)
 *   GDBF nodes             src/decodeGDBF.cpp:517-633      |
 *   RNGDBF phase loop      src/RNGDBF.cpp:277-404          |
 *   NGDBFhw nodes          src/NGDBFhw.cpp:546-677         |
 *   DD-BMP nodes           src/decodeDDBMP.cpp:350-423    /
 *   countDecisionErrors + accounting  decodeMinSum.cpp:270-288,382-393   ldpc_gpu_counters
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success
 * or a negative LDPC_GPU_ERR_* code (never aborts); ldpc_gpu_last_error() gives
 * the message of the last failure on the calling thread.  Handles are opaque,
 * created/destroyed by the library; all I/O buffers are caller-owned.  One
 * decoder handle per host thread / GPU; calls on different handles may run
 * concurrently.  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with LDPC_GPU_ERR_CUDA.
 */
#ifndef LDPC_GPU_H
#define LDPC_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------- */
#define LDPC_GPU_OK                 0
#define LDPC_GPU_ERR_INVALID_ARG   -1
#define LDPC_GPU_ERR_CUDA          -2
#define LDPC_GPU_ERR_NOMEM         -3
#define LDPC_GPU_ERR_UNSUPPORTED   -4
#define LDPC_GPU_ERR_COMM          -5
#define LDPC_GPU_ERR_IO            -6
#define LDPC_GPU_ERR_BAD_CODE      -7

/* ---- decoder kinds: one per reference translation unit ----------------- */
#define LDPC_GPU_KIND_MINSUM    0   /* src/decodeMinSum.cpp  (plain / offset / normalised by flags) */
#define LDPC_GPU_KIND_BP        1   /* src/decodeBP.cpp */
#define LDPC_GPU_KIND_GDBF      2   /* src/decodeGDBF.cpp, src/RNGDBF.cpp (LDPC_GPU_F_REDECODE) */
#define LDPC_GPU_KIND_NGDBF_HW  3   /* src/NGDBFhw.cpp */
#define LDPC_GPU_KIND_DDBMP     4   /* src/decodeDDBMP.cpp */
#define LDPC_GPU_KIND_NGDBF_SC 5 /* NGDBF as the reference's SystemC model runs it (SystemC/NGDBF/inc/nodes.h:76-138,
                                   inc/decoder.h:183-254): Nq = 2^Q level quantiser with end points, local
                                   threshold adapted both ways (theta /= lambda on a flip, *= lambda otherwise), w = alpha*Ymax/dv,
                                   one quantised Gaussian per clock shifted node to node, smoothing over the last `windowsize`
                                   (32 in the model) iterations, one clock of pipeline delay.  Parameters: num_iterations, theta,
                                   lambda, Ymax, Q (the model's `precision`), alpha, windowsize, noiseScale (1 in the model).
                                   batch.noise: [F][noise_rows] raw N(0,1) values, noise_rows >= N + T + 1.  No runnable reference
                                   (SystemC absent): parity unpinned, checked against the C restatement in oracle/. */

/* ---- flag bits: one per reference -D macro (Makefile:24-71) ------------ */
#define LDPC_GPU_F_QUANTIZE_SAMPLES       (1u << 0)   /* -D quantizeSamples  */
#define LDPC_GPU_F_SATURATE_SAMPLES       (1u << 1)   /* -D saturateSamples  */
#define LDPC_GPU_F_NORMALIZED_MS          (1u << 2)   /* -D normalizedMS     */
#define LDPC_GPU_F_OFFSET_MS              (1u << 3)   /* -D offsetMS         */
#define LDPC_GPU_F_SEQUENTIALMODE         (1u << 4)   /* -D sequentialmode   */
#define LDPC_GPU_F_MODESWITCHING          (1u << 5)   /* -D modeswitching    */
#define LDPC_GPU_F_ADDNOISE               (1u << 6)   /* -D addNoise         */
#define LDPC_GPU_F_WEIGHTSYNDROMES        (1u << 7)   /* -D weightSyndromes  */
#define LDPC_GPU_F_OUTPUTSMOOTHING        (1u << 8)   /* -D outputSmoothing  */
#define LDPC_GPU_F_THRESHOLDADAPTATION    (1u << 9)   /* -D thresholdAdaptation */
#define LDPC_GPU_F_UNIFORMNOISE           (1u << 10)  /* -D uniformNoise     */
#define LDPC_GPU_F_NOISESHAPING           (1u << 11)  /* -D noiseShaping     */
#define LDPC_GPU_F_QUANTIZEPROBABILITIES  (1u << 12)  /* -D quantizeProbabilities */
#define LDPC_GPU_F_REDECODE               (1u << 13)  /* -D redecode: RNGDBF.cpp conventions
                                                         (phase loop, w = alpha*Ymax/dv, clip y before copy) */

#define LDPC_GPU_F_CERT_STOP              (1u << 14)  /* no reference macro.  LDPC_GPU_PREC_F16X2 on an exact lattice only (csrc/ldpc_ms_x2.cuh):
                                                         stop iterating a frame once its decisions are certified final -- the reported decisions,
                                                         iteration counts (T) and counters are still those of T full iterations; out_soft is refused */

/* ---- arithmetic of the message / metric path --------------------------- */
#define LDPC_GPU_PREC_F64   0   /* parity instantiation: IEEE double, the reference's operation order */
#define LDPC_GPU_PREC_F32   1   /* throughput instantiation: fp32 messages, same dataflow */
#define LDPC_GPU_PREC_F16X2 2   /* min-sum family, binary16 messages.  On an exact lattice (plain / offset min-sum, quantised samples with a
                                   power-of-two quantiser unit) the results are the reference's bit for bit: two frames per lane on the 802.3an
                                   H (csrc/ldpc_ms_x2.cuh), 64-frame binary16 message tiles in HBM for codes beyond one SM with dc, dv <= 8
                                   (csrc/ldpc_ms_tileh.cuh, DVB-S2); frames that leave the exactly representable range are re-decoded by the
                                   fp64 instantiation (ldpc_gpu_decoder_stats).  Off the lattice (802.3an H only): v2c clamped to +-512, NOT the
                                   reference's arithmetic: decisions match on converging frames, BER/FER within confidence intervals */

#define LDPC_GPU_DT_QP   4   /* the same quantiser levels, bit-packed: Q bits per sample (Q = cfg.Q, 2 <= Q <= 8), sample i of a frame in bits
                              * [iQ, (i+1)Q) of the frame's N*Q/8 bytes (little-endian bit order; N*Q must be a multiple of 32).  Code =
                              * (negative << (Q-1)) | (min(k, 2^(Q-1)) - 1) with k the Q8 level magnitude (saturation = all ones).  A host that
                              * streams samples to several GPUs is bound by its own memory bandwidth: 5 bits instead of 8 per sample. */

/* ---- on-device noise generator (Philox4x32-10 keyed by (seed, frame id) either way) ---- */
#define LDPC_GPU_CHANNEL_EXACT 0   /* Box-Muller in explicitly rounded fp32 polynomials: reproducible bit for bit on a CPU (oracle/) */
#define LDPC_GPU_CHANNEL_FAST  1   /* Box-Muller on the SFU approximations (lg2 / sqrt / sin / cos .approx) and y = x fma(sigma, n, 1) in
                                      fp32: deterministic per GPU architecture, ~5x cheaper; the samples are available through
                                      ldpc_gpu_channel_dump only */

/* ---- where a caller buffer lives --------------------------------------- */
#define LDPC_GPU_MEM_HOST    0
#define LDPC_GPU_MEM_DEVICE  1

/* ---- element type of caller sample / soft buffers ---------------------- */
#define LDPC_GPU_DT_F64  0
#define LDPC_GPU_DT_F32  1
#define LDPC_GPU_DT_F16  2   /* IEEE binary16 samples in (halves the host->device bytes); out_soft is then fp32 */
#define LDPC_GPU_DT_Q8   3   /* int8 quantiser levels in, for the min-sum family with LDPC_GPU_F_QUANTIZE_SAMPLES: the samples as a
                              * Q-bit converter delivers them.  k = +-max(1, floor(|y| (Nq-1) / (2 Ymax))), Nq = 2^Q, with the sign of y,
                              * +-32 for |y| > Ymax; the decoder uses k * 2 Ymax / (Nq-1) (or +-Ymax), which is what quantize()
                              * (src/decodeMinSum.cpp:480-489) makes of y.  Needs Q <= 6.  out_soft is then fp32. */

/* Length of NGDBFhw's per-frame noise buffer (src/NGDBFhw.cpp:151-152). */
#define LDPC_GPU_HW_QBUF  2648

typedef struct ldpc_gpu_code    ldpc_gpu_code;
typedef struct ldpc_gpu_decoder ldpc_gpu_decoder;

/* Replaces the -D macro set plus the mutable parameter globals
 * (src/decodeGDBF.cpp:48-56, src/decodeMinSum.cpp:35, src/NGDBFhw.cpp:48-57).
 * Field names are the reference's variable names. */
typedef struct ldpc_gpu_decoder_cfg {
    int32_t  kind;            /* LDPC_GPU_KIND_*                                      */
    uint32_t flags;           /* LDPC_GPU_F_*                                         */
    int32_t  precision;       /* LDPC_GPU_PREC_*                                      */
    int32_t  num_iterations;  /* T                                                    */
    double   Ymax;            /* clip / quantiser range                               */
    int32_t  Q;               /* MS/DDBMP quantiser bits: Nq = 2^Q (decodeMinSum.cpp:125) */
    int32_t  NQ;              /* GDBF quantiser bits (decodeGDBF.cpp:56); NGDBFhw: fixed 5 */
    double   alpha;           /* normalised-MS divisor, or GDBF syndrome weight       */
    double   delta;           /* offset-MS offset                                     */
    double   theta;           /* GDBF flip threshold                                  */
    double   lambda;          /* GDBF threshold adaptation                            */
    double   noiseScale;      /* perturbation sigma = channel sigma * noiseScale      */
    int32_t  windowsize;      /* output smoothing window                              */
    int32_t  maxphase;        /* RNGDBF maxphase / NGDBFhw maxPhases                  */
    int32_t  Tswitch;         /* mode-switching start (decodeGDBF.cpp:51), default 0  */
    int32_t  channel_mode;    /* LDPC_GPU_CHANNEL_*: noise generator of ldpc_gpu_simulate / _channel_dump / _redecode_stats */
    double   w;               /* NGDBFhw syndrome weight (NGDBFhw.cpp:50)             */
    double   theta0;          /* NGDBFhw threshold before quantisation (:57)          */
    double   MAXLLR;          /* BP message clip (decodeBP.cpp:58), default 20        */
} ldpc_gpu_decoder_cfg;

/* Channel operating point: N0 = 10^(-snr_db/10)/R, sigma = sqrt(N0/2)
 * (src/decodeMinSum.cpp:146-147). */
typedef struct ldpc_gpu_channel {
    double snr_db;   /* Eb/N0 in dB */
    double R;        /* code rate   */
} ldpc_gpu_channel;

/* The accumulators every reference main() keeps (src/decodeMinSum.cpp:166-173,
 * src/decodeGDBF.cpp:195-198, src/RNGDBF.cpp:193).  Histograms are optional
 * caller-owned HOST arrays; NULL skips them. */
typedef struct ldpc_gpu_counters {
    int64_t errors;
    int64_t uncodedErrors;
    int64_t totalBits;
    int64_t totalWords;
    int64_t wordErrors;
    int64_t totalIterations;
    int64_t smoothingUsed;
    int64_t undetectedWords;        /* e>0 with all checks satisfied ("All checks satisfied.", decodeGDBF.cpp:383-387) */
    int64_t *error_weight_hist;     /* [N]              error_weight_hist[e-1]++          */
    int64_t *iter_hist;             /* [ldpc_gpu_iter_hist_len(cfg)] iterations used per frame (itdist source, NGDBFhw.cpp:420-421):
                                       T+1 entries, or T*maxphase+1 for the GDBF family with LDPC_GPU_F_REDECODE and maxphase > 1
                                       (RNGDBF.cpp:394 adds `it` per phase) */
    int64_t *phase_hist;            /* [maxphase]       phase_hist[phase-1]++ (RNGDBF.cpp:403) */
} ldpc_gpu_counters;

/* One batch of caller-supplied frames for the parity entry. */
typedef struct ldpc_gpu_batch {
    int64_t      n_frames;
    int32_t      mem;           /* LDPC_GPU_MEM_*: where ALL pointers below live        */
    int32_t      y_dtype;       /* LDPC_GPU_DT_*: element type of y and out_soft (F16 in -> F32 soft out) */
    const void  *y;             /* [F][N] raw channel samples y = x(1+sigma n), before clip/quantise */
    const double*noise;         /* optional raw RNG outputs consumed in reference order:
                                   GDBF addNoise: rann() values, [F][noise_rows][N], one row per executed
                                   flip step (phases concatenated); uniformNoise / quantizeProbabilities:
                                   ranu() values, same layout; NGDBFhw: rann() values [F][LDPC_GPU_HW_QBUF] */
    int64_t      noise_rows;    /* rows per frame in `noise` (GDBF); ignored otherwise  */
    const uint8_t *codeword;    /* optional [F][N] bytes 0/1 (bit 1 <-> x=-1); NULL = all-zero word */
    const int32_t *qpointer0;   /* optional [F] NGDBFhw noise-window start per frame (:356-358); NULL = 0 */
    uint8_t     *out_bits;      /* optional [F][ceil(N/8)] hard decisions, bit (i%8) of byte i/8; 1 <-> d=-1 */
    int32_t     *out_iters;     /* optional [F] `it` as the reference accounts it        */
    void        *out_soft;      /* optional [F][N] final a-posteriori sum (MS/BP/DDBMP), y_dtype */
    int32_t     *out_errors;    /* optional [F] Hamming distance to the codeword         */
    uint8_t     *out_flags;     /* optional [F] bit0 satisfied, bit1 smoothing applied, bits 4.. phases used */
} ldpc_gpu_batch;

/* Arguments of the throughput entry (on-device Philox channel). */
typedef struct ldpc_gpu_sim_args {
    uint64_t seed;              /* Philox key                                            */
    int64_t  frame_begin;       /* first global frame id of this call                    */
    int64_t  n_frames;          /* frames to run (upper bound when a stop rule is set)   */
    int64_t  stop_errors;       /* reference stop rule: run while errors < stop_errors || */
    int64_t  stop_word_errors;  /*   wordErrors < stop_word_errors; both 0 = fixed count  */
    int64_t  poll_frames;       /* frames per launch between stop-rule polls (0 = default) */
} ldpc_gpu_sim_args;

/* ---- library ------------------------------------------------------------ */
int  ldpc_gpu_version(void);
const char *ldpc_gpu_last_error(void);
/* Select the CUDA devices this process drives (one per rank in the usual
 * one-process-per-GPU layout).  n = 0 selects device 0. */
int  ldpc_gpu_init(const int *device_ordinals, int n);
int  ldpc_gpu_shutdown(void);
int  ldpc_gpu_device_count(void);

/* ---- parity-check matrix ------------------------------------------------ */
/* Takes exactly the arrays loadFile() produces (inc/alist.h:21-36): 1-based
 * indices, rows zero-padded to biggest_num_n / biggest_num_m. */
int  ldpc_gpu_code_create(int N, int M, int biggest_num_n, int biggest_num_m,
                          const int *num_nlist, const int *num_mlist,
                          const int *nlist_flat, const int *mlist_flat,
                          ldpc_gpu_code **out);
/* Host alist parser (padded or unpadded rows, src/alist.cpp:22-95 semantics). */
int  ldpc_gpu_code_load_alist(const char *path, ldpc_gpu_code **out);
/* The SystemC trees store H transposed (first header number = checks; SystemC/NGDBF/src/ldpcsim.cpp:107-110 reads
 * p.N = alist.M, p.M = alist.N and walks `mlist` per symbol): same parser, roles of the two lists swapped. */
int  ldpc_gpu_code_load_alist_transposed(const char *path, ldpc_gpu_code **out);
int  ldpc_gpu_code_dims(const ldpc_gpu_code *code, int *N, int *M, int *E, int *dv_max, int *dc_max);
int  ldpc_gpu_code_destroy(ldpc_gpu_code *code);
/* SURVEY.md 8(f) N1 -- codewords for codes that ship without a data.enc (the reference lists
 * codes/802_3/data_c.enc as missing, /.MISSING_LARGE_BLOBS:1): H is brought to reduced row-echelon
 * form over GF(2) on the host (redundant rows allowed), the non-pivot columns carry Philox-random
 * information bits.  bits01 is HOST [n][N] bytes 0/1, the layout ldpc_gpu_decoder_set_codewords and
 * ldpc_gpu_batch.codeword take.  *rank (optional) receives rank(H).  Dense elimination: codes with
 * M*N above 2^28 bits return LDPC_GPU_ERR_UNSUPPORTED. */
int  ldpc_gpu_code_random_codewords(ldpc_gpu_code *code, uint64_t seed, int64_t n, uint8_t *bits01, int *rank);

/* ---- decoder ------------------------------------------------------------ */
int  ldpc_gpu_decoder_cfg_default(int kind, ldpc_gpu_decoder_cfg *cfg);
/* Number of entries a caller-owned ldpc_gpu_counters.iter_hist must hold for this configuration. */
int  ldpc_gpu_iter_hist_len(const ldpc_gpu_decoder_cfg *cfg);
int  ldpc_gpu_decoder_create(const ldpc_gpu_code *code, const ldpc_gpu_decoder_cfg *cfg, int device,
                             ldpc_gpu_decoder **out);
int  ldpc_gpu_decoder_destroy(ldpc_gpu_decoder *dec);
/* Codeword table used by ldpc_gpu_simulate: frame f sends row f % n (the
 * reference reads data.enc cyclically, src/decodeMinSum.cpp:195-211).
 * bits01 is HOST [n][N] bytes 0/1; n = 0 restores the all-zero word. */
int  ldpc_gpu_decoder_set_codewords(ldpc_gpu_decoder *dec, const uint8_t *bits01, int64_t n);

/* Parity entry: decode caller-supplied samples exactly like the reference
 * loop body; optionally accumulate the reference's counters. */
int  ldpc_gpu_decode_batch(ldpc_gpu_decoder *dec, const ldpc_gpu_channel *ch,
                           const ldpc_gpu_batch *batch, ldpc_gpu_counters *counters /* optional */);

/* Throughput entry: Philox4x32-10 channel keyed by (seed, frame id), decode,
 * count.  Counters are ADDED to *counters (zero them first). */
int  ldpc_gpu_simulate(ldpc_gpu_decoder *dec, const ldpc_gpu_channel *ch,
                       const ldpc_gpu_sim_args *args, ldpc_gpu_counters *counters);

/* Re-decode statistics (replaces src/redecodeStatistics.cpp:262-395, the GSL-free program of
 * the Makefile goal `redecodeStatistics`): every frame of [frame_begin, frame_begin + n_frames) is decoded
 * n_redecodes times from the same Philox channel samples, each time with fresh decoder noise (GDBF family with
 * addNoise / quantizeProbabilities, NGDBFhw), every decode run to its own end as the reference does.
 * outcomes: HOST [n_frames][n_redecodes] error weights (0 = decoded), the rows the reference appends to its log.
 * counters (optional) accumulate over all n_frames * n_redecodes decodes. */
int  ldpc_gpu_redecode_stats(ldpc_gpu_decoder *dec, const ldpc_gpu_channel *ch, const ldpc_gpu_sim_args *args,
                             int32_t n_redecodes, int32_t *outcomes, ldpc_gpu_counters *counters);

/* Replay of ONE frame of the throughput entry, addressed by (seed, frame_id), with a per-iteration trace (SURVEY.md 8(f) N2;
 * replaces the trace files of src/replayGDBF.cpp:312-314,368-373 and NGDBFhw's LOG_PROCESSING dump, src/NGDBFhw.cpp:304-335):
 * row t (t = 0 .. *n_rows - 1) = the hard decisions after iteration t + 1 (trace_d, HOST [max_rows][ceil(N/8)], packed like
 * out_bits) and the syndromes that iteration started from, i.e. of the decisions after iteration t, row 0: of the channel's
 * hard decisions (trace_syn, optional HOST [max_rows][ceil(M/8)], bit 1 = check unsatisfied).  *n_rows = iterations executed
 * (the bit-flipping decoders stop at the first all-satisfied syndrome; rows beyond max_rows are counted, not stored).
 * Output smoothing and re-decoding phases are left out of the trace, as in the reference's trace files.  O(T) small launches. */
int  ldpc_gpu_replay_frame(ldpc_gpu_decoder *dec, const ldpc_gpu_channel *ch, uint64_t seed, int64_t frame_id, int32_t max_rows,
                           uint8_t *trace_d, uint8_t *trace_syn, int32_t *n_rows, int32_t *final_errors);

/* The exact samples ldpc_gpu_simulate feeds its decoder for frames
 * [frame_begin, frame_begin+n_frames): y is HOST [F][N] doubles.  noise, if
 * non-NULL, receives the decoder-side raw RNG outputs in ldpc_gpu_batch layout
 * ([F][noise_rows][N], or [F][LDPC_GPU_HW_QBUF] for NGDBFhw). */
int  ldpc_gpu_channel_dump(ldpc_gpu_decoder *dec, const ldpc_gpu_channel *ch, uint64_t seed,
                           int64_t frame_begin, int64_t n_frames,
                           double *y, double *noise, int64_t noise_rows);

/* Raw Philox4x32-10 block, for known-answer tests. */
int  ldpc_gpu_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Timing of the last decode/simulate call on this decoder: kernel time from
 * CUDA events on the launching stream, and the number of kernel launches. */
int  ldpc_gpu_last_timing(const ldpc_gpu_decoder *dec, double *kernel_ms, int64_t *launches);
/* LDPC_GPU_PREC_F16X2 on an exact lattice (csrc/ldpc_ms_x2.cuh, csrc/ldpc_ms_tileh.cuh): *exact_lattice = 1 when the decoder runs
 * one of the binary16 kernels whose results are the reference's bit for bit, and *redo_frames = frames (since creation) it could
 * not vouch for and handed to the fp64 instantiation.  Both optional. */
int  ldpc_gpu_decoder_stats(const ldpc_gpu_decoder *dec, int64_t *redo_frames, int32_t *exact_lattice);
/* Launch geometry the library chose (for DESIGN/roofline reporting). */
int  ldpc_gpu_decoder_geometry(const ldpc_gpu_decoder *dec, int *grid, int *block, int *smem_bytes,
                               int *frames_per_cta);

/* ---- multi-GPU ---------------------------------------------------------- */
/* The path's only collective: sum the counter struct (and any non-NULL
 * histograms, lengths given) over all ranks.  The communicator is NCCL,
 * bootstrapped from a 128-byte unique id the caller distributes. */
int  ldpc_gpu_comm_unique_id(uint8_t id[128]);
int  ldpc_gpu_comm_init(const uint8_t id[128], int rank, int nranks, int device);
int  ldpc_gpu_comm_destroy(void);
/* Lengths are taken from the decoder configuration the counters were filled under:
 * error_weight_hist [N], iter_hist [ldpc_gpu_iter_hist_len(cfg)], phase_hist [max(1, cfg->maxphase)]. */
int  ldpc_gpu_allreduce_counters(ldpc_gpu_counters *counters, int N, const ldpc_gpu_decoder_cfg *cfg);

/* ---- non-binary GF(q) codes (BASELINE.json configs[4]) --------------------
 * The reference's SystemC/NB-LDPC tree does not compile and decodes nothing (SURVEY.md "Read this first" 6): PARITY UNPINNED.
 * Kept from it: the code format (SystemC/NB-LDPC/src/alist.cpp:23-56,97-124: header `N M q`, then the binary alist layout with
 * every index followed by its GF value, `0 0` padding) and the symbol <-> bit mapping (bit b of the integer value, least
 * significant first, inc/nodes.h:104-108).  The decoder is textbook min-max (csrc/ldpc_nb_kernel.cuh), q in {2, 4, 8, 16},
 * GF(2^m) in the polynomial basis modulo x^2+x+1, x^3+x+1, x^4+x+1; stop at the first all-zero syndrome or after T iterations.
 * Samples: y [F][N*m] doubles, bit b of symbol i at [i*m + b], BPSK bit 0 -> +1.  All-zero codeword in ldpc_gpu_nb_simulate
 * (Philox channel keyed by (seed, frame id) as everywhere else).  Counters: errors = bit errors, smoothingUsed = symbol errors. */
typedef struct ldpc_gpu_nb_code ldpc_gpu_nb_code;
typedef struct ldpc_gpu_nb_decoder ldpc_gpu_nb_decoder;
int  ldpc_gpu_nb_code_create(int N, int M, int q, int biggest_num_n, int biggest_num_m, const int *num_nlist, const int *num_mlist,
                             const int *nlist_flat, const int *nvals_flat, const int *mlist_flat, const int *mvals_flat,
                             ldpc_gpu_nb_code **out);
int  ldpc_gpu_nb_code_load_alist(const char *path, ldpc_gpu_nb_code **out);
int  ldpc_gpu_nb_code_dims(const ldpc_gpu_nb_code *code, int *N, int *M, int *q, int *E);
int  ldpc_gpu_nb_code_destroy(ldpc_gpu_nb_code *code);
int  ldpc_gpu_nb_decoder_create(const ldpc_gpu_nb_code *code, int num_iterations, int device, ldpc_gpu_nb_decoder **out);
int  ldpc_gpu_nb_decoder_destroy(ldpc_gpu_nb_decoder *dec);
int  ldpc_gpu_nb_decode_batch(ldpc_gpu_nb_decoder *dec, const ldpc_gpu_channel *ch, int64_t n_frames, const double *y,
                              uint8_t *out_symbols, int32_t *out_iters, ldpc_gpu_counters *counters);
int  ldpc_gpu_nb_simulate(ldpc_gpu_nb_decoder *dec, const ldpc_gpu_channel *ch, uint64_t seed, int64_t frame_begin, int64_t n_frames,
                          ldpc_gpu_counters *counters, double *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_GPU_H */
