/* ldpc_oracle.h -- CPU restatement of the reference decode loop (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may link or call this.  The product (ldpcsimulation_b200/csrc) never does.
 *
 * Parity status: PINNED.  Every decoder here is checked against the reference's own
 * object code (oracle/_ref, built by oracle/build_ref.py from /root/reference) in
 * tests/test_oracle_vs_ref.py, and against the committed outputs of that object code in
 * tests/golden/ (the reference ships no golden vectors of its own).
 *
 * Shares the configuration / batch / counter structs of include/ldpc_gpu.h so that a
 * parity test feeds the oracle and the CUDA path the very same argument blocks.
 */
#ifndef LDPC_ORACLE_H
#define LDPC_ORACLE_H

#include <stdint.h>
#include "../include/ldpc_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_code {
    int N, M, E;
    int dv_max, dc_max;
    int *col_deg;    /* [N]                       num_nlist                                  */
    int *row_deg;    /* [M]                       num_mlist                                  */
    int *nlist;      /* [N*dv_max] 0-based check of slot s of variable i, -1 padded          */
    int *mlist;      /* [M*dc_max] 0-based variable of slot k of check j, -1 padded          */
    int *vn_slot;    /* [M*dc_max] for edge (j,k): slot of check j inside variable's nlist   */
    int *cn_slot;    /* [N*dv_max] for edge (i,s): slot of variable i inside check's mlist   */
} oracle_code;

oracle_code *oracle_code_create(int N, int M, int biggest_num_n, int biggest_num_m,
                                const int *num_nlist, const int *num_mlist,
                                const int *nlist_flat, const int *mlist_flat);
oracle_code *oracle_code_load_alist(const char *path);
void         oracle_code_free(oracle_code *c);
const char  *oracle_last_error(void);

/* Decode caller-supplied frames exactly as the reference loop body would.
 * All pointers in `batch` are HOST pointers; y must be LDPC_GPU_DT_F64. */
int oracle_decode_batch(const oracle_code *code, const ldpc_gpu_decoder_cfg *cfg,
                        const ldpc_gpu_channel *ch, const ldpc_gpu_batch *batch,
                        ldpc_gpu_counters *counters);

/* The framework's own counter-based channel (not reference behaviour; the reference
 * uses libc random()).  Restated here independently of the CUDA code so the two can
 * be compared bit for bit. */
void oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_normal4(uint64_t seed, uint64_t frame, uint32_t block, uint32_t row, uint32_t stream, float out[4]);
int  oracle_channel_dump(const oracle_code *code, const ldpc_gpu_decoder_cfg *cfg, const ldpc_gpu_channel *ch,
                         uint64_t seed, int64_t frame_begin, int64_t n_frames,
                         const uint8_t *codewords, int64_t n_codewords,
                         double *y, double *noise, int64_t noise_rows);
int  oracle_simulate(const oracle_code *code, const ldpc_gpu_decoder_cfg *cfg, const ldpc_gpu_channel *ch,
                     const ldpc_gpu_sim_args *args, const uint8_t *codewords, int64_t n_codewords,
                     ldpc_gpu_counters *counters);

/* Scalar helpers exposed for unit tests (reference line in ldpc_oracle.c). */
double oracle_quantize_ms(double x, double Ymax, double Nq);
double oracle_quantize_gdbf(double x, double Ymax, int NQ);
int    oracle_hw_pack(double ymod, double Ymax, double w);
int    oracle_hw_unpack(int code);

#ifdef __cplusplus
}
#endif
#endif
