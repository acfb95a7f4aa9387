/*
 * qamrecon.h -- C ABI of libqamrecon.so, the B200 (sm_100a) implementation of the
 * reverse-reconciliation hot path of moriglia/qam-reconciliation.
 *
 * The reference has no FFI: its boundary is the Python API of its Cython extension
 * classes (qamreconciliation/__init__.py:1-4).  Each entry point below names the
 * reference method it stands behind (file:line relative to the reference checkout);
 * the Python classes in qam-reconciliation_b200/qamreconciliation/ bind these through
 * ctypes with the reference's names and argument order.
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer, `h_` = host pointer.
 *   - every array is caller-allocated; the library never frees caller memory.
 *   - batched arrays are row-major [frames][per-frame length].
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); work is
 *     enqueued on it and NOT synchronised unless stated.
 *   - return value: QR_OK or an error code; qr_last_error() gives the message of the
 *     last failure on the calling thread.
 *   - dtype codes: QR_F32 / QR_F64 for LLR-like arrays.
 *   - there is no CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef QAMRECON_H
#define QAMRECON_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QR_OK 0
#define QR_ERR_INVALID 1  /* bad argument or size mismatch (Python: ValueError)   */
#define QR_ERR_GRAPH 2    /* malformed Tanner graph            (Python: ValueError)   */
#define QR_ERR_CUDA 3     /* CUDA runtime failure              (Python: RuntimeError) */
#define QR_ERR_NOMEM 4    /* allocation failure                (Python: MemoryError)  */

#define QR_F32 32
#define QR_F64 64

/* demap modes for qr_demap_lappr */
#define QR_DEMAP_EXACT 0      /* replays the reference's 1e-9 bisection (noisemapper.pyx:310-345) */
#define QR_DEMAP_FAST 1       /* safeguarded Newton on F_Y, then the same dyadic cell as the bisection */
#define QR_DEMAP_CORRECTED 2  /* flag bit: divide the k<j exponent by 2*sigma^2 too (fixes noisemapper.pyx:503-507) */
#define QR_DEMAP_F32GRADE 4   /* flag bit, with QR_DEMAP_FAST: LLRs good to float precision (~1e-6 relative): no 2^-30 cell
                                 replay, MUFU-based exp / log / reciprocals.  What the fp32 decoder mode is fed. */

/* decoder schedules (qr_decoder_set_schedule) */
#define QR_SCHED_PERSISTENT 0 /* one cooperative kernel per batch, grid barriers between phases */
#define QR_SCHED_LAUNCH 1     /* one check + one variable kernel launch per iteration */
#define QR_SCHED_FUSED 2      /* check update + variable sums + syndrome test in ONE pass per iteration, two
                                 message buffers, L2-sized lane tiles, tile-pipelined bookkeeping and refills;
                                 needs check degrees <= 8, variable degrees 1..64 and fewer than 2^27 variables
                                 (QR_ERR_INVALID otherwise); same results */
#define QR_SCHED_AUTO 3       /* the default: QR_SCHED_FUSED when every variable has degree 3 and a 32-lane tile of
                                 the code fits L2 (config 2), else QR_SCHED_PERSISTENT, or QR_SCHED_LAUNCH on graphs
                                 of more than 2 M edges */

typedef struct qr_graph qr_graph;
typedef struct qr_decoder qr_decoder;
typedef struct qr_mapper qr_mapper;

int qr_abi_version(void);
const char *qr_last_error(void);
int qr_device_count(int *count);

/* ---------------------------------------------------------------- Tanner graph
 * Replaces Decoder.__cinit__ (decoder.pyx:93-146) and Matrix.__cinit__ (matrix.pyx:21-38):
 * vid/cid are the reference's e_to_v / e_to_c edge arrays (host, int64, length n_edges).
 * O(E) counting sort instead of the reference's O(nodes*E) scan (decoder.pyx:60-89);
 * per-node edge lists keep the reference's ascending-edge-id order (:73-76).
 * Unlike the reference it validates: ids >= 0, every check degree >= 2, no id gaps.
 * device < 0 builds the host tables only (no CUDA call), for tests. */
int qr_graph_create(const int64_t *h_vid, const int64_t *h_cid, int64_t n_edges, int device,
                    qr_graph **out);
/* Matrix.__cinit__ (matrix.pyx:21-38) accepts ANY edge list -- checks of degree 0 or 1, unused check ids, degrees
 * above 64 -- and Matrix.eval_syndrome works on it; so does this variant (ids must still be >= 0).  A graph that
 * does not meet the decoder's requirements is refused by qr_decoder_create and the single-node entry points. */
int qr_graph_create_any(const int64_t *h_vid, const int64_t *h_cid, int64_t n_edges, int device,
                        qr_graph **out);
void qr_graph_destroy(qr_graph *g);
/* *eligible = 1 when QR_SCHED_FUSED can run this graph (what QR_SCHED_AUTO goes by) */
int qr_graph_fused_eligible(const qr_graph *g, int *eligible);
/* Decoder.cnum/vnum/ednum (decoder.pyx:157-172), Matrix.cnum/vnum/ednum (matrix.pxd:24-27) */
int qr_graph_info(const qr_graph *g, int64_t *n_vars, int64_t *n_checks, int64_t *n_edges,
                  int32_t *max_check_degree, int32_t *max_var_degree);
/* host copies of the device tables (any pointer may be NULL):
 *   chk_order[C]   internal check slot -> original check id (slots sorted by degree, then id)
 *   slot_edge[E]   CSR slot -> original edge id (slots grouped by internal check, ascending edge id)
 *   slot_var[E]    CSR slot -> variable id
 *   var_ptr[N+1], var_slot[E]  per variable: CSR slots of its edges in ascending edge id */
int qr_graph_export(const qr_graph *g, int32_t *chk_order, int32_t *slot_edge, int32_t *slot_var,
                    int32_t *var_ptr, int32_t *var_slot);

/* ---------------------------------------------------------------- syndrome
 * Matrix.eval_syndrome (matrix.pyx:55-60), batched: synd[b][c] = XOR of word[b][v] over the
 * check's edges (whole bytes are XOR-ed, as the reference does). */
int qr_eval_syndrome(const qr_graph *g, const uint8_t *d_word, uint8_t *d_synd, int64_t frames,
                     void *stream);
/* Decoder.check_word (decoder.pyx:220-232) / check_lappr (decoder.pyx:260-281), batched:
 * ok[b] = 1 iff every check sees synd XOR parity == 0; lappr bit = (lappr < 0) (decoder.pyx:244). */
int qr_check_word(const qr_graph *g, const uint8_t *d_word, const uint8_t *d_synd, int64_t frames,
                  uint8_t *d_ok, void *stream);
int qr_check_lappr(const qr_graph *g, const void *d_lappr, int dtype, const uint8_t *d_synd,
                   int64_t frames, uint8_t *d_ok, void *stream);
/* utils.count_errors_from_lappr (utils.pyx:27-40), batched over the first `k` bits of each frame:
 * errors[b] = #{i<k : (lappr[b][i] >= 0 ? 0 : 1) != word[b][i]}. */
int qr_count_errors(const void *d_lappr, int dtype, const uint8_t *d_word, int64_t frames,
                    int64_t frame_len, int64_t k, int32_t *d_errors, void *stream);

/* ---------------------------------------------------------------- decoder
 * Decoder._decode / decode (decoder.pyx:391-455), batched over independent frames.
 * precision QR_F64: the reference's box-plus recursion in double, same association order;
 * precision QR_F32: fast mode (exp-domain forward/backward products, float messages).
 * `lanes` = frames resident at once (rounded up to a multiple of 32; 0 = library default: 512, fewer if device
 * memory is short).  A batch uses min(lanes, frames) lanes.  One lane per frame of the largest batch is the fastest
 * setting wherever the workspace fits -- (2 E + 2 N) * sizeof(message) + C bytes per lane with the fused schedule --
 * because no lane is then refilled mid-launch; the Python layer sizes it that way. */
int qr_decoder_create(const qr_graph *g, int precision, int64_t lanes, qr_decoder **out);
void qr_decoder_destroy(qr_decoder *d);
int qr_decoder_set_schedule(qr_decoder *d, int schedule);
/* llr [frames][N] (QR_F32|QR_F64), synd [frames][C]; outputs success[frames] (u8), iters[frames]
 * (i32), post [frames][N] (QR_F32|QR_F64, may be NULL).  Semantics per frame are the reference's:
 * (1,0,input) if the input already satisfies the syndrome; (1,t) with the posteriors of iteration t
 * on success; (0,max_iterations) with the last posteriors otherwise. */
int qr_decode_batch(qr_decoder *d, const void *d_llr, int llr_dtype, const uint8_t *d_synd,
                    int64_t frames, int32_t max_iterations, uint8_t *d_success, int32_t *d_iters,
                    void *d_post, int post_dtype, void *stream);
/* counters of the LAST qr_decode_batch on this handle (synchronises the stream):
 * flooding iterations summed over frames, and schedule steps executed */
int qr_decoder_last_stats(qr_decoder *d, int64_t *frame_iterations, int64_t *steps);

/* Single-node entry points the reference keeps for its unit tests (decoder.pyx:190-217,
 * 301-319, 372-388); device arrays are fp64 and updated in place like the caller's numpy
 * arrays are in the reference.  They run the same device functions as the QR_F64 decoder. */
int qr_process_check_node(const qr_graph *g, int64_t check, const uint8_t *d_synd, double *d_c2v,
                          const double *d_v2c, void *stream);
int qr_process_var_node(const qr_graph *g, int64_t var, const double *d_llr, const double *d_c2v,
                        double *d_v2c, double *d_post, void *stream);
int qr_check_synd_node(const qr_graph *g, int64_t check, const uint8_t *d_word,
                       const uint8_t *d_synd, uint8_t *d_ok, void *stream);

/* ---------------------------------------------------------------- noise mapper
 * NoiseMapper.__cinit__ tables (noisemapper.pyx:103-236) for a PAM alphabet
 * (alphabet.pyx:35-76): constellation[order], thresholds[order+1], probabilities[order],
 * sign_config[order] (host arrays).  Tables are computed on the device with the same erf
 * the kernels use. */
int qr_mapper_create(int bits_per_symbol, const double *h_constellation, const double *h_thresholds,
                     const double *h_probabilities, double noise_var, const uint8_t *h_sign_config,
                     int device, qr_mapper **out);
void qr_mapper_destroy(qr_mapper *m);
/* Symbol / region indices handed to the entry points below must lie in [0, order); the reference's bounds-checked
 * Cython raises IndexError otherwise.  Kernels cannot raise: they count out-of-range indices (and compute on index
 * 0 instead of reading out of bounds).  This call synchronises `stream`, returns the count since the last call
 * and resets it; the Python classes raise IndexError when it is non-zero. */
int qr_mapper_index_errors(qr_mapper *m, int64_t *count, void *stream);
/* host copies: F_Y_thresholds[order+1], delta_F_Y[order], fwrd[order*order], back[order*order],
 * bare_llr_table[order*bps], inf_erf_table[order*order] (noisemapper.pxd:19-35); NULL to skip */
int qr_mapper_tables(const qr_mapper *m, double *F_Y_thresholds, double *delta_F_Y, double *fwrd,
                     double *back, double *bare_llr_table, double *inf_erf_table);

/* NoiseMapper.hard_decide_index (noisemapper.pyx:349-359) */
int qr_hard_decide_index(const qr_mapper *m, const double *d_y, int64_t n, int64_t *d_index,
                         void *stream);
/* PAMAlphabet.demap_symbols_to_bits (alphabet.pyx:98-107): bits[i*bps+k] */
int qr_symbols_to_bits(const qr_mapper *m, const int64_t *d_index, int64_t n, uint8_t *d_bits,
                       void *stream);
/* NoiseMapper.map_noise (noisemapper.pyx:373-388) */
int qr_map_noise(const qr_mapper *m, const double *d_y, const int64_t *d_index, int64_t n,
                 double *d_n_hat, void *stream);
/* the three above fused in one pass over y (any output may be NULL) */
int qr_front_end(const qr_mapper *m, const double *d_y, int64_t n, int64_t *d_index,
                 double *d_n_hat, uint8_t *d_bits, void *stream);
/* NoiseMapper.demap_lappr_array (noisemapper.pyx:544-559), output scaled by alpha
 * (sims/reconciliation.pyx:144-145); llr [n*bps] as QR_F32 or QR_F64 */
int qr_demap_lappr(const qr_mapper *m, const double *d_n_hat, const int64_t *d_tx_index, int64_t n,
                   int mode, double alpha, void *d_llr, int llr_dtype, void *stream);
/* NoiseMapper.g_inv_search (noisemapper.pyx:310-345) for n values and one region index each */
int qr_g_inv_search(const qr_mapper *m, const double *d_n_hat, const int64_t *d_region, int64_t n,
                    int mode, double *d_y_hat, void *stream);
/* NoiseMapper.bare_llr (noisemapper.pyx:423-432) */
int qr_bare_llr(const qr_mapper *m, const int64_t *d_tx_index, int64_t n, void *d_llr,
                int llr_dtype, void *stream);
/* direct-reconciliation LLR (sims/reconciliation.pyx:25-72); two_variance = Es*10^(-snr/10) */
int qr_direct_llr(const qr_mapper *m, const double *d_y, int64_t n, double two_variance,
                  void *d_llr, int llr_dtype, void *stream);

/* ---------------------------------------------------------------- rest of the NoiseMapper surface
 * Sign rule of g / map_noise / g_inv / demap_noise / the two formulations below: NoiseMapper uses its
 * sign_config (the default after qr_mapper_create); NoiseMapperFlipSign (noisemapper.pyx:775-797) uses 1 on the
 * lower half of the alphabet, NoiseMapperAntiFlipSign (:798-816) 1 on the upper half.  g_inv_search and
 * demap_lappr keep the constructor's sign_config, as in the reference (the subclasses do not override them). */
int qr_mapper_set_g_sign(qr_mapper *m, const uint8_t *h_sign_g);
/* the dense grid of noisemapper.pyx:135-144: y_range = numpy.linspace(y_low, y_high, n_points), F_Y on it
 * (NoiseMapper.F_Y, :264-275: uniform weights).  Needed by qr_demap_noise / qr_demap_lappr_variant. */
int qr_mapper_build_grid(qr_mapper *m, double y_low, double y_high, int64_t n_points);
/* NoiseMapper.y_range / F_Y_values (noisemapper.pyx:254-261); host arrays of n_points, NULL to skip */
int qr_mapper_grid(const qr_mapper *m, int64_t *n_points, double *h_y_range, double *h_F_Y);
/* NoiseMapper.F_Y (noisemapper.pyx:264-275) and the module function F_Z (:70-80) */
int qr_F_Y(const qr_mapper *m, const double *d_y, int64_t n, double *d_out, void *stream);
int qr_F_Z(const double *d_z, int64_t n, double mu, double sigma, double *d_out, void *stream);
/* NoiseMapper.demap_noise = g_inv per element (noisemapper.pyx:391-404, :295-307, __interp :47-63) */
int qr_demap_noise(const qr_mapper *m, const double *d_n_hat, const int64_t *d_symb, int64_t n, double *d_y_hat,
                   void *stream);
/* variant 1: demap_lappr_simplified_array (noisemapper.pyx:563-621); 2: demap_lappr_sofisticated_array (:624-766) */
int qr_demap_lappr_variant(const qr_mapper *m, int variant, const double *d_n_hat, const int64_t *d_tx_index,
                           int64_t n, double *d_llr, void *stream);

/* mutual_information.montecarlo_information (mutual_information.pyx:212-300) after its sample draw: for the
 * samples (x_ind[n] Alice's symbols, y[n] Bob's channel outputs, device arrays) ADDS to d_sums[3] the sums over
 * the samples of the three per-sample terms -- log2(P(xhat)/P(xhat|x)) (:257-258), log2 of the I(X;Y) kernel
 * (:262-268), and minus log2 of the I(X,N;Xhat) kernel (:272-290); the caller zeroes d_sums and divides by n
 * (:293-295).  which: bit 0/1/2 selects the term (the reference's `which` array); d_p_Xhat[order] = P_xhat(nm)
 * (:29-39); demap_mode: QR_DEMAP_EXACT or QR_DEMAP_FAST for the one g_inv_search per sample (:280). */
int qr_information_sums(const qr_mapper *m, const double *d_p_Xhat, const int64_t *d_x_ind, const double *d_y,
                        int64_t n, int which, int demap_mode, double *d_sums, void *stream);

/* ---------------------------------------------------------------- whole path, device buffers
 * The chain of sims/reconciliation.pyx:129-153 (mode 0 soft reverse, 1 hard reverse :300-308,
 * 2 soft direct :214-227) for `frames` frames whose channel outputs d_y [frames][S] and Alice's
 * symbols d_tx_index [frames][S] are already in device memory.  Outputs (device; d_post, d_word,
 * d_synd, d_bit_errors may be NULL): success, iters, post [frames][N], word [frames][N] (the bits the
 * syndrome was taken of), synd [frames][C], bit_errors[frames] over the first k_info bits.
 * One call, all kernels enqueued back to back on `stream`; results are identical to calling
 * qr_front_end / qr_eval_syndrome / qr_demap_lappr / qr_decode_batch / qr_count_errors one after the
 * other (intermediate arrays live in scratch owned by the decoder handle). */
int qr_reconcile_device(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                        const double *d_y, const int64_t *d_tx_index, int64_t frames,
                        int32_t max_iterations, int64_t k_info, uint8_t *d_success, int32_t *d_iters,
                        void *d_post, int post_dtype, uint8_t *d_word, uint8_t *d_synd,
                        int32_t *d_bit_errors, void *stream);

/* ---------------------------------------------------------------- whole path, host buffers
 * One reverse-reconciliation pass over `frames` frames held in HOST memory (pinned memory makes
 * the copies asynchronous), the chain of sims/reconciliation.pyx:129-153:
 *   y -> hard decision -> softening metric, bits -> syndrome -> Alice's LLR (from her symbols
 *   h_tx_index) * alpha -> syndrome decoding.
 * reconciliation mode: 0 soft reverse, 1 hard reverse (bare_llr), 2 soft direct.
 * Outputs (host, any may be NULL): success[frames], iters[frames], post[frames][N] (post_dtype),
 * word[frames][N] (the bits the syndrome was taken of), bit_errors[frames] over the first k_info bits.
 * Copies, kernels and the final synchronisation all happen inside the call. */
int qr_reconcile_host(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                      const double *h_y, const int64_t *h_tx_index, int64_t frames,
                      int32_t max_iterations, int64_t k_info, uint8_t *h_success, int32_t *h_iters,
                      void *h_post, int post_dtype, uint8_t *h_word, int32_t *h_bit_errors,
                      void *stream);

/* The same pass over a COMPACT WIRE FORMAT (an extension: the reference's arrays are float64 / int64): channel
 * outputs as float32, Alice's symbols as bytes (alphabets of at most 256 points), and instead of the final LLRs the
 * decoder's hard decisions (decoder.pyx:244: bit = lappr < 0) packed 8 per byte, bit (i & 7) of byte i >> 3 of a
 * frame's row of ceil(N / 8) bytes.  5 bytes per symbol in and N / 8 bytes per frame out instead of 16 and 4 N:
 * what a deployment whose host link is the limit would ship (bench.py reports it as `e2e_compact`).  Widening and
 * packing happen on the device; everything in between is the chain of qr_reconcile_host, unchanged. */
int qr_reconcile_host_compact(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                              const float *h_y32, const uint8_t *h_tx8, int64_t frames, int32_t max_iterations,
                              int64_t k_info, uint8_t *h_success, int32_t *h_iters, uint8_t *h_decisions_packed,
                              int32_t *h_bit_errors, void *stream);

/* The chunk boundaries qr_reconcile_host (compact = 0) / qr_reconcile_host_compact (compact = 1) use for a batch of
 * `frames` frames on a decoder with `lanes` resident frames (host logic only, no device work): cuts[0] = 0 < cuts[1] < ... < cuts[n_cuts - 1] = frames.
 * Chunks flow through upload / kernels / download on three streams; a chunk never exceeds the lanes when a quarter
 * of the batch fits them (no lane is refilled), and the batch ramps up and down (reference types: 64, 192, 576, ...,
 * 320, 64 frames; compact: one piece of at most 256 frames at either end) so that only those copies are exposed.  (The reference has no counterpart: sims/reconciliation.pyx:127-161
 * handles one frame at a time in host memory.) */
int qr_host_chunk_cuts(int64_t frames, int64_t lanes, int compact, int64_t *cuts, int32_t max_cuts, int32_t *n_cuts);

#ifdef __cplusplus
}
#endif
#endif /* QAMRECON_H */
