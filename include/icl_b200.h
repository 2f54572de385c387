/* icl_b200.h -- C-ABI of libicl_b200.so: the B200 (sm_100a) implementation of the neural hot path of
 * cmcervantes/ImageCaptionLearn_py.
 *
 * The reference has no FFI: its seam is the Python function layer of nn_utils/core.py in front of
 * `sess.run(op, feed_dict)` (nn_utils/core.py:625).  Each entry point below replaces one piece of that seam;
 * the Python shim `imagecaptionlearn_py_b200/core.py` binds them with ctypes and re-exposes the reference's
 * own function names (setup_bidirectional_lstm, setup_core_architecture, add_train_op, run_op,
 * get_pred_scores_mcc).  Plain pointers and sizes only -- no torch types.
 *
 * Conventions: every function returns 0 on success, non-zero on error (text via icl_last_error()); nothing
 * throws across the boundary.  The caller owns every input/output buffer; the library owns parameters,
 * optimizer state and workspaces inside the handle.  One handle per device per process; calls on a handle are
 * not re-entrant.  There is NO CPU fallback: icl_create fails if no sm_100 device is present.
 */
#ifndef ICL_B200_H
#define ICL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICL_MAX_HEADS 8
#define ICL_MAX_LAYERS 8
#define ICL_N_INDEX 12

/* task / encoding / activation enums follow the strings of nn_utils/core.py:452-453,178-185 */
enum { ICL_TASK_NONVIS = 0, ICL_TASK_CARD = 1, ICL_TASK_REL_INTRA = 2, ICL_TASK_REL_CROSS = 3, ICL_TASK_AFFINITY = 4 };
enum { ICL_ENC_FIRST_LAST_MENTION = 0, ICL_ENC_FIRST_LAST_SENTENCE = 1 };
enum { ICL_ACT_NONE = 0, ICL_ACT_SIGMOID = 1, ICL_ACT_TANH = 2, ICL_ACT_RELU = 3, ICL_ACT_LEAKY_RELU = 4 };
enum { ICL_F32 = 0, ICL_F64 = 1, ICL_I32 = 2, ICL_I64 = 3 };
/* index matrices, in the order of nn_utils/data.py:409-413 */
enum { ICL_FIRST_I_BW = 0, ICL_FIRST_I_FW, ICL_LAST_I_FW, ICL_LAST_I_BW, ICL_SENT_LAST_I_FW, ICL_SENT_FIRST_I_BW,
       ICL_FIRST_J_BW, ICL_LAST_J_FW, ICL_FIRST_J_FW, ICL_LAST_J_BW, ICL_SENT_LAST_J_FW, ICL_SENT_FIRST_J_BW };
/* what icl_run computes -- the `op` argument of run_op (nn_utils/core.py:517) */
enum { ICL_OP_PREDICT = 0,     /* predicted_proba / pred (+ loss, accuracy when labels are given) */
       ICL_OP_GRADS = 1,       /* forward + backward; gradients left in the flat gradient buffer (DP: all-reduce it) */
       ICL_OP_TRAIN = 2 };     /* train_op: forward + backward + clip_by_global_norm + Adam */
enum { ICL_GEMM_TCGEN05_TF32 = 0, ICL_GEMM_SIMT_FP32 = 1 };

/* one classification head = one setup_core_architecture() call (nn_utils/core.py:443-514) */
typedef struct icl_head_config {
  int32_t task, encoding, batch_size, n_classes;
  int32_t n_feats;          /* n_mention_feats */
  int32_t box_width;        /* box_embedding_width (affinity) or 0 */
  int32_t n_box_feats;      /* n_box_feats or 0 */
  int32_t n_hidden;         /* number of hidden layers = len(get_widths(start, depth)) */
  int32_t widths[ICL_MAX_LAYERS];
  int32_t activation, weighted_classes;
  char scope[32];           /* variable-scope prefix ("" single-task; task name in icl_multitask_lstm.py:50-82) */
} icl_head_config;

/* setup_bidirectional_lstm (core.py:271) + heads + add_train_op (core.py:74) */
typedef struct icl_config {
  int32_t embed_width, lstm_hidden, data_norm;
  int32_t max_seqs, max_seq_len;         /* capacity: sequences per call, padded length T */
  int32_t n_heads;
  icl_head_config heads[ICL_MAX_HEADS];
  float learn_rate, adam_epsilon, clip_norm;   /* clip_norm <= 0: no clipping (clip_norm=None) */
  float beta1, beta2;
  int32_t device, gemm_mode;
} icl_config;

/* the per-head part of the batch_tensors dict (nn_utils/data.py:349-528) */
typedef struct icl_head_batch {
  const void* idx[ICL_N_INDEX];          /* [B,3] rows [dir,sent,word]; NULL when unused by the head */
  int32_t idx_dtype;
  const void* feats;  int32_t feats_dtype;      /* m_feats / ij_feats [B,F] */
  const void* box;    int32_t box_dtype;        /* box_embeddings [B,box_width] or NULL */
  const void* bfeats; int32_t bfeats_dtype;     /* b_feats or NULL */
  const void* labels; int32_t labels_dtype;     /* one-hot [B,C] or NULL (include_labels=False) */
  int32_t sent_offset;                           /* added to the `sent` column (multi-head shared encoder pass) */
  const int32_t* box_rows;                       /* or (box == NULL): [B] rows of the device-resident box table (icl_set_box_table) */
  int32_t inactive;                              /* 1: this head is not fed in this call (TF evaluates only the fetched
                                                    task's subgraph, icl_multitask_lstm.py:327-334); its outputs are NaN */
  /* multi-head calls (icl_multitask_lstm.py:213-358): every task's load_batch builds its OWN padded 'sentences' tensor.  When
     icl_batch.sentences and token_rows are both NULL the library reads each fed head's tensor in place -- its sequences are
     numbered [sent_offset, sent_offset + n_seqs) -- instead of making the caller concatenate them (184 MB per step for C5). */
  const void* sentences;                         /* padded [n_seqs, padded_T, E], dtype = icl_batch.sent_dtype */
  int32_t n_seqs, padded_T;
} icl_head_batch;

typedef struct icl_batch {
  const void* sentences;  int32_t sent_dtype;   /* padded [S,T,E] (sent_packed=0) or packed [sum(len),E] (=1) */
  int32_t sent_packed;
  const int32_t* token_rows;                    /* or (sentences == NULL): packed caption-major [sum(len)] row numbers into the
                                                   device-resident token table (icl_set_token_table) -- 4 bytes per token on the wire */
  const void* seq_lengths; int32_t len_dtype;   /* [S] */
  int32_t n_seqs, padded_T;
  int64_t seq_gid_offset, ex_gid_offset;        /* global ids of row 0 (dropout RNG is keyed on global ids) */
  int32_t n_heads;
  icl_head_batch heads[ICL_MAX_HEADS];
} icl_batch;

typedef struct icl_head_out {
  float* proba;      /* [B,C] or NULL */
  int64_t* pred;     /* [B]   or NULL */
  float loss, accuracy;
} icl_head_out;

typedef struct icl_model icl_model;

const char* icl_last_error(void);
int icl_version(void);
int icl_create(const icl_config* cfg, icl_model** out);          /* replaces graph construction, core.py:271-514,74-106 */
void icl_destroy(icl_model* m);
/* corpus cache: the embedding rows of every caption token, concatenated, stay resident in HBM;
   nn_utils/data.py:397-403 copies them row by row into a fresh [S,T,300] tensor per batch instead */
int icl_set_token_table(icl_model* m, const float* table_rows_by_E, int64_t n_rows);
int icl_set_box_table(icl_model* m, const float* table_rows_by_W, int64_t n_rows, int32_t box_width);   /* affinity box features, once */
uint32_t icl_crc32c(const void* data, uint64_t n, uint32_t crc0);  /* host: CRC-32C of TF Saver-V2 checkpoint tensors (tf_checkpoint.py) */
int icl_set_stream(icl_model* m, void* cuda_stream);             /* cudaStream_t of the caller (e.g. torch's current) */

/* parameters are named like the TF variables so checkpoints map 1:1 (tf.train.Saver, icl_core_lstm.py:107,155) */
int icl_param_count(icl_model* m);
int icl_param_info(icl_model* m, int i, const char** name, int32_t* rows, int32_t* cols, int64_t* offset);
int icl_get_tensor(icl_model* m, int kind, const char* name, float* host);  /* kind: 0 param, 1 grad, 2 adam m, 3 adam v */
int icl_set_tensor(icl_model* m, int kind, const char* name, const float* host);
int icl_get_step(icl_model* m, int64_t* t);
int icl_set_step(icl_model* m, int64_t t);

/* sess.run(op, feed_dict) (core.py:625): host buffers in, host results out, copies inside */
int icl_run(icl_model* m, int op, const icl_batch* b, float keep_in, float keep, uint64_t seed, icl_head_out* out);
/* split form: stage the batch in HBM once, then run on resident data (bench `value`, DP overlap).
   Wire format of the sentence rows: the valid tokens are packed (and float64 -> float32 converted, run_op's feed conversion,
   core.py:558-561) into pinned memory by a small host thread pool (ICL_HOST_THREADS; default up to 4 per process, the ranks of a node
   split the host cores)
   with non-temporal stores (ICL_PACK_NT=0: memcpy).  In ICL_GEMM_TCGEN05_TF32 mode without data_norm the rows are rounded to fp16
   while they are packed (ICL_WIRE_FP16=0: keep fp32): the device's next step rounds the prepared inputs to 10 mantissa bits for
   the tensor cores anyway, so the copy carries 2 bytes per element instead of 4.  The fp32 validation mode and data_norm models
   always upload fp32. */
int icl_upload(icl_model* m, const icl_batch* b);
int icl_run_resident(icl_model* m, int op, float keep_in, float keep, uint64_t seed);
int icl_fetch(icl_model* m, icl_head_out* out);
/* pipelined training (the reference's loop calls run_op(train_op) per batch and discards the result, icl_core_lstm.py:138-150):
   inputs are double-buffered, icl_upload copies on its own stream while the previous step still computes, and
   icl_train_async returns once the step is enqueued.  `prev` (may be NULL) receives loss / accuracy of the PREVIOUS
   icl_poll_stats / icl_train_async call (NaN when there is none); icl_poll_stats is the read-back half for callers that
   compose upload / run_resident / all-reduce / apply_update themselves (data parallel). */
int icl_train_async(icl_model* m, const icl_batch* b, float keep_in, float keep, uint64_t seed, icl_head_out* prev);
int icl_poll_stats(icl_model* m, icl_head_out* prev);
/* data-parallel: flat fp32 gradient buffer on the device (all-reduce SUM it), then apply clip + Adam */
int icl_grad_buffer(icl_model* m, void** dev_ptr, int64_t* n_floats);
int icl_param_buffer(icl_model* m, void** dev_ptr, int64_t* n_floats);
/* all-reduce overlapped with the backward pass: the buffer is [LSTM | heads]; the heads' gradients are final before the BPTT
   starts.  icl_wait_head_grads makes the given stream wait for them, so a collective on floats [first_head_float, n) enqueued
   there overlaps the BPTT + weight-gradient GEMMs of the compute stream; the LSTM part is reduced after icl_run_resident. */
int icl_grad_split(icl_model* m, int64_t* first_head_float);
int icl_wait_head_grads(icl_model* m, void* cuda_stream);
/* LSTM slice of the gradient: floats [0, first_bw_float) belong to the forward direction, whose weight-gradient GEMM finishes first;
   icl_wait_fw_lstm_grads makes the given stream wait for it (its collective then overlaps the backward direction's GEMM). */
int icl_grad_split_lstm(icl_model* m, int64_t* first_bw_float);
/* Joins work queued on the library's side streams behind the last call (the fp16 repack of the LSTM weights after an update; it
   normally overlaps the next step's input preparation) into the main stream, so that events around ONE step time all of its work. */
int icl_join_side_work(icl_model* m);
/* Data parallel over NVSwitch (no counterpart in the reference, which has no multi-GPU path): the library writes its gradients into a
   buffer the caller allocated in symmetric / multicast-mapped memory, and icl_nvls_allreduce sums that buffer over the ranks inside the
   switch (multimem.ld_reduce + multimem.st, each rank one slice).  The caller provides the cross-rank barriers around it. */
int icl_adopt_grad_buffer(icl_model* m, void* symmetric_buffer, int64_t n_floats);
int icl_nvls_allreduce(icl_model* m, void* multicast_ptr, int32_t rank, int32_t world);
/* Test hook: poisons the fp16 operand rows of the forward recurrence (65504.0); a correct run never reads a row before it is published. */
int icl_debug_poison_recurrence(icl_model* m);
int icl_wait_fw_lstm_grads(icl_model* m, void* cuda_stream);
int icl_apply_update(icl_model* m);
/* Adam state selection: one (m, v, beta-power) set per tf.train.AdamOptimizer instance -- the `alternate` multitask scheme has
   one per task (icl_multitask_lstm.py:387-393).  Slot 0 exists from the start; others are created zeroed on first use.
   Parameters of heads that were not fed (icl_head_batch.inactive) are never touched by an update. */
/* weighted_joint (icl_multitask_lstm.py:248-255): joint = sum_j (sum_t loss_t W[t,j] + b_j) with a trainable 5x5 W kept by the
   caller.  icl_set_loss_weights gives every head its d joint / d loss_t (row sums of W; NULL = all 1); icl_apply_update_ex adds
   the caller's squared gradient norm (of W, b) to clip_by_global_norm and returns the global norm used. */
int icl_set_loss_weights(icl_model* m, const float* w /*[n_heads]*/);
int icl_apply_update_ex(icl_model* m, double extra_sumsq, float* gnorm_out);
int icl_set_optimizer_slot(icl_model* m, int slot);
int icl_optimizer_slots(icl_model* m);
int icl_sync(icl_model* m);

/* test / profiling hooks */
int icl_get_lstm_outputs(icl_model* m, int dir, float* host_STH);        /* pre-dropout outputs, padded [S,T,H] */
int icl_get_batch_input(icl_model* m, int head, float* host_BD);         /* concat input of head [B,D0] */
/* Affinity layer 1 runs factorised when a batch repeats mentions and boxes (z1 = U[mention] + V[box] + b1; the concat it
   replaces: nn_utils/core.py:421-433,439): factorised / distinct mentions / distinct boxes of the resident batch, totals[3] over all
   uploads = {batches factorised, their distinct mentions, their distinct boxes}.  ICL_AFF_FACTOR=0 disables, =2 forces it. */
int icl_head_factor_stats(icl_model* m, int head, int32_t* factorised, int32_t* n_mentions, int32_t* n_boxes, int64_t* totals);
/* host-only (testable without a GPU): groups of byte-identical rows in order of first appearance, as icl_upload finds the distinct box rows */
int icl_group_rows(const void* rows, int64_t n_rows, int64_t row_bytes, int32_t* group_of, int32_t* n_groups);
int icl_get_activation(icl_model* m, int head, int layer, float* host_BW);  /* hidden layer output (post-dropout) [B,w] */
int icl_rec_trace(icl_model* m, int cta, long long* host);               /* bring-up trace of one persistent-kernel CTA */
int icl_debug_mask(icl_model* m, uint64_t seed, uint32_t stream, int64_t first_idx, int64_t n, float keep, float* host);
int icl_gemm(icl_model* m, int mode, int a_mn_major, int b_mn_major, int M, int N, int K,
             const float* A, const float* B, float* C, int splits);     /* C[M,N] = A*B on the device (validation) */
int icl_kernel_launches(icl_model* m, int64_t* n);                       /* launches since create */
int icl_last_step_ms(icl_model* m, float* ms);                           /* device time of last run_resident (NaN unless icl_set_phase_timing is on) */
#define ICL_N_PHASES 8   /* prep, input-projection GEMM, recurrence fwd, heads fwd, heads bwd, BPTT, weight grads, clip+Adam */
int icl_set_phase_timing(icl_model* m, int on);                          /* record the per-phase timing events (off by default: they serialise the stream, ~50 us per step) */
int icl_phase_ms(icl_model* m, float* ms /*[ICL_N_PHASES]*/);            /* per-phase device time of the last run_resident */
int icl_copy_bytes(icl_model* m, int64_t* h2d, int64_t* d2h);            /* bytes copied by the last upload / fetch */
/* the host conversion icl_upload applies to sentence rows (run_op's float64 -> float32 feed conversion, core.py:558-561), exposed
   so that it is testable without a GPU: n elements of src (ICL_F32 / ICL_F64) -> float32 (half = 0) or fp16 bits (half = 1) at dst */
int icl_pack_rows(const void* src, int32_t src_dtype, int64_t n, void* dst, int32_t half);
int icl_batch_stats(icl_model* m, int64_t* n_seqs, int64_t* n_tokens, int32_t* t_max);   /* of the resident batch */

#ifdef __cplusplus
}
#endif
#endif
