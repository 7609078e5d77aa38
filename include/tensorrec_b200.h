/*
 * tensorrec_b200.h -- C ABI of the B200-native predict / predict_rank hot path of jfkirk/tensorrec.
 *
 * The reference (pure Python over TensorFlow 1.x, commit 80690737) has no FFI of its own; the boundary a
 * maintainer would bind is the set of TF ops its graph evaluates on this path.  Each entry point below
 * replaces one of those graph nodes (reference file:line given per function; paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the caller owns every buffer;
 *   - matrices are dense row-major; indices are int32; values float32;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous;
 *   - return value: 0 = ok, negative = TRK_ERR_*; trk_last_error() gives the text for this thread;
 *   - nothing here allocates persistent device memory; workspace sizes are queried and caller-provided;
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns TRK_ERR_CUDA.
 */
#ifndef TENSORREC_B200_H_
#define TENSORREC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRK_OK 0
#define TRK_ERR_ARG (-1)         /* bad argument (null pointer, size, alignment, unsupported shape) */
#define TRK_ERR_CUDA (-2)        /* a CUDA runtime / driver call failed */
#define TRK_ERR_UNSUPPORTED (-3) /* shape outside what the fused tensor-core kernel supports */

/* ABI version of this header (major*1000 + minor). */
int trk_version(void);

/* Text of the last error raised on the calling thread ("" if none). */
const char* trk_last_error(void);

/* ------------------------------------------------------------------------------------------------------
 * K1  sparse features -> dense representation (CSR x dense gather-reduce)
 *
 * replaces tf.sparse_tensor_dense_matmul in LinearRepresentationGraph.connect_representation_graph
 * (tensorrec/representation_graphs.py:32-43, the matmul is :40) and, with n_normalize >= 1, the
 * tf.nn.l2_normalize of NormalizedLinearRepresentationGraph (:53-58, :57) and of relative_cosine
 * (tensorrec/recommendation_graphs.py:119-120).
 *
 *   out[r, :] = sum over p in [indptr[r], indptr[r+1])  val[p] * weights[col[p], :]     (fp32, CSR order)
 *   then n_normalize times:  out[r, :] *= rsqrt(max(sum(out[r, :]^2), 1e-12))
 *
 * CSR entries of one row are accumulated sequentially in storage order (duplicates are summed), so the
 * result is run-to-run bit-identical (reference requirement: test/test_tensorrec.py:418-458).
 *
 * Outputs (either may be NULL, not both):
 *   out_f32   [rows, d]            the representation as the reference returns it;
 *   out_split [rows, 2*d_pad] f16  the operand layout of the tensor-core score kernel: columns [0,d_pad) hold
 *                                  hi = fp16(x * 2^e_r), columns [d_pad, 2*d_pad) hold lo = fp16(x * 2^e_r - hi),
 *                                  zero padded from d to d_pad (d_pad a multiple of 64);
 *   out_scale [rows]               2^-e_r, the exact power of two that undoes the per-row scaling
 *                                  (required when out_split is given);
 *   out_norm  [rows] (optional)    |out[r, :]|_2 inflated by 2^-9: an UPPER bound, the factor of the filter's error bound;
 *   stats     [3]    (optional)    zeroed by this call, then stats[0] = max out_norm, stats[1] = max out_scale over the
 *                                  non-zero rows (atomic max on the float bits: order independent); stats[2] is left
 *                                  for trk_pack_item_bias.  The row is still in registers when these are formed: the
 *                                  separate pass of trk_operand_stats over the operand is not needed after K1.
 * ---------------------------------------------------------------------------------------------------- */
int trk_csr_gather_reduce_f32(const int32_t* indptr, const int32_t* col, const float* val,
                              const float* weights, int64_t rows, int32_t n_features, int32_t d,
                              int32_t n_normalize, float* out_f32, void* out_split, int32_t d_pad,
                              float* out_scale, float* out_norm, float* stats, void* stream);

/* Converts an existing dense fp32 representation [rows, d] into the split-fp16 operand + scales
 * (same layout as above).  Used when a representation comes from a user-defined plugin graph. */
int trk_split_f32_to_f16x2(const float* repr, int64_t rows, int32_t d, int32_t n_normalize, void* out_split,
                           int32_t d_pad, float* out_scale, void* stream);

/* L2-normalises rows of a dense fp32 matrix in place: tf.nn.l2_normalize(x, 1)
 * (tensorrec/recommendation_graphs.py:119-120). */
int trk_l2_normalize_rows_f32(float* x, int64_t rows, int32_t d, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * project_biases (tensorrec/recommendation_graphs.py:4-19):
 *   out[r] = sum over the row's entries  val[p] * feature_biases[col[p]]      (fp32, CSR order)
 * ---------------------------------------------------------------------------------------------------- */
int trk_csr_project_biases_f32(const int32_t* indptr, const int32_t* col, const float* val,
                               const float* feature_biases, int64_t rows, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K2 (exact fp32, CUDA cores)  dense prediction for every user x item pair, any shape
 *
 * replaces tf.matmul(user, item, transpose_b=True) of DotProductPredictionGraph.connect_dense_prediction_graph
 * (tensorrec/prediction_graphs.py:49-50) / relative_cosine (tensorrec/recommendation_graphs.py:121, inputs
 * pre-normalised by K1), followed by collapse_mixture_of_tastes without attention (max over tastes,
 * tensorrec/recommendation_graphs.py:107) and bias_prediction_dense (:41):
 *
 *   out[u, i] = max_t ( sum_k user_repr[t, u, k] * item_repr[i, k] )  + user_bias[u] + item_bias[i]
 *
 * user_repr is [n_tastes, n_users, d]; user_bias / item_bias may be NULL (unbiased model).
 * mode: 0 = dot product, 1 = negative euclidean distance (EuclideanSimilarityPredictionGraph,
 * tensorrec/prediction_graphs.py:84-100).
 * ---------------------------------------------------------------------------------------------------- */
int trk_score_f32(const float* user_repr, const float* item_repr, const float* user_bias,
                  const float* item_bias, float* out, int64_t n_users, int64_t n_items, int32_t d,
                  int32_t n_tastes, int32_t mode, void* stream);

/* Attention variant of the taste collapse (tensorrec/recommendation_graphs.py:96-103):
 *   out[u,i] = sum_t softmax_t(att[t,u,i]) * pred[t,u,i]  (+ biases), pred/att = dot products of
 *   user_repr[t] / attention_repr[t] with item_repr. */
int trk_score_attention_f32(const float* user_repr, const float* attention_repr, const float* item_repr,
                            const float* user_bias, const float* item_bias, float* out, int64_t n_users,
                            int64_t n_items, int32_t d, int32_t n_tastes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K3 (full)  rank_predictions (tensorrec/recommendation_graphs.py:73-82): the reference's double
 * tf.nn.top_k(k = n_items) == for every user row
 *     rank[u, i] = 1 + #{j : s[u,j] > s[u,i]} + #{j < i : s[u,j] == s[u,i]}          (int32, 1-based)
 * computed by a per-row sort of (score descending, index ascending) keys.
 * workspace: trk_rank_full_workspace_bytes(n_users, n_items) bytes of device memory.
 * ---------------------------------------------------------------------------------------------------- */
size_t trk_rank_full_workspace_bytes(int64_t n_users, int64_t n_items);
int trk_rank_full(const float* scores, int32_t* ranks, int64_t n_users, int64_t n_items, void* workspace,
                  size_t workspace_bytes, void* stream);
/* order[ranks[i] - 1] = i for ONE row of ranks: the items of that row listed by reference rank (tf.nn.top_k order).
 * trk_rank_full on the 1 x n_items row of item biases followed by this call is the stable descending sort that fixes the
 * filter kernel's processing order (no library sort on the predict path). */
int trk_order_from_ranks(const int32_t* ranks, int64_t n, int32_t* order, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K2+K3 fused (tensor cores, sm_100a)  scores and per-user top-k without materialising [n_users, n_items]
 *
 * replaces the chain  tf.matmul (prediction_graphs.py:49-50)  ->  bias_prediction_dense
 * (recommendation_graphs.py:41)  ->  rank_predictions (recommendation_graphs.py:73-82) restricted to the
 * entries with rank <= k (the only ones tensorrec/eval.py:23,49,68-69 ever reads).
 *
 * Operands are the split-fp16 layout produced by K1 (hi/lo halves, per-row power-of-two scale); the score is
 *   s[u,i] = (hi_u.hi_i + hi_u.lo_i + lo_u.hi_i) * scale_u * scale_i + user_bias[u] + item_bias[i]
 * accumulated in fp32 in tensor memory (three tcgen05.mma passes per k-block; relative error vs an fp32 dot
 * product <= 2^-21 of |u|.|i|, and exact for integer-valued representations).
 *
 * The item axis is cut into n_splits contiguous ranges (parallelism when n_users is small; shards when the
 * item axis is distributed over GPUs).  For each (user, split) the kernel emits the k best candidates ordered by
 * (score descending, item id ascending):
 *   cand_score [n_users, n_splits, k] f32,  cand_item [n_users, n_splits, k] i32 (GLOBAL ids = local + item_id_offset;
 *   unused slots: score = -inf, id = INT32_MAX).
 * item_meta [n_items_padded256, 2] f32 = {item scale, item bias} per item, rows beyond n_items = {0, -inf}
 * (build with trk_pack_item_meta).  user_bias may be NULL.
 * Constraints: d_pad in {64, 128}; 1 <= k <= trk_score_topk_max_k(d_pad).
 * n_users_live (device int32, may be NULL): only the first *n_users_live user rows hold work -- user blocks beyond
 * them are skipped on the device.  This is how the rows the filter's certificate rejects are re-scored without a
 * host round trip (their count exists only on the device, see trk_select_flagged_rows).
 * ---------------------------------------------------------------------------------------------------- */
int trk_score_topk_max_k(int32_t d_pad);
int trk_pack_item_meta(const float* item_scale, const float* item_bias, int64_t n_items, float* item_meta,
                       int64_t n_items_padded, void* stream);
int trk_score_topk_f16x3(const void* user_split, const float* user_scale, const float* user_bias,
                         const void* item_split, const float* item_meta, int64_t n_users, int64_t n_items,
                         int32_t d_pad, int32_t k, int32_t n_splits, int32_t item_id_offset,
                         float* cand_score, int32_t* cand_item, const int32_t* n_users_live, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K2+K3 fused, FILTER form (the throughput path of predict_rank(k)): one tensor-core pass over the fp16 "hi"
 * halves gives approximate scores with a proven error bound m = 1.5*2^-10 * |u|_2 * max_j |i_j|_2 (+ bias
 * rounding); per user the kernel keeps every item whose approximate score is within 2.25 m of the running k-th best.
 * trk_rescore_topk_split then scores the survivors from the full split operands (22-bit operands, fp32 accumulate: the
 * arithmetic of trk_score_topk_f16x3; reference chain tensorrec/prediction_graphs.py:49-50 and
 * recommendation_graphs.py:41), ranks them in tf.nn.top_k order (recommendation_graphs.py:81) and verifies the
 * bound; users it flags are re-run through trk_score_topk_f16x3 (device-side routing below).  Same reference chain
 * as trk_score_topk_f16x3, one third of its tensor work.
 *
 * Preparation (all device-side, no host sync):
 *   K1 (trk_csr_gather_reduce_f32) already yields the user norms and the item statistics; for operands that do not
 *   come from K1 (user-defined representation graphs):
 *   trk_operand_stats      norm[r] = |row r|_2 (upper bound) of a split operand; stats[0] = max norm, stats[1] = max
 *                          row scale, both by atomic max (stats[3] must be zeroed by the caller; either output may be
 *                          NULL)
 *   trk_rescale_hi_global  item "hi" half re-expressed with ONE scale for the whole matrix and laid out in PROCESSING
 *                          order: out_hi[p, :] (f16 [rows, d_pad]) = hi[perm[p], :] * (scale[perm[p]] / stats[1]) (exact
 *                          power-of-two factors; perm NULL = identity)
 *   trk_pack_item_bias     out[p] = bias[perm[p]], padded with -inf to n_padded (multiple of 256) entries; stats[2] =
 *                          max |bias|; block_max[b] / block_min[b] = max / min bias of positions [128 b, 128 b + 128)
 *                          (min = -inf as soon as the block holds padding; block_min may be NULL)
 * Processing order: the host sorts the items by DESCENDING bias (perm = stable argsort) so that the biases inside a
 * 128-item block are nearly equal and the running k-th best rises early; the kernel's hot loop then bounds
 * acc_j + bias_j / c by max_j acc_j + block_max / c and touches neither the biases nor an FFMA per score.  With
 * block_bias_min the kernel also starts every (user, split) sweep from a threshold derived from the first tile
 * (k-th largest of 16 group maxima + block minimum) instead of -inf.  Candidate ids are reported in ORIGINAL numbering.
 * Filter outputs, one list per (user, split): cand_* [n_users, n_splits, 16] (approximate score, global id;
 * unused = (-inf, INT32_MAX)), row_theta [n_users, n_splits].
 * Constraints: d_pad in {64, 128}; 1 <= k <= trk_score_filter_max_k().
 * ---------------------------------------------------------------------------------------------------- */
int trk_score_filter_max_k(void);
int trk_score_filter_list_width(void); /* candidates per list: 16 (one list per (user, split)) */
int trk_operand_stats(const void* split, const float* scale, int64_t rows, int32_t d_pad, float* out_norm,
                      float* stats, void* stream);
int trk_rescale_hi_global(const void* split, const float* scale, const float* stats, const int32_t* perm,
                          int64_t rows, int32_t d_pad, void* out_hi, void* stream);
int trk_pack_item_bias(const float* item_bias, const int32_t* perm, int64_t n_items, float* out,
                       int64_t n_items_padded, float* stats, float* block_max, float* block_min, void* stream);
int trk_score_filter_f16(const void* user_split, const float* user_scale, const float* user_bias,
                         const float* user_norm, const void* item_hi_global, const float* item_stats,
                         const float* item_bias_padded, const float* block_bias_max, const float* block_bias_min,
                         const int32_t* item_perm, int64_t n_users, int64_t n_items, int32_t d_pad, int32_t k,
                         int32_t n_splits, int32_t item_id_offset, float* cand_score, int32_t* cand_item,
                         float* row_theta, void* stream);
/* item_split / item_scale / item_bias hold the rows of THIS shard: global id g lives at row g - item_id_offset.
 * n_lists = n_splits, list_width = 16.  Row u of the result is written at out_score + u * out_row_stride and
 * out_item + u * out_row_stride (both may point into one [n_users, 2k] exchange buffer: stride 2k, out_item =
 * out_score + k).  out_flag[u] = 1 -> user u must be re-run through the exact kernel. */
int trk_rescore_topk_split(const void* user_split, const float* user_scale, const void* item_split,
                           const float* item_scale, const float* user_bias, const float* item_bias,
                           const int32_t* cand_item, const float* row_theta, const float* user_norm,
                           const float* item_stats, int64_t n_users, int64_t n_items_local, int32_t d_pad,
                           int32_t n_lists, int32_t list_width, int32_t k, int32_t item_id_offset, float* out_score,
                           int32_t* out_item, int64_t out_row_stride, int32_t* out_flag, void* stream);

/* Device-side routing of the flagged users (no host round trip):
 *   trk_select_flagged_rows  idx[0 .. min(count, capacity)) = rows with flags != 0 (any order), counters[0] = count
 *                            (counters: int32[4], zeroed by this call; count > capacity = overflow, the host layer
 *                            checks it at its next natural synchronisation and re-runs the whole batch exactly);
 *   trk_gather_operand_rows  sub_split / sub_scale / sub_bias [capacity, ...] = the selected rows of a split operand;
 *                            also sets the live counts of the two re-scoring tiers: counters[2] = count when
 *                            count <= small_capacity (else 0), counters[3] = min(count, capacity) otherwise (else 0);
 *   trk_score_topk_f16x3 + trk_topk_merge with n_users_live = &counters[2] over the first small_capacity rows and MANY
 *                            item splits (a handful of user blocks still fills the machine), and with n_users_live =
 *                            &counters[3] over all capacity rows and few splits: exactly one tier does work;
 *   trk_scatter_topk_rows    out[idx[i]] = sub[i] for i < min(count, capacity) (row i of sub_* at i * sub_row_stride). */
int trk_select_flagged_rows(const int32_t* flags, int64_t n, int32_t* idx, int32_t capacity, int32_t* counters,
                            void* stream);
int trk_gather_operand_rows(const int32_t* idx, int32_t* counters, int32_t capacity, int32_t small_capacity,
                            const void* split, const float* scale, const float* bias, int32_t d_pad, void* sub_split,
                            float* sub_scale, float* sub_bias, void* stream);
int trk_scatter_topk_rows(const int32_t* idx, const int32_t* counters, int32_t capacity, const float* sub_score,
                          const int32_t* sub_item, int64_t sub_row_stride, int32_t k, float* out_score,
                          int32_t* out_item, int64_t out_row_stride, void* stream);

/* Tensor-core dense prediction with the same operands, writing the full fp32 matrix out[n_users, n_items]
 * (predict(); tensorrec/tensorrec.py:636-664).  HBM-write bound. */
int trk_score_dense_f16x3(const void* user_split, const float* user_scale, const float* user_bias,
                          const void* item_split, const float* item_meta, int64_t n_users, int64_t n_items,
                          int32_t d_pad, float* out, int64_t out_row_stride, void* stream);

/* Merges n_lists candidate lists per user (each sorted by (score desc, id asc), k_in entries) into the global
 * top k_out per user, same order.  Lists are the n_splits of one GPU and/or the shards received from the other GPUs
 * (item-axis sharding; the exchange itself is one NCCL all-to-all done by the host layer).
 * Entry j of list l of user u is read at cand_*[u * user_stride + l * list_stride + j]:
 *   one GPU, [n_users, n_lists, k_in]:                user_stride = n_lists * k_in, list_stride = k_in;
 *   exchange receive buffer [n_lists, n_users, 2k]:   user_stride = 2k, list_stride = n_users * 2k, cand_item = cand_score + k.
 * Row u of the result goes to out_*[u * out_row_stride ...].  n_users_live: see trk_score_topk_f16x3.
 * dedup != 0: the lists of a user may name the same item -- the per-TASTE top-k lists of a mixture-of-tastes model,
 * whose prediction is the maximum over the tastes (collapse_mixture_of_tastes, tensorrec/recommendation_graphs.py:107):
 * every item is emitted once, with its best score (k_out <= 32).  The top-k of max_t s_t(u, .) is contained in the union
 * of the per-taste top-k lists, so n_tastes fused top-k sweeps + this merge give the model's top-k without the
 * [n_users, n_items] matrix. */
int trk_topk_merge(const float* cand_score, const int32_t* cand_item, int64_t n_users, int32_t n_lists,
                   int32_t k_in, int32_t k_out, int64_t user_stride, int64_t list_stride, float* out_score,
                   int32_t* out_item, int64_t out_row_stride, const int32_t* n_users_live, int32_t dedup, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * The sampled-rank training step: everything of one Adam step of
 * LinearRepresentationGraph x DotProductPredictionGraph x WMRBLossGraph / BalancedWMRBLossGraph that is not a
 * sparse x dense product (those are trk_csr_gather_reduce_f32 on the CSR of the features and of their transpose).
 *
 * trk_sample_items   replaces sample_items behind tf.py_func (tensorrec/util.py:12-21, tensorrec/tensorrec.py:298-302):
 *                    out[u, j] (int32 [n_users, n_sampled]) = item ids drawn for user u, with replacement (uniform) or
 *                    without (Floyd's algorithm: a uniformly random n_sampled-subset; n_sampled <= 4096), from the
 *                    counter-based Philox4x32-10 stream (seed, step, user, draw).  trk_sample_stream_u64 exposes that
 *                    stream on the host (tests reproduce a device sample from it).
 * trk_wmrb_step      forward and backward between the representations and the loss, one warp per user:
 *                      pred(u, i) = (sum_k user_repr[u, k] * item_repr[i, k] + user_bias[u]) + item_bias[i]
 *                                   (tensorrec/prediction_graphs.py:52-55, recommendation_graphs.py:44-57)
 *                      for every stored interaction n = (u, i, val) (CSR by user, reference COO order) with val > 0:
 *                        loss[n] = log(n_items / n_sampled * sum_j max(0, 1 - pred(u, i) + pred(u, samples[u, j]))
 *                                      [* val / item_weight_sum[i]]  + 1)     (tensorrec/loss_graphs.py:153-180, 190-227)
 *                      loss[n] = 0 for val <= 0; pred_serial[n] = pred(u, i) for every n;
 *                      gradients of sum_n loss[n]: d_user_repr [n_users, d] and d_user_bias [n_users] are written,
 *                      d_item_repr [n_items, d] and d_item_bias [n_items] are ADDED to with red.global.add (zero them
 *                      first; the order of the floating-point additions is not fixed, as in tf.gather's GPU gradient).
 *                    Representations are fp32 or bf16 (repr_is_bf16; BASELINE config #4 allows bf16), arithmetic and
 *                    gradients fp32.  coef [nnz] is scratch.  Constraints: d % 4 == 0, d <= 512, n_sampled <= 2048.
 * trk_f32_to_bf16    round-to-nearest-even conversion of a representation for the bf16 form.
 * trk_adam_step_f32  tf.train.AdamOptimizer on (grad + l2 * w) (tensorrec/tensorrec.py:487-489):
 *                      m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  w -= lr_t m / (sqrt(v) + epsilon),
 *                    lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) formed by the caller.
 * ---------------------------------------------------------------------------------------------------- */
int trk_sample_items(int64_t n_users, int64_t n_items, int32_t n_sampled, int32_t replace, uint64_t seed,
                     uint32_t step, int32_t* out, void* stream);
uint64_t trk_sample_stream_u64(uint64_t seed, uint32_t step, uint32_t user, uint32_t draw);
int trk_wmrb_step(const void* user_repr, const void* item_repr, int32_t repr_is_bf16, const float* user_bias,
                  const float* item_bias, const int32_t* inter_indptr, const int32_t* inter_item,
                  const float* inter_val, const float* item_weight_sum, const int32_t* samples, int64_t n_users,
                  int64_t n_items, int32_t d, int32_t n_sampled, float* loss, float* pred_serial, float* coef,
                  float* d_user_repr, float* d_user_bias, float* d_item_repr, float* d_item_bias, void* stream);
int trk_f32_to_bf16(const float* x, int64_t n, void* out, void* stream);
int trk_adam_step_f32(float* w, const float* grad, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
                      float epsilon, float l2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TENSORREC_B200_H_ */
