/*
 * laff_b200 — C ABI of the B200-native (sm_100a) LAFF retrieval hot path.
 *
 * This is the drop-in boundary.  The reference (ruc-aimc-lab/LAFF) has no FFI layer: its hot path is Python calling
 * PyTorch/numpy ops.  Each entry point below replaces the reference call sites cited next to it (paths relative to
 * the reference repo); the Python host side (laff_b200/*.py) mirrors the reference's module/function names and
 * binds these symbols with ctypes (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch types.  All pointers are DEVICE pointers unless named host_*.
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without synchronising.
 *   - no allocation inside: ops that need scratch take (workspace, workspace_bytes); query the size with the
 *     matching *_workspace_bytes().  Distinct streams need distinct workspaces.
 *   - return value: 0 = ok, < 0 = LAFF_E* (bad argument / unsupported), > 0 = a cudaError_t.
 *     laff_last_error() returns a thread-local message for the last failure.
 *   - 16-bit operand dtype codes are the tcgen05 kind::f16 format codes: 0 = fp16, 1 = bf16.
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute call fails.
 */
#ifndef LAFF_B200_H_
#define LAFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAFF_OK 0
#define LAFF_EINVAL (-1)   /* bad argument */
#define LAFF_ENOTSUP (-2)  /* unsupported shape / option */
#define LAFF_ENODEV (-3)   /* no sm_100 device / driver entry point missing */
#define LAFF_EWORKSPACE (-4)

#define LAFF_F16 0
#define LAFF_BF16 1
#define LAFF_F32 2

#define LAFF_MAX_FEATURES 8
#define LAFF_MAX_TOPK 16
#define LAFF_FUSE_MAX_FC 4
#define LAFF_FUSE_MAX_TILED 2

#define LAFF_ACT_NONE 0
#define LAFF_ACT_TANH 1
#define LAFF_ACT_RELU 2
#define LAFF_ACT_SIGMOID 3

#define LAFF_DIR_T2I 0
#define LAFF_DIR_I2T 1
#define LAFF_DIR_BIDIR 2

const char* laff_last_error(void);
int laff_abi_version(void);

/* Number of CUDA kernels this library has launched since the last reset (bench.py's gpu_launches). */
long long laff_launch_count(int reset);

/* Tuning knobs of the tcgen05 GEMM engine (0 keeps the current value).
 *   cta_group   1: one CTA per 128x256 tile, 2: CTA pair per 256x256 tile (cta_group::2 MMA).
 *   chunk_tiles number of consecutive 256-column gallery tiles one work unit sweeps with per-row top-k state.
 *   m_group     number of query row-tiles that sweep the gallery together (L2 residency of the query block). */
int laff_set_tuning(int cta_group, int chunk_tiles, int m_group);
/* SM budget of the persistent kernels launched after the call (GEMM engine, fused fusion kernel, grid-stride kernels):
 * at most `sms` SMs are filled; 0 = the whole device.  Returns the previous value.  The retrieval pipeline gives the
 * sweep all but a couple of SMs and runs the next batch's query fusion (and the NCCL kernels of its collectives) on
 * those, so the two overlap instead of queueing behind each other.  Process-wide; set by the launching host thread. */
int laff_set_sm_limit(int sms);
int laff_get_tuning(int* cta_group, int* chunk_tiles, int* m_group);
/* laff_sim_rank_topk (and laff_sim_collect) cut a sweep into back-to-back launches: one per group of m_group row tiles and
 * per ~480 gallery column tiles, which keeps the CTA pairs that share a gallery tile aligned (10 % faster than one launch,
 * same results).  Diagnostics, environment variables read once: LAFF_SWEEP_SPLIT=0 (no row-group split),
 * LAFF_SWEEP_COLSPLIT=n (n launches per gallery pass; 1 = no column split). */

/* Which variant of the single-kernel fusion (laff_fuse_forward) runs: 0 = chosen per call from the row count,
 * 1 = cta_group::1 MMAs in 2-CTA clusters (all SMs), 2 = cta_group::2 MMA pairs in 4-CTA clusters (fewer bytes per
 * SM per k-block; 33 clusters fit a 148-SM B200).  Results are identical; only the speed differs. */
int laff_set_fuse_variant(int cta_group);
/* Debug: cycle counters of the fused kernel's epilogue phases (zeros unless built with -DLAFF_FUSE_PROFILE). */
int laff_debug_fuse_profile(unsigned long long* out8, int reset);
int laff_get_fuse_variant(void);

/* ---------------------------------------------------------------------------------------------------------------
 * S1  loss.l2norm (loss.py:8-13) applied per head, then rounding to the tensor-core operand type.
 *     x [rows, heads*head_dim] fp32, row stride ldx.  out[r, h, :] = x[r, h, :] / (||x[r, h, :]||_2 + eps)
 *     with eps = the reference's (eps + 1e-14) passed as a double; out_dtype LAFF_F16 / LAFF_BF16 / LAFF_F32.
 *     ld_out in elements of out_dtype.  eps < 0 skips the normalisation (pure cast). */
int laff_l2norm_quantize(const float* x, long long rows, int heads, int head_dim, long long ldx, double eps,
                         int out_dtype, void* out, long long ld_out, void* stream);

/* Split an fp32 matrix into 3-term 16-bit operands for near-fp32 tensor-core products:
 *   x ~= hi + lo with hi = round16(x), lo = round16(x - hi).
 *   side 0 (left operand):  out[r] = [hi | lo | hi]      side 1 (right operand): out[r] = [hi | hi | lo]
 *   so that dot(out_left[i], out_right[j]) = hi.hi + lo.hi + hi.lo.  cols_pad >= cols (zero filled), ld_out >= 3*cols_pad. */
int laff_split3_16(const float* x, long long rows, int cols, long long ldx, int side, int out_dtype, void* out,
                   int cols_pad, long long ld_out, void* stream);

/* fp32 -> 16-bit cast with zero padding of columns [cols, cols_pad) (TMA needs 16-byte aligned row pitches). */
int laff_cast_pad_16(const float* x, long long rows, int cols, long long ldx, int out_dtype, void* out, int cols_pad,
                     long long ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * S2  W2VVPP.get_txt2vis_matrix / compute_sim (model/model.py:1003-1016, :1567-1578) -> loss.cosine_sim's
 *     query.mm(retrio.t()) (loss.py:34) on already normalised + rounded operands, mean over heads = scale 1/H.
 *     out[i, j] = scale * sum_k q[i, k] * g[j, k]      q [Q, D] ldq, g [V, D] ldg (16-bit, D % 8 == 0), out fp32.
 *     tcgen05 GEMM, fp32 accumulate in TMEM. */
int laff_sim_dense(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                   float scale, float* out, long long ld_out, void* stream);

/* N1 at gallery scale (predictor.py:53-88: the top-2000 / top-500 lists per query) without the dense Q x V matrix:
 *     the same tcgen05 sweep as laff_sim_dense, but only the scores s = scale * <q_i, g_j> with s >= thr[i] are kept,
 *     appended unordered to query i's candidate list: cand_val / cand_idx [Q, cap] (global index = j + col_offset; unused
 *     slots -inf / -1), count[i] = number of such scores (may exceed cap: the list is then truncated).  The call
 *     initialises count and the lists.  The k best of a list are the k best of the row whenever k <= count[i] <= cap. */
int laff_sim_collect(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                     float scale, const float* thr, int col_offset, int cap, int32_t* count, float* cand_val,
                     int32_t* cand_idx, void* stream);
/* laff_sim_collect that ALSO counts the rank of every query's ground truth in the same sweep (sgt_raw / gt_global as in
 * laff_sim_rank_topk; rank_count[i] = number of local videos ranked above the ground truth, zeroed by the call), for a
 * caller that writes the long lists and the metrics of one query set (predictor.py:232-259): one pass, not two. */
int laff_sim_collect_rank(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                          float scale, const float* thr, int col_offset, int cap, int32_t* count, float* cand_val,
                          int32_t* cand_idx, const float* sgt_raw, const int32_t* gt_global, int32_t* rank_count,
                          void* stream);

/* Diagnostics: run the GEMM mainloop of laff_sim_* with a null epilogue (mode 1 drains TMEM, mode < 0 only sets the
 * TMA L2 eviction hints used by later laff_sim_* calls; hint codes 0 default, 1 normal, 2 evict-first, 3 evict-last). */
int laff_debug_gemm(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                    int mode, int hint_a, int hint_b, float* sink, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * E2  rank extraction: np.argsort(t2i_matrix, axis=1) + the per-query ground-truth search
 *     (predictor.py:232-244, trainer.py:584-594, evaluation.py:64-79) fused into the similarity GEMM; the Q x V
 *     matrix is never written.
 *
 *     Tie rule (documented, = np.argsort(kind='stable')[::-1]): order by (score desc, index desc);
 *       rank0[i] = #{j != gt: s_ij > s_i,gt} + #{j > gt: s_ij == s_i,gt}.
 *
 *  step 1  laff_sim_gt_scores: sgt_raw[i] = sum_k q[i,k] * g[gt_local[i], k] computed by the same MMA sequence as the
 *          sweep (unscaled fp32 accumulator).  gt_local[i] < 0 (ground truth lives on another shard) -> 0.
 *          Multi-GPU: all_reduce(sum) sgt_raw across gallery shards before step 2.
 *  step 2  laff_sim_rank_topk: one sweep over the local gallery shard g [V, D] whose column j has global index
 *          col_offset + j.  count[i] (int32) = local contribution to rank0[i]  (all_reduce(sum) across shards);
 *          topk_val/topk_idx [Q, k]: local top-k (scaled scores, global indices) ordered by the tie rule
 *          (entries beyond the shard size: -inf / -1).
 *  step 3  laff_topk_merge: merge n_lists ordered lists per query (all_gather'ed shards) into the global top-k. */
size_t laff_sim_gt_workspace_bytes(int Q, int D);
int laff_sim_gt_scores(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                       const int32_t* gt_local, float* sgt_raw, void* workspace, size_t workspace_bytes, void* stream);

size_t laff_sim_rank_workspace_bytes(int Q, int V, int D);
int laff_sim_rank_topk(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                       float scale, const float* sgt_raw, const int32_t* gt_global, int col_offset, int k,
                       int32_t* count, float* topk_val, int32_t* topk_idx, void* workspace, size_t workspace_bytes,
                       void* stream);

/* vals/idx: n_lists lists per query, list l of query i at vals + l*list_stride + i*k_in, each ordered by the tie rule.
 * out_val[i, 0..k_out) = in_scale * merged values. */
int laff_topk_merge(const float* vals, const int32_t* idx, int n_lists, int Q, int k_in, long long list_stride,
                    int k_out, float in_scale, float* out_val, int32_t* out_idx, void* stream);

/* Rank / top-k of an already materialised fp32 score matrix (the small-config predict() path and the integer
 * parity check of the fused kernel): same tie rule as above. */
int laff_rank_from_scores(const float* scores, int Q, int V, long long ld, const int32_t* gt, int k, int32_t* rank0,
                          float* topk_val, int32_t* topk_idx, void* stream);

/* Ranked lists for the result writers (predictor.py:53-88 txt2video_write_to_file: `inds[index][::-1][0:TopK]`,
 * TopK = 2000 for id.sent.score.txt, 500 for t2v.pkl): top-k of every row of a dense fp32 matrix scores[rows, cols]
 * (row pitch ld), 1 <= k <= LAFF_MAX_TOPK_DENSE, ordered by the tie rule above (score desc, index desc).
 *   idx_in == NULL: candidate j of a row has index j (any cols < 2^31).
 *   idx_in != NULL: int32 [rows, cols] (pitch ld_idx) indices carried by the candidates, -1 = empty slot; cols <= 16384.
 *                   Used to merge the all_gather'ed per-shard lists of a sharded gallery.
 * out_val[i, r] = scale * score, out_idx[i, r] = index; slots past the number of candidates: -inf / -1. */
#define LAFF_MAX_TOPK_DENSE 2048
int laff_topk_dense(const float* scores, long long ld, const int32_t* idx_in, long long ld_idx, int rows, long long cols,
                    int k, float scale, float* out_val, int32_t* out_idx, void* stream);

/* Video -> text direction (predictor.py:262-270): a row has several ground-truth columns (the captions of the video),
 * given as CSR lists gt_cols[gt_offsets[i] .. gt_offsets[i+1]).  rank0[e] = 0-based rank of ground truth e in its row
 * under the tie rule (-1 for a column outside [0, cols)). */
int laff_rank_multi_gt(const float* scores, long long ld, int rows, long long cols, const long long* gt_offsets,
                       const int32_t* gt_cols, int32_t* rank0, void* stream);

/* evaluation.eval (evaluation.py:92-109) from laff_rank_multi_gt's ranks: first[i] = rank of the best-ranked ground
 * truth of row i, ap[i] = mean_t (t + 1) / (r_(t) + 1) over its ground truths sorted by rank; out8 as laff_rank_metrics
 * with [6] = mAP = mean(ap).  Every row must own at least one ground truth (the reference raises IndexError). */
int laff_multi_gt_metrics(const int32_t* rank0, const long long* gt_offsets, int rows, int32_t* first, double* ap,
                          double* out8, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * E3  evaluation.eval_qry2retro metrics (evaluation.py:81-89) / evaluation.eval (evaluation.py:105-109) on device.
 *     rank0 int32 [Q] (0-based rank of the first ground truth).  out (device, 8 doubles):
 *       [0] R@1  [1] R@5  [2] R@10  (percent)  [3] MedR = floor(median(rank0)) + 1  [4] MeanR = mean(rank0) + 1
 *       [5] MIR = mean(1 / (rank0 + 1))  [6] mAP (= MIR for a single ground truth)  [7] Q. */
int laff_rank_metrics(const int32_t* rank0, int Q, double* out8, void* stream);

/* evaluation.eval(label_matrix) (evaluation.py:92-109) for the caller that still builds the 0/1 label matrix
 * (predictor.py:236-246): label uint8 [Q, V] ld, column p = p-th ranked item.  rank0[i] = 0-based position of the
 * first 1 (-1 if none), ap[i] = average precision; out8 as laff_rank_metrics with [6] = mAP = mean(ap). */
int laff_label_metrics(const uint8_t* label, int Q, int V, long long ld, int32_t* rank0, double* ap, double* out8,
                       void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * N2  Text front-end after tokenisation (model/model.py:322-434, txt2vec.py:49-109).  Token lists are CSR:
 *     ids[offsets[i] .. offsets[i+1]) belong to caption i; ids outside the table are skipped.
 *   laff_bow_counts   BowVec._encoding: out[i, id] += 1 (rows are zeroed first).  out fp32 [rows, ndims].
 *   laff_bow_project  BoWTxtEncoder + its TransformNet without the dense count vector (model/model.py:399-417 then
 *       :257-276; txt2vec.py:56-63): y[i, :] = BN(act(row_scale[i] * sum_t wt[id_t, :] + bias)), wt = the FC weight
 *       transposed, fp32 [vocab, D]; a repeated token adds its row twice (its count).  Token t of caption i is
 *       tok_ids[tok_offsets[i] - id_base + ...] (id_base lets a caller pass a slice of the id array with the
 *       offsets of the whole batch).  row_scale (optional): 1 / norm of the count vector (Txt2Vec.do_norm).
 *       The result enters laff_fuse_forward as a tiled feature with in_dim = D and no BatchNorm.
 *   laff_gather_mean  W2Vec._encoding: mean of table rows, summed in float64 in list order, rounded to fp32 once
 *                     (numpy's np.array(vectors).mean(axis=0) on float64); no valid id -> zeros.  The caller passes the
 *                     ids de-duplicated and sorted, as BigFile.read returns them (bigfile.py:204-211).
 *   laff_gather_rows  nn.Embedding lookup: out[i] = table[ids[i]] (zeros for an id outside the table).
 *   laff_gru_cell     one nn.GRU step for packed sequences: gi [B, 3H] = W_ih x_t + b_ih, gh [B, 3H] = W_hh h + b_hh
 *                     (gate order r, z, n); sequences with lengths[b] <= t keep h.  sum += h_t (for 'mean' pooling) and
 *                     last = h at t == len - 1 are optional outputs (model/model.py:361-383).
 *   laff_mean_over_length  x[b, :] /= lengths[b]. */
/* Host half of the front-end (no GPU work): batch tokenisation + vocabulary lookup, the equivalent of
 * TextTool.tokenize(clean=True, 'en') (textlib.py:27-47) followed by the per-word dict lookups of BowVec / W2Vec / IndexVec.
 *   laff_vocab_create   words: concatenated UTF-8 bytes with n_words + 1 offsets; ids: id per word or NULL (= position).
 *                       Returns NULL on error (laff_last_error).  Also used for the stop-word set.
 *   laff_tokenize_lookup  captions: concatenated UTF-8 bytes with n_caps + 1 offsets; stopwords may be NULL.
 *       mode 0 (IndexVec): start_id, every token (unknown -> unk_id), end_id   (start/end skipped when negative)
 *       mode 1 (BowVec):   in-vocabulary tokens in order
 *       mode 2 (W2Vec):    distinct in-vocabulary ids, ascending (bigfile.py:204-211)
 *     Writes CSR offsets [n_caps + 1] and up to `capacity` ids; returns the number of ids (> capacity: call again). */
typedef struct laff_vocab laff_vocab;
laff_vocab* laff_vocab_create(const char* words_blob, const long long* offsets, const int32_t* ids, int n_words);
void laff_vocab_destroy(laff_vocab* v);
long long laff_tokenize_lookup(const char* text_blob, const long long* cap_offsets, int n_caps, const laff_vocab* vocab,
                               const laff_vocab* stopwords, int mode, int unk_id, int start_id, int end_id,
                               long long* out_offsets, int32_t* out_ids, long long capacity);

int laff_bow_project(const long long* tok_offsets, const int32_t* tok_ids, long long id_base, int rows, int vocab,
                     const float* wt, long long ld_wt, int D, const float* bias, int activation, const float* bn_scale,
                     const float* bn_shift, const float* row_scale, float* y, long long ld_y, void* stream);
int laff_bow_counts(const long long* tok_offsets, const int32_t* tok_ids, int rows, int ndims, float* out, long long ld,
                    void* stream);
int laff_gather_mean(const float* table, long long ld_table, long long n_table, const long long* offsets,
                     const int32_t* ids, int rows, int dim, float* out, long long ld, void* stream);
int laff_gather_rows(const float* table, long long ld_table, long long n_table, const int32_t* ids, long long n, int dim,
                     float* out, long long ld, void* stream);
int laff_gru_cell(const float* gi, long long ld_gi, const float* gh, long long ld_gh, const float* h_prev,
                  const int32_t* lengths, int t, int B, int H, float* h_out, float* sum, float* last, void* stream);
int laff_mean_over_length(float* x, const int32_t* lengths, int B, int H, void* stream);
/* Training the GRU sentence encoder (backward through time; the reference leaves it to autograd, model/model.py:340-387):
 *   laff_gru_cell_backward  one step of BPTT.  dh_carry [B, H] (in/out): gradient reaching h_t through the z * h path;
 *                           dh_gemm (nullable): the part through W_hh (dGh_{t+1} @ W_hh, a GEMM of the caller);
 *                           dmean / dlast (nullable): d loss / d pooled output for 'mean' (divided by the length here)
 *                           and 'last' pooling.  Writes dgi_t, dgh_t [B, 3H]; sequences with lengths[b] <= t pass dh on.
 *   laff_scatter_add_rows   nn.Embedding backward: table_grad[ids[i]] += dx[i] (caller zeroes table_grad).
 *   laff_column_sum         out[c] = sum_r x[r, c] in a fixed order (bias gradients). */
int laff_gru_cell_backward(const float* gi, long long ld_gi, const float* gh, long long ld_gh, const float* h_prev,
                           const float* dmean, const float* dlast, const float* dh_gemm, const int32_t* lengths, int t, int B,
                           int H, float* dh_carry, float* dgi, long long ld_dgi, float* dgh, long long ld_dgh, void* stream);
int laff_scatter_add_rows(const float* dx, long long ld, const int32_t* ids, long long n, int dim, long long n_table,
                          float* table_grad, long long ld_table, void* stream);
int laff_column_sum(const float* x, long long ld, long long rows, int cols, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * F1  TransformNet.forward (model/model.py:257-276): y = BN(act(x W^T + b)), eval mode (dropout = identity,
 *     BatchNorm1d running stats, eps bn_eps).  x16 [rows, K] ldx, w16 [D, K] ldw (16-bit, K-major, pitches % 8 == 0).
 *     Eval-mode BatchNorm is the per-column affine map y*bn_scale + bn_shift produced by laff_bn_fold.
 *     bias / bn_scale / bn_shift may be NULL (stage absent).  y fp32 [rows, D] ldy.  tcgen05 GEMM, epilogue fused. */
int laff_project(const void* x16, const void* w16, long long rows, int K, int D, long long ldx, long long ldw,
                 int dtype, const float* bias, int activation, const float* bn_scale, const float* bn_shift,
                 float* y, long long ldy, void* stream);

/* nn.BatchNorm1d in eval mode (model/model.py:232, :273-274): scale = weight / sqrt(running_var + eps),
 * shift = bias - running_mean * scale.  weight / bias may be NULL (affine=False). */
int laff_bn_fold(const float* weight, const float* bias, const float* running_mean, const float* running_var,
                 double eps, int D, float* scale, float* shift, void* stream);

/* F2-F6  VisMutiTransformNet(AddAttnetion) / MultiScaleTxtEncoderAttention (model/model.py:1807-1876, :1663-1705),
 *     Multi_head_MyApply_Attention + Attention_1 (model/Attention.py:508-531, :78-105), loss.l2norm(eps=0).
 *     Per feature l the source is either a projected y_l [rows, D] (kind 0, from laff_project) or a "no-transform"
 *     raw feature x_l [rows, in_dim] that is tiled D/in_dim times across the D axis and passed through BN only
 *     (kind 1; model/model.py:1822-1823, :1675-1676, :1804-1805).
 *     Per head h: e_l = w_h . Y[l,h,:] + c_h ; a = softmax_l(e) ; out_h = l2norm(sum_l (a_l [+ omega]) Y[l,h,:]).
 *     mul: logits taken on Y[l,h,:] * mean_l Y[.,h,:] (Attention.py:83-86).  with_ave: + omega * mean-pool
 *     (Attention.py:94-99; omega = global_emb_weight_net.weight).
 *     out fp32 [rows, H*dh] ld_out (may be NULL), out16 (may be NULL) the same values rounded to out16_dtype,
 *     att (may be NULL) fp32 [rows, H, L] the attention weights the reference keeps in self.weights. */
typedef struct {
  int kind;            /* 0 = projected fp32 [rows, D]; 1 = raw fp32 [rows, in_dim] tiled + BN */
  int in_dim;          /* kind 1 only */
  const float* src;
  long long ld;
  const float* bn_scale; /* kind 1: folded BatchNorm1d(D) (laff_bn_fold), may be NULL = identity */
  const float* bn_shift;
} laff_pool_source;

typedef struct {
  int n_features;
  int heads;
  int head_dim;
  int with_ave;
  int mul;
  float omega;
  double norm_eps;        /* added to the L2 norm: the reference's l2norm(eps=0) adds 1e-14 */
  const float* att_weight; /* [heads, head_dim]  attention_layer.<h>.embedding_common.0.weight */
  const float* att_bias;   /* [heads] */
  laff_pool_source src[LAFF_MAX_FEATURES];
} laff_pool_desc;

int laff_attention_pool(const laff_pool_desc* desc, long long rows, float* out, long long ld_out, void* out16,
                        int out16_dtype, long long ld_out16, float* att, void* stream);

/* F1-F6 in ONE kernel: every FC projection (tcgen05 GEMM) + bias + activation + eval-BN + per-head attention logits +
 *     softmax over features + weighted sum + L2 normalise; the projected features never reach HBM.  Covers the shipped
 *     LAFF / LAFF-ml settings: head_dim = 512, with_ave = mul = False (Attention.py:78-105 with the softmax denominator
 *     cancelling under the final l2norm), up to LAFF_FUSE_MAX_FC projected and LAFF_FUSE_MAX_TILED "no-transform" features.
 *     fc[l].x16 [rows, K] 16-bit (pitch % 8 == 0, zero padded), fc[l].w16 [H*512, K]; tiled[l].x fp32 [rows, in_dim].
 *     out fp32 [rows, H*512] and / or out16; at least one.  Other settings: laff_project + laff_attention_pool. */
typedef struct {
  int n_fc;
  int n_tiled;
  int heads;
  int head_dim;   /* must be 512 */
  int dtype;      /* LAFF_F16 / LAFF_BF16 operands of the projection GEMMs */
  double norm_eps; /* added to the L2 norm (reference: 1e-14) */
  const float* att_weight; /* [heads, 512] */
  const float* att_bias;   /* [heads] */
  struct {
    const void* x16;
    long long ldx;
    const void* w16;
    long long ldw;
    int K;
    int activation;
    const float* bias;
    const float* bn_scale;
    const float* bn_shift;
  } fc[LAFF_FUSE_MAX_FC];
  struct {
    const float* x;
    long long ld;
    int in_dim;
    const float* bn_scale;
    const float* bn_shift;
  } tiled[LAFF_FUSE_MAX_TILED];
} laff_fuse_desc;

int laff_fuse_forward(const laff_fuse_desc* desc, long long rows, float* out, long long ld_out, void* out16,
                      int out16_dtype, long long ld_out16, void* stream);

/* F7  frame-level LAFF (model/model.py:2160-2173 -> Attention_1(dim), model/Attention.py:78-105):
 *     frames fp32 [B, F, dim] (zero padded frames take part in the softmax exactly like the reference),
 *     out fp32 [B, dim] unit norm. */
int laff_frame_pool(const float* frames, long long B, int F, int dim, const float* att_weight, float att_bias,
                    int with_ave, int mul, float omega, double norm_eps, float* out, long long ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * L1/L2  sum over heads of MarginRankingLoss.forward(s = txt[:, h, :], im = vis[:, h, :])
 *     (loss.py:95-135, model/model.py:852-862, :2036-2038), forward and backward in one call.
 *     txt, vis fp32 [B, H, dh] contiguous.  loss: device scalar.  d_txt / d_vis (may be NULL): dLoss/dtxt, dLoss/dvis.
 * L3     MarginRankingLossWithScore.forward(score) (loss.py:161-200): score fp32 [B, B] ld; d_score may be NULL. */
size_t laff_mrl_workspace_bytes(int B, int H, int dh);
int laff_mrl_forward_backward(const float* txt, const float* vis, int B, int H, int dh, float margin,
                              int max_violation, int direction, int cost_mean, float* loss, float* d_txt,
                              float* d_vis, void* workspace, size_t workspace_bytes, void* stream);
int laff_mrl_score_forward_backward(const float* score, int B, long long ld, float margin, int max_violation,
                                    int direction, int cost_mean, float* loss, float* d_score, void* stream);

/* DualSoftmaxLoss (loss.py:291-310; config.loss == 'dsl'), summed over heads like the margin loss:
 *   sim = cosine_sim(s, im);  f(A) = -sum_i log_softmax(A * softmax(A / temp, dim=0) * len(A), dim=-1)[i, i];
 *   loss = (f(sim) + f(sim^T)) / 2.   B <= 1024.  Workspace: laff_mrl_workspace_bytes(B, H, dh). */
int laff_dsl_forward_backward(const float* txt, const float* vis, int B, int H, int dh, float temp, float* loss, float* d_txt,
                              float* d_vis, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * T1 / N4  Training step (model/model.py:964-1001).  The projections and the weight-gradient products run on the
 *     GEMM entry points above (laff_project, laff_sim_dense); these are the stages around them.
 *
 * laff_transform_train_forward — TransformNet.forward in train mode after the activation (model/model.py:268-274):
 *     dropout (inverted, keep-probability 1 - p_drop, counter-based mask from `seed` (+ the device word *seed_dev when
 *     given: a step counter that lets a captured CUDA graph draw a fresh mask per replay), written to mask uint8 [B, D]) and
 *     BatchNorm1d with batch statistics (biased variance normalises, unbiased updates running_var, momentum as torch).
 *     src fp32 [B, src_cols]: the activated projection (src_cols == D) or a raw "no-transform" feature tiled D / src_cols
 *     times (model/model.py:1822-1823).  save_mean / save_invstd [D] feed the backward.  running_* may be NULL.
 * laff_transform_train_backward — its backward down to the GEMM output: dz = BN'(dy) * mask / (1 - p) * act'(a)
 *     (a = activated projection saved by the forward), dgamma / dbeta / dbias [D].  For a tiled feature pass tiled_x
 *     instead of a (its input is a leaf: dz must be NULL).
 * laff_attention_pool_backward — backward of Multi_head_MyApply_Attention + Attention_1 (with_ave / mul / omega as in
 *     laff_attention_pool; omega itself gets no gradient: the reference reads it with .item(), model/Attention.py:96):
 *     ys / dys: HOST arrays of n_features device pointers to y_l / dy_l fp32 [rows, H*d_h] (pitches lds, host array);
 *     dw [H, d_h], dc [H]: gradients of the per-head logit weights / biases; dw_part [rows*H*d_h], dc_part [rows*H]
 *     scratch (per-row partials, reduced in a fixed order: deterministic).
 * laff_transpose_16 — fp32 [rows, cols] -> 16-bit [cols, pad8(rows) * terms] (terms 1: rounding; 3: 3-term split, side
 *     0 = [hi|lo|hi], 1 = [hi|hi|lo]): K-major operands of dW = dZ^T x for laff_sim_dense.
 * laff_optimizer_step — clip_grad_norm_(params, max_grad_norm) (coef = max / (norm + 1e-6), clamped to 1, applied to
 *     the gradients in place when grad_out == grad) followed by torch.optim.RMSprop (kind 0: alpha, eps) or Adam (kind 1:
 *     beta1, beta2, eps, bias correction at `step`), all tensors in one pass, no host synchronisation.  Tensors with
 *     grad == NULL are skipped like parameters without .grad.  blk_* come from laff_optimizer_blocks (host helper: call
 *     with NULL outputs for the count, then with buffers of that capacity).  step_dev / lr_dev (optional device words)
 *     override `step` / `lr`, so a captured CUDA graph of the whole training step can be replayed.
 * laff_optimizer_step_scaled — the same step under the reference's float16 branch (model/model.py:970-989:
 *     scaler.scale(loss).backward(); clip_grad_norm_ on the SCALED gradients; scaler.step(); scaler.update()), with the
 *     GradScaler state in a device struct: clip coefficient min(1, max / (S*||g|| + 1e-6)); the step is skipped
 *     when a gradient is non-finite or S*max|g| >= overflow_limit (65520: it would be inf in the reference's fp16
 *     backward); then S *= backoff after a skipped step, S *= growth after growth_interval good ones.  step_dev is
 *     advanced here, and only on a good step.  total_norm_dev receives S*||g|| (what clip_grad_norm_ returns there);
 *     ctl_dev = {coef, skipped}.  No host synchronisation: graph-capturable. */
typedef struct {
  float scale;        /* S, torch default init 65536 */
  int growth_tracker; /* good steps since the last change of S */
  int found_inf;      /* 1 if the last step was skipped */
  int skipped;        /* skipped steps so far */
} laff_scaler_state;

typedef struct {
  float* param;
  const float* grad;
  float* grad_out; /* clipped gradient destination (may alias grad) or NULL */
  float* state1;   /* RMSprop square_avg / Adam exp_avg */
  float* state2;   /* Adam exp_avg_sq */
  long long n;
} laff_opt_tensor;

int laff_transform_train_forward(const float* src, long long ld_src, int src_cols, int B, int D, float p_drop,
                                 unsigned long long seed, const unsigned long long* seed_dev, const float* gamma,
                                 const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                 int use_bn, float* y, long long ld_y, uint8_t* mask, float* save_mean, float* save_invstd,
                                 void* stream);
int laff_transform_train_backward(const float* dy, long long ld_dy, const float* a, long long ld_a, const float* tiled_x,
                                  long long ld_x, int in_dim, const uint8_t* mask, float p_drop, int activation, int use_bn,
                                  const float* gamma, const float* save_mean, const float* save_invstd, int B, int D,
                                  float* dz, long long ld_dz, float* dgamma, float* dbeta, float* dbias, void* stream);
int laff_attention_pool_backward(const float* const* ys, const long long* lds, int n_features, int heads, int head_dim,
                                 const float* att_weight, const float* att_bias, const float* dout, long long ld_dout,
                                 long long rows, float norm_eps, int with_ave, int mul, float omega, float* const* dys,
                                 float* dw_part, float* dc_part, float* dw, float* dc, void* stream);
int laff_transpose_16(const float* x, long long ld, int rows, int cols, int dtype, int terms, int side, void* out16,
                      long long ld_out, void* stream);
/* LAFF-ml (model/model.py:2147-2190): gradient reaching a tiled "no-transform" feature, dx[b, j] = sum_h dz[b, h*in_dim + j]
 * (backward of x.repeat(1, heads)); and the backward of the frame-level Attention_1 block that produced it — gradients
 * of its logit weight [dim] / bias [1] (frames fp32 [B, F, dim] are leaves; F <= 128; dw_part [B*dim], dc_part [B] scratch). */
int laff_fold_tiles(const float* dz, long long ld_dz, int B, int D, int in_dim, float* dx, long long ld_dx, void* stream);
int laff_frame_pool_backward(const float* frames, long long B, int F, int dim, const float* att_weight, const float* dout,
                             long long ld_dout, double norm_eps, float* dw_part, float* dc_part, float* dw, float* dc,
                             void* stream);
int laff_optimizer_blocks(const long long* sizes, int n_tensors, int* blk_tensor, long long* blk_start, int capacity);
int laff_optimizer_step(const laff_opt_tensor* tensors_dev, const int* blk_tensor_dev, const long long* blk_start_dev,
                        int n_blocks, int kind, float lr, float alpha_or_beta1, float beta2, float eps, long long step,
                        float max_grad_norm, double* partial_dev, double* total_norm_dev, const long long* step_dev,
                        const float* lr_dev, void* stream);
int laff_optimizer_step_scaled(const laff_opt_tensor* tensors_dev, const int* blk_tensor_dev, const long long* blk_start_dev,
                               int n_blocks, int kind, float alpha_or_beta1, float beta2, float eps, float max_grad_norm,
                               double* partial_dev, float* partial_max_dev, double* total_norm_dev, long long* step_dev,
                               const float* lr_dev, laff_scaler_state* scaler_dev, float growth_factor, float backoff_factor,
                               int growth_interval, float overflow_limit, float* ctl_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAFF_B200_H_ */
