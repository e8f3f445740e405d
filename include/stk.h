/* stk.h — C ABI of libstk.so, the sm_100a kernel library behind stonkgs_b200.
 *
 * The reference (stonkgs/stonkgs) is pure Python on top of HuggingFace BERT and has no FFI layer;
 * its hot path is `STonKGsForPreTraining.forward` (src/stonkgs/models/stonkgs_model.py:149-258) and
 * the extraction loop of `get_stonkgs_embeddings` (src/stonkgs/models/stonkgs_for_embeddings.py:176-184).
 * Each entry point below replaces one piece of that path; the reference lines it replaces are cited.
 * "HF" = transformers/models/bert/modeling_bert.py (the third-party arithmetic the reference calls).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch tensors); the library never
 *    allocates or frees caller-visible memory, workspaces are passed in;
 *  - every call takes (device ordinal, cudaStream_t as void*) and is asynchronous on that stream;
 *  - return value: 0 = ok, negative = STK_ERR_*; stk_last_error() gives the text (thread-local);
 *  - bf16 tensors are passed as void* (raw uint16 storage), row-major, leading dimension in ELEMENTS;
 *  - no exceptions cross the boundary; functions are re-entrant (backward runs on the autograd thread).
 */
#ifndef STK_H_
#define STK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STK_OK 0
#define STK_ERR_BAD_ARG (-1)
#define STK_ERR_CUDA (-2)
#define STK_ERR_UNSUPPORTED (-3)

#define STK_VERSION 103

/* library version (STK_VERSION) */
int stk_version(void);
/* copies the calling thread's last error text into buf (NUL terminated); returns its length */
int stk_last_error(char* buf, size_t n);
/* number of kernel launches issued by this library in the calling process so far */
long long stk_launch_count(void);
/* Leave n SMs of `device` free in every persistent kernel launched from now on (tcgen05 GEMM, attention): room for the
 * CTAs of a concurrent collective (the data-parallel gradient all-reduce, stonkgs_pretraining.py:147-168).  Returns the
 * previous value (>= 0) or STK_ERR_BAD_ARG.  0 = use every SM (default). */
int stk_set_sm_reserve(int device, int n);
/* Tile scheduling of stk_gemm (all epilogues except the LayerNorm-fused ones): 0 = static stride over a persistent grid,
 * 1 = dynamic through cluster launch control (one CTA pair per tile is launched; running pairs cancel pending ones and
 * take over their tiles), which keeps a GEMM at full speed on the SMs it gets when a concurrent kernel holds some of
 * them.  Returns the previous setting.  Default: env STK_GEMM_DYNAMIC, else 0. */
int stk_set_gemm_dynamic(int on);

/* ------------------------------------------------------------------------------------------------
 * Embedding stages
 * ---------------------------------------------------------------------------------------------- */

/* A1 — input stage of the frozen LM backbone: out = LN(word[id] + type[0] + pos[p]).
 * Replaces `self.lm_backbone(input_ids[:, :256])` embeddings (stonkgs_model.py:178 -> HF:72-112,
 * called with ids only, so token type is 0 and there is no mask).
 * ids: int64 [B, S] with row pitch ids_pitch (elements), so the text half of a [B,512] batch can be
 * passed in place.  word/pos/type tables and LayerNorm gain/bias are fp32.  out: bf16 [B*S, 768].
 * err_flag (device int, may be NULL): bit 0 (value 1) is set if an id is outside [0, vocab).  All err_flag arguments of
 * this library are bit sets (atomic OR): 1 = input id out of range, 2 = label out of range, 4 = label capacity. */
int stk_embed_text_ln_fwd(int device, void* stream, const int64_t* ids, int64_t ids_pitch, int B, int S,
                          const float* word, int vocab, const float* pos, const float* type_emb,
                          const float* gamma, const float* beta, void* out_bf16, int* err_flag);

/* A3+A4 — joint embedding stage: for t < 256 the row is the LM-backbone hidden state, for t >= 256
 * it is the KG table row T[input_ids[b, t]] (dense table restating stonkgs_model.py:123-141,182-189);
 * out = LN((row + type[token_type]) + pos[t])   (stonkgs_model.py:193-210 -> HF:102-112).
 * input_ids/token_type_ids: int64 [B, 512] (token_type_ids may be NULL = 0 for t<256, 1 otherwise).
 * lm_hidden: bf16 [B, 256, 768]; kg_table: fp32 [table_rows, 768].
 * out: bf16 [B*512, 768]; mean/rstd: fp32 [B*512] saved for backward (may be NULL);
 * inputs_embeds_out: optional fp32 [B*512, 768] copy of the un-normalised gathered rows (bit-exact
 * gather check / API parity).  err_flag set to 1 on an id outside [0, table_rows) (reference: KeyError). */
int stk_embed_joint_ln_fwd(int device, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,
                           int B, const void* lm_hidden_bf16, const float* kg_table, int64_t table_rows,
                           const float* pos, const float* type_emb, const float* gamma, const float* beta,
                           void* out_bf16, float* mean, float* rstd, float* inputs_embeds_out, int* err_flag);

/* Backward of the joint embedding stage (the gathered rows are frozen: no gradient flows to them).
 * dy: bf16 [B*512, 768].  Accumulates (+=) into fp32 dpos [512,768], dtype [2,768], dgamma, dbeta [768]. */
int stk_embed_joint_ln_bwd(int device, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,
                           int B, const void* lm_hidden_bf16, const float* kg_table, int64_t table_rows,
                           const float* pos, const float* type_emb, const float* gamma, const float* mean,
                           const float* rstd, const void* dy_bf16, float* dpos, float* dtype, float* dgamma,
                           float* dbeta);

/* Shape-generic forms of the two calls above for the TransE variant (transestonkgs_model.py:44,93: 256 text tokens +
 * 4 KG tokens, max_position_embeddings = 260).  text_len = T, seq_len = S: input_ids / token_type_ids are [B, S],
 * lm_hidden is [B, T, 768], pos is [S, 768].  Activations (out, mean, rstd, inputs_embeds_out, dy) use seq_pad >= S rows
 * per pair; rows t >= S are written as zeros by the forward (the caller masks them out as attention keys) and ignored
 * by the backward.  (256, 512, 512) is exactly stk_embed_joint_ln_fwd / _bwd. */
int stk_embed_joint_ln_fwd_shape(int device, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,
                                 int B, int text_len, int seq_len, int seq_pad, const void* lm_hidden_bf16,
                                 const float* kg_table, int64_t table_rows, const float* pos, const float* type_emb,
                                 const float* gamma, const float* beta, void* out_bf16, float* mean, float* rstd,
                                 float* inputs_embeds_out, int* err_flag);
int stk_embed_joint_ln_bwd_shape(int device, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,
                                 int B, int text_len, int seq_len, int seq_pad, const void* lm_hidden_bf16,
                                 const float* kg_table, int64_t table_rows, const float* pos, const float* type_emb,
                                 const float* gamma, const float* mean, const float* rstd, const void* dy_bf16,
                                 float* dpos, float* dtype, float* dgamma, float* dbeta);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over rows of 768 (HF:294-298, 352-356, 481-485; eps = 1e-12)
 * ---------------------------------------------------------------------------------------------- */
int stk_layernorm_fwd(int device, void* stream, const void* x_bf16, int M, const float* gamma, const float* beta,
                      void* y_bf16, float* mean, float* rstd);
/* dx: bf16 [M,768]; dgamma/dbeta: fp32 [768], accumulated (+=).  x is the forward INPUT. */
int stk_layernorm_bwd(int device, void* stream, const void* dy_bf16, const void* x_bf16, int M, const float* gamma,
                      const float* mean, const float* rstd, void* dx_bf16, float* dgamma, float* dbeta);
/* The same, fused with what follows it in the backward of BertSelfOutput / BertOutput (HF:294-298, 352-356): dxm (bf16,
 * NULL when thr == 0) = dx through the dropout mask of the dense output (seed, site, thr as in stk_dropout_fwd) and
 * dbias (fp32 [768], may be NULL) += column sums of dxm (of dx when thr == 0), i.e. the dense layer's bias gradient. */
int stk_layernorm_bwd_fused(int device, void* stream, const void* dy_bf16, const void* x_bf16, int M, const float* gamma,
                            const float* mean, const float* rstd, void* dx_bf16, float* dgamma, float* dbeta,
                            void* dxm_bf16, float* dbias, uint32_t seed, uint32_t site, uint32_t thr);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 GEMM:  C[M,N] = epilogue( A[M,K] * B[N,K]^T )      (bf16 in, fp32 accumulate in TMEM)
 * Replaces every nn.Linear on the path (HF:158-160,179-181,287-298,330-356,456-468,471-485;
 * stonkgs_model.py:62-73) and their autograd backward.
 *
 * a_major / b_major: 0 = K-major (A stored [M][K], B stored [N][K], i.e. nn.Linear weight layout),
 *                    1 = MN-major (A stored [K][M], B stored [K][N]).
 *   forward  y = x W^T      : A = x (K-major), B = W  (K-major)
 *   dgrad    dx = dy W      : A = dy (K-major), B = W (MN-major: W is [N_out=K][N_in=N])
 *   wgrad    dW = dy^T x    : A = dy (MN-major: dy is [tokens=K][M]), B = x (MN-major)
 * lda/ldb/ldc: row pitch in elements of the stored matrices.
 * ---------------------------------------------------------------------------------------------- */
enum {
  STK_EPI_BIAS = 0,          /* C(bf16) = acc + bias[n]                    (bias may be NULL)         */
  STK_EPI_BIAS_GELU = 1,     /* C(bf16) = gelu_erf(acc + bias)                                        */
  STK_EPI_BIAS_GELU_SAVE = 2,/* C(bf16) = gelu_erf(u), C2(bf16) = u = acc + bias  (training forward)  */
  STK_EPI_BIAS_RESID = 3,    /* C(bf16) = acc + bias + R[m,n]              (R bf16, pitch ldr)        */
  STK_EPI_BIAS_TANH_F32 = 4, /* C(fp32) = tanh(acc + bias)                 (pooler, HF:456-468)       */
  STK_EPI_DGELU = 5,         /* C(bf16) = acc * gelu_erf'(R[m,n])          (R = saved pre-activation) */
  STK_EPI_F32_ADD = 6,       /* C(fp32) += acc   (TMA reduce-add; split-K and grad accumulation)      */
  STK_EPI_F32 = 7,           /* C(fp32) = acc                                                          */
  STK_EPI_CE_STATS = 8,      /* no C: per-row (max, sum exp) partials + target logit (A9+A10 fwd)     */
  STK_EPI_CE_DLOGIT = 9,     /* C(bf16) = (exp(acc - lse[m]) - [n+n_offset == label[m]]) * *scale_dev   */
  STK_EPI_BIAS_RESID_LN = 10,/* z = acc + bias + R;  C(bf16) = LayerNorm_768(z) * gamma + beta  (HF:294-298,352-356);
                                N must be 768 (rows are normalised across a 3-CTA cluster through distributed
                                shared memory); optional C2(bf16) = z and ln_mean/ln_rstd[m] for the backward */
  STK_EPI_BIAS_GELU_SAVE_GRAD = 11, /* C(bf16) = gelu_erf(u), C2(bf16) = gelu_erf'(u), u = acc + bias: the training forward of
                                       BertIntermediate saves the derivative, so that the backward epilogue is one multiply */
  STK_EPI_MUL = 12,                 /* C(bf16) = acc * R[m,n]   (R = the saved derivative; backward of the GELU)              */
  STK_EPI_BIAS_DROP_RESID_LN = 13   /* train() form of 10: z = drop(acc + bias) + R with the keep decisions of csrc/stk_rng.cuh for
                                       (drop_seed, drop_site, row m, column n), survivors scaled by 128 / (128 - drop_thr)
                                       (HF:296-298, 354-356: dropout of the dense output BEFORE the residual sum); the mask is
                                       regenerated by stk_layernorm_bwd_fused, never stored */
};

typedef struct StkGemmEpilogue {
  const float* bias;      /* [N] fp32 or NULL */
  const void* resid;      /* bf16 [M, ldr]: residual (BIAS_RESID), saved pre-activation (DGELU) or saved derivative (MUL) */
  int64_t ldr;
  void* c2;               /* second output (BIAS_GELU_SAVE): bf16 [M, ldc2] */
  int64_t ldc2;
  const int32_t* labels;  /* CE: [M] target column in the FULL vocabulary, or -1 */
  const float* lse;       /* CE_DLOGIT: [M] log-sum-exp of the full row */
  const float* scale_dev; /* CE_DLOGIT: device scalar, gradient scale (upstream grad / labelled-row count) */
  float* ce_partial;      /* CE_STATS: fp32 [M, ce_pitch, 2] (max, sumexp) per 128-column slab */
  int64_t ce_pitch;       /* number of slabs in the full vocabulary = 2 * ceil(N_full / 256) */
  float* tgt_logit;       /* CE_STATS: [M] logit of the target column (written by the owning slab) */
  int32_t n_offset;       /* CE: first vocabulary column of this call's B block (multiple of 256) */
  const float* ln_gamma;  /* BIAS_RESID_LN: [768] LayerNorm weight */
  const float* ln_beta;   /* BIAS_RESID_LN: [768] LayerNorm bias */
  float* ln_mean;         /* BIAS_RESID_LN: optional [M] row mean of z (saved for the backward) */
  float* ln_rstd;         /* BIAS_RESID_LN: optional [M] 1/sqrt(var + 1e-12) */
  uint32_t drop_seed;     /* BIAS_DROP_RESID_LN: dropout seed of this step */
  uint32_t drop_site;     /* BIAS_DROP_RESID_LN: site id (encoder, layer, dense output) */
  uint32_t drop_thr;      /* BIAS_DROP_RESID_LN: round(128 p) in [1, 127] */
} StkGemmEpilogue;

/* split_k: number of k-splits of an STK_EPI_F32_ADD GEMM (each split reduce-adds its partial tile); 0 = chosen by the
 * library so that the persistent grid is full; ignored (1) for every other epilogue. */
int stk_gemm(int device, void* stream, int a_major, int b_major, const void* A_bf16, int64_t lda,
             const void* B_bf16, int64_t ldb, int M, int N, int K, int epilogue, void* C, int64_t ldc,
             const StkGemmEpilogue* epi, int split_k);

/* ------------------------------------------------------------------------------------------------
 * MLM / ELM heads as single calls (stonkgs_model.py:62-73 decoders without bias, :229-245 mean cross-entropies)
 * ---------------------------------------------------------------------------------------------- */
/* Workspace bytes of the calls that take one; negative = STK_ERR_*.
 *   STK_WS_ATTN_BWD       a = B, b = S            (stk_attn_bwd / stk_attn_bwd_dropout)
 *   STK_WS_LINEAR_CE_FWD  a = rows, b = vocabulary (stk_linear_ce_fwd)
 *   STK_WS_LINEAR_CE_BWD  a = rows, b = vocabulary (stk_linear_ce_bwd) */
enum { STK_WS_ATTN_BWD = 0, STK_WS_LINEAR_CE_FWD = 1, STK_WS_LINEAR_CE_BWD = 2 };
int64_t stk_query_workspace(int op, int64_t a, int64_t b);

/* Label selection on the device (replaces `labels != -100` boolean indexing, stonkgs_model.py:229-245, and the host
 * sync it costs): labels int64 [B, width] (-100 = ignore).  Writes, in row-major order of the labelled positions,
 * rows_out[i] = b * row_pitch + col_offset + t (row of the [B * row_pitch, 768] sequence output) and labels_out[i];
 * entries [count, capacity) are padding (row -1, label -1: gathered as zeros, no loss, zero gradient), so every
 * consumer runs at the fixed size `capacity` and nothing is read back.  count_out[0] = number of labelled positions.
 * err_flag (device int, may be NULL): bit 1 (value 2) = a label outside [0, vocab) (torch raises IndexError; the row
 * is skipped), bit 2 (value 4) = more labelled positions than capacity (the surplus is dropped: caller must raise). */
int stk_compact_labels(int device, void* stream, const int64_t* labels, int B, int width, int row_pitch, int col_offset,
                       int vocab, int capacity, int32_t* rows_out, int32_t* labels_out, int32_t* count_out,
                       int* err_flag);

/* Fused decoder GEMM + cross-entropy forward over R rows: t bf16 [R, 768] (head-transform output of the labelled rows),
 * w bf16 [V, 768] (text_decoder / entity_decoder weight, no bias), labels int32 [R] (target column, < 0 = padding row).
 * Outputs: lse [R] (log-sum-exp of the full row of logits), row_loss [R] (lse - target logit; 0 for padding rows; may be
 * NULL), loss_count [2] (may be NULL) = { mean of row_loss over the rows with label >= 0 — NaN if there are none, like
 * torch — , that count as a float }.  The [R, V] logits are never materialised. */
int stk_linear_ce_fwd(int device, void* stream, const void* t_bf16, const void* w_bf16, int R, int V,
                      const int32_t* labels, void* workspace, int64_t workspace_bytes, float* lse, float* row_loss,
                      float* loss_count);
/* Its backward: dlogit = (softmax - onehot) * *scale_dev (scale_dev: device scalar = upstream gradient / labelled-row
 * count; padding rows get 0), recomputed per vocabulary chunk into the workspace; dT fp32 [R, 768] and dW fp32 [V, 768]
 * are accumulated (+=). */
int stk_linear_ce_bwd(int device, void* stream, const void* t_bf16, const void* w_bf16, int R, int V,
                      const int32_t* labels, const float* lse, const float* scale_dev, void* workspace,
                      int64_t workspace_bytes, float* dT, float* dW);

/* ------------------------------------------------------------------------------------------------
 * Fused masked-softmax attention (HF:115-140 eager / 192-205 sdpa; additive key mask HF:666-672)
 * qkv: bf16 [B*S, 2304] = [Q | K | V], head h at columns h*64 of each third.  S in {128, 256, 384, 512}.
 * key_bias: fp32 [B, S] additive bias per key (0 or finfo.min), NULL = no mask (LM backbone).
 * out: bf16 [B*S, 768].  lse: optional fp32 [B, 12, S] (row log-sum-exp of the scaled, biased scores).
 * ---------------------------------------------------------------------------------------------- */
int stk_attn_fwd(int device, void* stream, const void* qkv_bf16, const float* key_bias, int B, int S,
                 void* out_bf16, float* lse);
/* Same, for the first q_rows query rows of every sequence only (q_rows a multiple of 128, <= S); the other rows of out /
 * lse are not written.  The extraction path needs only row 0 of the last layer: pooler_output reads hidden[:, 0]
 * (HF:456-468 BertPooler; stonkgs_for_embeddings.py:180). */
int stk_attn_fwd_qrows(int device, void* stream, const void* qkv_bf16, const float* key_bias, int B, int S, int q_rows,
                       void* out_bf16, float* lse);
/* dqkv: bf16 [B*S, 2304].  dout/out: bf16 [B*S, 768].
 * workspace: fp32 [B*S*768 + B*12*S] (dQ accumulator over key blocks, then row dot(dO,O)). */
int stk_attn_bwd(int device, void* stream, const void* qkv_bf16, const float* key_bias, int B, int S,
                 const void* out_bf16, const void* dout_bf16, const float* lse, float* workspace,
                 void* dqkv_bf16);

/* Training-mode variants with dropout of the attention probabilities (HF:132).  keep(seed, site, row, key) is the
 * counter-based decision of csrc/stk_rng.cuh with row = (b*12 + h)*S + query; thr = round(128 p) in [0, 127]
 * (thr = 0: no dropout); survivors are scaled by 128 / (128 - thr).  lse stays that of the full softmax. */
int stk_attn_fwd_dropout(int device, void* stream, const void* qkv_bf16, const float* key_bias, int B, int S,
                         void* out_bf16, float* lse, uint32_t seed, uint32_t site, uint32_t thr);
int stk_attn_bwd_dropout(int device, void* stream, const void* qkv_bf16, const float* key_bias, int B, int S,
                         const void* out_bf16, const void* dout_bf16, const float* lse, float* workspace,
                         void* dqkv_bf16, uint32_t seed, uint32_t site, uint32_t thr);

/* ------------------------------------------------------------------------------------------------
 * Hidden dropout (training mode; HF:110, 297, 355).  Same decision function, row = token row, column = hidden index.
 * ---------------------------------------------------------------------------------------------- */
/* y = drop(x) over bf16 [M,768] (in place allowed).  Also the backward of every hidden-dropout site. */
int stk_dropout_fwd(int device, void* stream, const void* x_bf16, int M, uint32_t seed, uint32_t site, uint32_t thr,
                    void* y_bf16);
/* z = drop(x) + resid ; y = LayerNorm(z) * gamma + beta   (BertSelfOutput / BertOutput in train()).
 * z_out (bf16, may be NULL) and mean / rstd (may be NULL) are what stk_layernorm_bwd needs. */
int stk_dropout_resid_ln_fwd(int device, void* stream, const void* x_bf16, const void* resid_bf16, int M,
                             const float* gamma, const float* beta, uint32_t seed, uint32_t site, uint32_t thr,
                             void* z_out_bf16, void* y_bf16, float* mean, float* rstd);

/* ------------------------------------------------------------------------------------------------
 * Small fused helpers
 * ---------------------------------------------------------------------------------------------- */
/* int64 attention_mask [B,S] (1 = attend) -> fp32 additive key bias (0 / finfo.min), HF:666-672 */
int stk_mask_to_bias(int device, void* stream, const int64_t* mask, int64_t n, float* bias);
/* fp32 -> bf16 cast of n elements (weights after load / optimizer step) */
int stk_cast_f32_to_bf16(int device, void* stream, const float* src, void* dst_bf16, int64_t n);
/* gather rows: dst[i,:] = src[idx[i],:] for bf16 rows of 768 (labelled-row compaction for the heads); idx[i] < 0 = zeros */
int stk_gather_rows(int device, void* stream, const void* src_bf16, const int32_t* idx, int n_rows, void* dst_bf16);
/* scatter-add rows: dst[idx[i],:] += src[i,:]  (bf16 += bf16; non-negative idx unique, idx[i] < 0 skipped) — head gradient
 * back to the sequence */
int stk_scatter_add_rows(int device, void* stream, const void* src_bf16, const int32_t* idx, int n_rows,
                         void* dst_bf16);
/* column sums: out[n] (+)= sum_m x[m,n], x bf16 [M, ld]; bias gradients */
int stk_colsum(int device, void* stream, const void* x_bf16, int64_t ld, int M, int N, float* out, int accumulate);
/* CE finalisation (stonkgs_model.py:229-245): reduce the slab partials to lse[m] and
 * loss_sum += sum_m (lse[m] - tgt[m]); count is the number of rows. */
int stk_ce_finalize(int device, void* stream, const float* ce_partial, int64_t ce_pitch, const float* tgt_logit,
                    int M, float* lse, float* row_loss);
/* A8+A11: pooled -> Linear(768->2) -> CE (HF:528-533); pooled fp32 [B,768]; logits fp32 [B,2];
 * row_loss fp32 [B] (may be NULL when labels is NULL); a label outside {0, 1} gives a NaN row loss and sets bit 1
 * (value 2) of err_flag (may be NULL) — torch raises IndexError there */
int stk_nsp_head_fwd(int device, void* stream, const float* pooled, int B, const float* w, const float* b,
                     const int64_t* labels, float* logits, float* row_loss, int* err_flag);
/* Mean-pooled extraction output (BASELINE north_star "mean-pooled embedding extraction path"; an EXTRA beside the
 * reference's pooler_output, stonkgs_for_embeddings.py:180, never a replacement): out[b, :] = mean over the attended
 * tokens (attention_mask[b, t] != 0, t < seq_len; NULL = all) of the last hidden state x bf16 [B*seq_pad, 768]; out fp32
 * [B, 768]. */
int stk_masked_mean_pool(int device, void* stream, const void* x_bf16, const int64_t* attention_mask, int B, int seq_len,
                         int seq_pad, float* out);
/* head_mask of the reference forward (stonkgs_model.py:158,209 -> HF attention: probabilities * head_mask[layer, head]):
 * a per-head scale of the attention probabilities is a per-head scale of the context columns.
 * y[m, h*64 + d] = x[m, h*64 + d] * scales[h]  over bf16 [M, 768] (in place allowed); scales fp32 [12]. */
int stk_scale_heads(int device, void* stream, const void* x_bf16, int M, const float* scales, void* y_bf16);

/* dx = dy * gelu_erf'(pre), n bf16 elements (n % 8 == 0): backward of the head transform's GELU (HF:481-485) */
int stk_gelu_bwd(int device, void* stream, const void* dy_bf16, const void* pre_bf16, int64_t n, void* dx_bf16);
/* Backward of A11+A8: NSP cross-entropy -> seq_relationship Linear -> pooler tanh.
 * scale_dev: device scalar = upstream grad / B.  dw [2,768], db [2] are accumulated (+=);
 * dpre: bf16 [B,768] gradient w.r.t. the pooler's pre-activation. */
int stk_nsp_pool_bwd(int device, void* stream, const float* pooled, const float* logits, const int64_t* labels,
                     int B, const float* scale_dev, const float* w, float* dw, float* db, void* dpre_bf16);

/* Sequence-classification head of fine-tuned checkpoints (reference stonkgs_finetuning.py:237-346:
 * dropout -> Linear(768 -> num_labels) on the pooled output, single-label cross-entropy; num_labels <= 32).
 * pooled fp32 [B,768]; w [num_labels,768]; logits fp32 [B,num_labels]; labels may be NULL (inference);
 * err_flag (may be NULL) is set to 1 on a label outside [0, num_labels). */
int stk_cls_head_fwd(int device, void* stream, const float* pooled, int B, int num_labels, const float* w,
                     const float* b, const int64_t* labels, float* logits, float* row_loss, int* err_flag);
/* Backward of the head above and of the pooler's tanh.  scale_dev: device scalar = upstream grad / B.
 * dlogit_ws: fp32 workspace [B,num_labels]; dw [num_labels,768], db [num_labels] accumulated (+=);
 * dpre: bf16 [B,768] gradient w.r.t. the pooler's pre-activation. */
int stk_cls_pool_bwd(int device, void* stream, const float* pooled, const float* logits, const int64_t* labels,
                     int B, int num_labels, const float* scale_dev, const float* w, float* dlogit_ws, float* dw,
                     float* db, void* dpre_bf16);

/* ------------------------------------------------------------------------------------------------
 * Input pipeline for pre-tokenised pairs
 * ---------------------------------------------------------------------------------------------- */
/* Joint 512-token inputs of n pairs (reference stonkgs_for_embeddings.py:100-135): text ids / padding mask
 * int32 [n,256] (mask may be NULL = all ones), node indices int32 [n] into the random-walk table
 * walks int32 [num_nodes, walk_len] (index outside [0,num_nodes) = unknown node -> UNK walk), KG half =
 * walk(src) SEP walk(tgt) SEP.  Outputs int64 [n,512]: input_ids, attention_mask, token_type_ids. */
int stk_assemble_pairs(int device, void* stream, const int32_t* text_ids, const int32_t* text_mask,
                       const int32_t* src_node, const int32_t* tgt_node, const int32_t* walks, int num_nodes,
                       int walk_len, int unk_id, int sep_id, int n, int64_t* input_ids, int64_t* attention_mask,
                       int64_t* token_type_ids);
/* replace_mlm_tokens (reference indra_for_pretraining.py:33-77) on both halves of input_ids int64 [n,512], in
 * place: n_pick (= int(256*0.15) = 38) distinct positions per half, 80 % mask_id / 10 % unchanged / 10 % random
 * id below vocab_len (text) or kg_vocab_len (KG); labels int64 [n,256] = original id at picked positions, -100
 * elsewhere.  Philox4x32-10, counter (position, first_row + row, half, step), key = seed. */
int stk_mask_tokens(int device, void* stream, int64_t* input_ids, int64_t* mlm_labels, int64_t* elm_labels, int n,
                    int vocab_len, int kg_vocab_len, int mask_id, int n_pick, uint64_t seed, uint32_t step,
                    int64_t first_row);

/* Data-parallel bucket helpers (the producer / consumer kernels around the NCCL all-reduce that
 * replaces torch DDP's Reducer, reference stonkgs_pretraining.py:147-168,215-223):
 *   pack   = stk_cast_f32_to_bf16 on a slice of the flat gradient buffer (bf16 on the wire)
 *   unpack = dst[i] = float(src[i]) * scale   (scale = 1 / world size: gradient mean) */
int stk_unpack_scale(int device, void* stream, const void* src_bf16, float* dst, int64_t n, float scale);

/* ------------------------------------------------------------------------------------------------
 * Fused clip + AdamW (HF Trainer defaults reached from stonkgs_pretraining.py:171-193)
 * ---------------------------------------------------------------------------------------------- */
/* out[0] += sum(x^2): global gradient norm (caller zeroes out first) */
int stk_sumsq(int device, void* stream, const float* x, int64_t n, float* out);
/* out[0] += sum((scale * x)^2) over a bf16 buffer: the norm of the MEAN gradient taken straight from the all-reduced
 * (summed) bf16 wire buffer of the data-parallel path, scale = 1 / world */
int stk_sumsq_bf16(int device, void* stream, const void* x_bf16, int64_t n, float scale, float* out);

typedef struct StkAdamSeg {
  void* p;        /* fp32 parameter [n] (updated in place) */
  const void* g;  /* fp32 gradient [n] */
  void* m;        /* fp32 exp_avg [n] */
  void* v;        /* fp32 exp_avg_sq [n] */
  void* w16;      /* optional bf16 copy of the updated parameter [n] (GEMM operand), or NULL */
  void* p32_copy; /* optional second fp32 destination (fused q|k|v bias vector), or NULL */
  int64_t n;
  const void* g16; /* optional bf16 gradient [n]: read INSTEAD of g (the all-reduced wire buffer, see grad_scale), or NULL */
} StkAdamSeg;

/* One multi-tensor AdamW step.  segs_dev: device array of segments; chunk_seg_dev / chunk_off_dev:
 * per 65 536-element chunk the segment index and the element offset inside it (n_chunks blocks).
 * sumsq_dev: device scalar with the squared global grad norm (NULL = no clipping);
 * clip coefficient = min(1, max_grad_norm / (sqrt(*sumsq_dev) + 1e-6)) like torch clip_grad_norm_.
 * grad_scale multiplies every gradient first (1 for local gradients; 1 / world when the segments read the summed bf16
 * wire buffer of the data-parallel all-reduce through g16, which fuses DDP's "unpack + mean" into this pass). */
int stk_adamw_step(int device, void* stream, const StkAdamSeg* segs_dev, const int32_t* chunk_seg_dev,
                   const int64_t* chunk_off_dev, int n_chunks, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float bias_correction1, float bias_correction2, const float* sumsq_dev,
                   float max_grad_norm, float grad_scale);

#ifdef __cplusplus
}
#endif
#endif /* STK_H_ */
