/*
 * maven_sm100.h -- C ABI of libmaven_sm100.so: the sm_100a (B200) kernels behind the CLIP contrastive
 * training step of multimodal-supernovae.
 *
 * The reference has no FFI at this boundary: the path sits behind Python nn.Module classes
 * (src/transformer_utils.py, src/models_multimodal.py) and two functions (src/loss.py).  Each entry point
 * below names the reference lines whose arithmetic it replaces.  The drop-in Python modules in
 * multimodal-supernovae_b200/ call these through ctypes from torch.autograd.Function.forward/backward.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; buffers are owned by the caller
 *     (PyTorch), valid for the duration of the call, contiguous, 16-byte aligned.  Kernels never allocate.
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue work and return, they never synchronise.
 *   - return value: 0 success; >0 a cudaError_t; <0 an MVN_E_* code.  mvn_last_error() gives the text
 *     (thread-local).  There is no CPU fallback and no backend dispatch.
 *   - ragged sequences are processed as a PACKED token stream: row m of a [M_cap,E] activation is one token,
 *     sequence b owns rows cu_seqlens[b] .. cu_seqlens[b+1]-1.  The live row count lives on the device
 *     (cu_seqlens[B]); `n_rows_dev` arguments point at it so no host read is needed (CUDA-graph safe).
 *   - `prec`: 0 = fp32 FFMA arithmetic (parity tier 1e-5); 1 = TF32 tensor-core contraction via tcgen05 with
 *     fp32 accumulate (parity tier 1e-3); 2 (whole-encoder calls) = tier 1 arithmetic with the fused block kernels that
 *     keep per-layer intermediates on chip (per-op entry points treat 2 like 1).  Storage is fp32 in all tiers.
 *   - masks are the uint8 storage of torch.bool tensors (0/1).
 */
#ifndef MAVEN_SM100_H
#define MAVEN_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVN_E_BADARG      (-1)   /* null pointer / non-positive size / misaligned */
#define MVN_E_UNSUPPORTED (-2)   /* shape outside what the kernels are built for   */
#define MVN_E_WORKSPACE   (-3)   /* workspace too small                            */

#define MVN_ACT_NONE 0
#define MVN_ACT_RELU 1
#define MVN_ACT_GELU 2           /* exact erf GELU (nn.GELU default) */

#define MVN_AGG_MEAN 0
#define MVN_AGG_MAX  1
#define MVN_AGG_NONE 2           /* agg="pretraining": return the zero-padded (B,T,E) token tensor */

const char* mvn_last_error(void);
int         mvn_abi_version(void);
int         mvn_num_sms(void);   /* SM count of the current device (148 on B200) */
int         mvn_num_slabs(void); /* row slabs every partial-sum workspace is split into (sizes the *_workspace_bytes results) */
long long   mvn_launch_count(void);   /* kernels this process has enqueued through the library */
/* Which arithmetic tier the GEMM-class / attention launches of this process actually ran on: 0 FFMA kernels, 1 tcgen05
 * (tc_gemm / tc_wgrad), 2 warp-MMA attention, 3 fused block kernels (ffn_fused / attn_fused).  A shape the tensor-core kernels
 * do not cover runs on the FFMA kernel and is counted there, so a test can assert that nothing fell back. */
long long   mvn_tier_count(int tier);
void        mvn_tier_reset(void);
/* Per-kernel-class device timing for the roofline leg of bench.py: while a class bit is set, every launch of
 * that class is bracketed by CUDA events on its own stream; mvn_prof_read sums and clears them.
 * classes: 0 GEMM (fwd + input-grad), 1 weight-grad GEMM, 2 attention fwd, 3 attention bwd, 4 row kernels
 * (embed/LN-bwd/pool/reduce), 5 CLIP loss, 6 optimizer, 7 ConvMixer. */
/* Programmatic dependent launch of the persistent kernels on (default) / off; the environment variable MVN_PDL overrides.  Host
 * policy (models_multimodal.py): off when two sequence-encoder chains run concurrently on two streams. */
void        mvn_set_pdl(int on);
void        mvn_prof_enable(unsigned class_mask);
int         mvn_prof_read(int cls, double* total_ms, long long* count);

/* ------------------------------------------------------------------------------------------------
 * Ragged packing.  Replaces the implicit "compute every padded position" of the reference and its
 * per-forward host-built band index (src/transformer_utils.py:219-231).  Bit-exact integer work.
 *   valid_only=1: tokens = positions with mask!=0, in (b,t) order; keyvalid[m]=1.
 *   valid_only=0: tokens = all B*T positions; keyvalid[m]=mask (1 if mask==NULL).
 * cu_seqlens[B+1], tok_src[B*T] (flat b*T+t of each packed row; rows >= cu_seqlens[B] are -1), keyvalid[B*T]. */
int mvn_pack_plan(const uint8_t* mask, int B, int T, int valid_only,
                  int32_t* cu_seqlens, int32_t* tok_src, uint8_t* keyvalid, void* stream);

/* A1+A2: time sin/cos embedding + Linear(1->E) + band embedding, written packed.
 * src/transformer_utils.py:166-176 (TimePositionalEncoding) and :214-231.
 * div_term[E/2] is the fp32 buffer exp(arange(0,E,2)*(-ln(norm)/E)) formed by the caller with the reference
 * expression (parity trap: one ulp in it moves the output). band = t_index / (T/nband). */
int mvn_embed_fwd(const float* x, const float* t, const int32_t* cu_seqlens, const int32_t* tok_src,
                  const float* div_term, const float* w, const float* b, const float* band_emb,
                  int B, int T, int E, int nband, float* out, void* stream);
/* grads of embedding_mag.{weight,bias} and band_emb.weight; overwrites dw[E], db[E], dband[nband*E]. */
int mvn_embed_bwd(const float* x, const int32_t* cu_seqlens, const int32_t* tok_src, const float* dout,
                  int B, int T, int E, int nband, float* dw, float* db, float* dband,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Linear layers (nn.Linear: Y = X W^T + b, W is [N,K] row-major) with fused epilogues.
 * Used for tokeys/toqueries/tovalues (:45-47 as one [3E,E] GEMM), ff.0+ReLU (:101-105), the projection heads
 * (:251; src/models_multimodal.py:278,285,292) and the ConvMixer MLP head (src/models_multimodal.py:82-89). */
int mvn_linear_fwd(const float* X, const float* W, const float* bias, float* Y,
                   const int32_t* n_rows_dev, int M_cap, int N, int K, int act, int prec, void* stream);
/* Y = LayerNorm(X W^T + b + R) * gamma + beta  -- unifyheads+residual+norm1 (:89,:111) and ff.2+residual+norm2
 * (:113-114).  Also writes xhat (normalised, pre-affine) and rstd[M] for the backward.  N in {16,32,64,128}. */
int mvn_linear_res_ln_fwd(const float* X, const float* W, const float* bias, const float* R,
                          const float* gamma, const float* beta, float* Y, float* xhat, float* rstd,
                          const int32_t* n_rows_dev, int M_cap, int N, int K, float eps, int prec, void* stream);
/* dX[M,K] = dY[M,N] W[N,K]  (+ addend[M,K])  then, by dact: 1: *= (act_src>0) ; 2: *= gelu'(act_src). */
int mvn_linear_bwd_input(const float* dY, const float* W, float* dX, const float* addend, const float* act_src,
                         int dact, const int32_t* n_rows_dev, int M_cap, int N, int K, int prec, void* stream);
/* dW[N,K] = dY^T X, db[N] = colsum(dY) (db may be NULL).  accumulate!=0 adds into dW/db. */
int mvn_linear_bwd_weight(const float* dY, const float* X, float* dW, float* db,
                          const int32_t* n_rows_dev, int M_cap, int N, int K, int accumulate,
                          void* workspace, size_t workspace_bytes, int prec, void* stream);
size_t mvn_linear_bwd_weight_workspace_bytes(int M_cap, int N, int K);

/* Feed-forward half of a TransformerBlock in ONE kernel (src/transformer_utils.py:101-105,113-115):
 *   Y = dropout( LayerNorm( X + ff.2( relu( ff.0(X) ) ) ) * gamma + beta )        W1 [4E,E], W2 [E,4E]
 * The hidden activation never reaches HBM.  Also writes xhat (normalised, pre-affine) and rstd[M] for the backward.
 * TF32 warp-MMA contraction, fp32 accumulate.  E in {32, 64}, ff_mult == 4 (MVN_E_UNSUPPORTED otherwise).
 * Dropout: keep-factor of mvn_dropout_scale(seed, site, p) applied to Y (xhat stays pre-dropout); p == 0 switches it off. */
int mvn_ffn_fused_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2,
                      const float* gamma, const float* beta, float* Y, float* xhat, float* rstd,
                      const int32_t* n_rows_dev, int M_cap, int E, int ff_mult, float eps,
                      float dropout_p, uint64_t seed, int site, void* stream);
/* Its backward in ONE kernel (+ the fixed-order reduction of the per-SM partial sums): dropout and LayerNorm backward from
 * (dY, xhat, rstd), h recomputed from X on chip, dX = dz + (dz W2 * relu') W1, and dW1, db1, dW2, db2, dgamma, dbeta. */
size_t mvn_ffn_fused_bwd_workspace_bytes(int E, int ff_mult);
int mvn_ffn_fused_bwd(const float* dY, const float* xhat, const float* rstd, const float* X,
                      const float* W1, const float* b1, const float* W2, const float* gamma,
                      float* dX, float* dW1, float* db1, float* dW2, float* db2, float* dgamma, float* dbeta,
                      const int32_t* n_rows_dev, int M_cap, int E, int ff_mult,
                      float dropout_p, uint64_t seed, int site,
                      void* workspace, size_t workspace_bytes, void* stream);

/* dPre = dY * (H > 0): ReLU backward from the saved activation (MLP heads, src/models_multimodal.py:846-850). */
int mvn_relu_bwd(const float* dY, const float* H, int64_t n, float* dPre, void* stream);

/* LayerNorm backward from saved (xhat, rstd): dZ, and overwrites dgamma[E], dbeta[E]. */
int mvn_layernorm_bwd(const float* dY, const float* xhat, const float* rstd, const float* gamma,
                      float* dZ, float* dgamma, float* dbeta, const int32_t* n_rows_dev, int M_cap, int E,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A3: padding-masked multi-head self-attention on the packed stream.  src/transformer_utils.py:49-87.
 * qkv is [M_cap,3E] = q|k|v; scores = scale * q.k with scale = 1/sqrt(E) passed by the caller (the reference
 * divides q and k by E**0.25 each); masked keys take the reference's -1e7 fill, i.e. contribute exactly 0 when
 * any key is valid and a uniform distribution when none is.  lse is [M_cap,H]. Streaming softmax: the TxT score
 * matrix is never written. */
int mvn_attention_fwd(const float* qkv, const int32_t* cu_seqlens, const uint8_t* keyvalid,
                      float* out, float* lse, int B, int E, int H, float scale, int prec, void* stream);
int mvn_attention_bwd(const float* qkv, const int32_t* cu_seqlens, const uint8_t* keyvalid,
                      const float* out, const float* lse, const float* dout, float* dqkv,
                      int B, int E, int H, float scale, int prec, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A6: masked pooling over each sequence's rows.  src/transformer_utils.py:234-240.
 * mean divides by the valid count (0 valid -> NaN like the reference); max also sees the zeroed padded rows
 * when the sequence is shorter than T.  argmax[B,E] (packed row or -1) is written for agg=max. */
int mvn_pool_fwd(const float* X, const int32_t* cu_seqlens, const uint8_t* keyvalid, int B, int T, int E, int agg,
                 float* pooled, int32_t* argmax, void* stream);
int mvn_pool_bwd(const float* dpooled, const int32_t* cu_seqlens, const uint8_t* keyvalid, const int32_t* argmax,
                 int B, int T, int E, int agg, float* dX, void* stream);
/* packed rows -> zero-padded [B*T,E] (agg="pretraining", :248) and its adjoint gather. */
int mvn_unpack_rows(const float* X, const int32_t* tok_src, const uint8_t* keyvalid, const int32_t* n_rows_dev,
                    int BT, int E, float* out, void* stream);
int mvn_pack_rows(const float* dense, const int32_t* tok_src, const uint8_t* keyvalid, const int32_t* n_rows_dev,
                  int BT, int E, float* X, void* stream);

/* agg="attn" in closed form (src/transformer_utils.py:202-207,241-247): nn.MultiheadAttention(emb, H heads, batch_first) with ONE
 * learnable query over the zero-padded token tensor and NO key mask.  All padded rows are the same key (k = b_k, v = b_v), so a CTA
 * per sequence reads only the valid rows of x [B,T,E] (mask = uint8 view of the bool [B,T]) and accounts for the T - n padded rows as
 * one virtual key of that multiplicity; q / k / v / out projections are applied to the query and to the pooled vectors, never to the
 * B*T tokens.  in_w [3E,E] / in_b [3E] = agg_attn.in_proj_{weight,bias}; out_w / out_b = agg_attn.out_proj.  `saved` (caller-owned,
 * mvn_attn_pool_saved_bytes) carries q, u, c, lse, xbar, o to the backward, which overwrites dx [B,T,E] and every parameter gradient. */
size_t mvn_attn_pool_saved_bytes(int B, int E, int H);
size_t mvn_attn_pool_bwd_workspace_bytes(int B, int E, int H);
int mvn_attn_pool_fwd(const float* x, const unsigned char* mask, const float* query, const float* in_w, const float* in_b,
                      const float* out_w, const float* out_b, int B, int T, int E, int H, float* out, float* saved, size_t saved_bytes,
                      void* stream);
int mvn_attn_pool_bwd(const float* x, const unsigned char* mask, const float* query, const float* in_w, const float* in_b,
                      const float* out_w, const float* saved, const float* dout, int B, int T, int E, int H, float* dx, float* dquery,
                      float* d_in_w, float* d_in_b, float* d_out_w, float* d_out_b, void* workspace, size_t workspace_bytes, void* stream);
/* A6, agg="attn": nn.MultiheadAttention(E, H) with one learnable query per sequence over the zero-padded tokens, no key
 * mask.  src/transformer_utils.py:241-247.  q[E] = in_proj(query) (shared by all sequences), kv[B*T,2E] = k|v after
 * in_proj (padded rows therefore carry the biases, as in the reference); score scale 1/sqrt(E/H).
 * fwd writes out[B,E] (heads concatenated, before out_proj) and probs[B,H,T]; bwd writes dkv[B*T,2E] and dq[E]
 * (summed over the batch); workspace >= B*E floats. */
int mvn_query_pool_fwd(const float* q, const float* kv, int B, int T, int E, int H, float* out, float* probs, void* stream);
int mvn_query_pool_bwd(const float* q, const float* kv, const float* probs, const float* dout, int B, int T, int E, int H,
                       float* dkv, float* dq, void* workspace, size_t workspace_bytes, void* stream);

/* A7: x / ||x||_2 per row, no epsilon.  src/models_multimodal.py:279,286,293. */
int mvn_l2norm_fwd(const float* X, float* Y, float* norm, int B, int D, void* stream);
int mvn_l2norm_bwd(const float* dY, const float* Y, const float* norm, float* dX, int B, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole sequence encoder (A1-A7) as one call: pack, embed, depth x block, pool, projection,
 * <modality>_projection, optional L2-norm.  TransformerWithTimeEmbeddings.forward (:209-253) +
 * {lightcurve,spectral}_embeddings_with_projection (src/models_multimodal.py:281-293).
 * Flat fp32 parameter layout (same layout for `grads`):
 *   emb_w[E] emb_b[E] band_emb[nband*E if nband>1]
 *   depth x { Wq[E*E] Wk[E*E] Wv[E*E] Wu[E*E] bu[E] ln1_g[E] ln1_b[E] W1[4E*E] b1[4E] W2[E*4E] b2[E] ln2_g[E] ln2_b[E] }
 *   proj_w[n_out*E] proj_b[n_out] mproj_w[enc_dim*n_out] mproj_b[enc_dim]
 * Output: [B,enc_dim] (agg mean/max) or [B*T,E] (MVN_AGG_NONE, no projections). */
typedef struct {
    int32_t B, T, E, H, depth, nband, n_out, enc_dim;
    int32_t agg;          /* MVN_AGG_*                                             */
    int32_t normalize;    /* 1: L2-normalise the output rows (CLIP branch)          */
    int32_t prec;         /* 0 fp32, 1 tf32 tensor cores, 2 tf32 + fused block kernels */
    int32_t ff_mult;      /* 4 (Transformer default, src/transformer_utils.py:124)  */
    float   ln_eps;       /* 1e-5                                                   */
    float   dropout_p;    /* nn.Dropout p of Transformer/TransformerBlock (0 = off); counter-based in-kernel mask,
                             regenerated (not stored) by the backward from the same seed */
    uint64_t seed;        /* dropout seed of THIS call (ignored when dropout_p==0)  */
} mvn_seq_cfg;

size_t mvn_seq_param_count(const mvn_seq_cfg* cfg);
size_t mvn_seq_workspace_bytes(const mvn_seq_cfg* cfg);
int mvn_seq_encoder_fwd(const mvn_seq_cfg* cfg, const float* params, const float* div_term,
                        const float* x, const float* t, const uint8_t* mask, float* out,
                        void* workspace, size_t workspace_bytes, void* stream);
int mvn_seq_encoder_bwd(const mvn_seq_cfg* cfg, const float* params, const float* x, const float* dout,
                        float* grads, void* workspace, size_t workspace_bytes, void* stream);

/* The keep/scale factor (0 or 1/(1-p)) mvn_seq_encoder_* applies at dropout site `site` to element [row, col] of the
 * packed [rows, cols] activation: site 0 = transformer input, 1+2l / 2+2l = after norm1 / norm2 of layer l
 * (src/transformer_utils.py:147,112,115).  Lets a caller reproduce or test a dropped-out forward exactly. */
int mvn_dropout_scale(uint64_t seed, int site, float p, int rows, int cols, float* out, void* stream);
/* nn.Dropout as a standalone op (MLP heads src/models_multimodal.py:834-857; TransformerBlock / Transformer used on their own,
 * src/transformer_utils.py:112,115,147): y[r,c] = x[r,c] * keep-factor(seed, site, r, c).  The backward is the same call on the
 * gradient (the mask is regenerated, never stored).  x == y (in place) is allowed. */
int mvn_dropout_apply(const float* x, float* y, int rows, int cols, uint64_t seed, int site, float p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A10/A11: symmetric InfoNCE, streamed; the N x N logits are never written.  src/loss.py:14-38.
 * Z_ij = (e2_i . e1_j) * exp(logit_scale) + logit_bias.   Local rows n (this rank) against all N columns;
 * row_offset is the global index of local row 0 (0 and n==N on one GPU).
 * fwd writes lse_row[n] (LogSoftmax dim=1, rows e2_local), lse_col[n] (dim=0, columns e1_local) and this
 * rank's loss share  (sum_i (lse_row_i - Z_ii) + sum_j (lse_col_j - Z_jj)) / (2N)  to loss_out[0]. */
size_t mvn_clip_loss_workspace_bytes(int n, int N, int D);
int mvn_clip_loss_fwd(const float* e1_local, const float* e2_local, const float* e1_all, const float* e2_all,
                      int n, int N, int D, int row_offset, const float* logit_scale, const float* logit_bias,
                      float* loss_out, float* lse_row, float* lse_col,
                      void* workspace, size_t workspace_bytes, int prec, void* stream);
/* bwd needs both LSE vectors for all N (all-gathered): G = (P_row + P_col - 2I)/(2N);
 * d_e2_local = g*s*G E1, d_e1_local = g*s*G^T E2, d_logit_scale[0] = g * sum_{i local,j} G_ij (Z_ij - b);
 * grad_out points at the upstream scalar g on the device (NULL = 1). */
int mvn_clip_loss_bwd(const float* e1_local, const float* e2_local, const float* e1_all, const float* e2_all,
                      int n, int N, int D, int row_offset, const float* logit_scale, const float* logit_bias,
                      const float* lse_row_all, const float* lse_col_all, const float* grad_out,
                      float* d_e1_local, float* d_e2_local, float* d_logit_scale,
                      void* workspace, size_t workspace_bytes, int prec, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A8/A9: ConvMixer image encoder.  src/models_multimodal.py:38-95.
 *   patch conv (no bias) -> GELU -> BN ; depth x [ x + BN(GELU(dwconv kxk 'same')) ; BN(GELU(conv1x1)) ] ;
 *   global avg-pool -> Linear(dim,1024) -> GELU -> Linear(1024,n_out) -> image_projection Linear(n_out,enc_dim)
 *   -> optional L2 norm.  BatchNorm uses batch statistics when training!=0 (and updates running stats with
 *   momentum 0.1, unbiased variance) and running statistics otherwise.
 * Flat parameter layout (grads use the same):
 *   patch_w[dim*C*p*p] bn0_g[dim] bn0_b[dim]
 *   depth x { dw_w[dim*k*k] dw_b[dim] bnA_g[dim] bnA_b[dim] pw_w[dim*dim] pw_b[dim] bnB_g[dim] bnB_b[dim] }
 *   fc1_w[1024*dim] fc1_b[1024] fc2_w[n_out*1024] fc2_b[n_out] iproj_w[enc_dim*n_out] iproj_b[enc_dim]
 * Flat running-stat layout: (1+2*depth) x { mean[dim] var[dim] } in module order.
 * bn_sync (nullable) lets a data-parallel caller all-reduce the batch statistics between the two halves of
 * a layer: see mvn_convmixer_* in DESIGN.md; single-GPU callers pass world=1. */
typedef struct {
    int32_t B, C, H, W, dim, depth, kernel_size, patch_size, n_out, enc_dim, hidden;
    int32_t normalize, training, prec;
    float   bn_eps, bn_momentum;
    int64_t global_count;   /* B*Hp*Wp summed over ranks (== local count on one GPU) */
    float   dropout_p;      /* nn.Dropout p of ConvMixer (src/models_multimodal.py:62-77,85-87); applied when training!=0.
                               Site s in 1..2*depth follows BatchNorm s, site 1+2*depth follows the head's GELU; the mask of
                               element [row, col] at a site is mvn_dropout_scale(seed, site, p, ...) and is regenerated in the backward */
    uint64_t seed;          /* dropout seed of THIS call (ignored when dropout_p==0) */
} mvn_conv_cfg;

size_t mvn_conv_param_count(const mvn_conv_cfg* cfg);
size_t mvn_conv_workspace_bytes(const mvn_conv_cfg* cfg);
int    mvn_conv_num_bn(const mvn_conv_cfg* cfg);
/* The forward is split at every BatchNorm so a DP caller can all-reduce stats[2*dim] (sum, sumsq) in between:
 * stage s in [0, num_bn] ; stage s computes up to the pre-BN activation of BN #s and its local statistics into
 * stats_local (double[2*dim]); the caller reduces them (identity on one GPU) and passes them as stats_global to
 * stage s+1.  Stage num_bn finishes the head and writes out[B,enc_dim]. */
int mvn_convmixer_fwd_stage(const mvn_conv_cfg* cfg, int stage, const float* params, const float* img,
                            float* running_stats, double* bn_stats /*[num_bn][2*dim] reduced sums*/,
                            float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Backward, split the same way (stage num_bn first, down to 0); bn_stats_bwd[num_bn][2*dim] carries the
 * (sum dy*g, sum dy*g*xhat) reductions a DP caller all-reduces between stages. */
int mvn_convmixer_bwd_stage(const mvn_conv_cfg* cfg, int stage, const float* params, const float* img,
                            const double* bn_stats, double* bn_stats_bwd, const float* dout, float* grads,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1 (next row): NoisyDataLoader.__iter__ on the device.  src/dataloader.py:88-287.
 *   mvn_augment_seq:    out = x + noise * err * level        (:125 "mag + randn_like(mag) * magerr * noise_level_mag", same for spectra)
 *   mvn_image_noise_range: range = max_noise_intensity * std(imgs) (unbiased std over the whole batch, :93)
 *   mvn_augment_images: out[b] = rot90^{rot_k[b]}( img[b] + (2*noise_u - 1) * range )   (:96-112; RandomRotation([a,a]) with
 *                       a = 90*k is exactly the counter-clockwise index permutation torch.rot90(img, k, (1,2)))
 * `noise` / `noise_u` NULL: draw N(0,1) / U[0,1) in-kernel from `seed`; given: the result is bit-identical to the torch
 * expression.  `img` is fp32 [B,C,H,W] (is_u8 = 0) or the raw 8-bit pixels, planar [B,C,H,W] (is_u8 = 1) or in the decoder's
 * [B,H,W,C] order (is_u8 = 2; mvn_augment_images only), converted as float(v)/255 like load_images (:326-331). */
int mvn_augment_seq(const float* x, const float* err, const float* noise, float level, int64_t n, uint64_t seed, float* out, void* stream);
size_t mvn_image_noise_range_workspace_bytes(void);
int mvn_image_noise_range(const void* img, int is_u8, int64_t n, float max_noise_intensity, float* range_out, void* workspace,
                          size_t workspace_bytes, void* stream);
int mvn_augment_images(const void* img, int is_u8, const float* noise_u, const int32_t* rot_k, const float* range_dev, uint64_t seed,
                       int B, int C, int H, int W, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A12 heads: weighted cross-entropy (src/models_multimodal.py:335-349) and MSE (:326).
 * labels are int64 (torch .long()).  loss = sum_i w[y_i] nll_i / sum_i w[y_i]. */
int mvn_weighted_ce_fwd(const float* logits, const int64_t* labels, const float* class_w, int B, int C,
                        float* loss_buf /* [2+2B]: [0]=loss, [1]=sum of weights, rest scratch */, void* stream);
int mvn_weighted_ce_bwd(const float* logits, const int64_t* labels, const float* class_w, int B, int C,
                        const float* loss_buf, const float* grad_out, float* dlogits, void* stream);
int mvn_mse_fwd(const float* pred, const float* target, int n, float* loss, void* stream);
int mvn_mse_bwd(const float* pred, const float* target, int n, const float* grad_out, float* dpred, void* stream);
/* masked-light-curve pretraining objective (src/models_pretraining.py:183-226): nn.MSELoss()(x[mask_pred], x_pred[mask_pred]) as a
 * masked mean over the (B*T) positions; mask = uint8 view of the bool tensor; loss_buf = [loss, number of selected positions]. */
int mvn_masked_mse_fwd(const float* pred, const float* target, const unsigned char* mask, int n, float* loss_buf, void* stream);
int mvn_masked_mse_bwd(const float* pred, const float* target, const unsigned char* mask, int n, const float* loss_buf,
                       const float* grad_out, float* dpred, void* stream);

/* "meta" modality input (src/models_multimodal.py:295-304): out[b] = [ class_emb[cls_b] (half floats) | redshift_b repeated half times ];
 * the backward is the gradient of the embedding table, dclass_emb[c] = sum of dout[b, :half] over rows with cls_b == c (overwritten). */
int mvn_meta_input_fwd(const float* class_emb, const int64_t* cls, const float* redshift, int B, int half, int n_classes,
                       float* out, void* stream);
int mvn_meta_input_bwd(const float* dout, const int64_t* cls, int B, int half, int n_classes, float* dclass_emb, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A13: torch.optim.RAdam step on a flat buffer (coupled L2 weight decay).  src/models_multimodal.py:306-310.
 * Host passes the per-step scalars: bias_correction1 = 1-beta1^t, sqrt_bias_correction2 = sqrt(1-beta2^t),
 * rect = the variance-rectification factor, or a negative number while rho_t <= 5 (un-adapted update). */
int mvn_radam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay,
                   float bias_correction1, float sqrt_bias_correction2, float rect, void* stream);
/* CUDA-graph form of the same update: the step count lives on the device (`step_dev`, incremented by this call) and the
 * step-dependent scalars (bias corrections, rectification) are computed from it on the device into `scalars_dev[3]`, so a
 * captured graph performs step t+1's update at its next replay.  All `n` elements share one step count.  A step over several
 * disjoint segments (parameters without a gradient are skipped, like torch does) passes `step_dev` for the first segment only
 * and NULL for the rest, which then reuse the scalars. */
int mvn_radam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, uint32_t* step_dev, float* scalars_dev, void* stream);
/* Registers a device-resident step counter (or NULL) that every dropout site mixes into its mask hash, in addition to the
 * per-call seed: a CUDA graph freezes the seed argument, the counter (advanced by mvn_radam_step_dev) still changes the masks
 * from replay to replay.  mvn_dropout_scale reads the same counter, so masks stay reproducible for tests. */
void mvn_set_step_counter(const uint32_t* dev_counter);

/* N3: retrieval rank of the true partner, count_i[cos(e1_i,e2_j) > cos(e1_j,e2_j)].  src/utils.py:380-426. */
int mvn_retrieval_ranks(const float* e1, const float* e2, int N, int D, int32_t* ranks, void* stream);
/* N3: the ROC-like curve of get_ROC_data (src/utils.py:395-413): counts[t] = #{sources j : ranks[j] < k_thr[t]}, i.e. how many true
 * partners sit inside the top int(threshold_t * N) of their source's similarity ranking.  k_thr[n_thr] are those integer
 * cut-offs (formed by the caller exactly like the reference: int(threshold * N) in double precision).  counts is overwritten. */
int mvn_retrieval_curve(const int32_t* ranks, int N, const int32_t* k_thr, int n_thr, int32_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAVEN_SM100_H */
