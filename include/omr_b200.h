/*
 * omr_b200.h -- C ABI of libomr_b200.so: the sm_100a kernels behind the
 * image+audio -> **kern transcription model hot path.
 *
 * The reference (mariaalfaroc/omr_a2s_multimodal_transformer) has no FFI or plugin
 * registry: every operator on its hot path is a PyTorch library op called from
 * src/transformer/{encoder,decoder,model}.py.  Each entry point below therefore
 * replaces ONE library-op call site (cited as path:line in the reference tree);
 * the Python module surface that sits on top lives in
 * omr_a2s_multimodal_transformer_b200/{encoder,decoder,model}.py and binds these
 * symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless marked "host";
 *     nothing here allocates, frees or synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); all entry
 *     points are CUDA-graph capturable;
 *   - `dt` selects the storage type of activations/weights: OMR_F32 or OMR_BF16;
 *     accumulation, statistics, losses and parameter gradients are always fp32;
 *   - image-like activations are NHWC ([N,H,W,C], C innermost), sequences are
 *     [B,T,D] row-major;
 *   - return value 0 = ok, negative = error; omr_last_error() gives the text.
 */
#ifndef OMR_B200_H_
#define OMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMR_F32 0
#define OMR_BF16 1

#define OMR_OK 0
#define OMR_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define OMR_ERR_CUDA (-2)    /* a CUDA runtime call failed       */

typedef void* omr_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------- */
int omr_abi_version(void);
const char* omr_last_error(void);
/* 1 when the tcgen05/TMA kernels are compiled in and enabled (env OMR_FORCE_SIMT=1 disables). */
int omr_tensor_core_path_enabled(void);
void omr_set_tensor_core_path(int enabled);
/* number of C-ABI calls served by a tcgen05/TMA kernel since load (tests assert the tensor-core path ran). */
long long omr_tc_call_count(void);
/* number of kernels this library launched since load (bench.py's gpu_launches). */
long long omr_launch_count(void);

/* ---- elementwise / layout ------------------------------------------------------------ */
/* dst[i] = (dst_dt) src[i] */
int omr_cast(int src_dt, int dst_dt, const void* src, void* dst, long long n, omr_stream_t stream);
/* y = relu'(y_saved) * dy  (in place on dy allowed): backward of the fused ReLUs
 * (encoder.py:162-176,220-228; FFN ReLU of nn.TransformerDecoderLayer) */
int omr_relu_bwd(int dt, const void* y, const void* dy, void* dx, long long n, omr_stream_t stream);
/* out = a + b (residual add of the DSC blocks, encoder.py:287-289) */
int omr_add(int dt, const void* a, const void* b, void* out, long long n, omr_stream_t stream);
/* nn.Dropout / nn.Dropout2d (MixDropout, encoder.py:87-104; decoder dropouts p=0.1):
 * y[i] = keep(seed, key(i)) ? x[i] / (1-p) : 0 with key(i) = i (element-wise) or, when channelwise,
 * (i / per_sample) * C + (i % C) (one decision per (sample, channel) of an NHWC tensor).
 * The mask is a pure function of (seed, key): calling it again on dy is the backward.
 * seed_offset (device int32 scalar, may be NULL) is mixed into the seed on the device: a CUDA-graph-captured
 * training step passes its step counter here so that every replay draws fresh masks. */
int omr_dropout(int dt, const void* x, void* y, long long n, int C, long long per_sample, float p, long long seed,
                int channelwise, const int* seed_offset, omr_stream_t stream);
/* Conv2d weight [Co,Ci,3,3] fp32 -> kernel layout in dt (tap-major, channels innermost):
 * transpose == 0: [Co,3,3,Ci] (forward operand); transpose == 1: [Ci,3,3,Co] (data-gradient operand) */
int omr_pack_conv_weight(int dt, const float* w, void* out, int Co, int Ci, int transpose, omr_stream_t stream);
/* depthwise weight [C,1,3,3] fp32 -> [3,3,C] in dt */
int omr_pack_dw_weight(int dt, const float* w, void* out, int C, omr_stream_t stream);

/* ---- convolutional encoders (nn.Conv2d call sites encoder.py:132-150, 56-70) ---------- */
/* y[N,Ho,Wo,Co] = act(conv3x3(x[N,H,W,Ci], w[Co,3,3,Ci], pad 1, stride (sh,sw)) + bias)
 * Ho = ceil(H/sh), Wo = ceil(W/sw); relu != 0 fuses the activation.
 * in_sums (may be NULL): fp64 [N][Co][2] RECEIVES (sum y, sum y^2) over the pixels of every (sample, channel) of the STORED
 * output -- the statistics of the nn.InstanceNorm2d that follows conv2 (encoder.py:151-156,174); pass them to
 * omr_instnorm_fwd(..., sums_ready = 1).  Accumulated in the convolution's epilogue (bf16, Co <= 64), else by one
 * extra pass over y. */
int omr_conv3x3_fwd(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Ci,
                    int Co, int sh, int sw, int relu, double* in_sums, omr_stream_t stream);
/* dx[N,H,W,Ci] = conv3x3 data gradient of dy[N,Ho,Wo,Co].
 * mask (same shape/type as dx, may be NULL): fused backward of the ReLU (and dropout) that PRODUCED this conv's
 * input: dx = mask > 0 ? dx * mask_scale : 0, with mask = the conv's own forward input (a ReLU output, possibly
 * passed through dropout, whose zeros cover both the inactive and the dropped elements) and mask_scale the
 * dropout's 1/(1-p) (1 without dropout).  Saves the separate relu_bwd / dropout-backward passes.
 * colsum (may be NULL): fp32 [Ci] += column sums of the stored dx = the BIAS gradient of the convolution whose ReLU
 * output this conv consumed (its omr_conv3x3_wgrad is then called with db = NULL).
 * in_x + in_bsums (may be NULL; not together with colsum): dx is the gradient of an InstanceNorm OUTPUT, in_x that
 * norm's input (laid out like dx); fp64 in_bsums [N][Ci][2] RECEIVES the raw backward sums (sum dx, sum dx * in_x) for
 * omr_instnorm_bwd(..., sums_ready = 1).  Both are accumulated in the epilogue (bf16, Ci <= 64), else by an extra pass. */
int omr_conv3x3_dgrad(int dt, const void* dy, const void* w, void* dx, int N, int H, int W, int Ci, int Co, int sh,
                      int sw, const void* mask, float mask_scale, float* colsum, const void* in_x, double* in_bsums,
                      omr_stream_t stream);
/* dw[Co,Ci,3,3] (fp32, torch layout) and db[Co] (fp32, may be NULL) ; accumulate != 0 adds to dw/db.
 * ws (may be NULL): caller-owned fp32 scratch of 9*Co*Ci floats, 16-byte aligned; when given, the wide layers reduce their
 * per-CTA partial sums into it with vector reductions and add it to dw in a second small kernel. */
int omr_conv3x3_wgrad(int dt, const void* x, const void* dy, float* dw, float* db, int N, int H, int W, int Ci, int Co,
                      int sh, int sw, int accumulate, float* ws, omr_stream_t stream);
/* depthwise 3x3, stride 1, pad 1 (DepthSepConv2D.depth_conv, encoder.py:56-64) ; w [3,3,C] */
int omr_dwconv3x3_fwd(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int C,
                      omr_stream_t stream);
int omr_dwconv3x3_dgrad(int dt, const void* dy, const void* w, void* dx, int N, int H, int W, int C,
                        omr_stream_t stream);
/* dw [C,1,3,3] fp32, db [C] fp32 */
int omr_dwconv3x3_wgrad(int dt, const void* x, const void* dy, float* dw, float* db, int N, int H, int W, int C,
                        int accumulate, omr_stream_t stream);

/* nn.InstanceNorm2d(eps, affine=False) (encoder.py:151-156, 210-215) on NHWC.
 * stats[N,C,2] fp32 receives (mean, rstd); ws: fp64 scratch of N*C*2 doubles (the plane sums are
 * combined in double so that E[x^2]-E[x]^2 and the backward's mean subtractions do not cancel).
 * sums_ready != 0: ws already holds (sum x, sum x^2) per (n, c) (omr_conv3x3_fwd's in_sums): the statistics pass over x is skipped. */
int omr_instnorm_fwd(int dt, const void* x, void* y, float* stats, double* ws, int N, int HW, int C, float eps,
                     int sums_ready, omr_stream_t stream);
/* dx from dy, the saved INPUT x and stats; ws: fp64 scratch of N*C*2 doubles.
 * relu_mask != 0: x is a ReLU output (possibly after dropout); the result is additionally multiplied by
 * (x > 0 ? mask_scale : 0), i.e. the backward of that ReLU/dropout is fused (see omr_conv3x3_dgrad).
 * sums_ready != 0: ws already holds the RAW sums (sum dy, sum dy * x) per (n, c) (omr_conv3x3_dgrad's in_bsums): the
 * reduction pass over dy and x is skipped.  colsum (may be NULL): fp32 [C] += column sums of the stored dx (the bias
 * gradient of the convolution in front of the norm's ReLU). */
int omr_instnorm_bwd(int dt, const void* dy, const void* x, const float* stats, void* dx, double* ws, int N, int HW,
                     int C, int relu_mask, float mask_scale, int sums_ready, float* colsum, omr_stream_t stream);

/* PositionalEncoding2D + flatten/permute + concat (model.py:45-48, 498, 506, 654):
 * out[b, row_off + p, c] = x[b, p, c] + pe[(p / w) * pe_w + (p % w), c]   for p < h*w
 * x [B,h*w,C] (NHWC encoder output), pe [pe_h,pe_w,C] fp32 (NHWC copy of the `pe` buffer),
 * out has `out_rows` rows per sample (the fused memory [B,S,C]). */
int omr_pe2d_add(int dt, const void* x, const float* pe, void* out, int B, int h, int w, int C, int pe_w, int out_rows,
                 int row_off, omr_stream_t stream);
/* strided row-block copy: dst[b, r, :] = src[b, src_off + r, :], r < rows (backward of concat) */
int omr_copy_rows(int dt, const void* src, void* dst, int B, int rows, int C, int src_rows, int src_off,
                  omr_stream_t stream);

/* ---- masks (decoder.py:150-189, 219-254; model.py:654-672) ---------------------------- */
/* bias[b, seg_off + j] = (j >= lens[b]) ? value : 0   for j < seg_len ; bias is [B,S] fp32.
 * value = -inf reproduces a bool key-padding mask, +1.0 the reference's float 0/1 mask. */
int omr_key_bias_from_lengths(float* bias, const int* lens, int B, int S, int seg_off, int seg_len, float value,
                              omr_stream_t stream);
/* bias[b,t] = (tokens[b,t] == pad_id) ? value : 0  (tgt_pad_mask = (tgt == 0).float()) */
int omr_key_bias_from_tokens(float* bias, const long long* tokens, long long n, long long pad_id, float value,
                             omr_stream_t stream);

/* ---- decoder ------------------------------------------------------------------------- */
/* nn.Embedding + PositionalEncoding1D (decoder.py:73-82,124,29-32):
 * out[b,t,:] = table[tok[b,t],:] + pe[pos0 + t,:] ; table in dt, pe fp32 [max_len,D].
 * pos_dev (device int32 scalar, may be NULL) overrides pos0 so that a decode step can be replayed
 * from a CUDA graph; the same convention holds for omr_kv_append / omr_attn_decode / omr_argmax_step. */
int omr_embed_pe_fwd(int dt, const long long* tokens, const void* table, const float* pe, void* out, int B, int T, int D,
                     int pos0, const int* pos_dev, omr_stream_t stream);
/* dtable[tok,:] += dout[b,t,:] for tok != padding_idx ; dtable fp32 [V,D] (must be pre-zeroed or accumulated) */
int omr_embed_bwd(int dt, const long long* tokens, const void* dout, float* dtable, long long rows, int D,
                  long long padding_idx, omr_stream_t stream);

/* General (batched) GEMM: C[b] = act(opA(A[b]) * opB(B[b]) + bias) (+ C[b] if accumulate)
 *   opA(A) is M x K: transA == 0 -> A[m*lda + k], else A[k*lda + m]
 *   opB(B) is K x N: transB == 0 -> B[k*ldb + n], else B[n*ldb + k]
 *   in_dt: storage type of A and B; out_dt: storage type of C (OMR_F32 for parameter gradients)
 *   bias fp32; bias_mode 0 none, 1 per column n, 2 per row m.
 * Replaces nn.Linear / packed in-proj / Conv2d 1x1 / Conv1d k=1 (decoder.py:86-102,
 * encoder.py:65-70) and their backward GEMMs. */
int omr_gemm(int in_dt, int out_dt, int transA, int transB, int M, int N, int K, const void* A, long long lda,
             long long strideA, const void* B, long long ldb, long long strideB, void* C, long long ldc,
             long long strideC, int batch, const float* bias, int bias_mode, int relu, int accumulate,
             omr_stream_t stream);
/* out[n] (+)= sum_r x[r, n] ; x [rows, ld] in dt, out fp32 (bias gradients) */
int omr_colsum(int dt, const void* x, long long rows, int N, long long ld, float* out, int accumulate,
               omr_stream_t stream);

/* Scaled-dot-product attention with the reference's mask algebra
 * (torch F.multi_head_attention_forward as driven by decoder.py:128-142 and model.py:327):
 *   S = scale * Q K^T + key_bias[b, k] ; excluded pairs get -inf ; P = softmax(S) ; O = P V
 *   causal != 0: key k visible to query t iff k <= t + (Tk - Tq); window > 0 additionally requires
 *   k >= t + (Tk - Tq) - window (create_variable_window_mask, decoder.py:191-217).
 *   key_bias: fp32 [B, Tk] or NULL: 0, +1.0 (the reference's additive float padding masks) or -inf
 *   (its bool masks).
 *   q_len / kv_len (int32 [B]) or NULL: CrossAttention.create_attention_mask (model.py:329-355):
 *   pair (t,k) excluded iff t >= q_len[s] && k >= kv_len[s], with s = b when quirk_mod == 0 and
 *   s = (b*H + h) % quirk_mod otherwise -- the reference tiles its [B,Tq,Tk] mask head-major
 *   (`repeat(num_heads,1,1)`) while torch indexes it batch-major, so quirk_mod = B reproduces it.
 * Layout: element (b, t, h, d) of Q lives at q[b*q_bs + t*q_rs + h*hd + d] (same for k, v, o with
 * their strides), so packed in-proj outputs are consumed in place.
 * lse [B,H,Tq] fp32 receives log-sum-exp (natural log) for the backward. */
/* Attention-probability dropout (the `dropout` of nn.MultiheadAttention as built at decoder.py:86-95, and of the
 * mixers' nn.MultiheadAttention, model.py:292-297; train mode only): arms the NEXT omr_attn_fwd or omr_attn_bwd call
 * (one-shot), which then computes O = (P o M / (1-p)) V with a keep mask M that is a pure function of (seed [+ the
 * device int32 *seed_off], batch*head, query, key) -- the backward regenerates it from the same arguments.  p in [0,1).
 * Threading contract: the armed state is PER CALLING THREAD (thread_local); arm and consume on the same thread.  */
int omr_attn_next_dropout(float p, unsigned int seed, const int* seed_off);
int omr_attn_fwd(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                 const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse,
                 const float* key_bias, int B, int H, int Tq, int Tk, int hd, float scale, int causal, int window,
                 const int* q_len, const int* kv_len, int quirk_mod, omr_stream_t stream);
/* Backward: dq/dk/dv use the q/k/v strides of their own (dq_bs, ...).  delta_ws: fp32 scratch of
 * B*H*Tq*65 + 4 floats (delta [B,H,Tq], then the fp32 dQ accumulators [B,H,Tq,64] of the tensor-core
 * kernel, which adds the contributions of the key tiles with vector atomics).  dq/dk/dv are written
 * (not accumulated). */
int omr_attn_bwd(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                 const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                 const void* dout, long long do_bs, long long do_rs, const float* lse, void* dq, long long dq_bs,
                 long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv, long long dv_bs,
                 long long dv_rs, float* delta_ws, const float* key_bias, int B, int H, int Tq, int Tk, int hd,
                 float scale, int causal, int window, const int* q_len, const int* kv_len, int quirk_mod,
                 omr_stream_t stream);

/* s = x + res ; y = LayerNorm(s) * gamma + beta (post-norm residual blocks of
 * nn.TransformerDecoderLayer, eps 1e-5).  s_out (dt) and stats[rows,2] (mean, rstd; fp32) are
 * saved for the backward; res may be NULL. */
int omr_add_layernorm_fwd(int dt, const void* x, const void* res, const float* gamma, const float* beta, void* s_out,
                          void* y, float* stats, long long rows, int D, float eps, omr_stream_t stream);
/* ds (gradient wrt s = x + res), dgamma/dbeta fp32 [D] (accumulated into, caller zeroes) */
int omr_layernorm_bwd(int dt, const void* dy, const void* s, const float* stats, const float* gamma, void* ds,
                      float* dgamma, float* dbeta, long long rows, int D, omr_stream_t stream);

/* Opt-in fused forms of the residual blocks of nn.TransformerDecoderLayer in TRAIN mode (decoder.py:86-95:
 * x = norm(x + dropout(sublayer(x)))), shortening the decoder's kernel chain (OMR_FUSE_DECODER_LINKS=1):
 *   omr_dropout_add_layernorm_fwd : s = dropout(x; p, seed) + res ; y = LayerNorm(s)        (= omr_dropout + omr_add_layernorm_fwd)
 *   omr_layernorm_bwd_dropout     : ds as omr_layernorm_bwd ; da = dropout'(ds; p, seed)    (= omr_layernorm_bwd + omr_dropout);
 *                                   dbias (fp32 [D], may be NULL) += column sums of da = the bias gradient of the linear
 *                                   layer whose output went through that dropout (ABI 3)
 *   omr_mask_scale                : dx = (mask > 0 ? scale : 0) * dx  -- backward of relu followed by dropout, with
 *                                   mask = the dropped ReLU output (FFN of the decoder layer)
 * The keep decision is exactly omr_dropout's element-wise one (same seed / seed_offset convention). */
int omr_dropout_add_layernorm_fwd(int dt, const void* x, const void* res, const float* gamma, const float* beta, void* s_out,
                                  void* y, float* stats, long long rows, int D, float eps, float p, long long seed,
                                  const int* seed_offset, omr_stream_t stream);
int omr_layernorm_bwd_dropout(int dt, const void* dy, const void* s, const float* stats, const float* gamma, void* ds,
                              void* da, float* dgamma, float* dbeta, long long rows, int D, float p, long long seed,
                              const int* seed_offset, float* dbias, omr_stream_t stream);
int omr_mask_scale(int dt, void* dx, const void* mask, float scale, long long n, omr_stream_t stream);

/* Softmax cross-entropy over the vocabulary (CrossEntropyLoss(ignore_index), model.py:109,444):
 * logits [rows, ld] in dt (class index innermost), targets int64 [rows].
 * row_loss[r] = lse_r - logit[r, target_r] (0 for ignored rows); row_lse[r] = lse_r. */
int omr_ce_fwd(int dt, const void* logits, long long ld, const long long* targets, long long rows, int V,
               long long ignore_index, float* row_loss, float* row_lse, omr_stream_t stream);
/* loss_out[0] = sum(row_loss over valid rows) / n_valid ; loss_out[1] = n_valid */
int omr_ce_reduce(const float* row_loss, const long long* targets, long long rows, long long ignore_index,
                  float* loss_out, omr_stream_t stream);
/* dlogits[r, v] = (softmax(logits)[r,v] - [v == target_r]) * gscale[0] / n_valid  (0 for ignored rows);
 * may be in place (dlogits == logits).  gscale: device fp32 scalar (upstream gradient);
 * loss_out as written by omr_ce_reduce supplies n_valid. */
int omr_ce_bwd(int dt, const void* logits, long long ld, const long long* targets, const float* row_lse,
               const float* loss_out, const float* gscale, void* dlogits, long long rows, int V,
               long long ignore_index, omr_stream_t stream);

/* Classifier FUSED with the softmax cross-entropy (out_layer of decoder.py:145-146 + CrossEntropyLoss of
 * model.py:109,166,444,588) -- the [rows, V] logits are never written (bf16 tensor-core path, D = 256 only):
 * x [rows, D] (row stride x_ld), w [V, D] (row stride w_ld), bias fp32 [V] or NULL, targets int64 [rows].
 * omr_proj_ce_supported: 1 if the fused kernels serve (dt, D) on this build / device setting, else 0 (the caller then
 * uses omr_gemm + omr_ce_fwd / omr_ce_bwd).  forward: row_loss / row_lse as omr_ce_fwd, follow with omr_ce_reduce.
 * backward: dx [rows, D] = dL/dx (dt) on one call; dw [V, D] / db [V] (fp32, ACCUMULATED into, db may be NULL) on the
 * other -- two entry points so that the weight gradient can run on a side stream. */
int omr_proj_ce_supported(int dt, int D);
int omr_proj_ce_fwd(int dt, const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                    const long long* targets, long long rows, int V, int D, long long ignore_index, float* row_loss,
                    float* row_lse, omr_stream_t stream);
int omr_proj_ce_bwd_dx(int dt, const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                       const long long* targets, const float* row_lse, const float* loss_out, const float* gscale,
                       long long rows, int V, int D, long long ignore_index, void* dx, long long dx_ld,
                       omr_stream_t stream);
int omr_proj_ce_bwd_dw(int dt, const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                       const long long* targets, const float* row_lse, const float* loss_out, const float* gscale,
                       long long rows, int V, int D, long long ignore_index, float* dw, float* db, omr_stream_t stream);

/* ---- optimizer (torch.optim.Adam(lr=1e-4), model.py:134-139,475-483) ------------------- */
/* One fused multi-tensor Adam step.  `table` is a device array of n_tensors omr_adam_entry;
 * `step` is a device int32 scalar that the kernel reads (the host increments it via
 * omr_adam_tick so the whole step is graph-capturable). */
typedef struct omr_adam_entry {
  float* param;      /* fp32 master [n]                                                   */
  const float* grad; /* fp32 [n] (NULL -> tensor skipped)                                 */
  float* exp_avg;    /* fp32 [n]                                                          */
  float* exp_avg_sq; /* fp32 [n]                                                          */
  void* shadow;      /* bf16 working copy in kernel layout `layout`, or NULL              */
  void* shadow2;     /* second bf16 working copy in `layout2`, or NULL                    */
  long long n;
  int layout;        /* 0 same order; 1 [Co,Ci,3,3]->[Co,3,3,Ci]; 2 [C,1,3,3]->[3,3,C];
                        3 [Co,Ci,3,3]->[Ci,3,3,Co]; 4 [R,C]->[C,R]                        */
  int layout2;
  int d0;            /* Co (layouts 1,3) / C (layout 2) / R (layout 4)                    */
  int d1;            /* Ci (layouts 1,3) / C (layout 4)                                   */
} omr_adam_entry;
int omr_adam_tick(int* step, omr_stream_t stream); /* *step += 1 */
int omr_adam_step(const omr_adam_entry* table, int n_tensors, long long max_n, const int* step, double lr,
                  double beta1, double beta2, double eps, double grad_scale, omr_stream_t stream);

/* ---- greedy decode helpers (model.py:170-199, 592-617) -------------------------------- */
/* first-max argmax over V of logits [B, ld] (dt); tok int64 [B], val fp32 [B] (the raw logit that
 * get_pred_seq_and_pred_prob_seq reports, model.py:253-256).
 * If finished != NULL: rows with finished[b] != 0 emit `pad_id` and are left finished; rows that
 * emit `eos_id` become finished.  out_tokens[b, step] = tok when out_tokens != NULL. */
int omr_argmax_step(int dt, const void* logits, long long ld, int B, int V, long long* tok, float* val, int* finished,
                    long long eos_id, long long pad_id, long long* out_tokens, float* out_vals, int out_ld, int step,
                    const int* step_dev, omr_stream_t stream);
/* copy the new key/value rows of a decode step into the KV cache:
 * cache[b, pos, :] = src[b, :] for `width` elements ; src row stride src_rs, cache [B, Tmax, width] */
int omr_kv_append(int dt, const void* src, long long src_rs, void* cache, int B, int Tmax, int width, int pos,
                  const int* pos_dev, omr_stream_t stream);
/* Single-query attention over a KV cache (one decode step; the KV-cached equivalent of re-running
 * the decoder on the growing prefix, model.py:184-186): for every (b,h)
 *   o[b, h*hd:] = softmax(scale * q[b,h] . K[b, j, h] + key_bias[b, j]) V[b, j, h],  j in [j_lo, Tk)
 * with j_lo = max(0, Tk-1-window) when window > 0 (decoder.py:191-217), else 0.
 * q element (b,h,d) at q[b*q_bs + h*hd + d]; K element (b,j,h,d) at k[b*k_bs + j*k_rs + h*hd + d].
 * pos_dev != NULL: the live key count is *pos_dev + 1 and Tk only bounds it (graph replay).
 * ws: fp32 scratch, at least B*H*nsplit*(hd+2) floats with nsplit <= max(1, ceil(1184/(B*H))). */
int omr_attn_decode(int dt, const void* q, long long q_bs, const void* k, long long k_bs, long long k_rs,
                    const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, const float* key_bias,
                    long long kb_bs, float* ws, long long ws_floats, int B, int H, int Tk, int hd, float scale,
                    int window, const int* pos_dev, omr_stream_t stream);

/* The whole greedy decoder as ONE persistent kernel (a 4-CTA cluster per sample, CTA r = head r): `nsteps` complete
 * decode steps (embedding + PE, L post-norm layers with KV-cached self-attention and cross-attention over the
 * pre-projected memory, classifier, first-max argmax, EOS bookkeeping) per launch, three cluster exchanges per layer.
 * ABI v5: the K/V caches are HEAD-MAJOR, [B][H][2][rows][D/H]: the K rows of one head are one contiguous block, its V rows
 * the next (a CTA streams exactly one head).  self_kv must be zero-initialised.
 * ABI v4: w_o, wc_o and w2 -- the projections that follow an attention head / the FFN quarter and are split along
 * their reduction index inside the kernel -- are passed as COLUMN SLICES [4][D][D/4] (slice r = W[:, r*D/4:(r+1)*D/4],
 * contiguous).  dt = fp32: slices and all other matrices row-major.  dt = bf16: every matrix (w_in, wc_q, w1, w_out and
 * each column slice) in mma.sync A-FRAGMENT ORDER: rows padded with zeros to a multiple of 32, then
 * [rows/16][K/32][2][8][4][8] = (16-row tile, 32-wide k-block, row half hf, row g, lane quarter t, element e) holding
 * W[16*tile + 8*hf + g][32*kb + 8*t + e] (K = D, or D/4 for a slice).
 * Replaces the per-token Python loop of model.py:184-193 / 602-611.  `layers` is a DEVICE array of L
 * omr_decode_layer records (weights in `dt`, biases / LayerNorm affine in fp32).  State (tok, val, finished,
 * out_tokens/out_vals [B,out_ld], *pos) stays on the device across launches; *pos advances by the steps executed.
 * scratch: fp32, at least omr_decode_persistent_scratch_floats(B, H, D, V) floats.  B <= 64, D = 256, head dim 64. */
typedef struct omr_decode_layer {
  const void* w_in;  const float* b_in;   /* self-attn packed in-proj [3D,D], [3D]              */
  const void* w_o;   const float* b_o;    /* self-attn out-proj, column slices [4][D][D/4], [D] */
  const void* wc_q;  const float* bc_q;   /* cross-attn query rows of the packed in-proj [D,D]   */
  const void* wc_o;  const float* bc_o;   /* cross-attn out-proj, column slices [4][D][D/4]     */
  const void* w1;    const float* b1;     /* linear1 [D,D] (ff_dim == D)                        */
  const void* w2;    const float* b2;     /* linear2, column slices [4][D][D/4]                 */
  const float *g1, *be1, *g2, *be2, *g3, *be3; /* norm1..3 weight / bias                        */
  void* self_kv;                          /* [B,H,2,Tmax,D/H] head-major cache in dt (written)  */
  const void* cross_kv;                   /* [B,H,2,S,D/H] head-major projected memory in dt    */
} omr_decode_layer;
long long omr_decode_persistent_scratch_floats(int B, int H, int D, int V);
int omr_decode_persistent(int dt, const void* layers, int L, const void* emb, const float* pe, const void* w_out,
                          const float* b_out, int B, int H, int D, int V, int S, int Tmax, int nsteps, int window,
                          long long* tok, float* val, int* finished, long long* out_tokens, float* out_vals, int out_ld,
                          int* pos, long long eos, long long pad, const float* mem_bias, long long mem_bias_bs,
                          float ln_eps, float* scratch, long long scratch_floats, long long* timing,
                          omr_stream_t stream);
/* timing (device int64 [16], may be NULL): SM-clock cycles spent by CTA 0 per phase kind, accumulated over the steps
 * (0 embed, 1 qkv, 2 self-attn, 3 out-proj, 4 cross-q, 5 cross-attn, 6 cross-out, 7 ffn1, 8 ffn2, 9 classifier, 10 argmax). */

/* ---- the steps either side of the model ------------------------------------------ */

/* Batch collation on the device (reference pad_batch_inputs, src/data/preprocessing.py:55-74, and get_number_of_frames,
 * src/data/ar_dataset.py:439-442).  The B ragged single-channel fp32 samples lie back to back in `flat`
 * (sample b = heights[b] x widths[b] row-major at flat + offsets[b]); dst [B,1,Hmax,Wmax] fp32 receives the sample
 * in its top-left corner and pad_value elsewhere (1.0 for score images, 0.0 for spectrograms,
 * preprocessing.py:118-136).  n_frames (may be NULL): n_frames[b] = ceil(h/height_reduction) * ceil(w/width_reduction). */
int omr_pad_collate(const float* flat, const long long* offsets, const int* heights, const int* widths, float* dst, int B,
                    int Hmax, int Wmax, float pad_value, int* n_frames, int height_reduction, int width_reduction,
                    omr_stream_t stream);
/* pad_batch_transcripts of transcript[:-1] / transcript[1:] (preprocessing.py:77-100): transcript b =
 * flat[offsets[b] .. offsets[b+1]) (offsets has B+1 entries); y_in, y_out int64 [B,T] with T = max length - 1. */
int omr_pad_transcripts(const long long* flat, const long long* offsets, int B, int T, long long* y_in, long long* y_out,
                        long long pad_id, omr_stream_t stream);
/* One step of token-level late fusion (weighted_prediction, src/multimodal/weighted_multimodal/test.py:47-70):
 * p = alpha * softmax(logits_a[b]) + (1 - alpha) * softmax(logits_b[b]) (fp32), tok[b] = first-max argmax of p,
 * val[b] = that probability; finished / eos / pad / out_* / step_dev exactly as omr_argmax_step. */
int omr_mix_argmax_step(int dt, const void* logits_a, long long lda, const void* logits_b, long long ldb, int B, int V,
                        float alpha, long long* tok, float* val, int* finished, long long eos_id, long long pad_id,
                        long long* out_tokens, float* out_vals, int out_ld, int step, const int* step_dev,
                        omr_stream_t stream);
/* Levenshtein distance (unit costs) of `pairs` (truth, hypothesis) token-id sequences, sequence p =
 * tokens[offsets[p] .. offsets[p+1]) (compute_ed_metrics, src/utils/metrics.py:52-88).  ed_out int32 [pairs];
 * sums (may be NULL; must be zeroed by the caller) accumulates {sum of distances, sum of truth lengths, number of pairs
 * with distance > 0}: Sym-ER = 100*sums[0]/sums[1], Seq-ER = 100*sums[2]/pairs.  max_len >= every truth length. */
int omr_levenshtein(const long long* truth, const long long* truth_offsets, const long long* hyp,
                    const long long* hyp_offsets, int pairs, int max_len, int* ed_out, unsigned long long* sums,
                    omr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OMR_B200_H_ */
