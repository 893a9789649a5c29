// dispatch.cu -- C-ABI entry points that choose between the tcgen05/TMA kernels (bf16, eligible
// shapes) and the exact-fp32 CUDA-core kernels.  There is no CPU fallback anywhere.
#include <stdlib.h>

#include "common.cuh"
#include "attn_drop.cuh"
#include "kernels.h"

static int g_tc_enabled = -1;
static bool tc_enabled() {
  if (g_tc_enabled < 0) {
    const char* e = getenv("OMR_FORCE_SIMT");
    g_tc_enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return g_tc_enabled == 1;
}
static long long g_tc_calls = 0;
extern "C" long long omr_tc_call_count(void) { return g_tc_calls; }
#define TC_TRY(call)                        \
  do {                                      \
    int rc__ = (call);                      \
    if (rc__ != OMR_TC_NOT_ELIGIBLE) {      \
      if (rc__ == OMR_OK) ++g_tc_calls;     \
      return rc__;                          \
    }                                       \
  } while (0)
extern "C" int omr_tensor_core_path_enabled(void) { return tc_enabled() ? 1 : 0; }
extern "C" void omr_set_tensor_core_path(int enabled) { g_tc_enabled = enabled ? 1 : 0; }

extern "C" int omr_gemm(int in_dt, int out_dt, int transA, int transB, int M, int N, int K, const void* A,
                        long long lda, long long strideA, const void* B, long long ldb, long long strideB, void* C,
                        long long ldc, long long strideC, int batch, const float* bias, int bias_mode, int relu,
                        int accumulate, omr_stream_t stream) {
  OMR_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "omr_gemm: negative dimension");
  if (M == 0 || N == 0 || batch == 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  if (tc_enabled() && in_dt == OMR_BF16) {
    if (batch == 1) {
      TC_TRY(omr_gemm_tc(out_dt, transA, transB, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, 1, bias,
                           bias_mode, relu, accumulate, st));
    } else {
      // batched problems (the [B,V,T] logits of the public forward(), reference decoder.py:145-146): one tensor-core launch
      // per batch element; element 0 decides eligibility for all (same shape, strides are multiples of the alignment or not)
      const size_t esz_c = out_dt == OMR_F32 ? 4 : 2;
      int rc0 = OMR_TC_NOT_ELIGIBLE;
      const bool aligned = ((strideA * 2) % 16 == 0) && ((strideB * 2) % 16 == 0);
      if (aligned)
        rc0 = omr_gemm_tc(out_dt, transA, transB, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, bias, bias_mode, relu, accumulate, st);
      if (rc0 != OMR_TC_NOT_ELIGIBLE) {
        if (rc0) return rc0;
        for (int i = 1; i < batch; ++i) {
          int rc = omr_gemm_tc(out_dt, transA, transB, M, N, K, (const char*)A + (size_t)i * strideA * 2, lda, 0,
                               (const char*)B + (size_t)i * strideB * 2, ldb, 0, (char*)C + (size_t)i * strideC * esz_c, ldc, 0, 1, bias,
                               bias_mode, relu, accumulate, st);
          if (rc) {
            OMR_REQUIRE(rc != OMR_TC_NOT_ELIGIBLE, "omr_gemm: batch element %d not eligible for the tensor-core kernel after element 0 was", i);
            return rc;
          }
        }
        ++g_tc_calls;
        return OMR_OK;
      }
    }
  }
  return omr_gemm_simt(in_dt, out_dt, transA, transB, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch,
                       bias, bias_mode, relu, accumulate, st);
}

// Fused outputs (in_sums / colsum / in_bsums) are produced in the tcgen05 epilogue when the kernel takes them (bf16, <= 64
// output channels) and by a separate pass over the stored result otherwise -- same contents either way.
static int conv_fwd_plain(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Ci, int Co,
                          int sh, int sw, int relu, cudaStream_t st) {
  if (Ci == 1) {  // first layer: K = 9, output-write bound streaming kernel (both dtypes)
    int rc1 = omr_conv3x3_fwd_c1(dt, x, w, bias, y, N, H, W, Co, sh, sw, relu, st);
    if (rc1 != OMR_TC_NOT_ELIGIBLE) return rc1;
  }
  if (tc_enabled() && dt == OMR_BF16) {
    TC_TRY(omr_conv3x3_fwd_tc(x, w, bias, y, N, H, W, Ci, Co, sh, sw, relu, nullptr, st));
  }
  return omr_conv3x3_fwd_simt(dt, x, w, bias, y, N, H, W, Ci, Co, sh, sw, relu, st);
}

extern "C" int omr_conv3x3_fwd(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W,
                               int Ci, int Co, int sh, int sw, int relu, double* in_sums, omr_stream_t stream) {
  OMR_REQUIRE(N >= 0 && H > 0 && W > 0 && Ci > 0 && Co > 0 && sh > 0 && sw > 0, "omr_conv3x3_fwd: bad shape");
  cudaStream_t st = as_stream(stream);
  if (!in_sums) return conv_fwd_plain(dt, x, w, bias, y, N, H, W, Ci, Co, sh, sw, relu, st);
  const int Ho = (H + sh - 1) / sh, Wo = (W + sw - 1) / sw;
  if (tc_enabled() && dt == OMR_BF16 && Ci > 1 && Co <= 64) {
    OMR_CUDA(cudaMemsetAsync(in_sums, 0, sizeof(double) * (size_t)N * Co * 2, st));
    TC_TRY(omr_conv3x3_fwd_tc(x, w, bias, y, N, H, W, Ci, Co, sh, sw, relu, in_sums, st));
  }
  int rc = conv_fwd_plain(dt, x, w, bias, y, N, H, W, Ci, Co, sh, sw, relu, st);
  if (rc) return rc;
  return omr_in_partial_sums(dt, 0, y, nullptr, in_sums, N, Ho * Wo, Co, st);
}

extern "C" int omr_conv3x3_dgrad(int dt, const void* dy, const void* wT, void* dx, int N, int H, int W, int Ci, int Co,
                                 int sh, int sw, const void* mask, float mask_scale, float* colsum, const void* in_x,
                                 double* in_bsums, omr_stream_t stream) {
  OMR_REQUIRE(N >= 0 && H > 0 && W > 0 && Ci > 0 && Co > 0 && sh > 0 && sw > 0, "omr_conv3x3_dgrad: bad shape");
  OMR_REQUIRE(!(colsum && in_bsums), "omr_conv3x3_dgrad: colsum and in_bsums are mutually exclusive");
  OMR_REQUIRE(!in_bsums || in_x, "omr_conv3x3_dgrad: in_bsums needs in_x");
  cudaStream_t st = as_stream(stream);
  const bool fused = (colsum || in_bsums);
  if (tc_enabled() && dt == OMR_BF16) {
    if (fused) {  // sums from the kernel's own epilogue when it takes them ...
      if (in_bsums) OMR_CUDA(cudaMemsetAsync(in_bsums, 0, sizeof(double) * (size_t)N * Ci * 2, st));
      TC_TRY(omr_conv3x3_dgrad_tc(dy, wT, dx, N, H, W, Ci, Co, sh, sw, mask, mask_scale, colsum, in_x, in_bsums, st));
    }
    // ... else (or without sums) the plain tensor-core data gradient, followed by one pass over the stored result
    int rc = omr_conv3x3_dgrad_tc(dy, wT, dx, N, H, W, Ci, Co, sh, sw, mask, mask_scale, nullptr, nullptr, nullptr, st);
    if (rc != OMR_TC_NOT_ELIGIBLE) {
      if (rc) return rc;
      ++g_tc_calls;
      if (colsum) return omr_colsum(dt, dx, (long long)N * H * W, Ci, Ci, colsum, 1, stream);
      if (in_bsums) return omr_in_partial_sums(dt, 2, dx, in_x, in_bsums, N, H * W, Ci, st);
      return OMR_OK;
    }
  }
  int rc = omr_conv3x3_dgrad_simt(dt, dy, wT, dx, N, H, W, Ci, Co, sh, sw, st);
  if (rc) return rc;
  if (mask) {
    rc = omr_relu_mask_scale(dt, dx, mask, mask_scale, (long long)N * H * W * Ci, st);  // CUDA-core path: separate pass
    if (rc) return rc;
  }
  if (colsum) return omr_colsum(dt, dx, (long long)N * H * W, Ci, Ci, colsum, 1, stream);
  if (in_bsums) return omr_in_partial_sums(dt, 2, dx, in_x, in_bsums, N, H * W, Ci, st);
  return OMR_OK;
}

extern "C" int omr_conv3x3_wgrad(int dt, const void* x, const void* dy, float* dw, float* db, int N, int H, int W,
                                 int Ci, int Co, int sh, int sw, int accumulate, float* ws, omr_stream_t stream) {
  OMR_REQUIRE(N >= 0 && H > 0 && W > 0 && Ci > 0 && Co > 0 && sh > 0 && sw > 0, "omr_conv3x3_wgrad: bad shape");
  cudaStream_t st = as_stream(stream);
  int Ho = (H + sh - 1) / sh, Wo = (W + sw - 1) / sw;
  if (db) {
    int rc = omr_colsum(dt, dy, (long long)N * Ho * Wo, Co, Co, db, accumulate, stream);
    if (rc) return rc;
  }
  if (Ci == 1) {  // first layer: K = 9, HBM-bound streaming kernel (both dtypes)
    int rc1 = omr_conv3x3_wgrad_c1(dt, x, dy, dw, N, H, W, Co, sh, sw, accumulate, st);
    if (rc1 != OMR_TC_NOT_ELIGIBLE) return rc1;
  }
  if (tc_enabled() && dt == OMR_BF16) {
    static int small_path = -1;  // narrow stride-1 layers: ldmatrix + mma.sync kernel (OMR_WGRAD_SMALL=0 disables)
    if (small_path < 0) {
      const char* e = getenv("OMR_WGRAD_SMALL");
      small_path = (e && e[0] == '0') ? 0 : 1;
    }
    if (small_path) TC_TRY(omr_conv3x3_wgrad_small(x, dy, dw, N, H, W, Ci, Co, sh, sw, accumulate, st));
    TC_TRY(omr_conv3x3_wgrad_tc(x, dy, dw, N, H, W, Ci, Co, sh, sw, accumulate, ws, st));
  }
  return omr_conv3x3_wgrad_simt(dt, x, dy, dw, N, H, W, Ci, Co, sh, sw, accumulate, st);
}

// ---- attention-probability dropout: one-shot state consumed by the next omr_attn_fwd / omr_attn_bwd OF THE CALLING
// THREAD (thread_local: autograd runs backward on another thread than forward, and a second model / a validation
// loop in another thread must not steal or mis-arm the (p, seed) pair; contract stated in include/omr_b200.h) ----
namespace {
thread_local AttnDrop g_drop_next = {0, nullptr, 0, 1.f};
thread_local AttnDrop g_drop_cur = {0, nullptr, 0, 1.f};
void take_dropout() {
  g_drop_cur = g_drop_next;
  g_drop_next = AttnDrop{0, nullptr, 0, 1.f};
}
}  // namespace
const AttnDrop& omr_attn_cur_dropout() { return g_drop_cur; }
extern "C" int omr_attn_next_dropout(float p, unsigned int seed, const int* seed_off) {
  OMR_REQUIRE(p >= 0.f && p < 1.f, "omr_attn_next_dropout: p must be in [0, 1) (got %f)", (double)p);
  unsigned int thr = (unsigned int)(p * 32768.f + 0.5f);  // 15-bit uniforms (attn_drop.cuh)
  if (thr > 32767u) thr = 32767u;
  g_drop_next = AttnDrop{seed, seed_off, thr, thr ? 32768.f / (float)(32768u - thr) : 1.f};
  return OMR_OK;
}

extern "C" int omr_attn_fwd(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs,
                            long long k_rs, const void* v, long long v_bs, long long v_rs, void* o, long long o_bs,
                            long long o_rs, float* lse, const float* key_bias, int B, int H, int Tq, int Tk, int hd,
                            float scale, int causal, int window, const int* q_len, const int* kv_len, int quirk_mod,
                            omr_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  take_dropout();
  if (tc_enabled() && dt == OMR_BF16) {
    TC_TRY(omr_attn_fwd_tc(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, lse, key_bias, B, H, Tq, Tk, hd,
                             scale, causal, window, q_len, kv_len, quirk_mod, st));
  }
  return omr_attn_fwd_simt(dt, q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, lse, key_bias, B, H, Tq, Tk,
                           hd, scale, causal, window, q_len, kv_len, quirk_mod, st);
}

extern "C" int omr_attn_bwd(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs,
                            long long k_rs, const void* v, long long v_bs, long long v_rs, const void* o,
                            long long o_bs, long long o_rs, const void* dout, long long do_bs, long long do_rs,
                            const float* lse, void* dq, long long dq_bs, long long dq_rs, void* dk, long long dk_bs,
                            long long dk_rs, void* dv, long long dv_bs, long long dv_rs, float* delta_ws,
                            const float* key_bias, int B, int H, int Tq, int Tk, int hd, float scale, int causal,
                            int window, const int* q_len, const int* kv_len, int quirk_mod, omr_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  take_dropout();
  if (tc_enabled() && dt == OMR_BF16) {
    TC_TRY(omr_attn_bwd_tc(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, dout, do_bs, do_rs, lse, dq, dq_bs, dq_rs,
                           dk, dk_bs, dk_rs, dv, dv_bs, dv_rs, delta_ws, key_bias, B, H, Tq, Tk, hd, scale, causal, window,
                           q_len, kv_len, quirk_mod, st));
  }
  return omr_attn_bwd_simt(dt, q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, dout, do_bs, do_rs, lse, dq,
                           dq_bs, dq_rs, dk, dk_bs, dk_rs, dv, dv_bs, dv_rs, delta_ws, key_bias, B, H, Tq, Tk, hd, scale,
                           causal, window, q_len, kv_len, quirk_mod, st);
}

// ---- classifier fused with the cross-entropy (projce_tc.cu): tensor-core path only; the caller asks first ----
extern "C" int omr_proj_ce_supported(int dt, int D) { return (tc_enabled() && dt == OMR_BF16 && D == 256) ? 1 : 0; }

static int proj_ce_rc(int rc, const char* what) {
  if (rc == OMR_TC_NOT_ELIGIBLE) {
    omr_set_error("%s: shape / alignment not served by the fused kernels (check omr_proj_ce_supported, 16-byte aligned rows)", what);
    return OMR_ERR_INVALID;
  }
  if (rc == OMR_OK) ++g_tc_calls;
  return rc;
}

extern "C" int omr_proj_ce_fwd(int dt, const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                               const long long* targets, long long rows, int V, int D, long long ignore_index,
                               float* row_loss, float* row_lse, omr_stream_t stream) {
  OMR_REQUIRE(omr_proj_ce_supported(dt, D), "omr_proj_ce_fwd: not supported for dt %d, D %d", dt, D);
  if (rows <= 0) return OMR_OK;
  return proj_ce_rc(omr_proj_ce_fwd_tc(x, x_ld, w, w_ld, bias, targets, rows, V, D, ignore_index, row_loss, row_lse, as_stream(stream)),
                    "omr_proj_ce_fwd");
}

extern "C" int omr_proj_ce_bwd_dx(int dt, const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                                  const long long* targets, const float* row_lse, const float* loss_out, const float* gscale,
                                  long long rows, int V, int D, long long ignore_index, void* dx, long long dx_ld,
                                  omr_stream_t stream) {
  OMR_REQUIRE(omr_proj_ce_supported(dt, D), "omr_proj_ce_bwd_dx: not supported for dt %d, D %d", dt, D);
  OMR_REQUIRE(dx != nullptr, "omr_proj_ce_bwd_dx: dx is NULL");
  if (rows <= 0) return OMR_OK;
  return proj_ce_rc(omr_proj_ce_bwd_tc(x, x_ld, w, w_ld, bias, targets, row_lse, loss_out, gscale, rows, V, D, ignore_index, dx, dx_ld,
                                       nullptr, nullptr, as_stream(stream), as_stream(stream)),
                    "omr_proj_ce_bwd_dx");
}

extern "C" int omr_proj_ce_bwd_dw(int dt, const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                                  const long long* targets, const float* row_lse, const float* loss_out, const float* gscale,
                                  long long rows, int V, int D, long long ignore_index, float* dw, float* db,
                                  omr_stream_t stream) {
  OMR_REQUIRE(omr_proj_ce_supported(dt, D), "omr_proj_ce_bwd_dw: not supported for dt %d, D %d", dt, D);
  OMR_REQUIRE(dw != nullptr, "omr_proj_ce_bwd_dw: dw is NULL");
  if (rows <= 0) return OMR_OK;
  return proj_ce_rc(omr_proj_ce_bwd_tc(x, x_ld, w, w_ld, bias, targets, row_lse, loss_out, gscale, rows, V, D, ignore_index, nullptr, 0, dw,
                                       db, as_stream(stream), as_stream(stream)),
                    "omr_proj_ce_bwd_dw");
}
