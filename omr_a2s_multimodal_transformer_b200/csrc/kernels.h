// kernels.h -- internal (non-ABI) entry points of the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#define OMR_TC_NOT_ELIGIBLE 1 /* tcgen05 path declined the shape; caller uses the CUDA-core kernel */

int omr_gemm_simt(int in_dt, int out_dt, int transA, int transB, int M, int N, int K, const void* A, long long lda,
                  long long strideA, const void* B, long long ldb, long long strideB, void* C, long long ldc,
                  long long strideC, int batch, const float* bias, int bias_mode, int relu, int accumulate,
                  cudaStream_t st);
int omr_gemm_tc(int out_dt, int transA, int transB, int M, int N, int K, const void* A, long long lda,
                long long strideA, const void* B, long long ldb, long long strideB, void* C, long long ldc,
                long long strideC, int batch, const float* bias, int bias_mode, int relu, int accumulate,
                cudaStream_t st);

int omr_conv3x3_fwd_simt(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Ci,
                         int Co, int sh, int sw, int relu, cudaStream_t st);
int omr_conv3x3_dgrad_simt(int dt, const void* dy, const void* wT, void* dx, int N, int H, int W, int Ci, int Co, int sh,
                           int sw, cudaStream_t st);
int omr_conv3x3_wgrad_simt(int dt, const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Co, int sh,
                           int sw, int accumulate, cudaStream_t st);
int omr_conv3x3_fwd_tc(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Ci, int Co,
                       int sh, int sw, int relu, double* in_sums, cudaStream_t st);
int omr_conv3x3_dgrad_tc(const void* dy, const void* wT, void* dx, int N, int H, int W, int Ci, int Co, int sh, int sw,
                         const void* mask, float mask_scale, float* colsum, const void* in_x, double* in_bsums,
                         cudaStream_t st);
int omr_in_partial_sums(int dt, int mode, const void* a, const void* xin, double* out, int N, int HW, int C, cudaStream_t st);
int omr_relu_mask_scale(int dt, void* dx, const void* mask, float scale, long long n, cudaStream_t st);
int omr_conv3x3_wgrad_tc(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Co, int sh, int sw,
                         int accumulate, float* ws, cudaStream_t st);

int omr_attn_fwd_simt(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs,
                      long long k_rs, const void* v, long long v_bs, long long v_rs, void* o, long long o_bs,
                      long long o_rs, float* lse, const float* key_bias, int B, int H, int Tq, int Tk, int hd,
                      float scale, int causal, int window, const int* q_len, const int* kv_len, int quirk_mod,
                      cudaStream_t st);
int omr_attn_bwd_simt(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs,
                      long long k_rs, const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs,
                      long long o_rs, const void* dout, long long do_bs, long long do_rs, const float* lse, void* dq,
                      long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv,
                      long long dv_bs, long long dv_rs, float* delta_ws, const float* key_bias, int B, int H, int Tq,
                      int Tk, int hd, float scale, int causal, int window, const int* q_len, const int* kv_len,
                      int quirk_mod, cudaStream_t st);
int omr_attn_fwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                    const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse,
                    const float* key_bias, int B, int H, int Tq, int Tk, int hd, float scale, int causal, int window,
                    const int* q_len, const int* kv_len, int quirk_mod, cudaStream_t st);
int omr_attn_bwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                    const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                    const void* dout, long long do_bs, long long do_rs, const float* lse, void* dq, long long dq_bs,
                    long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv, long long dv_bs, long long dv_rs,
                    float* ws, const float* key_bias, int B, int H, int Tq, int Tk, int hd, float scale, int causal,
                    int window, const int* q_len, const int* kv_len, int quirk_mod, cudaStream_t st);
int omr_conv3x3_wgrad_small(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Co, int sh, int sw,
                            int accumulate, cudaStream_t st);
int omr_conv3x3_fwd_c1(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Co, int sh,
                       int sw, int relu, cudaStream_t st);
int omr_conv3x3_wgrad_c1(int dt, const void* x, const void* dy, float* dw, int N, int H, int W, int Co, int sh, int sw,
                         int accumulate, cudaStream_t st);

int omr_proj_ce_fwd_tc(const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                       const long long* targets, long long M, int V, int D, long long ignore_index, float* row_loss,
                       float* row_lse, cudaStream_t st);
int omr_proj_ce_bwd_tc(const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                       const long long* targets, const float* row_lse, const float* loss_out, const float* gscale,
                       long long M, int V, int D, long long ignore_index, void* dx, long long dx_ld, float* dw, float* db,
                       cudaStream_t st_dx, cudaStream_t st_dw);
