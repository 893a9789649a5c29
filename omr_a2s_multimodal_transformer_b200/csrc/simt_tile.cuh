// simt_tile.cuh -- the fp32 FFMA register-tile micro-kernel shared by the CUDA-core GEMM,
// implicit-GEMM convolution and attention kernels.  These are the exact-fp32 paths (needed for
// the 1e-4 / token-identity parity mode, where tensor cores cannot be used) and the reference
// implementation the tcgen05 kernels are validated against.
#pragma once
#include "common.cuh"

constexpr int SIMT_BK = 16;

// acc[r][c] += sum_k As[k][m0 + r] * Bs[k][n0 + c]   (both tiles stored k-major, padded rows)
template <int LDA, int LDB, int KLEN>
__device__ __forceinline__ void simt_mma_4x4(const float* __restrict__ As, const float* __restrict__ Bs, int m0,
                                             int n0, float (&acc)[4][4]) {
#pragma unroll
  for (int k = 0; k < KLEN; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(As + k * LDA + m0);
    const float4 b = *reinterpret_cast<const float4*>(Bs + k * LDB + n0);
    acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
    acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
    acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
    acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
    acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
    acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
    acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
    acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
  }
}
