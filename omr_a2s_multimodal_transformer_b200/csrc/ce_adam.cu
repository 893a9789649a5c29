// ce_adam.cu -- vocabulary softmax cross-entropy (forward / backward, HBM-bound row kernels) and the
// fused multi-tensor Adam step that also refreshes the bf16 kernel-layout weight copies.
#include "common.cuh"

namespace {

// one block per row: lse = log sum exp(logits), loss = lse - logit[target]
template <typename T>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const T* __restrict__ logits, long long ld,
                                                     const long long* __restrict__ targets, int V,
                                                     long long ignore_index, float* __restrict__ row_loss,
                                                     float* __restrict__ row_lse) {
  omr_pdl_enter();
  __shared__ float sm[33];
  const long long r = blockIdx.x;
  const T* x = logits + r * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, to_f(x[i]));
  mx = block_max(mx, sm);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(to_f(x[i]) - mx);
  s = block_sum(s, sm);
  if (threadIdx.x == 0) {
    float lse = mx + logf(s);
    row_lse[r] = lse;
    long long t = targets[r];
    row_loss[r] = (t == ignore_index || t < 0 || t >= V) ? 0.f : lse - to_f(x[t]);
  }
}

// Vector variant (rows 16-byte aligned, V <= 4 * 256 * 16/sizeof(T)): a row is read ONCE with 16-byte loads and kept in
// registers for both the max and the sum; rows whose target is ignore_index (the padded tail of a batch, ~1/3 of the
// rows) are not read at all -- their loss is 0 and the backward writes zeros without the log-sum-exp.
template <typename T>
__global__ void __launch_bounds__(256) ce_fwd_vec_kernel(const T* __restrict__ logits, long long ld,
                                                         const long long* __restrict__ targets, int V,
                                                         long long ignore_index, float* __restrict__ row_loss,
                                                         float* __restrict__ row_lse) {
  omr_pdl_enter();
  constexpr int VEC = 16 / (int)sizeof(T);
  __shared__ float sm[33];
  const long long r = blockIdx.x;
  const long long t = targets[r];
  if (t == ignore_index || t < 0 || t >= V) {  // uniform per block
    if (threadIdx.x == 0) { row_loss[r] = 0.f; row_lse[r] = 0.f; }
    return;
  }
  const T* x = logits + r * ld;
  const int nv = (V + VEC - 1) / VEC;  // the padded row stride covers the last partial vector
  float v[4][VEC];
  float mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.x + u * 256;
    if (i < nv) {
      if constexpr (sizeof(T) == 2) {
        const uint4 raw = *reinterpret_cast<const uint4*>(x + (long long)i * VEC);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[u][2 * k] = __low2float(h[k]); v[u][2 * k + 1] = __high2float(h[k]); }
      } else {
        const float4 raw = *reinterpret_cast<const float4*>(x + (long long)i * VEC);
        v[u][0] = raw.x; v[u][1] = raw.y; v[u][2] = raw.z; v[u][3] = raw.w;
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (i * VEC + k >= V) v[u][k] = -INFINITY;  // padding columns of the row stride
        mx = fmaxf(mx, v[u][k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[u][k] = -INFINITY;
    }
  }
  mx = block_max(mx, sm);
  float sum = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < VEC; ++k) sum += expf(v[u][k] - mx);
  sum = block_sum(sum, sm);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(sum);
    row_lse[r] = lse;
    row_loss[r] = lse - to_f(x[t]);
  }
}

// single block: deterministic tree reduction of the per-row losses
__global__ void __launch_bounds__(1024) ce_reduce_kernel(const float* __restrict__ row_loss,
                                                         const long long* __restrict__ targets, long long rows,
                                                         long long ignore_index, float* __restrict__ out) {
  omr_pdl_enter();
  __shared__ float sm[33];
  float s = 0.f, n = 0.f;
  for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
    if (targets[i] != ignore_index) { s += row_loss[i]; n += 1.f; }
  }
  s = block_sum(s, sm);
  n = block_sum(n, sm);
  if (threadIdx.x == 0) {
    out[0] = n > 0.f ? s / n : 0.f;  // torch returns nan for an all-ignored batch; never on the hot path
    out[1] = n;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const T* __restrict__ logits, long long ld,
                                                     const long long* __restrict__ targets,
                                                     const float* __restrict__ row_lse,
                                                     const float* __restrict__ loss_out,
                                                     const float* __restrict__ gscale, T* __restrict__ dlogits, int V,
                                                     long long ignore_index) {
  omr_pdl_enter();
  const long long r = blockIdx.x;
  const long long t = targets[r];
  const T* x = logits + r * ld;
  T* d = dlogits + r * ld;
  if (t == ignore_index) {
    for (int i = threadIdx.x; i < V; i += blockDim.x) d[i] = from_f<T>(0.f);
    return;
  }
  const float n = loss_out[1];
  const float g = (gscale ? gscale[0] : 1.f) / (n > 0.f ? n : 1.f);
  const float lse = row_lse[r];
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float p = expf(to_f(x[i]) - lse);
    d[i] = from_f<T>((p - (i == t ? 1.f : 0.f)) * g);
  }
}

// Vector variant of the backward (16-byte loads / stores; the padded columns of the row stride are written as zeros)
template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_vec_kernel(const T* __restrict__ logits, long long ld,
                                                         const long long* __restrict__ targets,
                                                         const float* __restrict__ row_lse,
                                                         const float* __restrict__ loss_out,
                                                         const float* __restrict__ gscale, T* __restrict__ dlogits, int V,
                                                         long long ignore_index) {
  omr_pdl_enter();
  constexpr int VEC = 16 / (int)sizeof(T);
  const long long r = blockIdx.x;
  const long long t = targets[r];
  const T* x = logits + r * ld;
  T* d = dlogits + r * ld;
  const int nv = (V + VEC - 1) / VEC;
  if (t == ignore_index) {
    for (int i = threadIdx.x; i < nv; i += 256) *reinterpret_cast<uint4*>(d + (long long)i * VEC) = make_uint4(0, 0, 0, 0);
    return;
  }
  const float n = loss_out[1];
  const float g = (gscale ? gscale[0] : 1.f) / (n > 0.f ? n : 1.f);
  const float lse = row_lse[r];
  for (int i = threadIdx.x; i < nv; i += 256) {
    float v[VEC];
    if constexpr (sizeof(T) == 2) {
      const uint4 raw = *reinterpret_cast<const uint4*>(x + (long long)i * VEC);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[2 * k] = __low2float(h[k]); v[2 * k + 1] = __high2float(h[k]); }
    } else {
      const float4 raw = *reinterpret_cast<const float4*>(x + (long long)i * VEC);
      v[0] = raw.x; v[1] = raw.y; v[2] = raw.z; v[3] = raw.w;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = i * VEC + k;
      v[k] = c < V ? (expf(v[k] - lse) - (c == t ? 1.f : 0.f)) * g : 0.f;
    }
    if constexpr (sizeof(T) == 2) {
      uint4 o;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      *reinterpret_cast<uint4*>(d + (long long)i * VEC) = o;
    } else {
      *reinterpret_cast<float4*>(d + (long long)i * VEC) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// ---- Adam ------------------------------------------------------------------------------------
__global__ void adam_tick_kernel(int* step) {
  omr_pdl_enter(); *step += 1; }

__device__ __forceinline__ long long adam_shadow_index(long long i, int layout, int d0, int d1) {
  if (layout == 1) {  // [Co,Ci,3,3] -> [Co,3,3,Ci]   (d0 = Co, d1 = Ci)
    int tap = (int)(i % 9);
    long long r = i / 9;
    int ci = (int)(r % d1);
    long long co = r / d1;
    return (co * 9 + tap) * d1 + ci;
  }
  if (layout == 2) {  // [C,1,3,3] -> [3,3,C]   (d0 = C)
    int tap = (int)(i % 9);
    long long c = i / 9;
    return (long long)tap * d0 + c;
  }
  if (layout == 3) {  // [Co,Ci,3,3] -> [Ci,3,3,Co]  (data-gradient operand)
    int tap = (int)(i % 9);
    long long r = i / 9;
    int ci = (int)(r % d1);
    long long co = r / d1;
    return ((long long)ci * 9 + tap) * d0 + co;
  }
  if (layout == 4) {  // [R,C] -> [C,R] transpose  (d0 = R, d1 = C)
    long long r = i / d1;
    int c = (int)(i % d1);
    return (long long)c * d0 + r;
  }
  return i;
}

// grid (chunks of a tensor, tensors).  Round 2: the bias corrections (two double-precision pow() and a sqrt) are computed
// by ONE thread of a block that has work and broadcast through shared memory -- every one of the 5 M threads used to compute
// them itself --, blocks past the end of their tensor leave at once, and tensors whose arrays are 16-byte aligned are
// walked with float4 accesses (316 -> ~80 us for the 11.4 M parameters of the C3 model; floor: 28 bytes per parameter).
__global__ void __launch_bounds__(256) adam_kernel(const omr_adam_entry* __restrict__ table, const int* __restrict__ step,
                                                   double lr, double b1d, double b2d, double epsd, double gsd) {
  omr_pdl_enter();
  const omr_adam_entry e = table[blockIdx.y];
  if (e.grad == nullptr) return;
  const bool vec = (e.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(e.param) | reinterpret_cast<uintptr_t>(e.grad) |
                                       reinterpret_cast<uintptr_t>(e.exp_avg) | reinterpret_cast<uintptr_t>(e.exp_avg_sq)) & 15) == 0;
  const long long per_pass = (long long)blockDim.x * (vec ? 4 : 1);
  if ((long long)blockIdx.x * per_pass >= e.n) return;
  __shared__ float s_corr[2];
  if (threadIdx.x == 0) {
    const int t = *step;
    // bias corrections as torch.optim.Adam (single-tensor path): step_size = lr / (1 - b1^t),
    // denom = sqrt(v) / sqrt(1 - b2^t) + eps ; computed in double like the Python reference
    s_corr[0] = (float)(lr / (1.0 - pow(b1d, (double)t)));
    s_corr[1] = (float)sqrt(1.0 - pow(b2d, (double)t));
  }
  __syncthreads();
  const float step_size = s_corr[0], bc2s = s_corr[1];
  const float b1 = (float)b1d, b2 = (float)b2d, eps = (float)epsd, grad_scale = (float)gsd;
  bf16* sh0 = (bf16*)e.shadow;
  bf16* sh1 = (bf16*)e.shadow2;
  auto upd = [&](float g, float& m, float& v, float& p) {
    g *= grad_scale;
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    p = p - step_size * (m / (sqrtf(v) / bc2s + eps));
  };
  if (vec) {
    const long long n4 = e.n / 4;
    for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
      const float4 g4 = reinterpret_cast<const float4*>(e.grad)[i4];
      float4 m4 = reinterpret_cast<float4*>(e.exp_avg)[i4], v4 = reinterpret_cast<float4*>(e.exp_avg_sq)[i4];
      float4 p4 = reinterpret_cast<float4*>(e.param)[i4];
      upd(g4.x, m4.x, v4.x, p4.x); upd(g4.y, m4.y, v4.y, p4.y); upd(g4.z, m4.z, v4.z, p4.z); upd(g4.w, m4.w, v4.w, p4.w);
      reinterpret_cast<float4*>(e.exp_avg)[i4] = m4;
      reinterpret_cast<float4*>(e.exp_avg_sq)[i4] = v4;
      reinterpret_cast<float4*>(e.param)[i4] = p4;
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
      const long long i = i4 * 4;
      if (sh0) {
        if (e.layout == 0 && (reinterpret_cast<uintptr_t>(sh0) & 7) == 0) {
          uint2 w;
          __nv_bfloat162 lo = __floats2bfloat162_rn(pv[0], pv[1]), hi = __floats2bfloat162_rn(pv[2], pv[3]);
          w.x = *reinterpret_cast<uint32_t*>(&lo); w.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(sh0 + i) = w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) sh0[adam_shadow_index(i + k, e.layout, e.d0, e.d1)] = __float2bfloat16_rn(pv[k]);
        }
      }
      if (sh1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sh1[adam_shadow_index(i + k, e.layout2, e.d0, e.d1)] = __float2bfloat16_rn(pv[k]);
      }
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (long long)gridDim.x * blockDim.x) {
    float m = e.exp_avg[i], v = e.exp_avg_sq[i], p = e.param[i];
    upd(e.grad[i], m, v, p);
    e.exp_avg[i] = m;
    e.exp_avg_sq[i] = v;
    e.param[i] = p;
    if (sh0) sh0[adam_shadow_index(i, e.layout, e.d0, e.d1)] = __float2bfloat16_rn(p);
    if (sh1) sh1[adam_shadow_index(i, e.layout2, e.d0, e.d1)] = __float2bfloat16_rn(p);
  }
}

}  // namespace

extern "C" int omr_ce_fwd(int dt, const void* logits, long long ld, const long long* targets, long long rows, int V,
                          long long ignore_index, float* row_loss, float* row_lse, omr_stream_t stream) {
  if (rows <= 0) return OMR_OK;
  OMR_REQUIRE(V > 0, "omr_ce_fwd: empty vocabulary");
  {
    const int esz = dt == OMR_F32 ? 4 : 2, vec = 16 / esz;
    const long long nv = (V + vec - 1) / vec;
    if (nv * vec <= ld && nv <= 4 * 256 && (ld * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
      OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)rows, 256, 0, as_stream(stream))(ce_fwd_vec_kernel<T>, 
                                 (const T*)logits, ld, targets, V, ignore_index, row_loss, row_lse)));
      OMR_LAUNCHED();
      return OMR_OK;
    }
  }
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)rows, 256, 0, as_stream(stream))(ce_fwd_kernel<T>, 
                             (const T*)logits, ld, targets, V, ignore_index, row_loss, row_lse)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_ce_reduce(const float* row_loss, const long long* targets, long long rows, long long ignore_index,
                             float* loss_out, omr_stream_t stream) {
  OmrLaunch(1, 1024, 0, as_stream(stream))(ce_reduce_kernel, row_loss, targets, rows, ignore_index, loss_out);
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_ce_bwd(int dt, const void* logits, long long ld, const long long* targets, const float* row_lse,
                          const float* loss_out, const float* gscale, void* dlogits, long long rows, int V,
                          long long ignore_index, omr_stream_t stream) {
  if (rows <= 0) return OMR_OK;
  {
    const int esz = dt == OMR_F32 ? 4 : 2, vec = 16 / esz;
    const long long nv = (V + vec - 1) / vec;
    if (nv * vec <= ld && (ld * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) {
      OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)rows, 256, 0, as_stream(stream))(ce_bwd_vec_kernel<T>, 
                                 (const T*)logits, ld, targets, row_lse, loss_out, gscale, (T*)dlogits, V, ignore_index)));
      OMR_LAUNCHED();
      return OMR_OK;
    }
  }
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)rows, 256, 0, as_stream(stream))(ce_bwd_kernel<T>, 
                             (const T*)logits, ld, targets, row_lse, loss_out, gscale, (T*)dlogits, V, ignore_index)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_adam_tick(int* step, omr_stream_t stream) {
  OmrLaunch(1, 1, 0, as_stream(stream))(adam_tick_kernel, step);
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_adam_step(const omr_adam_entry* table, int n_tensors, long long max_n, const int* step, double lr,
                             double beta1, double beta2, double eps, double grad_scale, omr_stream_t stream) {
  if (n_tensors <= 0) return OMR_OK;
  long long bx = cdiv(max_n, 256 * 4 * 2);  // two float4 passes per thread on the largest tensor
  if (bx < 1) bx = 1;
  if (bx > 128) bx = 128;  // (blocks past the end of a small tensor exit at once, but still cost a launch slot)
  dim3 grid((unsigned)bx, (unsigned)n_tensors);
  OmrLaunch(grid, 256, 0, as_stream(stream))(adam_kernel, table, step, lr, beta1, beta2, eps, grad_scale);
  OMR_LAUNCHED();
  return OMR_OK;
}
