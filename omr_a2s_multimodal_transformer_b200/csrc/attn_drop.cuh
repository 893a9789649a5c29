// attn_drop.cuh -- attention-probability dropout (the `dropout` of nn.MultiheadAttention inside
// nn.TransformerDecoderLayer, reference decoder.py:86-95; torch applies it to softmax(S) before the P V product, train
// mode only).  The keep mask is a pure function of (seed, batch*head, query, key), so the forward kernels, the backward
// kernels and the CUDA-core twins all regenerate the same mask and nothing is stored:
//     O  = (P o M / (1-p)) V            dV = (P o M / (1-p))^T dO
//     dS = P o (M o dP / (1-p) - delta) with delta = rowsum(dO o O)  (O already carries the mask)
// One 32-bit hash yields the 16-bit uniforms of two neighbouring keys; a key is dropped iff its uniform < thr.
#pragma once
#include <stdint.h>

struct AttnDrop {
  uint32_t seed;        // host seed of this call
  const int* seed_off;  // optional device int32 mixed in on the GPU (fresh masks under CUDA-graph replay)
  uint32_t thr;         // round(p * 65536); 0 = dropout off
  float inv_keep;       // 1 / (1 - thr / 65536)
};
// the state set by omr_attn_next_dropout() for the attention call being dispatched (dispatch.cu)
const AttnDrop& omr_attn_cur_dropout();

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t attn_drop_stream(const AttnDrop& d, int bh) {
  uint32_t s = d.seed;
  if (d.seed_off) s += (uint32_t)(*d.seed_off) * 0x9E3779B9u;
  s ^= (uint32_t)bh * 0x632BE5ABu;
  s ^= s >> 15; s *= 0x2C1B3C6Du; s ^= s >> 12;
  return s;
}
// 2 x 16 random bits: low half for key 2*kpair, high half for key 2*kpair + 1 of query t (kp = key pairs per row)
__device__ __forceinline__ uint32_t attn_drop_bits(uint32_t stream, uint32_t t, uint32_t kpair, uint32_t kp) {
  uint32_t h = (t * kp + kpair) * 0x9E3779B1u ^ stream;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ bool attn_keep(uint32_t bits, int key, uint32_t thr) {
  return ((bits >> ((key & 1) * 16)) & 0xFFFFu) >= thr;
}
#endif
