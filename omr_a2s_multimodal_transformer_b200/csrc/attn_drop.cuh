// attn_drop.cuh -- attention-probability dropout (the `dropout` of nn.MultiheadAttention inside
// nn.TransformerDecoderLayer, reference decoder.py:86-95; torch applies it to softmax(S) before the P V product, train
// mode only).  The keep mask is a pure function of (seed, batch*head, query, key), so the forward kernels, the backward
// kernels and the CUDA-core twins all regenerate the same mask and nothing is stored:
//     O  = (P o M / (1-p)) V            dV = (P o M / (1-p))^T dO
//     dS = P o (M o dP / (1-p) - delta) with delta = rowsum(dO o O)  (O already carries the mask)
// One counter-based hash yields the 16-bit uniforms of a 2 x 2 block of (query, key) pairs; a pair is dropped iff its
// uniform < thr.
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
// 4 x 16 random bits for the 2 x 2 block (queries 2*tpair, 2*tpair+1) x (keys 2*kpair, 2*kpair+1): the block counter goes
// through a full-avalanche 32-bit mixer (two multiply / xor-shift rounds), a third round derives the second word --
// about 4 integer instructions per probability.  .x serves the even query, .y the odd one; in each word the low half is
// the even key, the high half the odd key.  The forward kernels (a thread owns a query row and walks key pairs) and the
// backward kernels (a thread owns a key row and walks query pairs) both consume one block per two probabilities.
// kp = key pairs per row.  (Checked offline: each 16-bit field is uniform to 1e-3 at thr = 0.1 / 0.25 and the fields,
// and neighbouring blocks, are uncorrelated to the sampling noise.)
__device__ __forceinline__ uint2 attn_drop_block(uint32_t stream, uint32_t tpair, uint32_t kpair, uint32_t kp) {
  uint32_t x = (tpair * kp + kpair) * 0x9E3779B1u ^ stream;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  uint32_t y = x * 0x9E3779B1u ^ (x >> 15);
  y *= 0x85EBCA6Bu; y ^= y >> 16;
  return make_uint2(x, y);
}
// the 16-bit uniform of pair (t, key) inside its block
__device__ __forceinline__ uint32_t attn_drop_u16(uint2 blk, int t, int key) {
  return (((t & 1) ? blk.y : blk.x) >> ((key & 1) * 16)) & 0xFFFFu;
}
#endif
