// attn_drop.cuh -- attention-probability dropout (the `dropout` of nn.MultiheadAttention inside
// nn.TransformerDecoderLayer, reference decoder.py:86-95; torch applies it to softmax(S) before the P V product, train
// mode only).  The keep mask is a pure function of (seed, batch*head, query, key), so the forward kernels, the backward
// kernels and the CUDA-core twins all regenerate the same mask and nothing is stored:
//     O  = (P o M / (1-p)) V            dV = (P o M / (1-p))^T dO
//     dS = P o (M o dP / (1-p) - delta) with delta = rowsum(dO o O)  (O already carries the mask)
// One counter-based hash yields the 15-bit uniforms of a 2 x 2 block of (query, key) pairs; a pair is dropped iff its
// uniform < thr (thr = round(p * 32768): p = 0.1 -> 0.100006, 0.25 and 0.5 exact).
#pragma once
#include <stdint.h>

struct AttnDrop {
  uint32_t seed;        // host seed of this call
  const int* seed_off;  // optional device int32 mixed in on the GPU (fresh masks under CUDA-graph replay)
  uint32_t thr;         // round(p * 32768), at most 32767; 0 = dropout off
  float inv_keep;       // 1 / (1 - thr / 32768)
};
// the state set by omr_attn_next_dropout() for the attention call being dispatched (dispatch.cu)
const AttnDrop& omr_attn_cur_dropout();

#ifdef __CUDACC__
// the call's effective seed (one global load) and the stream of one (batch, head) derived from it
__device__ __forceinline__ uint32_t attn_drop_seed(const AttnDrop& d) {
  uint32_t s = d.seed;
  if (d.seed_off) s += (uint32_t)(*d.seed_off) * 0x9E3779B9u;
  return s;
}
__device__ __forceinline__ uint32_t attn_drop_stream_of(uint32_t seed, int bh) {
  uint32_t s = seed ^ (uint32_t)bh * 0x632BE5ABu;
  s ^= s >> 15; s *= 0x2C1B3C6Du; s ^= s >> 12;
  return s;
}
__device__ __forceinline__ uint32_t attn_drop_stream(const AttnDrop& d, int bh) {
  return attn_drop_stream_of(attn_drop_seed(d), bh);
}
// 4 x 15 random bits for the 2 x 2 block (queries 2*tpair, 2*tpair+1) x (keys 2*kpair, 2*kpair+1).  The block counter is
// scrambled by 32 x 32 -> 64-bit multiplies whose halves are folded together (each fold brings the well-mixed high word
// down onto the weak low bits): one round for the state, one more per output word -- 8 integer instructions per block =
// 2 per probability (round 1 used a 17-instruction xorshift-multiply mixer; at head dim 64 the attention kernels are
// bound by their element-wise instruction count, and the mask was 44 % of it).  .x serves the even query, .y the odd
// one; in each word bits 0-14 are the even key, bits 16-30 the odd key (bits 15 / 31 are unused: the packed compare
// below needs them as guard bits).  The forward kernels (a thread owns a query row and walks key pairs) and the backward
// kernels (a thread owns a key row and walks query pairs) both consume one block per two probabilities.  kp = key pairs
// per row.  Checked offline (numpy, 8 streams x 512 x 2337 pairs, thr = 0.1 / 0.25): keep rate within 2.2 sigma, chi^2 of
// the field histogram 1.4 per degree of freedom at worst, row / column keep-rate variance 0.93-1.09 of binomial, 16
// lagged autocorrelations (incl. across blocks and the .x/.y words) within 3 sigma -- the same as the round-1 mixer.
__device__ __forceinline__ uint2 attn_drop_block(uint32_t stream, uint32_t tpair, uint32_t kpair, uint32_t kp) {
  const uint32_t x = (tpair * kp + kpair) * 0x9E3779B1u ^ stream;
  const uint64_t p0 = (uint64_t)x * 0x7FEB352Du;
  const uint32_t a = (uint32_t)(p0 >> 32) ^ (uint32_t)p0;
  const uint64_t p1 = (uint64_t)a * 0x846CA68Bu, p2 = (uint64_t)a * 0x85EBCA6Bu;
  return make_uint2((uint32_t)(p1 >> 32) ^ (uint32_t)p1, (uint32_t)(p2 >> 32) ^ (uint32_t)p2);
}
// the 15-bit uniform of pair (t, key) inside its block
__device__ __forceinline__ uint32_t attn_drop_u15(uint2 blk, int t, int key) {
  return (((t & 1) ? blk.y : blk.x) >> ((key & 1) * 16)) & 0x7FFFu;
}
// Packed compare of both fields of a word: bit 15 (even key) / bit 31 (odd key) of the result is set iff that field is KEPT
// (>= thr).  thr2 = thr * 0x00010001.  With the guard bits forced to one no borrow crosses a field.
__device__ __forceinline__ uint32_t attn_drop_flags(uint32_t w, uint32_t thr2) { return (w | 0x80008000u) - thr2; }
// flags -> 0xFFFF per kept half (AND mask for a packed bf16 pair: even key low, odd key high)
__device__ __forceinline__ uint32_t attn_drop_mask_bf16x2(uint32_t flags) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(flags));  // selector bit 3: replicate the byte's sign bit
  return m;
}
#endif
