// Device helpers shared by the token-stage kernels (forward: transformer.cu, backward:
// tokens_bwd.cu): mma.sync fragment utilities, LayerNorm as a GEMM prologue, GELU, reductions.
#pragma once
#include <math.h>
#include "vc_common.cuh"
#include "vc_tparams.h"

namespace vc {

__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
// four 8x8 b16 matrices in one instruction: lane l supplies the row address of row (l & 7) of
// matrix (l >> 3); thread (g, q) receives elements [g][2q..2q+1] of each matrix
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float ex2(float v) {  // 2^v, one MUFU (flush-to-zero on underflow)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float tanh_fast(float v) {   // one MUFU, max relative error 2^-11
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// NUMERICAL DEVIATION FROM THE REFERENCE (stated in DESIGN.md section 5): the reference's nn.GELU is the exact
// erf form (timm/layers/mlp.py:21); these kernels evaluate the TANH APPROXIMATION 0.5 v (1 + tanh(sqrt(2/pi)
// (v + 0.044715 v^3))) with the hardware tanh.approx.f32: |gelu_tanh(v) - gelu_erf(v)| <= 4.8e-4 for every v
// (plus 2^-11 relative from the MUFU), below the bf16 rounding of the stored activation (2^-9 relative) for
// |v| >= 0.25, for 6 instructions instead of ~20 (the erf polynomial was 25 % of the kernel).  Its effect on
// the logits alone is bounded in tests/test_oracle_blocks_cpu.py::test_tanh_gelu_deviation_is_bounded
// (<= 2e-3 relative in fp32) and it is part of every logits-vs-oracle test (the oracle uses the erf form).
__device__ __forceinline__ float gelu_tanh_approx(float v) {
  const float u = v * fmaf(0.0356774081f, v * v, 0.7978845608f);
  const float hv = 0.5f * v;
  return fmaf(hv, tanh_fast(u), hv);
}
// value and derivative (backward kernel); same approximation so forward and backward agree
__device__ __forceinline__ void gelu_tanh_approx_grad(float v, float& val, float& der) {
  const float v2 = v * v;
  const float t = tanh_fast(v * fmaf(0.0356774081f, v2, 0.7978845608f));
  const float hv = 0.5f * v;
  val = fmaf(hv, t, hv);
  der = fmaf(hv * (1.f - t * t), fmaf(0.1070322243f, v2, 0.7978845608f), fmaf(0.5f, t, 0.5f));
}

// ---- dropout (training only; vision_transformer.py:598-629 pos_drop, :57-105 proj_drop,
// mlp.py:41-47 drop after GELU and after fc2) -----------------------------------------------------
// Stateless masks: a 32-bit hash of (seed of this step, sample, site, token row, column pair) yields
// two 16-bit draws, an element is kept iff its draw >= thr = round(p * 65536) and scaled by 1/(1-p').
// Forward, recompute and backward evaluate the same function, so no mask is stored.
// Sites: 0 pos_drop; block l: 1+3l proj_drop, 2+3l drop after GELU, 3+3l drop after fc2.
struct DropCfg {
  uint32_t thr;       // 0 = dropout off
  uint32_t seed;
  float inv_keep;     // 65536 / (65536 - thr)
};
__device__ __forceinline__ uint32_t drop_key(int b, int site, int row, int colpair) {
  return (((uint32_t)b * 8u + (uint32_t)site) * 256u + (uint32_t)row) * 64u + (uint32_t)colpair;
}
__device__ __forceinline__ uint32_t drop_hash(uint32_t key, uint32_t seed) {   // two multiply-xorshift rounds
  uint32_t h = (key ^ seed) * 0x9E3779B1u;
  h ^= h >> 16;
  h *= 0x85EBCA77u;
  h ^= h >> 15;
  return h;
}
// columns (2 * colpair, 2 * colpair + 1) of one row
__device__ __forceinline__ void drop2(float& a, float& b, uint32_t key, const DropCfg& d) {
  const uint32_t h = drop_hash(key, d.seed);
  a = ((h & 0xFFFFu) >= d.thr) ? a * d.inv_keep : 0.f;
  b = ((h >> 16) >= d.thr) ? b * d.inv_keep : 0.f;
}
// one column (lane = column layouts): `odd` selects the column's half of its pair's hash
__device__ __forceinline__ float drop1(float a, uint32_t key, int odd, const DropCfg& d) {
  const uint32_t h = drop_hash(key, d.seed);
  const uint32_t draw = odd ? (h >> 16) : (h & 0xFFFFu);
  return (draw >= d.thr) ? a * d.inv_keep : 0.f;
}

// LayerNorm (eps 1e-6) of the two token rows this thread shares with its quad; result as the
// two K=16 A fragments of the following GEMM.
__device__ __forceinline__ void ln_to_afrag(const float (&x)[4][4], const float* gam, const float* bet, int q,
                                            uint32_t (&A)[2][4]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3]; }
  const float m0 = quad_sum(s0) * (1.f / kD), m1 = quad_sum(s1) * (1.f / kD);
  float v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float d;
    d = x[j][0] - m0; v0 += d * d;
    d = x[j][1] - m0; v0 += d * d;
    d = x[j][2] - m1; v1 += d * d;
    d = x[j][3] - m1; v1 += d * d;
  }
  const float rs0 = rsqrtf(quad_sum(v0) * (1.f / kD) + 1e-6f), rs1 = rsqrtf(quad_sum(v1) * (1.f / kD) + 1e-6f);
  float y[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 gg = *reinterpret_cast<const float2*>(gam + 8 * j + 2 * q);
    const float2 bb = *reinterpret_cast<const float2*>(bet + 8 * j + 2 * q);
    y[j][0] = (x[j][0] - m0) * rs0 * gg.x + bb.x;
    y[j][1] = (x[j][1] - m0) * rs0 * gg.y + bb.y;
    y[j][2] = (x[j][2] - m1) * rs1 * gg.x + bb.x;
    y[j][3] = (x[j][3] - m1) * rs1 * gg.y + bb.y;
  }
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    A[kk][0] = pack_bf16(y[2 * kk][0], y[2 * kk][1]);
    A[kk][1] = pack_bf16(y[2 * kk][2], y[2 * kk][3]);
    A[kk][2] = pack_bf16(y[2 * kk + 1][0], y[2 * kk + 1][1]);
    A[kk][3] = pack_bf16(y[2 * kk + 1][2], y[2 * kk + 1][3]);
  }
}

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float g_sum(float v) {  // sum over the 8 row groups of a warp (lane bits 2..4)
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

}  // namespace vc
