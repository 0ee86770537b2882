// Device helpers shared by the token-stage kernels (forward: transformer.cu, backward:
// tokens_bwd.cu): mma.sync fragment utilities, LayerNorm as a GEMM prologue, GELU, reductions.
#pragma once
#include <math.h>
#include "vc_common.cuh"
#include "vc_tparams.h"

namespace vc {

__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
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
// exact-erf GELU (nn.GELU default, mlp.py:21) with erf from Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, far below bf16 resolution): one RCP + one EX2 + a few FMAs.
__device__ __forceinline__ float gelu_erf(float v) {
  const float z = fabsf(v) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = p * t * ex2(-1.44269504088896340736f * z * z);   // 1 - erf(z)
  const float half_erfc = 0.5f * e;                                  // v>=0: Phi = 1 - e/2, v<0: Phi = e/2
  return v * (v >= 0.f ? 1.f - half_erfc : half_erfc);
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
