// Training-side kernels that are HBM-bound element-wise / reduction passes over SPS buffers or
// small per-step utilities:
//   * BatchNorm2d in training mode (batch statistics, running-stat update) + ReLU, forward and
//     backward -- the conv_bn_relu idiom of the stems (S2ENet bytecode, SURVEY.md App. A.2;
//     model/Multimodality_Mamba/Mutimodality_Mamba7.py:1035-1048) as autograd differentiates it;
//   * nn.CrossEntropyLoss(weight=w) forward + gradient (model_utils.py:63-66, 216, 929-936);
//   * optim.Adam step (model_utils.py:214-215);
//   * packing of fp32 master weights into the bf16 operand layouts of the tensor-core kernels.
#include <math.h>
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

__device__ __forceinline__ bool sps_row_valid(int R, int HALO, int PP, int PW, int P, int n) {
  const int r = R - HALO;
  if (r < 0) return false;
  const int b = r / PP;
  const int q = r - b * PP;
  const int i = q / PW, j = q - i * PW;
  return b < n && i < P && j < P;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xFFFF0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// block-wide sum of 16 per-thread values, then one double atomicAdd per value (256 threads)
__device__ __forceinline__ void block_reduce16_atomic(float (&acc)[16], double* dst0, double* dst1) {
  __shared__ float red[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[w][threadIdx.x];
    if (threadIdx.x < 8) atomicAdd(dst0 + threadIdx.x, s);
    else atomicAdd(dst1 + threadIdx.x - 8, s);
  }
}

// ---- BatchNorm forward -------------------------------------------------------------------------
// sums[c] += sum_rows y[c], sums[Cp + c] += sum_rows y[c]^2 (pad rows are zero and add nothing)
__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, long long RT, int Cp,
                                                       double* __restrict__ sums) {
  const int s = blockIdx.y;
  const uint4* src = reinterpret_cast<const uint4*>(y) + (long long)s * RT;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(__ldg(src + R), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] += v[k]; acc[8 + k] = fmaf(v[k], v[k], acc[8 + k]); }
  }
  block_reduce16_atomic(acc, sums + s * 8, sums + Cp + s * 8);
}

// mean / biased var -> scale = gamma * rstd, shift = beta - mean * scale; running stats (momentum,
// unbiased var) as nn.BatchNorm2d does in training mode; stats zeroed for the next step.
__global__ void bn_finalize_kernel(double* sums, double count, int C, int Cp, const float* gamma, const float* beta,
                                   float eps, float momentum, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float* scale, float* shift, float* mean, float* rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (c >= Cp) return;
  if (c < C) {
    const double m = sums[c] / count;
    double var = sums[Cp + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * rs;
    scale[c] = sc;
    shift[c] = beta[c] - (float)m * sc;
    mean[c] = (float)m;
    rstd[c] = rs;
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    scale[c] = 0.f; shift[c] = 0.f; mean[c] = 0.f; rstd[c] = 0.f;
  }
  sums[c] = 0.0;
  sums[Cp + c] = 0.0;
}

// z = relu(y * scale + shift) on valid cells, 0 elsewhere.  One thread per SPS row walking the
// slices: the row's validity is decoded once, per-channel constants come from shared memory.
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ z,
                                                       int S, long long RT, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, int relu, int P, int n) {
  __shared__ float sc_s[256], sh_s[256];
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) { sc_s[i] = scale[i]; sh_s[i] = shift[i]; }
  __syncthreads();
  const int HALO = sps_halo(P), PP = sps_pp(P), PW = P + 1;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    const bool valid = sps_row_valid((int)R, HALO, PP, PW, P, n);
    const uint4* src = reinterpret_cast<const uint4*>(y) + R;
    uint4* dst = reinterpret_cast<uint4*>(z) + R;
    if (!valid) {
      for (int s = 0; s < S; ++s) dst[(long long)s * RT] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      float v[8];
      unpack8(__ldg(src + (long long)s * RT), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = fmaf(v[k], sc_s[s * 8 + k], sh_s[s * 8 + k]);
        if (relu) v[k] = fmaxf(v[k], 0.f);
      }
      dst[(long long)s * RT] = pack8(v);
    }
  }
}

// ---- BatchNorm (+ReLU) backward ----------------------------------------------------------------
// g = dz * [y*scale+shift > 0];  sums[c] += sum g, sums[Cp+c] += sum g * xhat
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dz,
                                                            const __nv_bfloat16* __restrict__ y, long long RT, int Cp,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            int relu, double* __restrict__ sums) {
  const int s = blockIdx.y;
  const uint4* pdz = reinterpret_cast<const uint4*>(dz) + (long long)s * RT;
  const uint4* py = reinterpret_cast<const uint4*>(y) + (long long)s * RT;
  float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale[s * 8 + k]; sh[k] = shift[s * 8 + k]; mu[k] = mean[s * 8 + k]; rs[k] = rstd[s * 8 + k]; }
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    float g[8], v[8];
    unpack8(__ldg(pdz + R), g);
    unpack8(__ldg(py + R), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gg = (!relu || fmaf(v[k], sc[k], sh[k]) > 0.f) ? g[k] : 0.f;
      acc[k] += gg;
      acc[8 + k] = fmaf(gg, (v[k] - mu[k]) * rs[k], acc[8 + k]);
    }
  }
  block_reduce16_atomic(acc, sums + s * 8, sums + Cp + s * 8);
}

// dy = scale * (g - mean(g) - xhat * mean(g*xhat)) on valid cells (0 elsewhere); may run in place.
// One thread per SPS row walking the slices; per-channel constants (fp32) staged in shared memory:
//   dy = k0 * g - k1 - k2 * y   with  k0 = scale, k2 = scale * rstd * mean(g xhat),
//                                    k1 = scale * mean(g) - k2 * mean
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* dz, const __nv_bfloat16* __restrict__ y,
                                                           __nv_bfloat16* dy, int S, long long RT, int Cp,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           int relu, const double* __restrict__ sums, double inv_count,
                                                           int P, int n) {
  __shared__ float sc_s[256], sh_s[256], k1_s[256], k2_s[256];
  for (int c = threadIdx.x; c < S * 8; c += blockDim.x) {
    const float sc = scale[c];
    const float k2 = sc * rstd[c] * (float)(sums[Cp + c] * inv_count);
    sc_s[c] = sc;
    sh_s[c] = shift[c];
    k2_s[c] = k2;
    k1_s[c] = sc * (float)(sums[c] * inv_count) - k2 * mean[c];
  }
  __syncthreads();
  const int HALO = sps_halo(P), PP = sps_pp(P), PW = P + 1;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    const bool valid = sps_row_valid((int)R, HALO, PP, PW, P, n);
    const uint4* pdz = reinterpret_cast<const uint4*>(dz) + R;
    const uint4* py = reinterpret_cast<const uint4*>(y) + R;
    uint4* pdy = reinterpret_cast<uint4*>(dy) + R;
    if (!valid) {
      for (int s = 0; s < S; ++s) pdy[(long long)s * RT] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      float g[8], v[8];
      unpack8(pdz[(long long)s * RT], g);
      unpack8(__ldg(py + (long long)s * RT), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = s * 8 + k;
        const float gg = (!relu || fmaf(v[k], sc_s[c], sh_s[c]) > 0.f) ? g[k] : 0.f;
        g[k] = fmaf(sc_s[c], gg, -fmaf(k2_s[c], v[k], k1_s[c]));
      }
      pdy[(long long)s * RT] = pack8(g);
    }
  }
}

// dgamma = sum g*xhat, dbeta = sum g, dbias(conv) = 0 (BatchNorm removes the mean); clears sums
__global__ void bn_bwd_params_kernel(double* sums, int C, int Cp, float* dgamma, float* dbeta, float* dbias, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  if (c < C) {
    const float dg = (float)sums[Cp + c], db = (float)sums[c];
    dgamma[c] = accumulate ? dgamma[c] + dg : dg;
    dbeta[c] = accumulate ? dbeta[c] + db : db;
    if (dbias && !accumulate) dbias[c] = 0.f;
  }
}
__global__ void clear_f64_kernel(double* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

static inline int rows_grid(long long RT, int S) {
  long long per = (RT + 255) / 256;
  long long cap = (148LL * 8 + S - 1) / S;
  if (per > cap) per = cap;
  return (int)(per < 1 ? 1 : per);
}
static inline int flat_grid(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  return (int)(b < 1 ? 1 : b);
}

int bn_forward_launch(const void* y, void* z, int S, int C, int n_patches, int P, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, long long* nbt, double* sums,
                      float* scale, float* shift, float* mean, float* rstd, int relu, cudaStream_t st) {
  if (S < 1 || C > S * 8 || n_patches <= 0) return VC_ERR_ARG;
  const long long RT = sps_rows(n_patches, P);
  const int Cp = S * 8;
  bn_stats_kernel<<<dim3(rows_grid(RT, S), S), 256, 0, st>>>((const __nv_bfloat16*)y, RT, Cp, sums);
  bn_finalize_kernel<<<(Cp + 127) / 128, 128, 0, st>>>(sums, (double)n_patches * P * P, C, Cp, gamma, beta, eps, momentum,
                                                      running_mean, running_var, nbt, scale, shift, mean, rstd);
  bn_apply_kernel<<<flat_grid(RT), 256, 0, st>>>((const __nv_bfloat16*)y, (__nv_bfloat16*)z, S, RT, scale,
                                                                 shift, relu, P, n_patches);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

int bn_backward_launch(const void* dz, const void* y, void* dy, int S, int C, int n_patches, int P, const float* scale,
                       const float* shift, const float* mean, const float* rstd, int relu, double* sums, float* dgamma,
                       float* dbeta, float* dbias, int accumulate, cudaStream_t st) {
  if (S < 1 || C > S * 8 || n_patches <= 0) return VC_ERR_ARG;
  const long long RT = sps_rows(n_patches, P);
  const int Cp = S * 8;
  bn_bwd_reduce_kernel<<<dim3(rows_grid(RT, S), S), 256, 0, st>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, RT, Cp,
                                                                   scale, shift, mean, rstd, relu, sums);
  bn_bwd_apply_kernel<<<flat_grid(RT), 256, 0, st>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)y,
                                                                     (__nv_bfloat16*)dy, S, RT, Cp, scale, shift, mean, rstd,
                                                                     relu, sums, 1.0 / ((double)n_patches * P * P), P,
                                                                     n_patches);
  bn_bwd_params_kernel<<<(Cp + 127) / 128, 128, 0, st>>>(sums, C, Cp, dgamma, dbeta, dbias, accumulate);
  clear_f64_kernel<<<(2 * Cp + 127) / 128, 128, 0, st>>>(sums, 2 * Cp);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// ---- two-launch forms for the training step ----------------------------------------------------------
// bn_apply_kernel with the finalisation folded in: every block derives scale / shift from the batch sums itself,
// block 0 also records them (+ mean, rstd for the backward pass) and updates the running statistics.
__global__ void __launch_bounds__(256) bn_apply_finalize_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ z,
                                                                int S, long long RT, const double* __restrict__ sums, double count,
                                                                int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float eps, float momentum, float* running_mean, float* running_var,
                                                                long long* num_batches_tracked, float* scale, float* shift, float* mean,
                                                                float* rstd, int relu, int P, int n) {
  __shared__ float sc_s[256], sh_s[256];
  const int Cp = S * 8;
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float sc = 0.f, sh = 0.f, mf = 0.f, rs = 0.f;
    if (c < C) {
      const double m = sums[c] / count;
      double var = sums[Cp + c] / count - m * m;
      if (var < 0.0) var = 0.0;
      rs = (float)(1.0 / sqrt(var + (double)eps));
      sc = gamma[c] * rs;
      sh = beta[c] - (float)m * sc;
      mf = (float)m;
      if (blockIdx.x == 0 && running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mf;
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
    sc_s[c] = sc;
    sh_s[c] = sh;
    if (blockIdx.x == 0) { scale[c] = sc; shift[c] = sh; mean[c] = mf; rstd[c] = rs; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  __syncthreads();
  const int HALO = sps_halo(P), PP = sps_pp(P), PW = P + 1;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    const bool valid = sps_row_valid((int)R, HALO, PP, PW, P, n);
    const uint4* src = reinterpret_cast<const uint4*>(y) + R;
    uint4* dst = reinterpret_cast<uint4*>(z) + R;
    if (!valid) {
      for (int s = 0; s < S; ++s) dst[(long long)s * RT] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      float v[8];
      unpack8(__ldg(src + (long long)s * RT), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = fmaf(v[k], sc_s[s * 8 + k], sh_s[s * 8 + k]);
        if (relu) v[k] = fmaxf(v[k], 0.f);
      }
      dst[(long long)s * RT] = pack8(v);
    }
  }
}

int bn_forward_fused_launch(const void* y, void* z, int S, int C, int n_patches, int P, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, long long* nbt, double* sums,
                            float* scale, float* shift, float* mean, float* rstd, int relu, cudaStream_t st) {
  if (S < 1 || S > 32 || C > S * 8 || n_patches <= 0) return VC_ERR_ARG;
  const long long RT = sps_rows(n_patches, P);
  const int Cp = S * 8;
  bn_stats_kernel<<<dim3(rows_grid(RT, S), S), 256, 0, st>>>((const __nv_bfloat16*)y, RT, Cp, sums);
  bn_apply_finalize_kernel<<<flat_grid(RT), 256, 0, st>>>((const __nv_bfloat16*)y, (__nv_bfloat16*)z, S, RT, sums,
                                                          (double)n_patches * P * P, C, gamma, beta, eps, momentum, running_mean,
                                                          running_var, nbt, scale, shift, mean, rstd, relu, P, n_patches);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// bn_bwd_apply_kernel + bn_bwd_params_kernel: block 0 also writes dgamma / dbeta (and zeroes the conv bias gradient)
__global__ void __launch_bounds__(256) bn_bwd_apply_params_kernel(const __nv_bfloat16* dz, const __nv_bfloat16* __restrict__ y,
                                                                  __nv_bfloat16* dy, int S, long long RT, int C,
                                                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  int relu, const double* __restrict__ sums, double inv_count,
                                                                  float* dgamma, float* dbeta, float* dbias, int P, int n) {
  __shared__ float sc_s[256], sh_s[256], k1_s[256], k2_s[256];
  const int Cp = S * 8;
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    const float sc = scale[c];
    const float k2 = sc * rstd[c] * (float)(sums[Cp + c] * inv_count);
    sc_s[c] = sc;
    sh_s[c] = shift[c];
    k2_s[c] = k2;
    k1_s[c] = sc * (float)(sums[c] * inv_count) - k2 * mean[c];
    if (blockIdx.x == 0 && c < C) {
      dgamma[c] = (float)sums[Cp + c];
      dbeta[c] = (float)sums[c];
      if (dbias) dbias[c] = 0.f;
    }
  }
  __syncthreads();
  const int HALO = sps_halo(P), PP = sps_pp(P), PW = P + 1;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    const bool valid = sps_row_valid((int)R, HALO, PP, PW, P, n);
    const uint4* pdz = reinterpret_cast<const uint4*>(dz) + R;
    const uint4* py = reinterpret_cast<const uint4*>(y) + R;
    uint4* pdy = reinterpret_cast<uint4*>(dy) + R;
    if (!valid) {
      for (int s = 0; s < S; ++s) pdy[(long long)s * RT] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      float g[8], v[8];
      unpack8(pdz[(long long)s * RT], g);
      unpack8(__ldg(py + (long long)s * RT), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = s * 8 + k;
        const float gg = (!relu || fmaf(v[k], sc_s[c], sh_s[c]) > 0.f) ? g[k] : 0.f;
        g[k] = fmaf(sc_s[c], gg, -fmaf(k2_s[c], v[k], k1_s[c]));
      }
      pdy[(long long)s * RT] = pack8(g);
    }
  }
}

int bn_backward_fused_launch(const void* dz, const void* y, void* dy, int S, int C, int n_patches, int P, const float* scale,
                             const float* shift, const float* mean, const float* rstd, int relu, double* sums, float* dgamma,
                             float* dbeta, float* dbias, cudaStream_t st) {
  if (S < 1 || S > 32 || C > S * 8 || n_patches <= 0) return VC_ERR_ARG;
  const long long RT = sps_rows(n_patches, P);
  const int Cp = S * 8;
  bn_bwd_reduce_kernel<<<dim3(rows_grid(RT, S), S), 256, 0, st>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, RT, Cp,
                                                                   scale, shift, mean, rstd, relu, sums);
  bn_bwd_apply_params_kernel<<<flat_grid(RT), 256, 0, st>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, (__nv_bfloat16*)dy, S,
                                                            RT, C, scale, shift, mean, rstd, relu, sums,
                                                            1.0 / ((double)n_patches * P * P), dgamma, dbeta, dbias, P, n_patches);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// ---- weighted cross entropy ----------------------------------------------------------------------
// loss = sum_i w[y_i] * nll_i / sum_i w[y_i]   (nn.CrossEntropyLoss(weight=w), reduction 'mean',
// ignore_index -100); dlogits = grad_scale * w[y_i] * (softmax - onehot) / sum w.
// Pass 1: per-sample softmax (kept unnormalised in dlogits) + block sums -> double atomics;
// pass 2: scale by 1 / sum w.  acc: 2 doubles, zero on entry, zeroed again by pass 2.
__global__ void __launch_bounds__(256) ce_loss_pass1_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                            const float* __restrict__ weight, int n, int K,
                                                            float* __restrict__ dlogits, double* __restrict__ acc) {
  __shared__ double s_num[8], s_den[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double num = 0.0, den = 0.0;
  if (i < n) {
    const long long yl = labels[i];
    const bool ok = yl >= 0 && yl < K;
    const float w = ok ? (weight ? weight[yl] : 1.f) : 0.f;
    const float* l = logits + (long long)i * K;
    float m = l[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, l[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(l[k] - m);
    if (ok) { num = (double)(w * (logf(se) + m - l[yl])); den = (double)w; }
    if (dlogits) {
      const float is = w / se;
      float* d = dlogits + (long long)i * K;
      for (int k = 0; k < K; ++k) d[k] = is * expf(l[k] - m) - (k == yl ? w : 0.f);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  if ((threadIdx.x & 31) == 0) { s_num[threadIdx.x >> 5] = num; s_den[threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += s_num[w]; b += s_den[w]; }
    atomicAdd(acc, a);
    atomicAdd(acc + 1, b);
  }
}
__global__ void __launch_bounds__(256) ce_loss_pass2_kernel(int total, float grad_scale, float* __restrict__ loss_out,
                                                            float* __restrict__ dlogits, const double* __restrict__ acc) {
  const double den = acc[1];
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_out) { loss_out[0] = (float)(acc[0] / den); loss_out[1] = (float)den; }
  if (!dlogits) return;
  const float inv = (float)((double)grad_scale / den);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) dlogits[i] *= inv;
}
__global__ void ce_loss_clear_kernel(double* acc) { acc[0] = 0.0; acc[1] = 0.0; }

int ce_loss_launch(const float* logits, const long long* labels, const float* weight, int n, int K, float grad_scale,
                   float* loss_out, float* dlogits, double* acc, cudaStream_t st) {
  if (n <= 0 || K < 1 || !acc) return VC_ERR_ARG;
  ce_loss_clear_kernel<<<1, 1, 0, st>>>(acc);
  ce_loss_pass1_kernel<<<(n + 255) / 256, 256, 0, st>>>(logits, labels, weight, n, K, dlogits, acc);
  int blocks = (n * K + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  ce_loss_pass2_kernel<<<dlogits ? blocks : 1, 256, 0, st>>>(n * K, grad_scale, loss_out, dlogits, acc);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// ---- Adam (torch.optim.Adam: L2 weight decay into the gradient, no amsgrad) ------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

int adam_launch(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                float wd, int step, float grad_scale, cudaStream_t st) {
  if (n <= 0 || step < 1) return VC_ERR_ARG;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  adam_kernel<<<flat_grid(n), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, sqrtf(bc2), grad_scale);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// Graph-safe variant: hyper-parameters and the step counter live in device memory so that a
// captured CUDA graph replays correctly.  hyper: [lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, -]
__global__ void adam_prep_kernel(float* hyper, int* step) {
  const int t = *step + 1;
  *step = t;
  hyper[5] = 1.f - powf(hyper[1], (float)t);
  hyper[6] = sqrtf(1.f - powf(hyper[2], (float)t));
}
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, const float* __restrict__ hyper, float grad_scale) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], bc1 = hyper[5], bc2s = hyper[6];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - (lr / bc1) * (mi / (sqrtf(vi) / bc2s + eps));
  }
}
int adam_dev_launch(float* p, const float* g, float* m, float* v, long long n, float* hyper, int* step, float grad_scale,
                    cudaStream_t st) {
  if (n <= 0 || !hyper || !step) return VC_ERR_ARG;
  adam_prep_kernel<<<1, 1, 0, st>>>(hyper, step);
  adam_dev_kernel<<<flat_grid(n), 256, 0, st>>>(p, g, m, v, n, hyper, grad_scale);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// ---- weight packing ------------------------------------------------------------------------------
// torch conv weight fp32 [cout][cin][taps] -> bf16 [nsplit][taps][S_in][ncta][8] of conv_sps_tc.
// transpose = 0: the forward operand (in = cin, out = cout).
// transpose = 1: the data-gradient operand: a conv from cout channels back to cin channels with
//                flipped taps, W_d[o = ci][i = co][tap] = W[co][ci][taps-1-tap].
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int cout, int cin, int taps, int transpose, int S_in,
                                   int n_out, int nsplit, __nv_bfloat16* __restrict__ dst) {
  const int ncta = n_out / nsplit;
  const int total = nsplit * taps * S_in * ncta * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx & 7;
    int t = idx >> 3;
    const int nn = t % ncta; t /= ncta;
    const int s = t % S_in; t /= S_in;
    const int tap = t % taps;
    const int sp = t / taps;
    const int o = sp * ncta + nn, i = s * 8 + k;
    float v = 0.f;
    if (!transpose) {
      if (o < cout && i < cin) v = w[((long long)o * cin + i) * taps + tap];
    } else {
      if (o < cin && i < cout) v = w[((long long)i * cin + o) * taps + (taps - 1 - tap)];
    }
    dst[idx] = __float2bfloat16_rn(v);
  }
}

int pack_conv_w_launch(const float* w, int cout, int cin, int taps, int transpose, int S_in, int n_out, int nsplit,
                       void* dst, cudaStream_t st) {
  if (nsplit < 1 || n_out % nsplit || S_in < 1) return VC_ERR_ARG;
  const int total = taps * S_in * n_out * 8;
  pack_conv_w_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, cout, cin, taps, transpose, S_in, n_out, nsplit,
                                                          (__nv_bfloat16*)dst);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// The per-step chores in one launch (blockIdx.y = job): weight packing of every conv (forward and data-gradient
// operands), bias copies, zeroing of the atomically accumulated small gradients and of the BatchNorm sums.
__global__ void __launch_bounds__(256) train_prep_kernel(PrepTable t) {
  const PrepJob& j = t.job[blockIdx.y];
  const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j.type == 0) {
    const int ncta = j.n_out / j.nsplit, taps = j.taps, S_in = j.S_in, cin = j.cin, cout = j.cout;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j.dst);
    for (long long idx = i0; idx < j.n; idx += stride) {
      const int k = (int)(idx & 7);
      int q = (int)(idx >> 3);
      const int nn = q % ncta; q /= ncta;
      const int s = q % S_in; q /= S_in;
      const int tap = q % taps;
      const int sp = q / taps;
      const int o = sp * ncta + nn, i = s * 8 + k;
      float v = 0.f;
      if (!j.transpose) {
        if (o < cout && i < cin) v = j.src[((long long)o * cin + i) * taps + tap];
      } else {
        if (o < cin && i < cout) v = j.src[((long long)i * cin + o) * taps + (taps - 1 - tap)];
      }
      dst[idx] = __float2bfloat16_rn(v);
    }
  } else if (j.type == 1) {
    float* dst = reinterpret_cast<float*>(j.dst);
    for (long long idx = i0; idx < j.n; idx += stride) dst[idx] = j.src[idx];
  } else if (j.type == 2) {
    float* dst = reinterpret_cast<float*>(j.dst);
    for (long long idx = i0; idx < j.n; idx += stride) dst[idx] = 0.f;
  } else {
    double* dst = reinterpret_cast<double*>(j.dst);
    for (long long idx = i0; idx < j.n; idx += stride) dst[idx] = 0.0;
  }
}

int train_prep_launch(const PrepTable* t, cudaStream_t st) {
  if (!t || t->n <= 0) return VC_OK;
  if (t->n > kMaxPrepJobs) return VC_ERR_ARG;
  train_prep_kernel<<<dim3(24, t->n), 256, 0, st>>>(*t);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// Table-driven copy of fp32 parameter tensors into a packed blob: segment = rows x cols matrix
// (row-major, contiguous) written as rows of `pitch` elements, bf16 or fp32.
__global__ void pack_segments_kernel(const float* __restrict__ flat, uint8_t* __restrict__ blob,
                                     const long long* __restrict__ segs, int nsegs) {
  for (int sgi = blockIdx.x; sgi < nsegs; sgi += gridDim.x) {
    const long long* sg = segs + 6 * sgi;   // src_off, dst_off, rows, cols, pitch, is_bf16
    const float* src = flat + sg[0];
    const int rows = (int)sg[2], cols = (int)sg[3], pitch = (int)sg[4];
    if (sg[5]) {
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(blob + sg[1]);
      for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) d[(e / cols) * pitch + e % cols] = __float2bfloat16_rn(src[e]);
    } else {
      float* d = reinterpret_cast<float*>(blob + sg[1]);
      for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) d[(e / cols) * pitch + e % cols] = src[e];
    }
  }
}

int pack_segments_launch(const float* flat, void* blob, const long long* segs, int nsegs, cudaStream_t st) {
  if (nsegs <= 0) return VC_ERR_ARG;
  pack_segments_kernel<<<nsegs, 256, 0, st>>>(flat, (uint8_t*)blob, segs, nsegs);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
