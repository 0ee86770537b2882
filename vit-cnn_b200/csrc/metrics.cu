// Prediction post-processing on the device: the confusion matrix behind the reference's
// metrics() (utils.py:585-663: sklearn.metrics.confusion_matrix(target, prediction,
// labels=range(n_classes)) over the pixels whose TARGET is not an ignored label, :595-601).
// Integer work, bit-exact: per-block shared-memory histogram, one 64-bit atomic per non-zero
// cell per block.  OA / AA / kappa / F1 are a few flops on the K x K matrix and stay on the host.
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

__device__ __forceinline__ long long load_label(const void* p, int eb, long long i) {
  if (eb == 1) return reinterpret_cast<const unsigned char*>(p)[i];
  if (eb == 4) return reinterpret_cast<const int*>(p)[i];
  return reinterpret_cast<const long long*>(p)[i];
}

__global__ void __launch_bounds__(256) confusion_kernel(const void* pred, int peb, const void* target, int teb, long long n, int K,
                                                        unsigned long long ignored_mask, unsigned long long* cm) {
  extern __shared__ unsigned int hist[];   // [K*K]
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = load_label(target, teb, i);
    if (t >= 0 && t < 64 && ((ignored_mask >> t) & 1ull)) continue;     // ignored target label
    const long long p = load_label(pred, peb, i);
    if (t < 0 || t >= K || p < 0 || p >= K) continue;                      // outside labels=range(K): dropped by sklearn
    atomicAdd(&hist[(int)t * K + (int)p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

int confusion_launch(const void* pred, int peb, const void* target, int teb, long long n, int K, unsigned long long ignored_mask,
                     long long* cm, cudaStream_t stream) {
  if (n < 0 || K < 1 || K > 64 || (peb != 1 && peb != 4 && peb != 8) || (teb != 1 && teb != 4 && teb != 8) || !cm) return VC_ERR_ARG;
  if (n == 0) return VC_OK;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  confusion_kernel<<<(int)blocks, 256, (size_t)K * K * sizeof(unsigned int), stream>>>(pred, peb, target, teb, n, K, ignored_mask,
                                                                                    reinterpret_cast<unsigned long long*>(cm));
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc

// ------------------------------------------------------------------------------------------------
// Raster ingest: per-band min-max normalisation to [0, 1] (datasets.py:124-133 for the HSI cube, and
// the whole-array form used for the LiDAR raster), in place on the pixel-interleaved fp32 raster:
//   x <- (x - min_band) / (max_band - min_band)      (float32 subtract and divide: bit-exact with numpy)
namespace vc {

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) > v) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) < v) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}

// mm[c] = min, mm[Cs + c] = max over all pixels (Cs = C per band, 1 when global)
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ img, long long npix, int C, int per_band,
                                                     float* __restrict__ mm) {
  const int Cb = C < 256 ? C : 256;               // channels handled side by side
  const int lanes = 256 / Cb;                     // pixels handled side by side
  const int tc = threadIdx.x % Cb, tp = threadIdx.x / Cb;
  if (tp >= lanes) return;
  for (int c0 = 0; c0 < C; c0 += Cb) {
    const int c = c0 + tc;
    if (c >= C) continue;
    float lo = INFINITY, hi = -INFINITY;
    for (long long p = (long long)blockIdx.x * lanes + tp; p < npix; p += (long long)gridDim.x * lanes) {
      const float v = __ldg(img + p * C + c);
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    const int slot = per_band ? c : 0, Cs = per_band ? C : 1;
    if (lo <= hi) {
      atomic_min_f(mm + slot, lo);
      atomic_max_f(mm + Cs + slot, hi);
    }
  }
}
__global__ void minmax_init_kernel(float* mm, int Cs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Cs) { mm[i] = INFINITY; mm[Cs + i] = -INFINITY; }
}
__global__ void __launch_bounds__(256) minmax_apply_kernel(float* __restrict__ img, long long total, int C, int per_band,
                                                           const float* __restrict__ mm) {
  const int Cs = per_band ? C : 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = per_band ? (int)(i % C) : 0;
    const float lo = mm[c], hi = mm[Cs + c];
    img[i] = (img[i] - lo) / (hi - lo);
  }
}

int minmax_normalise_launch(float* img, long long npix, int C, int per_band, float* scratch, cudaStream_t stream) {
  if (npix <= 0 || C < 1 || !img || !scratch) return VC_ERR_ARG;
  const int Cs = per_band ? C : 1;
  minmax_init_kernel<<<(Cs + 255) / 256, 256, 0, stream>>>(scratch, Cs);
  minmax_kernel<<<148 * 4, 256, 0, stream>>>(img, npix, C, per_band, scratch);
  long long blocks = (npix * C + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  minmax_apply_kernel<<<(int)blocks, 256, 0, stream>>>(img, npix * C, C, per_band, scratch);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
