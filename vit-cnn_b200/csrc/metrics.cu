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
