// Second probe: which ingredient of the conv kernel's issue pattern doubles the per-MMA time
// relative to the plain back-to-back loop (53 cycles at M=128, N=64, no swizzle)?
// variant bits: 1 = A row shift changes every MMA (9 taps), 2 = B tile changes every MMA (9 taps),
// 4 = tcgen05.commit every 9 MMAs, 8 = three other warps stream tcgen05.ld concurrently,
// 16 = alternate between two accumulators every 81 MMAs, 32 = fence::after_thread_sync every 9 MMAs
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../vit-cnn_b200/csrc/vc_common.cuh"
using namespace vc;

constexpr int ROWS = 160, S_IN = 2;   // one conv stage: 2 slices x 160 rows x 16 B

__global__ void __launch_bounds__(128, 1) probe2(int variant, int N, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // variant 64: conv2's real layout (weights [9][16][64][16 B] = 147 KB first, stages after);
  // variant 128: stages first, big weight block after
  const bool big = (variant & (64 | 128)) != 0;
  uint8_t* As = (variant & 64) ? smem + 147456 : smem;               // 8 stages x 5 KB
  uint8_t* Bs = (variant & 64) ? smem : smem + 8 * 5120;             // weights
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (big ? (8 * 5120 + 147456) : (8 * 5120 + 9 * 2 * 256 * 16)) / 4; i += 128)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t idesc = umma_idesc_bf16(128, N);
  if (tid == 0) {
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    const uint64_t a_hi = umma_desc(0, ROWS * 16, 128) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = umma_desc(0, N * 16, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t a_l = (uint32_t)(umma_desc(0, ROWS * 16, 128) & 0xFFFF0000u);
    const uint32_t b_l = (uint32_t)(umma_desc(0, N * 16, 128) & 0xFFFF0000u);
    long long t0 = clock64();
    int n = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (((variant & 16) && (it & 1)) ? 256u : 0u);
      for (int ks = 0; ks < 9; ++ks) {
        const uint32_t a_lo = a_l | (((a0 + (ks & 7) * 5120) >> 4) + 16);
        const uint32_t b_lo = (b_l | (b0 >> 4)) + (big ? (uint32_t)((ks & 7) * 2 * N) : 0u);
        if (variant & 32) tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int shift = (variant & 1) ? ((tap / 3 - 1) * 12 + (tap % 3 - 1)) : 0;
          const uint32_t boff = (variant & 2) ? (uint32_t)(tap * (big ? 16 : 2) * N) : 0u;
          umma_bf16(d, a_hi | (uint64_t)(a_lo + (uint32_t)shift), b_hi | (uint64_t)(b_lo + boff), idesc, (ks | tap) != 0);
          ++n;
        }
        if (variant & 4) umma_commit(&bar2);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = (t1 - t0) / n;
    stop = 1;
  } else if (warp > 0 && (variant & 8)) {
    uint32_t v[16];
    while (!stop) {
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 384, v);
      tc_wait_ld();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* dC;
  cudaMalloc(&dC, 148 * 8);
  const int smem = 8 * 5120 + 147456 + 1024;
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int variants[] = {7, 64 | 7};
  for (int N : {32, 64, 128, 256})
    for (int iters : {4, 16, 64, 256, 1024})
    for (int v : variants) {
      probe2<<<148, 128, smem>>>(v, N, iters, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N %d variant %d: %s\n", N, v, cudaGetErrorString(e)); return 1; }
      std::vector<long long> c(148);
      cudaMemcpy(c.data(), dC, 148 * 8, cudaMemcpyDeviceToHost);
      std::sort(c.begin(), c.end());
      printf("N %3d variant %3d iters %4d : %lld cycles/MMA (median CTA; min %lld max %lld), ideal %d\n", N, v, iters, c[74], c[0], c[147], 128 * N / 256);
    }
  return 0;
}
