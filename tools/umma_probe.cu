// Micro-benchmark / correctness probe for tcgen05.mma shared-memory operand layouts on sm_100a.
// Question: which K-major layout lets the A operand be addressed at an arbitrary ROW SHIFT
// (the shifted-GEMM convolution trick) at full tensor-pipe rate?
//   mode 0: no swizzle, [k-slice][row][16 B]      (shift = start address + 16 B * s)
//   mode 1: SWIZZLE_128B, rows of 128 B (64 ch)    (shift = start + 128 B * s, base_offset = s % 8)
//   mode 2: SWIZZLE_32B,  rows of 32 B (16 ch)     (shift = start + 32 B * s)
//   mode 3: SWIZZLE_64B,  rows of 64 B (32 ch)     (shift = start + 64 B * s)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe.bin tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>
#include "../vit-cnn_b200/csrc/vc_common.cuh"

using namespace vc;

constexpr int ROWS_A = 192;   // rows resident for A (128 + shifts)
constexpr int KCH = 64;       // channels (4 K-steps of 16)

__host__ __device__ inline uint32_t a_offset(int mode, int rows, int r, int c) {
  switch (mode) {
    case 0: return ((c / 8) * rows + r) * 16 + (c % 8) * 2;
    case 1: return r * 128 + (((c / 8) ^ (r & 7)) * 16) + (c % 8) * 2;
    case 2: return (c / 16) * rows * 32 + r * 32 + ((((c % 16) / 8) ^ ((r >> 2) & 1)) * 16) + (c % 8) * 2;
    default: return (c / 32) * rows * 64 + r * 64 + ((((c % 32) / 8) ^ ((r >> 1) & 3)) * 16) + (c % 8) * 2;
  }
}

__device__ inline uint64_t make_desc(int mode, uint32_t addr, uint32_t lbo, uint32_t sbo, int bo_mode) {
  uint64_t d = umma_desc(addr, lbo, sbo);
  const uint64_t lt = mode == 0 ? 0 : mode == 1 ? 2 : mode == 2 ? 6 : 4;
  d |= lt << 61;
  uint32_t bo = 0;
  if (mode != 0 && bo_mode == 1) bo = (addr >> 7) & 7;
  d |= (uint64_t)bo << 49;
  return d;
}

struct Args {
  int mode, shift, N, iters, bo_mode;
  const __nv_bfloat16* A;  // [ROWS_A][KCH] logical
  const __nv_bfloat16* B;  // [N][KCH] logical
  float* D;                // [128][N]
  long long* cycles;       // per CTA
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* As = smem;                       // 192 rows * 128 B = 24 KB max
  uint8_t* Bs = smem + 32768;               // N rows * 128 B = up to 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < ROWS_A * KCH; i += 128) {
    const int r = i / KCH, c = i % KCH;
    *reinterpret_cast<__nv_bfloat16*>(As + a_offset(a.mode, ROWS_A, r, c)) = a.A[i];
  }
  for (int i = tid; i < a.N * KCH; i += 128) {
    const int r = i / KCH, c = i % KCH;
    *reinterpret_cast<__nv_bfloat16*>(Bs + a_offset(a.mode, a.N, r, c)) = a.B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t idesc = umma_idesc_bf16(128, a.N);
  if (tid == 0) {
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    long long t0 = clock64();
    for (int it = 0; it < a.iters; ++it) {
      for (int ks = 0; ks < KCH / 16; ++ks) {
        uint64_t ad, bd;
        if (a.mode == 0) {
          ad = make_desc(0, a0 + (2 * ks * ROWS_A + a.shift) * 16, ROWS_A * 16, 128, 0);
          bd = make_desc(0, b0 + (2 * ks * a.N) * 16, a.N * 16, 128, 0);
        } else if (a.mode == 1) {
          ad = make_desc(1, a0 + a.shift * 128 + ks * 32, 16, 1024, a.bo_mode);
          bd = make_desc(1, b0 + ks * 32, 16, 1024, a.bo_mode);
        } else if (a.mode == 2) {
          ad = make_desc(2, a0 + ks * ROWS_A * 32 + a.shift * 32, 16, 256, a.bo_mode);
          bd = make_desc(2, b0 + ks * a.N * 32, 16, 256, a.bo_mode);
        } else {
          ad = make_desc(3, a0 + (ks / 2) * ROWS_A * 64 + a.shift * 64 + (ks & 1) * 32, 16, 512, a.bo_mode);
          bd = make_desc(3, b0 + (ks / 2) * a.N * 64 + (ks & 1) * 32, 16, 512, a.bo_mode);
        }
        umma_bf16(tmem, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    a.cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  if (blockIdx.x == 0) {
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < a.N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tc_wait_ld();
      for (int k = 0; k < 16; ++k) a.D[row * a.N + c0 + k] = __uint_as_float(v[k]);
    }
  } else {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc_wait_ld();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const int only_mode = argc > 2 ? atoi(argv[2]) : -1;
  std::vector<__nv_bfloat16> hA(ROWS_A * KCH), hB(256 * KCH);
  std::vector<float> fA(ROWS_A * KCH), fB(256 * KCH);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 7) - 3); hA[i] = __float2bfloat16(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 5) - 2); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dD, 128 * 256 * 4);
  cudaMalloc(&dC, 1024 * 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int shifts[] = {0, 1, 3, 8, 13, 37};
  const int Ns[] = {32, 64, 128, 256};
  printf("mode shift N bo | max_err | cycles/MMA(K=16) median-CTA | ideal\n");
  for (int mode = 0; mode < 4; ++mode) {
    if (only_mode >= 0 && mode != only_mode) continue;
    for (int bo = 0; bo < (mode == 0 ? 1 : 2); ++bo)
      for (int N : Ns)
        for (int shift : shifts) {
          for (int pass = 0; pass < 2; ++pass) {   // pass 0: 1 iteration for the numeric check; pass 1: timing
            Args a{mode, shift, N, pass == 0 ? 1 : 256, bo, dA, dB, dD, dC};
            probe_kernel<<<pass == 0 ? 1 : grid, 128, 65536 + 1024>>>(a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d shift %d N %d: CUDA error %s\n", mode, shift, N, cudaGetErrorString(e)); return 1; }
            if (pass == 0) {
              std::vector<float> D(128 * N);
              cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
              double me = 0;
              for (int r = 0; r < 128; ++r)
                for (int n = 0; n < N; ++n) {
                  double s = 0;
                  for (int c = 0; c < KCH; ++c) s += (double)fA[(r + shift) * KCH + c] * fB[n * KCH + c];
                  me = fmax(me, fabs(s - D[r * N + n]));
                }
              printf("%d %2d %3d %d | %8.1f | ", mode, shift, N, bo, me);
            } else {
              std::vector<long long> c(grid);
              cudaMemcpy(c.data(), dC, grid * 8, cudaMemcpyDeviceToHost);
              std::sort(c.begin(), c.end());
              printf("%7.1f | %d\n", (double)c[grid / 2] / (256.0 * 4), 128 * N / 256);
            }
          }
        }
  }
  return 0;
}
