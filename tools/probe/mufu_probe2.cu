// Per-warp issue limits on sm_100a: MUFU.EX2 / FFMA / packed FFMA2 throughput as a function of resident warps per SM
// sub-partition, and an FMA-pipe exp2 (Cody-Waite + degree-3 polynomial) next to the hardware one.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe2.bin mufu_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ float ex2(float v) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float ex2_poly(float s) {
  const float x = fmaxf(s, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05587554f, 0.24229463f);
  p = fmaf(p, f, 0.69312726f);
  p = fmaf(p, f, 0.99994823f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <int MODE>
__global__ void k(float* out, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + 0.01f * i;
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) { for (int i = 0; i < 8; ++i) a[i] = ex2(a[i]) - 1.0f; }                         // 8 MUFU + 8 FADD
    if (MODE == 1) { for (int i = 0; i < 8; ++i) a[i] = ex2_poly(a[i]) - 1.0f; }                    // 8 poly
    if (MODE == 2) { for (int i = 0; i < 4; ++i) a[i] = ex2(a[i]) - 1.0f; for (int i = 4; i < 8; ++i) a[i] = ex2_poly(a[i]) - 1.0f; }
    if (MODE == 3) { for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f); }                // 8 FFMA (imm form)
    if (MODE == 4) { for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], a[(i + 1) & 7], a[(i + 3) & 7]); }   // 8 FFMA 3-reg
    if (MODE == 5) {                                                                                 // 4 FFMA2 = 8 fma
      for (int i = 0; i < 8; i += 2) {
        unsigned long long x, y, z;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(a[(i + 2) & 7]), "f"(a[(i + 3) & 7]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(z) : "l"(x), "l"(y));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(z));
      }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int warps_per_smsp) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int threads = warps_per_smsp * 4 * 32;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(d, 0.5f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148, threads>>>(d, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-26s warps/SMSP %d: %7.3f ms  %6.2f cycles per 8-element iteration per warp, %6.2f elements/clk/SM (%s)\n", name,
         warps_per_smsp, ms, cycles / ITERS, 8.0 * 32 * warps_per_smsp * 4 * ITERS / cycles, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  for (int w : {1, 2, 4, 8}) run<0>("ex2 MUFU (+fadd)", w);
  for (int w : {1, 2, 4, 8}) run<1>("ex2 poly (fma pipe)", w);
  for (int w : {1, 2, 4, 8}) run<2>("half MUFU / half poly", w);
  for (int w : {1, 2, 4, 8}) run<3>("ffma imm", w);
  for (int w : {1, 2, 4, 8}) run<4>("ffma 3-reg", w);
  for (int w : {1, 2, 4, 8}) run<5>("ffma2 packed", w);
  // accuracy of the polynomial
  return 0;
}
