// Throughput probe for the transcendental unit on sm_100a: which exp2 / tanh forms are native.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe.bin mufu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float seed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t u0 = __float_as_uint(a0) & 0x3bff3bffu, u1 = u0 ^ 0x00010001u, u2 = u0 ^ 0x00020002u, u3 = u0 ^ 0x00030003u;
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u3));
    } else if (MODE == 2) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u3));
    } else if (MODE == 3) {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 4) {
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u0)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u1));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u2)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u3));
    } else if (MODE == 5) {
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u0)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u1));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u2)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u3));
    } else if (MODE == 6) {   // FFMA reference
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(a1), "f"(a2)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(a2), "f"(a3));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(a3), "f"(a0)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(a0), "f"(a1));
    } else if (MODE == 7) {   // cvt pack f32 -> bf16x2
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u0) : "f"(a0), "f"(a1)); a0 += __uint_as_float(u0);
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u1) : "f"(a2), "f"(a3)); a2 += __uint_as_float(u1);
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u2) : "f"(a1), "f"(a2)); a1 += __uint_as_float(u2);
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u3) : "f"(a3), "f"(a0)); a3 += __uint_as_float(u3);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
}
template <int MODE>
void run(const char* name, int per_inst) {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 1024>>>(d, 0.5f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 8, 1024>>>(d, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = 148.0 * 8 * 1024 * ITERS * 4;   // thread-level instructions
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double per_clk_sm = inst / (ms * 1e-3) / (clk * 1e3) / 148;
  printf("%-22s %8.3f ms  %6.1f lane-inst/clk/SM  %6.1f results/clk/SM  (%s)\n", name, ms, per_clk_sm, per_clk_sm * per_inst, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<0>("ex2.f32", 1); run<1>("ex2.f16x2", 2); run<2>("ex2.bf16x2", 2); run<3>("tanh.f32", 1);
  run<4>("tanh.f16x2", 2); run<5>("tanh.bf16x2", 2); run<6>("ffma", 1); run<7>("cvt.bf16x2+fadd", 1);
  return 0;
}
