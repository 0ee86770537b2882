// Weight gradient of the thin 3x3 convs (LiDAR stem: <= 32 output x <= 16 input channels).
// On the tcgen05 kernel these layers pay a full M=128 instruction per 16 rows and tap for a
// 32 x 16 result (76 us each at batch 4096); here one WARP per tap runs mma.sync m16n8k16 with
// the ROW axis as K, both operands transposed on the fly by ldmatrix.trans straight from the
// SPS rows (8 channels = 16 bytes = one ldmatrix row), cp.async double-buffered 128-row chunks,
// and the same deterministic split-K reduction as the tensor-core kernel (wgrad_reduce_kernel).
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

struct WsArgs {
  const __nv_bfloat16* dy;   // [SA][RT][8]  M side (output channels), SA in {2, 4}
  const __nv_bfloat16* x;    // [2][RT][8]   N side (input channels), shifted per tap
  float* part;               // [gridDim.x][9][128][16]
  long long RT;
  int SA, ntiles, halo, pw;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

constexpr int kWsThreads = 288;   // 9 warps: one per tap

__global__ void __launch_bounds__(kWsThreads) wgrad_small_kernel(WsArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int HALO = a.halo, XR = 128 + 2 * HALO;
  const int MT = a.SA / 2;                                      // 16-channel output tiles
  const size_t buf_bytes = ((size_t)a.SA * 128 + 2 * (size_t)XR) * 16;
  const int shift = (warp / 3 - 1) * a.pw + (warp % 3 - 1);
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  auto issue = [&](int tile, int buf) {
    uint8_t* base = smem + (size_t)buf * buf_bytes;
    const long long R = HALO + (long long)tile * 128;           // first row of the chunk (with lead halo)
    const int n_dy = a.SA * 128, n_x = 2 * XR;
    for (int i = threadIdx.x; i < n_dy + n_x; i += kWsThreads) {
      if (i < n_dy) {
        const int s = i / 128, r = i - s * 128;
        cp_async16(base + (size_t)i * 16, a.dy + ((long long)s * a.RT + R + r) * 8);
      } else {
        const int k = i - n_dy, s = k / XR, r = k - s * XR;
        cp_async16(base + (size_t)i * 16, a.x + ((long long)s * a.RT + R - HALO + r) * 8);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int it = 0;
  const int first = blockIdx.x;
  if (first < a.ntiles) issue(first, 0);
  for (int tile = first; tile < a.ntiles; tile += gridDim.x, ++it) {
    const int next = tile + gridDim.x;
    if (next < a.ntiles) {
      issue(next, (it + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __nv_bfloat16* dy_s = reinterpret_cast<const __nv_bfloat16*>(smem + (size_t)(it & 1) * buf_bytes);
    const __nv_bfloat16* x_s = dy_s + (size_t)a.SA * 128 * 8;
    const int lm = lane >> 3, lr = lane & 7;
#pragma unroll 2
    for (int k0 = 0; k0 < 128; k0 += 16) {
      uint32_t B[4];   // both input-channel tiles: (slice 0, rows k0..k0+7), (slice 0, +8), (slice 1, ..), (slice 1, +8)
      ldsm4_t(B, x_s + ((size_t)(lm >> 1) * XR + HALO + k0 + shift + lr + 8 * (lm & 1)) * 8);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt < MT) {
          uint32_t A[4];   // (rows k0.., slice 2mt) (rows k0.., slice 2mt+1) (rows k0+8.., slice 2mt) (rows k0+8.., slice 2mt+1)
          ldsm4_t(A, dy_s + ((size_t)(2 * mt + (lm & 1)) * 128 + k0 + lr + 8 * (lm >> 1)) * 8);
          mma16816(acc[mt][0], A, B[0], B[1]);
          mma16816(acc[mt][1], A, B[2], B[3]);
        }
      }
    }
    __syncthreads();
  }
  // partial [tap][m][n] of this CTA in the layout wgrad_reduce_kernel reads ([9][128][16])
  float* dst = a.part + ((long long)blockIdx.x * 9 + warp) * 128 * 16;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
    if (mt < MT)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int m = 16 * mt + g, n = 8 * nt + 2 * q;
        *reinterpret_cast<float2*>(dst + m * 16 + n) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<float2*>(dst + (m + 8) * 16 + n) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      }
}

size_t wgrad_small_workspace_bytes() { return (size_t)148 * 2 * 9 * 128 * 16 * sizeof(float); }

// dW[co][ci][tap] (torch layout) of a 3x3 conv with cout <= 32 (SA = 2 or 4 slices of dy) and cin <= 16
// (x has 2 slices); returns VC_ERR_UNSUPPORTED for anything else.
int wgrad_small_launch(const void* dy, int SA, const void* x, int SB, int n_patches, int P, void* workspace, float* out,
                       int cout, int cin, cudaStream_t stream) {
  if ((SA != 2 && SA != 4) || SB != 2 || cout > SA * 8 || cin > 16 || n_patches <= 0 || P < 1) return VC_ERR_UNSUPPORTED;
  WsArgs a;
  a.dy = (const __nv_bfloat16*)dy;
  a.x = (const __nv_bfloat16*)x;
  a.part = (float*)workspace;
  a.RT = sps_rows(n_patches, P);
  a.SA = SA;
  a.ntiles = sps_tiles(n_patches, P);
  a.halo = sps_halo(P);
  a.pw = P + 1;
  const size_t smem = 2 * (((size_t)SA * 128 + 2 * (size_t)(128 + 2 * a.halo)) * 16);
  if (cudaFuncSetAttribute(wgrad_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VC_ERR_CUDA;
  int gx = 148 * 2;
  if (gx > a.ntiles) gx = a.ntiles;
  wgrad_small_kernel<<<gx, kWsThreads, smem, stream>>>(a);
  if (cudaGetLastError() != cudaSuccess) return VC_ERR_CUDA;
  return wgrad_reduce_launch(a.part, gx, 9, 16, cout, cin, out, (long long)cin * 9, 9, 1, -1, nullptr, 0, stream);
}

}  // namespace vc
