// LiDAR branch of the stem (eval mode) as ONE kernel: conv3x3(C2->8) -> conv3x3(8->16) ->
// conv3x3(16->32), each + folded BatchNorm + ReLU (S2ENet planes_b idiom, SURVEY.md App. A;
// model/Multimodality_Mamba/Mutimodality_Mamba7.py:1146).  1.4 MFLOP per window: far too thin
// for the tcgen05 pipeline (three latency-bound launches of N <= 32 MMAs cost 11 ms per Houston
// scene), so here one WARP owns one patch, keeps all three activation maps in shared memory
// in the same padded-grid row order as SPS (a tap is a row shift, zero pad cells give the
// 'same' padding) and runs the convs as mma.sync implicit GEMMs fed by ldmatrix:
//   * conv 1 and 2 have <= 8 input channels: two taps share one K=16 step (k 0-7 = tap 2p,
//     k 8-15 = tap 2p+1), 5 MMAs per output tile instead of 9;
//   * conv 3: K = 16 channels per tap, 4 output-channel tiles.
// Output goes straight into slices [slice_off, slice_off+4) of the fused feature buffer.
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

constexpr int kLwPitch = 24;                  // weight row pitch (elements): 48 B, conflict-free ldmatrix
constexpr int kLw1Rows = 5 * 8, kLw2Rows = 5 * 16, kLw3Rows = 9 * 32;
constexpr int kLwRows = kLw1Rows + kLw2Rows + kLw3Rows;
constexpr int kLidarMaxWarps = 16;

struct LidarArgs {
  const __nv_bfloat16* in;    // [S2][RT][8]   (slice 0 holds the <= 8 LiDAR bands)
  const uint8_t* blob;        // bf16 weights [kLwRows][24] then fp32 scale1[8] bias1[8] scale2[16] bias2[16] scale3[32] bias3[32]
  __nv_bfloat16* out;         // [.][RT][8]
  long long RT;
  int n_patches, P, out_slice_off;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}

__global__ void __launch_bounds__(kLidarMaxWarps * 32, 1) lidar_stem_kernel(LidarArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int P = a.P, PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P);
  const int MT = (PP + 15) / 16;
  const int rows = HALO + 16 * MT + HALO;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int kLidarWarps = blockDim.x >> 5;
  __nv_bfloat16* w_s = reinterpret_cast<__nv_bfloat16*>(smem);
  const float* aff = reinterpret_cast<const float*>(smem + kLwRows * kLwPitch * 2);
  const size_t wbytes = (size_t)kLwRows * kLwPitch * 2 + 112 * 4;
  // per-warp activation maps: in [rows][8], z1 [rows][8], z2 [rows][16]
  __nv_bfloat16* in_s = reinterpret_cast<__nv_bfloat16*>(smem + ((wbytes + 15) & ~size_t(15))) + (size_t)warp * rows * 32;
  __nv_bfloat16* z1_s = in_s + rows * 8;
  __nv_bfloat16* z2_s = z1_s + rows * 8;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < (int)(wbytes / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    uint4* z = reinterpret_cast<uint4*>(in_s);
    for (int i = lane; i < rows * 4; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);   // halos stay zero forever
  }
  __syncthreads();
  const __nv_bfloat16* w1 = w_s;
  const __nv_bfloat16* w2 = w_s + kLw1Rows * kLwPitch;
  const __nv_bfloat16* w3 = w2 + kLw2Rows * kLwPitch;
  const float *sc1 = aff, *bi1 = aff + 8, *sc2 = aff + 16, *bi2 = aff + 32, *sc3 = aff + 48, *bi3 = aff + 80;
  // ldmatrix lane roles: matrix m = lane / 8 supplies rows (lane % 8) (+8 for odd m); for the
  // paired-tap A operand matrices 2,3 are the second tap, otherwise the upper K half
  const int lm = lane >> 3, lr = (lane & 7) + 8 * (lm & 1);

  for (int b = blockIdx.x * kLidarWarps + warp; b < a.n_patches; b += gridDim.x * kLidarWarps) {
    const long long gbase = HALO + (long long)b * PP;
    for (int r = lane; r < PP; r += 32)
      *reinterpret_cast<uint4*>(in_s + (HALO + r) * 8) = __ldg(reinterpret_cast<const uint4*>(a.in + (gbase + r) * 8));
    __syncwarp();

    // ---- conv 1 (C2 -> 8) and conv 2 (8 -> 16): two taps per K=16 step ----
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      const __nv_bfloat16* src = layer ? z1_s : in_s;
      const __nv_bfloat16* w = layer ? w2 : w1;
      const int NTL = layer ? 2 : 1;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int p = 0; p < 5; ++p) {
          const int t0 = 2 * p, t1 = (2 * p + 1 < 9) ? 2 * p + 1 : 8;   // tap 9 does not exist: its weights are zero
          const int tap = (lm < 2) ? t0 : t1;
          const int shift = (tap / 3 - 1) * PW + (tap % 3 - 1);
          uint32_t A[4];
          ldsm_x4(A, src + (HALO + 16 * mt + lr + shift) * 8);
          if (layer == 0) {
            uint32_t b0, b1;
            ldsm_x2(b0, b1, w + (p * 8 + (lane & 7)) * kLwPitch + 8 * ((lane >> 3) & 1));
            mma16816(acc[0], A, b0, b1);
          } else {
            uint32_t B[4];
            ldsm_x4(B, w + (p * 16 + (lane & 7) + 8 * (lm >> 1)) * kLwPitch + 8 * (lm & 1));
            mma16816(acc[0], A, B[0], B[1]);
            mma16816(acc[1], A, B[2], B[3]);
          }
        }
        // epilogue: affine + ReLU, zero on pad cells, bf16 into the next map
        const float* sc = layer ? sc2 : sc1;
        const float* bi = layer ? bi2 : bi1;
        __nv_bfloat16* dst = layer ? z2_s : z1_s;
        const int pitch = layer ? 16 : 8;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = 16 * mt + g + 8 * hh;
          const int i = r / PW, j = r - i * PW;
          const bool valid = r < PP && i < P && j < P;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            if (nt < NTL) {
              const int c = 8 * nt + 2 * q;
              float y0 = fmaxf(acc[nt][2 * hh] * sc[c] + bi[c], 0.f), y1 = fmaxf(acc[nt][2 * hh + 1] * sc[c + 1] + bi[c + 1], 0.f);
              if (!valid) { y0 = 0.f; y1 = 0.f; }
              *reinterpret_cast<uint32_t*>(dst + (HALO + r) * pitch + c) = pack_bf16(y0, y1);
            }
          }
        }
      }
      __syncwarp();
    }

    // ---- conv 3 (16 -> 32): K = 16 channels per tap, 4 output-channel tiles ----
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int shift = (tap / 3 - 1) * PW + (tap % 3 - 1);
        uint32_t A[4], B0[4], B1[4];
        ldsm_x4(A, z2_s + (HALO + 16 * mt + lr + shift) * 16 + 8 * (lm >> 1));
        ldsm_x4(B0, w3 + (tap * 32 + (lane & 7) + 8 * (lm >> 1)) * kLwPitch + 8 * (lm & 1));
        ldsm_x4(B1, w3 + (tap * 32 + 16 + (lane & 7) + 8 * (lm >> 1)) * kLwPitch + 8 * (lm & 1));
        mma16816(acc[0], A, B0[0], B0[1]);
        mma16816(acc[1], A, B0[2], B0[3]);
        mma16816(acc[2], A, B1[0], B1[1]);
        mma16816(acc[3], A, B1[2], B1[3]);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = 16 * mt + g + 8 * hh;
        const int i = r / PW, j = r - i * PW;
        if (r >= PP) continue;
        const bool valid = i < P && j < P;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c = 8 * nt + 2 * q;
          float y0 = fmaxf(acc[nt][2 * hh] * sc3[c] + bi3[c], 0.f), y1 = fmaxf(acc[nt][2 * hh + 1] * sc3[c + 1] + bi3[c + 1], 0.f);
          if (!valid) { y0 = 0.f; y1 = 0.f; }
          *reinterpret_cast<uint32_t*>(a.out + ((long long)(a.out_slice_off + nt) * a.RT + gbase + r) * 8 + 2 * q) = pack_bf16(y0, y1);
        }
      }
    }
    __syncwarp();
  }
}

size_t lidar_blob_bytes() { return (size_t)kLwRows * kLwPitch * 2 + 112 * 4; }

int lidar_stem_launch(const void* in_sps, const void* blob, void* out_sps, int out_slice_off, int n_patches, int P,
                      cudaStream_t stream) {
  if (n_patches <= 0 || P < 1 || P > 15 || !blob) return VC_ERR_ARG;
  LidarArgs a;
  a.in = (const __nv_bfloat16*)in_sps;
  a.blob = (const uint8_t*)blob;
  a.out = (__nv_bfloat16*)out_sps;
  a.RT = sps_rows(n_patches, P);
  a.n_patches = n_patches;
  a.P = P;
  a.out_slice_off = out_slice_off;
  const int MT = (sps_pp(P) + 15) / 16, rows = 2 * sps_halo(P) + 16 * MT;
  static int max_smem = 0, num_sms = 0;
  if (!max_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const size_t wb = (lidar_blob_bytes() + 15) & ~size_t(15);
  int nwarps = (int)(((size_t)max_smem - wb) / ((size_t)rows * 64));
  if (nwarps > kLidarMaxWarps) nwarps = kLidarMaxWarps;
  if (nwarps < 4) return VC_ERR_UNSUPPORTED;
  const size_t smem = wb + (size_t)nwarps * rows * 64;
  if (cudaFuncSetAttribute(lidar_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VC_ERR_CUDA;
  int blocks = (n_patches + nwarps - 1) / nwarps;
  if (blocks > num_sms) blocks = num_sms;
  lidar_stem_kernel<<<blocks, nwarps * 32, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
