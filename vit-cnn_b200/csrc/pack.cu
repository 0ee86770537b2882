// Patch extraction kernels.
//  * gather_f32: the reference's per-pixel patch slice + HWC->CHW copy, bit-exact fp32
//    (datasets.py:550-593 MultiModalX.__getitem__ + default_collate; model_utils.py:1103-1112
//    batch assembly of test()).  HBM-bound; coalesced reads along channels, shared-memory
//    transpose, vectorised coalesced writes along pixels.
//  * pack_sps: the same gather fused with fp32->bf16 (RNE) conversion into the SPS layout the
//    tensor-core stem consumes (zero pad cells, zero halos).
//  * scene_index: window enumeration of utils.sliding_window (utils.py:357-401) for a chunk.
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

// ------------------------------------------------------------------------------------------
// One thread owns one SPS row (= one pixel of one patch, or a pad / halo row) and walks its
// channel slices: row decode and source address are computed once per row instead of once per
// 16 bytes, a warp writes 32 consecutive rows of a slice (512 contiguous bytes), and its reads
// are whole 32-byte sectors of the pixel's channel vector.
__global__ void __launch_bounds__(256) pack_sps_kernel(const float* __restrict__ src, long long sb, long long sc,
                                                       long long si, long long sj, const long long* __restrict__ patch_off,
                                                       int n_patches, int C, int P, __nv_bfloat16* __restrict__ sps, int S,
                                                       long long RT, int vec) {
  const int HALO = sps_halo(P), PP = sps_pp(P), PW = P + 1;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    const long long r = R - HALO;
    const float* p = nullptr;
    if (r >= 0) {
      const long long b = r / PP;
      const int q = (int)(r - b * PP);
      const int i = q / PW, j = q - i * PW;
      if (b < n_patches && i < P && j < P) p = src + (patch_off ? patch_off[b] : b * sb) + i * si + j * sj;
    }
    uint4* dst = reinterpret_cast<uint4*>(sps) + R;
    if (!p) {
      for (int s = 0; s < S; ++s) dst[(long long)s * RT] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const int full = vec ? C / 8 : 0;     // slices that are two aligned float4 loads
#pragma unroll 6
    for (int s = 0; s < full; ++s) {
      const float4 lo = __ldg(reinterpret_cast<const float4*>(p + s * 8));
      const float4 hi = __ldg(reinterpret_cast<const float4*>(p + s * 8 + 4));
      dst[(long long)s * RT] = make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
    }
    for (int s = full; s < S; ++s) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (s * 8 + k < C) ? __ldg(p + (long long)(s * 8 + k) * sc) : 0.f;
      dst[(long long)s * RT] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  }
}

int pack_sps_launch(const float* src, long long sb, long long sc, long long si, long long sj, const long long* patch_off,
                    int n_patches, int C, int P, void* sps, int S, cudaStream_t stream) {
  if (n_patches <= 0 || C <= 0 || S * 8 < C || P < 1) return VC_ERR_ARG;
  const long long RT = sps_rows(n_patches, P);
  const int vec = (sc == 1 && si % 4 == 0 && sj % 4 == 0 && sb % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0)
                      ? 1
                      : 0;  // patch_off entries are multiples of C in raster mode: 16-byte aligned iff C % 4 == 0
  long long blocks = (RT + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  pack_sps_kernel<<<(int)blocks, 256, 0, stream>>>(src, sb, sc, si, sj, patch_off, n_patches, C, P,
                                                   (__nv_bfloat16*)sps, S, RT, vec && (C % 4 == 0 || !patch_off));
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// Zero the lead / trailing halo rows of an SPS buffer (the conv epilogue writes the rest).
__global__ void zero_halo_kernel(__nv_bfloat16* sps, int S, long long RT, int HALO) {
  const int per = 2 * HALO;
  const int total = S * per;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int s = idx / per, k = idx - s * per;
    const long long R = k < HALO ? k : RT - 2 * HALO + k;
    *reinterpret_cast<uint4*>(sps + ((long long)s * RT + R) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

int zero_halo_launch(void* sps, int S, int n_patches, int P, cudaStream_t stream) {
  const int HALO = sps_halo(P);
  const int total = S * 2 * HALO;
  zero_halo_kernel<<<(total + 255) / 256, 256, 0, stream>>>((__nv_bfloat16*)sps, S, sps_rows(n_patches, P), HALO);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------
// Exact fp32 gather: out[b][c][i][j] = img[x0+i][y0+j][c].
__global__ void gather_f32_kernel(const float* __restrict__ img, int H, int W, int C, const int* __restrict__ xy, int n,
                                  int P, int center_mode, float* __restrict__ out, int pitch, int vec_in, int vec_out) {
  extern __shared__ float tile[];  // [P*P][pitch], pitch odd -> conflict-free transposed reads
  const int PP2 = P * P;
  const int rowlen = P * C;  // one patch row is contiguous in the raster
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    int x0 = xy[2 * b], y0 = xy[2 * b + 1];
    if (center_mode) { x0 -= P / 2; y0 -= P / 2; }
    __syncthreads();
    if (vec_in) {
      const int rl4 = rowlen / 4;
      for (int t = threadIdx.x; t < P * rl4; t += blockDim.x) {
        const int i = t / rl4, e = (t - i * rl4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(img + ((long long)(x0 + i) * W + y0) * C + e));
        const int j = e / C, c = e - j * C;  // C % 4 == 0: the four lanes stay inside one pixel
        float* d = tile + (i * P + j) * pitch + c;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      for (int t = threadIdx.x; t < P * rowlen; t += blockDim.x) {
        const int i = t / rowlen, e = t - i * rowlen;
        const int j = e / C, c = e - j * C;
        tile[(i * P + j) * pitch + c] = __ldg(img + ((long long)(x0 + i) * W + y0) * C + e);
      }
    }
    __syncthreads();
    float* o = out + (long long)b * C * PP2;
    const int total = C * PP2;
    if (vec_out) {
      for (int t = threadIdx.x * 4; t < total; t += blockDim.x * 4) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = (t + k) / PP2, pix = (t + k) - c * PP2;
          v[k] = tile[pix * pitch + c];
        }
        *reinterpret_cast<float4*>(o + t) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
      for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int c = t / PP2, pix = t - c * PP2;
        o[t] = tile[pix * pitch + c];
      }
    }
  }
}

int gather_f32_launch(const float* img, int H, int W, int C, const int* xy, int n, int P, int center_mode, float* out,
                      cudaStream_t stream) {
  if (n <= 0 || P < 1 || C < 1 || P > H || P > W) return VC_ERR_ARG;
  const int pitch = C | 1;
  const size_t smem = (size_t)P * P * pitch * sizeof(float);
  static int max_smem = 0;
  if (!max_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  if (smem > (size_t)max_smem) return VC_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(gather_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VC_ERR_CUDA;
  const int vec_in = (C % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0) ? 1 : 0;
  const int vec_out = (((long long)C * P * P) % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
  int blocks = n < 148 * 8 ? n : 148 * 8;
  gather_f32_kernel<<<blocks, 256, smem, stream>>>(img, H, W, C, xy, n, P, center_mode, out, pitch, vec_in, vec_out);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// Centre labels: labels[b] = int64(gt[x][y]) (datasets.py:573-581); gt is uint8 / int32 / int64.
__global__ void gather_labels_kernel(const void* gt, int eb, int H, int W, const int* xy, int n, int P, int center_mode,
                                     long long* labels) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    int x = xy[2 * b], y = xy[2 * b + 1];
    if (!center_mode) { x += P / 2; y += P / 2; }
    const long long e = (long long)x * W + y;
    long long v;
    if (eb == 1) v = reinterpret_cast<const unsigned char*>(gt)[e];
    else if (eb == 4) v = reinterpret_cast<const int*>(gt)[e];
    else v = reinterpret_cast<const long long*>(gt)[e];
    labels[b] = v;
  }
}

int gather_labels_launch(const void* gt, int eb, int H, int W, const int* xy, int n, int P, int center_mode,
                         long long* labels, cudaStream_t stream) {
  if (n <= 0 || (eb != 1 && eb != 4 && eb != 8)) return VC_ERR_ARG;
  gather_labels_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gt, eb, H, W, xy, n, P, center_mode, labels);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// Window enumeration for windows [first, first+count) of the scene, row-major over (xs, ys).
__global__ void scene_index_kernel(const int* xs, const int* ys, int nx, int ny, int first, int count, int W, int C1,
                                   int C2, int P, long long* off1, long long* off2, long long* out_idx, int* xy) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
    const int wi = first + t;
    const int x = xs[wi / ny], y = ys[wi % ny];
    const long long pix = (long long)x * W + y;
    if (off1) off1[t] = pix * C1;
    if (off2) off2[t] = pix * C2;
    if (out_idx) out_idx[t] = (long long)(x + P / 2) * W + (y + P / 2);
    if (xy) { xy[2 * t] = x; xy[2 * t + 1] = y; }
  }
}

int scene_index_launch(const int* xs, const int* ys, int nx, int ny, int first, int count, int W, int C1, int C2, int P,
                       int K, long long* off1, long long* off2, long long* out_idx, int* xy, cudaStream_t stream) {
  if (count <= 0 || first < 0 || (long long)first + count > (long long)nx * ny) return VC_ERR_ARG;
  (void)K;
  scene_index_kernel<<<(count + 255) / 256, 256, 0, stream>>>(xs, ys, nx, ny, first, count, W, C1, C2, P, off1, off2,
                                                             out_idx, xy);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// Element offsets of the top-left corner of the P x P patch centred on (x, y) in the two rasters
// (MultiModalX.__getitem__: x1 = x - P//2, y1 = y - P//2, datasets.py:551-556).
__global__ void center_offsets_kernel(const int* xy, int n, int W, int C1, int C2, int P, long long* off1, long long* off2) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const long long pix = (long long)(xy[2 * t] - P / 2) * W + (xy[2 * t + 1] - P / 2);
    off1[t] = pix * C1;
    off2[t] = pix * C2;
  }
}

int center_offsets_launch(const int* xy, int n, int W, int C1, int C2, int P, long long* off1, long long* off2,
                          cudaStream_t stream) {
  if (n <= 0) return VC_ERR_ARG;
  center_offsets_kernel<<<(n + 255) / 256, 256, 0, stream>>>(xy, n, W, C1, C2, P, off1, off2);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
