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
                                                       const unsigned char* __restrict__ ops, int n_patches, int C, int P,
                                                       __nv_bfloat16* __restrict__ sps, int S, long long RT, int vec) {
  const int HALO = sps_halo(P), PP = sps_pp(P), PW = P + 1;
  for (long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x; R < RT; R += (long long)gridDim.x * blockDim.x) {
    const long long r = R - HALO;
    const float* p = nullptr;
    if (r >= 0) {
      const long long b = r / PP;
      const int q = (int)(r - b * PP);
      const int i = q / PW, j = q - i * PW;
      if (b < n_patches && i < P && j < P) {
        int ii = i, jj = j;
        if (ops) dihedral_src(ops[b], P, i, j, ii, jj);     // flip / rot90 augmentation = index remap
        p = src + (patch_off ? patch_off[b] : b * sb) + ii * si + jj * sj;
      }
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
                    const unsigned char* ops, int n_patches, int C, int P, void* sps, int S, cudaStream_t stream) {
  if (n_patches <= 0 || C <= 0 || S * 8 < C || P < 1) return VC_ERR_ARG;
  const long long RT = sps_rows(n_patches, P);
  const int vec = (sc == 1 && si % 4 == 0 && sj % 4 == 0 && sb % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0)
                      ? 1
                      : 0;  // patch_off entries are multiples of C in raster mode: 16-byte aligned iff C % 4 == 0
  long long blocks = (RT + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  pack_sps_kernel<<<(int)blocks, 256, 0, stream>>>(src, sb, sc, si, sj, patch_off, ops, n_patches, C, P,
                                                   (__nv_bfloat16*)sps, S, RT, vec && (C % 4 == 0 || !patch_off));
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------
// Scene gather for stride-1 sliding windows (test(), model_utils.py:1086-1112): consecutive
// windows of one window row overlap in P-1 of their P columns, so a CTA stages the raster strip
// under `ws` consecutive windows (P rows x (ws+P-1) pixels x a part of the channels, fp32) in
// shared memory ONCE -- one bulk async copy (TMA unit) per pixel, padded pixel pitch so the
// 32-byte reads below are bank-conflict free -- and emits all ws patches from it.  Each raster
// pixel is then fetched from L2 ~(ws+P-1)/ws times per window row instead of P times, which
// leaves the kernel bound by its bf16 SPS writes.
struct StripArgs {
  const float* img;          // [H][W][C]
  const int* xs;             // window-row starts [nx]
  const int* ys;             // window-column starts [ny]
  __nv_bfloat16* sps;        // [S][RT][8]
  long long RT;
  int W, C, P, S, ny, first, count, ws, strips_per_row, gs_first, nparts, part_slices;
};

__global__ void __launch_bounds__(256) pack_strip_kernel(StripArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ int s_fallback;
  const int P = a.P, PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P);
  const int part = blockIdx.y;
  const int sl0 = part * a.part_slices;
  const int nsl = (a.S - sl0) < a.part_slices ? (a.S - sl0) : a.part_slices;     // slices of this part
  const int c0 = sl0 * 8;
  const int cn = (a.C - c0) < nsl * 8 ? (a.C - c0) : nsl * 8;                    // real channels of this part (may be <= 0)
  const int pitch = nsl * 32 + 16;                                              // bytes per staged pixel
  const int SW = a.ws + P - 1;                                                   // staged pixels per row
  const bool aligned = (a.C % 4 == 0) && cn == nsl * 8 && ((reinterpret_cast<uintptr_t>(a.img) & 15) == 0);
  short* rowtab = reinterpret_cast<short*>(smem);                                // [PP] pixel index i*SW+j or -1
  uint8_t* strip = smem + ((PP * 2 + 127) & ~127);
  for (int r = threadIdx.x; r < PP; r += blockDim.x) {
    const int i = r / PW, j = r - i * PW;
    rowtab[r] = (i < P && j < P) ? (short)(i * SW + j) : (short)-1;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  const int nstrips = gridDim.x;   // one strip per block in x; loop kept for generality
  for (int gs = a.gs_first + blockIdx.x; gs < a.gs_first + nstrips; gs += gridDim.x) {
    const int xr = gs / a.strips_per_row, m = gs - xr * a.strips_per_row;
    int w0 = xr * a.ny + m * a.ws, w1 = w0 + a.ws;                               // window index range of the strip
    const int row_end = (xr + 1) * a.ny;
    if (w1 > row_end) w1 = row_end;
    if (w0 < a.first) w0 = a.first;
    if (w1 > a.first + a.count) w1 = a.first + a.count;
    const int nw = w1 - w0;
    if (nw <= 0) continue;
    const int x = a.xs[xr], yi0 = w0 - xr * a.ny, y0 = a.ys[yi0];
    if (threadIdx.x == 0) s_fallback = (a.ys[yi0 + nw - 1] - y0 != nw - 1) ? 1 : 0;   // stride > 1: not contiguous
    __syncthreads();
    const bool fallback = s_fallback != 0;
    const int npx = nw + P - 1;
    if (!fallback && cn > 0) {
      if (aligned) {
        if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, (uint32_t)(P * npx) * (uint32_t)(cn * 4));
        __syncthreads();
        for (int t = threadIdx.x; t < P * npx; t += blockDim.x) {
          const int i = t / npx, j = t - i * npx;
          bulk_g2s(strip + (size_t)(i * SW + j) * pitch, a.img + ((long long)(x + i) * a.W + y0 + j) * a.C + c0,
                   (uint32_t)(cn * 4), &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
      } else {
        for (int t = threadIdx.x; t < P * npx * nsl * 8; t += blockDim.x) {
          const int c = t % (nsl * 8), pix = t / (nsl * 8);
          const int i = pix / npx, j = pix - i * npx;
          float v = 0.f;
          if (c < cn) v = __ldg(a.img + ((long long)(x + i) * a.W + y0 + j) * a.C + c0 + c);
          *reinterpret_cast<float*>(strip + (size_t)(i * SW + j) * pitch + c * 4) = v;
        }
        __syncthreads();
      }
    }
    // emit: one warp per (slice, window) pair, lanes walk the patch rows -> 512-byte contiguous
    // stores, no per-element index division
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int pair = threadIdx.x >> 5; pair < nsl * nw; pair += nwarps) {
      const int s = pair / nw, w = pair - s * nw;
      uint4* dst = reinterpret_cast<uint4*>(a.sps + ((long long)(sl0 + s) * a.RT + HALO + (long long)(w0 - a.first + w) * PP) * 8);
      const uint8_t* sbase = strip + (size_t)w * pitch + s * 32;
      for (int r = lane; r < PP; r += 32) {
        const int pix = rowtab[r];
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (pix >= 0 && cn > 0) {
          float4 lo, hi;
          if (!fallback) {
            const uint8_t* src = sbase + (size_t)pix * pitch;
            lo = *reinterpret_cast<const float4*>(src);
            hi = *reinterpret_cast<const float4*>(src + 16);
          } else {
            const int i = pix / SW, j = pix - i * SW;
            const float* src = a.img + ((long long)(x + i) * a.W + a.ys[yi0 + w] + j) * a.C + c0 + s * 8;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (s * 8 + k < cn) ? __ldg(src + k) : 0.f;
            lo = make_float4(v[0], v[1], v[2], v[3]);
            hi = make_float4(v[4], v[5], v[6], v[7]);
          }
          o = make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        }
        dst[r] = o;
      }
    }
    __syncthreads();   // the strip is overwritten by the next iteration
  }
}

// rows outside the patches (lead halo, tile padding, trailing halo) of every slice
__global__ void pack_strip_tail_kernel(__nv_bfloat16* sps, int S, long long RT, long long first_pad, int HALO) {
  const long long per = HALO + (RT - first_pad);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < per * S; idx += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(idx / per);
    const long long k = idx - (long long)s * per;
    const long long R = k < HALO ? k : first_pad + (k - HALO);
    *reinterpret_cast<uint4*>(sps + ((long long)s * RT + R) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

int pack_scene_launch(const float* img, int W, int C, const int* xs, const int* ys, int nx, int ny, int first, int count,
                      int P, void* sps, int S, cudaStream_t stream) {
  if (count <= 0 || first < 0 || (long long)first + count > (long long)nx * ny || S * 8 < C || P < 1) return VC_ERR_ARG;
  static int max_smem = 0;
  if (!max_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  StripArgs a;
  a.img = img; a.xs = xs; a.ys = ys; a.sps = (__nv_bfloat16*)sps;
  a.RT = sps_rows(count, P);
  a.W = W; a.C = C; a.P = P; a.S = S; a.ny = ny; a.first = first; a.count = count;
  // two CTAs per SM: split the channel slices into parts of <= 9 slices, 16 windows per strip
  a.part_slices = S <= 9 ? S : (S + 1) / 2 <= 9 ? (S + 1) / 2 : 9;
  a.nparts = (S + a.part_slices - 1) / a.part_slices;
  const int budget = 100 * 1024;
  const int pitch = a.part_slices * 32 + 16;
  int ws = budget / (P * pitch) - (P - 1);
  if (ws > 16) ws = 16;
  if (ws < 2) return VC_ERR_UNSUPPORTED;
  a.ws = ws;
  a.strips_per_row = (ny + ws - 1) / ws;
  const int xr0 = first / ny, xr1 = (first + count - 1) / ny;
  a.gs_first = xr0 * a.strips_per_row + (first - xr0 * ny) / ws;
  const int gs_last = xr1 * a.strips_per_row + (first + count - 1 - xr1 * ny) / ws;
  const size_t smem = ((sps_pp(P) * 2 + 127) & ~127) + (size_t)P * (ws + P - 1) * pitch;
  if (smem > (size_t)max_smem) return VC_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(pack_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VC_ERR_CUDA;
  dim3 grid(gs_last - a.gs_first + 1, a.nparts);
  pack_strip_kernel<<<grid, 256, smem, stream>>>(a);
  const int HALO = sps_halo(P);
  const long long first_pad = HALO + (long long)count * sps_pp(P);
  pack_strip_tail_kernel<<<8, 256, 0, stream>>>((__nv_bfloat16*)sps, S, a.RT, first_pad, HALO);
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
// Shared stem for dense sliding windows.  The output of a stack of d 3x3 pad-1 convs at pixel (i, j) of a
// window depends on the window only through the zero padding at the window's border: the value at scene
// pixel (y, x) is one of (2d+1)^2 variants, chosen by border_class(i) x border_class(j).  The variants are
// computed ONCE per scene on overlapping B x B blocks (abi.cu: conv 1 with the taps that leave the window
// dropped; deeper convs read, per tap, the variant plane of the neighbour's own class) and every window's
// stem output is a gather from them -- e.g. 121 / 9 times fewer conv-1 FLOPs at P = 11 -- with the same
// bits as the per-window convs (same accumulation order, dropped taps contribute exact zeros).
__global__ void block_offsets_kernel(int H, int W, int C, int B, int D, int nby, int nbx, long long* __restrict__ off) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nby * nbx) return;
  const int by = idx / nbx, bx = idx - by * nbx;
  off[idx] = ((long long)blk_origin(by, H, B, D) * W + blk_origin(bx, W, B, D)) * C;
}

int block_offsets_launch(int H, int W, int C, int B, int D, long long* off, cudaStream_t stream) {
  if (H < B || W < B || B <= 2 * D) return VC_ERR_ARG;
  const int nby = blk_count(H, B, D), nbx = blk_count(W, B, D);
  block_offsets_kernel<<<(nby * nbx + 255) / 256, 256, 0, stream>>>(H, W, C, B, D, nby, nbx, off);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// rowterm[y] = (block row of y) * blocks per row * rows per block + (y - block origin) * (B + 1); colterm[x] likewise for columns:
// the plane row of scene pixel (y, x) is sps_halo(B) + rowterm[y] + colterm[x] (what border_gather_kernel computes inline)
__global__ void scene_tables_kernel(int H, int W, int B, int D, int nbx, int* __restrict__ rowterm, int* __restrict__ colterm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int PPB = sps_pp(B);
  if (t < H) {
    const int ky = blk_index(t, H, B, D);
    rowterm[t] = ky * nbx * PPB + (t - blk_origin(ky, H, B, D)) * (B + 1);
  }
  if (t < W) {
    const int kx = blk_index(t, W, B, D);
    colterm[t] = kx * PPB + (t - blk_origin(kx, W, B, D));
  }
}

int scene_tables_launch(int H, int W, int B, int D, int* rowterm, int* colterm, cudaStream_t stream) {
  if (H < B || W < B || B <= 2 * D || !rowterm || !colterm) return VC_ERR_ARG;
  const int nbx = blk_count(W, B, D);
  if ((long long)blk_count(H, B, D) * nbx * sps_pp(B) >= (1LL << 31)) return VC_ERR_UNSUPPORTED;
  const int n = H > W ? H : W;
  scene_tables_kernel<<<(n + 255) / 256, 256, 0, stream>>>(H, W, B, D, nbx, rowterm, colterm);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

struct BorderGatherArgs {
  const __nv_bfloat16* v;     // [(2D+1)^2 variants][S][RTb][8]: depth-D stem variants over the scene blocks
  __nv_bfloat16* out;         // [>= S][RTo][8]: stem output of the chunk's windows (slices 0..S-1 are written)
  const int* xs;              // window-row starts [nx]
  const int* ys;              // window-column starts [ny]
  long long RTb, RTo;
  int S, ny, first, count, P, H, W, B, D, nbx, rows;
};

template <int S>    // slices per row: 16 / 8 / 4 for sharing depth 1 / 2 / 3 -- all S loads of a row are in flight together
__global__ void __launch_bounds__(256) border_gather_kernel(BorderGatherArgs a) {
  const int P = a.P, PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P), B = a.B, D = a.D;
  const int HB = sps_halo(B), PPB = sps_pp(B), NC = 2 * D + 1;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < a.rows; r += gridDim.x * blockDim.x) {
    const int b = r / PP, q = r - b * PP;
    const int i = q / PW, j = q - i * PW;
    const uint4* src = nullptr;
    if (b < a.count && i < P && j < P) {
      const int widx = a.first + b;
      const int ix = widx / a.ny, iy = widx - ix * a.ny;
      const int y = __ldg(a.xs + ix) + i, x = __ldg(a.ys + iy) + j;
      const int ky = blk_index(y, a.H, B, D), kx = blk_index(x, a.W, B, D);
      const long long srow = HB + (long long)(ky * a.nbx + kx) * PPB + (y - blk_origin(ky, a.H, B, D)) * (B + 1) +
                             (x - blk_origin(kx, a.W, B, D));
      src = reinterpret_cast<const uint4*>(a.v) + (long long)(border_class(i, P, D) * NC + border_class(j, P, D)) * S * a.RTb + srow;
    }
    uint4* dst = reinterpret_cast<uint4*>(a.out) + HALO + r;
    uint4 v[S];
#pragma unroll
    for (int s = 0; s < S; ++s) v[s] = src ? __ldg(src + (long long)s * a.RTb) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int s = 0; s < S; ++s) dst[(long long)s * a.RTo] = v[s];
  }
}

int border_gather_launch(const void* variants, int S, int B, int D, int H, int W, const int* xs, const int* ys, int ny, int first,
                         int count, int P, void* out, cudaStream_t stream) {
  if (count <= 0 || H < B || W < B || B <= 2 * D || (D == 1 ? P < 2 : P < 2 * D + 1)) return VC_ERR_ARG;
  const long long rows = (long long)sps_tiles(count, P) * 128;
  if (rows >= (1LL << 31) || (long long)first + count >= (1LL << 31)) return VC_ERR_ARG;
  BorderGatherArgs a;
  a.v = (const __nv_bfloat16*)variants;
  a.out = (__nv_bfloat16*)out;
  a.xs = xs; a.ys = ys;
  a.RTb = sps_rows(blk_count(H, B, D) * blk_count(W, B, D), B);
  a.RTo = sps_rows(count, P);
  a.S = S; a.ny = ny; a.first = first; a.count = count; a.P = P; a.H = H; a.W = W; a.B = B; a.D = D;
  a.nbx = blk_count(W, B, D);
  a.rows = (int)rows;
  long long blocks = (rows + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  if (S == 16) border_gather_kernel<16><<<(int)blocks, 256, 0, stream>>>(a);
  else if (S == 8) border_gather_kernel<8><<<(int)blocks, 256, 0, stream>>>(a);
  else if (S == 4) border_gather_kernel<4><<<(int)blocks, 256, 0, stream>>>(a);
  else return VC_ERR_ARG;
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------
// Exact fp32 gather: out[b][c][i][j] = img[x0+i][y0+j][c].
__global__ void gather_f32_kernel(const float* __restrict__ img, int H, int W, int C, const int* __restrict__ xy,
                                  const unsigned char* __restrict__ ops, int n, int P, int center_mode,
                                  float* __restrict__ out, int pitch, int vec_in, int vec_out) {
  extern __shared__ float tile[];  // [P*P][pitch], pitch odd -> conflict-free transposed reads
  const int PP2 = P * P;
  const int rowlen = P * C;  // one patch row is contiguous in the raster
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    int x0 = xy[2 * b], y0 = xy[2 * b + 1];
    if (center_mode) { x0 -= P / 2; y0 -= P / 2; }
    // memory safety only: callers validate the coordinates (ops.validate_xy); a window that leaves the raster
    // is moved inside instead of reading out of bounds
    x0 = min(max(x0, 0), H - P);
    y0 = min(max(y0, 0), W - P);
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
    const int op = ops ? ops[b] : 0;
    if (op) {           // augmented sample: output pixel (i, j) reads its source pixel from the staged patch
      for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int c = t / PP2, pix = t - c * PP2;
        int si_, sj_;
        dihedral_src(op, P, pix / P, pix % P, si_, sj_);
        o[t] = tile[(si_ * P + sj_) * pitch + c];
      }
    } else if (vec_out) {
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

int gather_f32_launch(const float* img, int H, int W, int C, const int* xy, const unsigned char* ops, int n, int P,
                      int center_mode, float* out, cudaStream_t stream) {
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
  gather_f32_kernel<<<blocks, 256, smem, stream>>>(img, H, W, C, xy, ops, n, P, center_mode, out, pitch, vec_in, vec_out);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// Centre labels: labels[b] = int64(gt[x][y]) (datasets.py:573-581); gt is uint8 / int32 / int64.
__global__ void gather_labels_kernel(const void* gt, int eb, int H, int W, const int* xy, const unsigned char* ops, int n, int P,
                                     int center_mode, long long* labels) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    int x = xy[2 * b], y = xy[2 * b + 1];
    if (!center_mode) { x += P / 2; y += P / 2; }
    if (ops && ops[b]) {      // label = transformed label window at [P/2, P/2] (datasets.py:557-581; matters for even P)
      int si_, sj_;
      dihedral_src(ops[b], P, P / 2, P / 2, si_, sj_);
      x += si_ - P / 2;
      y += sj_ - P / 2;
    }
    x = min(max(x, 0), H - 1);      // memory safety (see gather_f32_kernel)
    y = min(max(y, 0), W - 1);
    const long long e = (long long)x * W + y;
    long long v;
    if (eb == 1) v = reinterpret_cast<const unsigned char*>(gt)[e];
    else if (eb == 4) v = reinterpret_cast<const int*>(gt)[e];
    else v = reinterpret_cast<const long long*>(gt)[e];
    labels[b] = v;
  }
}

int gather_labels_launch(const void* gt, int eb, int H, int W, const int* xy, const unsigned char* ops, int n, int P,
                         int center_mode, long long* labels, cudaStream_t stream) {
  if (n <= 0 || (eb != 1 && eb != 4 && eb != 8)) return VC_ERR_ARG;
  gather_labels_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gt, eb, H, W, xy, ops, n, P, center_mode, labels);
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
__global__ void center_offsets_kernel(const int* xy, int n, int H, int W, int C1, int C2, int P, long long* off1, long long* off2) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int x0 = min(max(xy[2 * t] - P / 2, 0), H - P), y0 = min(max(xy[2 * t + 1] - P / 2, 0), W - P);   // memory safety
    const long long pix = (long long)x0 * W + y0;
    off1[t] = pix * C1;
    off2[t] = pix * C2;
  }
}

int center_offsets_launch(const int* xy, int n, int H, int W, int C1, int C2, int P, long long* off1, long long* off2,
                          cudaStream_t stream) {
  if (n <= 0 || P > H || P > W) return VC_ERR_ARG;
  center_offsets_kernel<<<(n + 255) / 256, 256, 0, stream>>>(xy, n, H, W, C1, C2, P, off1, off2);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
