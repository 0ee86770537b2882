// Exact fp32 patch gather on the TMA unit (tensor-map form): out[b][c][i][j] = img[x0 + i][y0 + j][c], bit-exact copies
// (MultiModalX.__getitem__ + default_collate, datasets.py:550-593; test()'s batch assembly, model_utils.py:1103-1112).
//
// The raster [H][W][C] fp32 is described by ONE 3-D tensor map (dims C, W, H; box 32 channels x P x P, SWIZZLE_128B),
// so a work unit = (patch, 32-channel group) is loaded by a single cp.async.bulk.tensor (SASS UTMALDG) into a
// [P*P pixels][32 channels] tile of 128-byte rows whose 16-byte chunks are XOR-swizzled by (row & 7); the threads only
// transpose it through shared memory (conflict-free 16-byte reads: a quarter warp covers the 8 chunks of a row; scalar
// writes with lanes 4 P^2 floats apart = 4 banks apart for odd P) into the [32][P*P] tile the output wants, and one
// 1-D bulk store (cp.async.bulk shared -> global, SASS UBLKCP) writes its valid channels as a contiguous run of the
// [C][P][P] patch.  Loads run kStages - 1 units ahead of the transposes, stores drain behind them: the SM's threads never
// wait on HBM in either direction.  Channel groups that reach past C are zero-filled by the TMA and not stored.
// Needs C % 4 == 0 and a 16-byte aligned raster / output (tensor-map strides are multiples of 16 bytes); the LiDAR raster
// (C2 = 1 or 2: rows of 4 / 8 bytes) and odd shapes stay on gather_f32_kernel (pack.cu).
#include <cuda.h>
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

namespace gt {
constexpr int kThreads = 256;
constexpr int kG = 32;                       // channels per unit = one 128-byte swizzled row
}  // namespace gt

struct GatherTmaArgs {
  const int* xy;
  const unsigned char* ops;
  float* out;
  int n, P, C, H, W, center_mode, ngroups;
  uint32_t a_bytes, b_bytes;                 // per-stage tile sizes (A padded to 1024 B)
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

template <int kStages>
__global__ void __launch_bounds__(gt::kThreads) gather_tma_kernel(const __grid_constant__ CUtensorMap tmap, GatherTmaArgs a) {
  using namespace gt;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  const int P = a.P, PP = P * P;
  uint8_t* A = smem;                                                   // [kStages][a_bytes]: [pixel][32 ch] swizzled rows
  uint8_t* B = smem + (size_t)kStages * a.a_bytes;                     // [kStages][b_bytes]: [32 ch][pixel]
  const long long nunits = (long long)a.n * a.ngroups;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](long long u, int s) {        // thread 0: one tensor-map load = the whole [P][P][32] box of unit u
    const int b = (int)(u / a.ngroups), g = (int)(u - (long long)b * a.ngroups);
    int x0 = a.xy[2 * b], y0 = a.xy[2 * b + 1];
    if (a.center_mode) { x0 -= P / 2; y0 -= P / 2; }
    x0 = min(max(x0, 0), a.H - P);                                     // memory safety (callers validate: ops.validate_xy)
    y0 = min(max(y0, 0), a.W - P);
    mbar_arrive_expect_tx(&full[s], (uint32_t)PP * 128u);
    tma_load_3d(A + (size_t)s * a.a_bytes, &tmap, g * kG, y0, x0, &full[s]);
  };

  const long long u0 = blockIdx.x, ustep = gridDim.x;
  if (threadIdx.x == 0)
    for (int k = 0; k < kStages - 1; ++k)
      if (u0 + k * ustep < nunits) issue(u0 + k * ustep, k);
  uint32_t phases = 0;          // bit s = parity the full barrier of stage s completes next
  int s = 0;
  for (long long u = u0, it = 0; u < nunits; u += ustep, ++it) {
    // stage (s + kStages - 1) % kStages was transposed one iteration ago (all threads are past that barrier): refill it
    if (threadIdx.x == 0) {
      const long long un = u + (long long)(kStages - 1) * ustep;
      if (un < nunits) issue(un, (s + kStages - 1) % kStages);
      bulk_wait_read<kStages - 1>();            // the bulk store that last read B[s] has finished reading shared memory
    }
    mbar_wait(&full[s], (phases >> s) & 1u);
    phases ^= 1u << s;
    __syncthreads();
    const int b = (int)(u / a.ngroups), g = (int)(u - (long long)b * a.ngroups);
    const int vc = min(kG, a.C - g * kG);                               // valid channels of this group (multiple of 4)
    const int op = a.ops ? a.ops[b] : 0;
    const uint8_t* At = A + (size_t)s * a.a_bytes;
    float* Bt = reinterpret_cast<float*>(B + (size_t)s * a.b_bytes);
    for (int idx = threadIdx.x; idx < PP * 8; idx += kThreads) {
      const int pix = idx >> 3, q = idx & 7;
      if (4 * q >= vc) continue;
      int r = pix;
      if (op) {                                                        // flip / rot90 augmentation = index remap of the source pixel
        int si, sj;
        dihedral_src(op, P, pix / P, pix % P, si, sj);
        r = si * P + sj;
      }
      const float4 v = *reinterpret_cast<const float4*>(At + (size_t)r * 128 + ((q ^ (r & 7)) << 4));
      float* d = Bt + (4 * q) * PP + pix;
      d[0] = v.x; d[PP] = v.y; d[2 * PP] = v.z; d[3 * PP] = v.w;
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_s2g(a.out + ((long long)b * a.C + (long long)g * kG) * PP, Bt, (uint32_t)vc * (uint32_t)PP * 4u);
      bulk_commit();
    }
    s = (s + 1) % kStages;
  }
  if (threadIdx.x == 0) bulk_wait_read<0>();    // shared memory must outlive the last stores' reads
  __syncthreads();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// VC_ERR_UNSUPPORTED: shape / alignment the tensor map cannot express -> the caller falls back to gather_f32_kernel
int gather_tma_launch(const float* img, int H, int W, int C, const int* xy, const unsigned char* ops, int n, int P, int center_mode,
                      float* out, cudaStream_t stream) {
  using namespace gt;
  if (n <= 0 || P < 1 || P > H || P > W) return VC_ERR_ARG;
  static const bool off = [] {
    const char* e = getenv("VITCNN_GATHER_TMA");
    return e && e[0] == '0';
  }();
  if (off || C < 4 || (C & 3) || P > 16 || (reinterpret_cast<uintptr_t>(img) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (long long)H * W * C >= (1LL << 40))
    return VC_ERR_UNSUPPORTED;
  static EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
  }();
  if (!encode) return VC_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t gstride[2] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kG, (cuuint32_t)P, (cuuint32_t)P};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(img), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return VC_ERR_UNSUPPORTED;
  GatherTmaArgs a;
  a.xy = xy; a.ops = ops; a.out = out;
  a.n = n; a.P = P; a.C = C; a.H = H; a.W = W; a.center_mode = center_mode;
  a.ngroups = (C + kG - 1) / kG;
  a.a_bytes = ((uint32_t)(P * P) * 128u + 1023u) & ~1023u;
  a.b_bytes = ((uint32_t)(P * P) * kG * 4u + 127u) & ~127u;
  static int max_smem = 0, num_sms = 0;
  if (!max_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // stages per CTA x CTAs per SM (VITCNN_GATHER_STAGES: 2 / 3 / 4 / 6).  Measured at the Houston shape, 65 536 dense windows
  // (profiles/r02_gather_f32.txt): 2 stages (3 CTAs per SM) 1.14 of the HBM copy peak on algorithmic bytes, 3 stages (2 CTAs)
  // 0.86, 4 / 6 stages (1 CTA) 0.48 -- resident CTAs hide each other's barriers, deeper pipelines in one CTA do not
  static const int want = [] {
    const char* e = getenv("VITCNN_GATHER_STAGES");
    return e ? atoi(e) : 2;
  }();
  const size_t per_stage = (size_t)a.a_bytes + a.b_bytes;
  int stages = want == 2 || want == 3 || want == 4 || want == 6 ? want : 2;
  while (stages > 2 && (size_t)stages * per_stage + 1024 > (size_t)max_smem) stages = stages == 6 ? 4 : stages - 1;
  const size_t smem = (size_t)stages * per_stage + 1024;
  if (smem > (size_t)max_smem) return VC_ERR_UNSUPPORTED;
  const long long nunits = (long long)n * a.ngroups;
  int occ = (int)((size_t)(max_smem + 1024) / (smem + 1024));      // + the 1 KB the runtime reserves per CTA
  if (occ < 1) occ = 1;
  if (occ > 4) occ = 4;
  long long blocks = (long long)num_sms * occ;
  if (blocks > nunits) blocks = nunits;
  auto go = [&](auto kernel) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return VC_ERR_CUDA;
    kernel<<<(int)blocks, kThreads, smem, stream>>>(tmap, a);
    return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
  };
  switch (stages) {
    case 2: return go(gather_tma_kernel<2>);
    case 3: return go(gather_tma_kernel<3>);
    case 4: return go(gather_tma_kernel<4>);
    default: return go(gather_tma_kernel<6>);
  }
}

}  // namespace vc
