// Weight-gradient GEMMs over the SPS activation layout on the 5th-gen tensor cores.
//
// dW[tap][m][n] = sum over rows r of  A[r + sa(tap)][m] * B[r + sb(tap)][n]
//
// i.e. autograd's conv2d weight gradient (torch: grad_weight of F.conv2d, stride 1, padding 1)
// for the reference's Conv2d(3x3, pad 1) idiom (S2ENet conv_bn_relu), and
// with taps == 1 the nn.Linear weight gradient of the token stage (vision_transformer.py:57-166).
// The reduction axis is the ROW axis, which in SPS ([slice][row][8 channels]) is the strided one:
// both operands are "MN-major" for tcgen05 (8 channels contiguous = 16 B, 8 consecutive rows =
// one 128-byte core matrix), so the very same buffers the forward pass streams are consumed
// without any transposition, and a 3x3 tap is again just a row shift of one operand.
//
// One CTA owns a group of taps (taps_per_cta * N fp32 columns of TMEM <= 512) and a strided set
// of 128-row tiles (split-K); its partial [tap][128][N] goes to a workspace and a second small
// kernel sums the partials in a fixed order (deterministic) and scatters into the torch layout.
#include <stdlib.h>
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

struct WgradArgs {
  const __nv_bfloat16* A;   // [SA][RT][8]   M side (SA*8 <= 128 channels; missing slices read as zero)
  const __nv_bfloat16* B;   // [SB][RT][8]   N side (N = SB*8, multiple of 16, <= 256)
  float* part;              // [gridDim.x][ntaps][128][N]
  long long RT;
  int SA, SB, ntaps, tpc, halo, pw, ntiles, shift_on_a, nstages, srows;
};

constexpr int kWgThreads = 192;    // warp0 producer, warp1 MMA issuer, warps 2..5 epilogue

// Shared-memory matrix descriptor, MN-major, no swizzle: canonical layout in 16-byte units
// ((1,n),(8,k)) : ((_,SBO),(1,LBO)) -- 8 consecutive K rows 16 B apart form a core matrix,
// LBO between 8-row K groups, SBO between 8-element MN groups (= slices).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return umma_desc(saddr, lbo_bytes, sbo_bytes);
}
__host__ __device__ inline uint32_t umma_idesc_bf16_mn(int M, int N) {
  return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

__global__ void __launch_bounds__(kWgThreads) wgrad_sps_tc_kernel(WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HALO = a.halo, PW = a.pw;
  const int N = a.SB * 8;
  const int kWgStageRows = a.srows;   // rows (= K) per pipeline stage: srows / 16 MMAs per tap
  const int rowsA = kWgStageRows + (a.shift_on_a ? 2 * HALO : 0);
  const int rowsB = kWgStageRows + (a.shift_on_a ? 0 : 2 * HALO);
  const uint32_t sliceA = (uint32_t)rowsA * 16u, sliceB = (uint32_t)rowsB * 16u;
  const uint32_t bytesA = 16u * sliceA;                 // all 16 M slices are addressable
  const uint32_t bytesB = (uint32_t)a.SB * sliceB;
  const uint32_t stage_bytes = bytesA + bytesB;
  const int tap0 = blockIdx.y * a.tpc;
  const int ntap = (a.ntaps - tap0) < a.tpc ? (a.ntaps - tap0) : a.tpc;

  uint8_t* stage_s = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + a.nstages;
  uint64_t* done = bars + 2 * a.nstages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  // M slices that are never loaded must read as zeros (their rows of D are discarded, but
  // they must not hold NaN patterns that a later reader of the workspace could trip over)
  if (a.SA < 16) {
    for (int st = 0; st < a.nstages; ++st) {
      uint4* z = reinterpret_cast<uint4*>(stage_s + (size_t)st * stage_bytes + (size_t)a.SA * sliceA);
      const int n16 = (16 - a.SA) * rowsA;
      for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nsteps = my_tiles * (128 / kWgStageRows);

  if (warp == 0) {
    // ===== producer: every lane issues its share of the per-slice bulk copies (a single issuing
    // thread was the bottleneck: up to 34 copies of 1-1.5 KB per stage) =====
    int st = 0;
    uint32_t ph = 0;
    const int ncopies = a.SA + a.SB;
    for (int s = 0; s < nsteps; ++s) {
      const int tile = blockIdx.x + (s / (128 / kWgStageRows)) * gridDim.x;
      // first row of this stage, as an index into the buffer WITH the lead halo
      const long long R = HALO + (long long)tile * 128 + (long long)(s % (128 / kWgStageRows)) * kWgStageRows;
      const long long RA = a.shift_on_a ? R - HALO : R, RB = a.shift_on_a ? R : R - HALO;
      if (lane == 0) {
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&full[st], (uint32_t)a.SA * sliceA + bytesB);
      }
      __syncwarp();
      uint8_t* dst = stage_s + (size_t)st * stage_bytes;
      for (int c = lane; c < ncopies; c += 32) {
        if (c < a.SA)
          bulk_g2s(dst + (size_t)c * sliceA, a.A + ((long long)c * a.RT + RA) * 8, sliceA, &full[st]);
        else
          bulk_g2s(dst + bytesA + (size_t)(c - a.SA) * sliceB, a.B + ((long long)(c - a.SA) * a.RT + RB) * 8, sliceB,
                   &full[st]);
      }
      if (++st == a.nstages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_mn(128, N);
    const uint64_t a_hi = umma_desc_mn(0, 128, sliceA) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = umma_desc_mn(0, 128, sliceB) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lbo = (uint32_t)(umma_desc_mn(0, 128, sliceA) & 0xFFFF0000u);
    const uint32_t b_lbo = (uint32_t)(umma_desc_mn(0, 128, sliceB) & 0xFFFF0000u);
    const uint32_t a_lo0 = a_lbo | (((smem_u32(stage_s) & 0x3FFFFu) >> 4) + (uint32_t)(a.shift_on_a ? HALO : 0));
    const uint32_t b_lo0 = b_lbo | ((((smem_u32(stage_s) + bytesA) & 0x3FFFFu) >> 4) + (uint32_t)(a.shift_on_a ? 0 : HALO));
    int st = 0;
    uint32_t ph = 0;
    for (int s = 0; s < nsteps; ++s) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t so = (uint32_t)st * (stage_bytes >> 4);
        const uint32_t acc = s != 0 ? 1u : 0u;
#pragma unroll 1
        for (int t = 0; t < ntap; ++t) {
          const int tap = tap0 + t;
          const int shift = (a.ntaps == 9) ? ((tap / 3 - 1) * PW + (tap % 3 - 1)) : 0;
          const uint32_t a_lo = a_lo0 + so + (uint32_t)(a.shift_on_a ? shift : 0);
          const uint32_t b_lo = b_lo0 + so + (uint32_t)(a.shift_on_a ? 0 : shift);
          const uint32_t d = tmem_base + (uint32_t)(t * N);
#pragma unroll 4
          for (int k = 0; k < kWgStageRows / 16; ++k)
            umma_bf16(d, a_hi | (uint64_t)(a_lo + 16u * k), b_hi | (uint64_t)(b_lo + 16u * k), idesc, k ? 1u : acc);
        }
        umma_commit(&empty[st]);
        if (s == nsteps - 1) umma_commit(done);
      }
      __syncwarp();
      if (++st == a.nstages) { st = 0; ph ^= 1u; }
    }
    if (nsteps > 0) {
      if (lane == 0) mbar_wait(done, 0);
      __syncwarp();
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 160;" ::: "memory");
  } else {
    // ===== epilogue (once): TMEM -> registers -> partial workspace =====
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    float* dst = a.part + ((long long)blockIdx.x * a.ntaps + tap0) * 128 * N + (long long)m * N;
    // block on a hardware barrier (no issue slots burnt while the main loop runs): the MMA warp
    // joins it once the last commit has landed
    asm volatile("bar.sync 1, 160;" ::: "memory");
    tc_fence_after();
    for (int t = 0; t < ntap; ++t) {
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        if (nsteps > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * N + c0), v);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = 0u;
        }
        float4* o = reinterpret_cast<float4*>(dst + (long long)t * 128 * N + c0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          o[k] = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                             __uint_as_float(v[4 * k + 3]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// out[m*sm + n*sn + tap*st] (+)= sum_p part[p][tap][m][n]   for m < M, n < Nr;
// column n == bias_col (if >= 0) goes to out_bias[m] instead (the "ones" slice of the B operand).
// A block owns 32 consecutive (tap, m, n) elements; its 8 warps sum interleaved subsets of the
// partials and are combined in a fixed order (deterministic, coalesced 128-byte reads).
__device__ __forceinline__ void wgrad_reduce_body(const float* __restrict__ part, int nparts, int ntaps, int N, int M, int Nr,
                                                  float* __restrict__ out, long long sm, long long sn, long long st, int bias_col,
                                                  float* __restrict__ out_bias, int accumulate, int block, int nblocks) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long total = (long long)ntaps * M * N;
  const long long pstride = (long long)ntaps * 128 * N;
  for (long long base = (long long)block * 32; base < total; base += (long long)nblocks * 32) {
    const long long idx = base + lane;
    int n = 0, m = 0, tap = 0;
    float s = 0.f;
    if (idx < total) {
      n = (int)(idx % N);
      m = (int)((idx / N) % M);
      tap = (int)(idx / ((long long)N * M));
      const float* p = part + ((long long)tap * 128 + m) * N + n;
      for (int k = w; k < nparts; k += 8) s += p[k * pstride];
    }
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && idx < total) {
      float t = red[0][lane];
#pragma unroll
      for (int k = 1; k < 8; ++k) t += red[k][lane];
      const bool is_bias = (n == bias_col);
      if (n < Nr || is_bias) {
        float* o = is_bias ? (out_bias ? out_bias + m : nullptr) : out + m * sm + n * sn + tap * st;
        if (o) *o = accumulate ? *o + t : t;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int ntaps, int N, int M,
                                                           int Nr, float* __restrict__ out, long long sm, long long sn,
                                                           long long st, int bias_col, float* __restrict__ out_bias,
                                                           int accumulate) {
  wgrad_reduce_body(part, nparts, ntaps, N, M, Nr, out, sm, sn, st, bias_col, out_bias, accumulate, blockIdx.x, gridDim.x);
}

// every deferred reduction of one backward pass in ONE launch: blockIdx.y = job (the training step at small per-GPU
// batches is bound by the number of launches, not by their work)
__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(WgradReduceTable t) {
  const WgradReduceJob& j = t.job[blockIdx.y];
  wgrad_reduce_body(j.part, j.nparts, j.ntaps, j.N, j.M, j.Nr, j.out, j.sm, j.sn, j.st, j.bias_col, j.out_bias, j.accumulate,
                    blockIdx.x, gridDim.x);
}

int wgrad_reduce_batched_launch(const WgradReduceTable* t, cudaStream_t stream) {
  if (!t || t->n <= 0) return VC_OK;
  wgrad_reduce_batched_kernel<<<dim3(148, t->n), 256, 0, stream>>>(*t);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

int wgrad_reduce_launch(const float* part, int nparts, int ntaps, int N, int M, int Nr, float* out, long long sm, long long sn,
                        long long st, int bias_col, float* out_bias, int accumulate, cudaStream_t stream) {
  const long long total = (long long)ntaps * M * N;
  int blocks = (int)((total + 31) / 32);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(part, nparts, ntaps, N, M, Nr, out, sm, sn, st, bias_col, out_bias, accumulate);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

static size_t wgrad_smem(int SA, int SB, int HALO, int shift_on_a, int nstages, int srows) {
  const size_t rowsA = srows + (shift_on_a ? 2 * HALO : 0), rowsB = srows + (shift_on_a ? 0 : 2 * HALO);
  (void)SA;
  return (size_t)nstages * (16 * rowsA * 16 + (size_t)SB * rowsB * 16) + (2 * nstages + 1) * 8 + 16;
}

int wgrad_grid_x(int SB, int ntaps) {
  const int N = SB * 8;
  int tpc = 512 / N;
  if (tpc > ntaps) tpc = ntaps;
  const int groups = (ntaps + tpc - 1) / tpc;
  int gx = 148 / groups;
  return gx < 1 ? 1 : gx;
}

size_t wgrad_workspace_bytes(int SB, int ntaps) {
  return (size_t)wgrad_grid_x(SB, ntaps) * ntaps * 128 * (SB * 8) * sizeof(float);
}

// rows mode (P == 0): plain [slice][ntiles*128][8] operands without halos (taps must be 1);
// n_patches then carries the number of 128-row tiles.
int wgrad_sps_launch(const void* A, int SA, const void* B, int SB, int n_patches, int P, int ntaps, int shift_on_a,
                     void* workspace, float* out, int M, int Nr, long long sm, long long sn, long long st,
                     int bias_col, float* out_bias, int accumulate, cudaStream_t stream, WgradReduceTable* defer) {
  if (SA < 1 || SA > 16 || SB < 2 || (SB & 1) || SB > 32 || (ntaps != 1 && ntaps != 9) || n_patches <= 0 || P < 0 ||
      M > SA * 8 || Nr > SB * 8 || !workspace || (P == 0 && ntaps != 1))
    return VC_ERR_ARG;
  WgradArgs a;
  a.A = (const __nv_bfloat16*)A;
  a.B = (const __nv_bfloat16*)B;
  a.part = (float*)workspace;
  a.RT = P ? sps_rows(n_patches, P) : (long long)n_patches * 128;
  a.SA = SA;
  a.SB = SB;
  a.ntaps = ntaps;
  a.halo = P ? sps_halo(P) : 0;
  a.pw = P + 1;
  a.ntiles = P ? sps_tiles(n_patches, P) : n_patches;
  a.shift_on_a = shift_on_a;
  const int N = SB * 8;
  int tpc = 512 / N;
  if (tpc > ntaps) tpc = ntaps;
  const int groups = (ntaps + tpc - 1) / tpc;
  tpc = (ntaps + groups - 1) / groups;   // balance the groups
  a.tpc = tpc;
  static int max_smem = 0;
  if (!max_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  // stage depth: 128-row stages (fewer barrier round trips, 2 KB+ bulk copies; measured 319 -> 218 us
  // on the conv-1 shape even with only 2 stages in flight) when at least 2 of them fit, else 64 rows.  VC_WGRAD_ROWS / VC_WGRAD_STAGES override for tuning runs.
  static int env_rows = -1, env_stages = -1;
  if (env_rows < 0) {
    const char* e = getenv("VC_WGRAD_ROWS");
    env_rows = e ? atoi(e) : 0;
    e = getenv("VC_WGRAD_STAGES");
    env_stages = e ? atoi(e) : 0;
  }
  int srows = 128;
  if (wgrad_smem(SA, SB, a.halo, shift_on_a, 2, 128) > (size_t)max_smem) srows = 64;
  if (env_rows == 64 || env_rows == 128) srows = env_rows;
  a.srows = srows;
  int nst = env_stages > 0 ? env_stages : 6;
  while (nst > 2 && wgrad_smem(SA, SB, a.halo, shift_on_a, nst, srows) > (size_t)max_smem) --nst;
  const size_t smem = wgrad_smem(SA, SB, a.halo, shift_on_a, nst, srows);
  if (smem > (size_t)max_smem) return VC_ERR_UNSUPPORTED;
  a.nstages = nst;
  if (cudaFuncSetAttribute(wgrad_sps_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VC_ERR_CUDA;
  int gx = wgrad_grid_x(SB, ntaps);
  dim3 grid(gx, groups);
  wgrad_sps_tc_kernel<<<grid, kWgThreads, smem, stream>>>(a);
  if (cudaGetLastError() != cudaSuccess) return VC_ERR_CUDA;
  if (defer && defer->n < kMaxWgradJobs) {      // the caller reduces every job of the pass in one launch (own workspace per job)
    WgradReduceJob& j = defer->job[defer->n++];
    j.part = a.part; j.out = out; j.out_bias = out_bias; j.sm = sm; j.sn = sn; j.st = st;
    j.nparts = gx; j.ntaps = ntaps; j.N = N; j.M = M; j.Nr = Nr; j.bias_col = bias_col; j.accumulate = accumulate;
    return VC_OK;
  }
  const long long total = (long long)ntaps * M * N;
  int blocks = (int)((total + 31) / 32);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(a.part, gx, ntaps, N, M, Nr, out, sm, sn, st, bias_col, out_bias,
                                                  accumulate);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
