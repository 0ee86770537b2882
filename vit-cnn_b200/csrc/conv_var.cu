// Shared stem of dense scene inference, conv L >= 2: ALL border-class variants of one row class in one pass.
//
// The output of conv L of the stem at window pixel (i, j) is one of (2L+1)^2 variants, chosen by
// border_class(i) x border_class(j); variant (cy, cx) reads, for tap (dy, dx), the conv L-1 variant plane
// (ry(cy, dy), rx(cx, dx)) of the neighbour pixel -- the row class of the neighbour depends on (cy, dy) only, its
// column class on (cx, dx) only (abi.cu neighbour_class).  Launching conv_sps_tc_kernel once per output variant
// stages up to nine input slabs per tile for nine taps' worth of tensor work and is bound by L2 -> shared-memory
// traffic (profiles/r02_SUMMARY.md: 41 B per clock and SM).  Here a work unit is (tile, cy): the <= 3 x (2L-1)
// input slabs of one K step are staged ONCE (with that K step's weights) and feed all 2L+1 column classes, i.e.
// up to 9 (2L+1) tensor-core instructions into 2L+1 accumulators side by side in TMEM: 4-5x less staging per
// instruction, which makes these launches tensor bound.  Per accumulator the instructions arrive in the plain
// conv's order (K step, then tap 0..8, dropped taps skipped), so the results are bit-identical to the per-window
// conv (tests/test_gpu_scene.py::test_shared_stem_is_bit_identical_to_the_per_window_path).
//
// Accumulators: TMEM is used as a RING of slots, each with its own full / empty barrier pair.  Where two whole units fit the
// 512 columns (conv 3: 7 x 32 = 224 each) a slot is a unit: classic double buffering.  Conv 2 of the HSI stem needs 5 x 64 = 320
// columns per unit, so there a slot is ONE variant (8 slots): the epilogue drains a variant as soon as its last instruction has
// landed while the issuer is already filling the following slots with the next variants / the next unit.
//
// Geometry: scene blocks of B = 31 / 63 / 95 pixels as SPS "patches" (pitch PW = B + 1 rows, lead halo B + 9 rows, both
// multiples of 8: vc_common.cuh), so every slab a tap row needs starts at a multiple of 8 rows = 128 bytes: slab(dy) =
// rows [R0 + PW dy - 8, R0 + PW dy + 136) of the input plane, the tap (dy, dx) reads it at row offset 8 + dx.
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

namespace cv {
constexpr int kThreads = 192;       // warp 0 producer, warp 1 MMA issuer, warps 2..5 epilogue
constexpr int kSlabRows = 144;
constexpr uint32_t kSlabBytes = kSlabRows * 16u;
constexpr int kMaxNC = 7;
}  // namespace cv

struct ConvVarArgs {
  const __nv_bfloat16* in;    // [NP*NP planes][S_in][RT][8]: conv L-1 variants
  const __nv_bfloat16* w;     // [9 taps][S_in][n][8] (nsplit = 1 packing of conv_sps)
  const float* scale;         // [n]
  const float* bias;          // [n]
  __nv_bfloat16* out;         // [NC*NC planes][n/8][RT][8]: conv L variants
  long long RT, in_plane, out_plane;     // rows per slice; elements per input / output plane
  int S_in, n, NC, NP, B, n_blocks, ntiles, relu, nstages, nslots, vps;   // vps: variants per ring slot (1, or NC = a whole unit)
  signed char ry[cv::kMaxNC][3];         // input row class of tap row dy for output row class cy (-1: outside the window)
  signed char rx[cv::kMaxNC][3];         // input column class of tap column dx for output column class cx
};

__global__ void __launch_bounds__(cv::kThreads) conv_var_kernel(ConvVarArgs a) {
  using namespace cv;
  extern __shared__ __align__(128) uint8_t smem[];
  // issue table: for (row class, column class, tap) the 16-byte-unit offset of the tap's A operand inside a stage, or
  // 0xFFFFFFFF when the tap leaves the window -- the issuing thread must spend as few instructions per MMA as possible
  __shared__ uint32_t tap_tab[kMaxNC * kMaxNC * 9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int PW = a.B + 1, HALO = sps_halo(a.B), PP = sps_pp(a.B);
  const int KS = a.S_in / 2, NC = a.NC, NP = a.NP, n = a.n;
  for (int t = threadIdx.x; t < NC * NC * 9; t += blockDim.x) {
    const int tap = t % 9, cx = (t / 9) % NC, cy = t / (9 * NC);
    const int d = tap / 3, e = tap % 3, rxp = a.rx[cx][e];
    tap_tab[t] = (a.ry[cy][d] < 0 || rxp < 0) ? 0xFFFFFFFFu : (uint32_t)((d * NP + rxp) * 2) * (kSlabBytes >> 4) + (uint32_t)(8 + e - 1);
  }
  const uint32_t a_bytes = 3u * (uint32_t)NP * 2u * kSlabBytes;          // input slabs of one K step: [dy][rx][2 slices]
  const uint32_t w_bytes = 9u * 2u * (uint32_t)n * 16u;                  // weights of one K step: [tap][2 slices][n][8]
  const uint32_t stage_bytes = a_bytes + w_bytes;
  const int nunits = a.ntiles * NC;

  uint8_t* stage_s = smem;
  float* sc_s = reinterpret_cast<float*>(smem + (size_t)a.nstages * stage_bytes);
  float* bi_s = sc_s + n;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bi_s + n) + 7) & ~uintptr_t(7));
  uint64_t* full = bars;
  uint64_t* empty = bars + a.nstages;
  uint64_t* tfull = bars + 2 * a.nstages;     // [nslots <= 32]
  uint64_t* tempty = tfull + 32;              // [nslots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 32);

  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sc_s[i] = a.scale[i];
    bi_s[i] = a.bias[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < a.nslots; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 128);
    }
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
  const int nslots = a.nslots, vps = a.vps;               // accumulator ring: slot = (variants issued so far) / vps % nslots
  const int spu = NC / vps;                               // slots per unit (NC, or 1)

  if (warp == 0) {
    // ===== producer: per (unit, K step) the slabs of every input plane the unit's taps read + the K step's weights =====
    int st = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int tile = u / NC, cy = u - tile * NC;
      const long long row0 = (long long)HALO + (long long)tile * 128 - 8;      // slab(dy = 0) starts 8 rows before the tile
      int ndy = 0;
      for (int d = 0; d < 3; ++d) ndy += a.ry[cy][d] >= 0 ? 1 : 0;
      for (int ks = 0; ks < KS; ++ks) {
        if (lane == 0) {
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&full[st], (uint32_t)(ndy * NP * 2) * kSlabBytes + w_bytes);
        }
        __syncwarp();
        uint8_t* dst = stage_s + (size_t)st * stage_bytes;
        // copies are spread over the lanes: job = (dy, rx, slice) for the slabs, then the nine taps' weights
        const int nslab = 3 * NP * 2;
        for (int job = lane; job < nslab + 9; job += 32) {
          if (job < nslab) {
            const int d = job / (NP * 2), rem = job - d * NP * 2, rxp = rem >> 1, h2 = rem & 1;
            const int ry = a.ry[cy][d];
            if (ry < 0) continue;
            const __nv_bfloat16* src = a.in + (long long)(ry * NP + rxp) * a.in_plane +
                                       ((long long)(2 * ks + h2) * a.RT + row0 + (long long)(d - 1) * PW) * 8;
            bulk_g2s(dst + (size_t)job * kSlabBytes, src, kSlabBytes, &full[st]);
          } else {
            const int tap = job - nslab;
            bulk_g2s(dst + a_bytes + (size_t)tap * 2u * n * 16u, a.w + ((long long)tap * a.S_in + 2 * ks) * n * 8, 2u * (uint32_t)n * 16u,
                     &full[st]);
          }
        }
        if (++st == a.nstages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint64_t a_hi = umma_desc(0, kSlabBytes, 128) & 0xFFFFFFFF00000000ull;          // LBO = next slice's slab
    const uint64_t b_hi = umma_desc(0, (uint32_t)n * 16u, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lbo_lo = (uint32_t)(umma_desc(0, kSlabBytes, 128) & 0xFFFF0000u);
    const uint32_t b_lbo_lo = (uint32_t)(umma_desc(0, (uint32_t)n * 16u, 128) & 0xFFFF0000u);
    const uint32_t s_lo0 = (smem_u32(stage_s) & 0x3FFFFu) >> 4;
    int st = 0;
    uint32_t ph = 0;
    int slot0 = 0;                        // ring slot of this unit's column class 0 ...
    uint32_t lap0 = 0;                    // ... and the parity of its lap around the ring (no divisions on the issue path)
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int cy = u % NC;
      for (int ks = 0; ks < KS; ++ks) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_lo0 = a_lbo_lo | (s_lo0 + (uint32_t)st * (stage_bytes >> 4));
        const uint32_t b_lo0 = b_lbo_lo | (s_lo0 + (uint32_t)st * (stage_bytes >> 4) + (a_bytes >> 4));
        if (ks == 0) {                    // the slots' previous tenants have been drained by the epilogue
          int slot = slot0;
          uint32_t lap = lap0;
          for (int k = 0; k < spu; ++k) {
            mbar_wait(&tempty[slot], lap ^ 1u);
            if (++slot == nslots) { slot = 0; lap ^= 1u; }
          }
          tc_fence_after();
        }
        if (elect_one()) {                // one warp-uniform region issues the whole K step: instruction issue must not stall
          int slot = slot0, sub = 0;
          const uint32_t* tab = tap_tab + cy * NC * 9;
          uint32_t d_col = tmem_base + (uint32_t)(slot * vps * n);
          for (int cx = 0; cx < NC; ++cx) {
            uint32_t go = ks != 0 ? 1u : 0u;        // the first instruction of a variant overwrites its accumulator
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t off = tab[cx * 9 + tap];
              if (off == 0xFFFFFFFFu) continue;                   // the tap leaves the window: zero padding
              umma_bf16(d_col, a_hi | (uint64_t)(a_lo0 + off), b_hi | (uint64_t)(b_lo0 + (uint32_t)tap * (uint32_t)(2 * n)), idesc, go);
              go = 1u;
            }
            d_col += (uint32_t)n;
            if (++sub == vps) {           // last variant of the slot: it is complete after the last K step
              if (ks + 1 == KS) umma_commit(&tfull[slot]);
              sub = 0;
              if (++slot == nslots) { slot = 0; d_col = tmem_base; }
            }
          }
          umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == a.nstages) { st = 0; ph ^= 1u; }
      }
      slot0 += spu;
      if (slot0 >= nslots) { slot0 -= nslots; lap0 ^= 1u; }
    }
  } else {
    // ===== epilogue: TMEM -> affine (+ReLU) -> bf16 -> the 2L+1 output planes of this row class =====
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    int slot = 0, sub = 0;
    uint32_t lap = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int tile = u / NC, cy = u - tile * NC;
      const long long r = (long long)tile * 128 + row_in_tile;
      const long long R = r + HALO;
      const long long b = r / PP;
      const int q = (int)(r - b * PP);
      const int i = q / PW, j = q - i * PW;
      const bool valid = (b < a.n_blocks) && (i < a.B) && (j < a.B);
      for (int cx = 0; cx < NC; ++cx) {
        if (sub == 0) {
          mbar_wait(&tfull[slot], lap);
          tc_fence_after();
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((slot * vps + sub) * n);
        __nv_bfloat16* op = a.out + (long long)(cy * NC + cx) * a.out_plane;
        for (int c0 = 0; c0 < n; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tc_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float y0 = __uint_as_float(v[2 * k]) * sc_s[c0 + 2 * k] + bi_s[c0 + 2 * k];
            float y1 = __uint_as_float(v[2 * k + 1]) * sc_s[c0 + 2 * k + 1] + bi_s[c0 + 2 * k + 1];
            if (a.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
            if (!valid) { y0 = 0.f; y1 = 0.f; }
            pk[k] = pack_bf16(y0, y1);
          }
          const int slice = c0 / 8;
          *reinterpret_cast<uint4*>(op + ((long long)slice * a.RT + R) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(op + ((long long)(slice + 1) * a.RT + R) * 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        if (++sub == vps) {
          tc_fence_before();
          mbar_arrive(&tempty[slot]);
          sub = 0;
          if (++slot == nslots) { slot = 0; lap ^= 1u; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// in: [NP^2][S_in][RT][8] planes of conv L-1 (NP = 2L-1), out: [NC^2][n_out/8][RT][8] planes of conv L (NC = 2L+1),
// w: conv_sps packing with nsplit = 1; ry / rx: [NC][3] class tables (abi.cu neighbour_class).  Returns
// VC_ERR_UNSUPPORTED when the geometry is not the B = 31 block geometry or the shapes do not fit.
int conv_var_launch(const void* in, int S_in, const void* w, const float* scale, const float* bias, void* out, int n_out, int L,
                    const signed char* ry, const signed char* rx, int n_blocks, int B, int relu, cudaStream_t stream) {
  using namespace cv;
  const int NC = 2 * L + 1, NP = 2 * L - 1;
  if (L < 2 || NC > kMaxNC || S_in <= 0 || (S_in & 1) || n_out % 16 || n_out < 16 || n_out > 128 || n_blocks <= 0) return VC_ERR_ARG;
  if ((B + 1) % 8 != 0 || NC * n_out > 512) return VC_ERR_UNSUPPORTED;   // slab alignment: PW = B + 1 and HALO = B + 9 multiples of 8
  static int max_smem = 0, num_sms = 0;
  if (!max_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  ConvVarArgs a;
  a.in = (const __nv_bfloat16*)in;
  a.w = (const __nv_bfloat16*)w;
  a.scale = scale;
  a.bias = bias;
  a.out = (__nv_bfloat16*)out;
  a.RT = sps_rows(n_blocks, B);
  a.in_plane = (long long)S_in * a.RT * 8;
  a.out_plane = (long long)(n_out / 8) * a.RT * 8;
  a.S_in = S_in; a.n = n_out; a.NC = NC; a.NP = NP; a.B = B; a.n_blocks = n_blocks;
  a.ntiles = sps_tiles(n_blocks, B);
  a.relu = relu;
  // accumulator ring (see the header): two whole units when they fit TMEM (one barrier pair per unit: measured faster for
  // conv 3, 3.17 vs 3.72 ms at the Houston shape), else one slot per variant (conv 2 of the HSI stem: 3.71 vs 4.38 ms)
  a.vps = 2 * NC * n_out <= 512 ? NC : 1;
  a.nslots = 512 / (a.vps * n_out);
  if (a.nslots > 32) a.nslots = 32;
  for (int c = 0; c < kMaxNC; ++c)
    for (int d = 0; d < 3; ++d) {
      a.ry[c][d] = c < NC ? ry[c * 3 + d] : -1;
      a.rx[c][d] = c < NC ? rx[c * 3 + d] : -1;
    }
  const size_t stage = (size_t)3 * NP * 2 * kSlabBytes + (size_t)9 * 2 * n_out * 16;
  const size_t fixed = (size_t)n_out * 8 + 8 + (2 * 4 + 64) * 8 + 16 + 128;
  int nst = 4;
  while (nst > 1 && nst * stage + fixed > (size_t)max_smem) --nst;
  if (nst < 2) return VC_ERR_UNSUPPORTED;
  a.nstages = nst;
  const size_t smem = nst * stage + fixed;
  if (cudaFuncSetAttribute(conv_var_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return VC_ERR_CUDA;
  int grid = num_sms;
  if (grid > a.ntiles * NC) grid = a.ntiles * NC;
  conv_var_kernel<<<grid, kThreads, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
