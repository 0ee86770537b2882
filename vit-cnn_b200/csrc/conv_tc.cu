// 3x3 'same' (or 1x1) convolution over the SPS activation layout as shifted GEMMs on the
// 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by the TMA unit
// with 1-D bulk copies).  Replaces the reference idiom Conv2d(k=3,pad=1) -> BatchNorm2d ->
// ReLU (S2ENet conv_bn_relu; model/Multimodality_Mamba/Mutimodality_Mamba7.py:1035-1048,
// 1152-1153) for inference (BN folded into a per-channel affine).
//
// GEMM view: M = stacked patch rows (128 per tile, tiles ignore patch boundaries),
// N = output channels handled by this CTA (16..128), K = taps x input channels.
// A tile for tap (dy,dx) = rows [R0+shift, R0+shift+128) of the input, a contiguous slab
// per 8-channel slice, so one resident stage [2 slices][128+2*halo rows][16 B] serves all
// nine taps by moving the descriptor start address.  Weights stay resident in shared memory
// for the life of the (persistent) CTA.
#include <stdlib.h>
#include "vc_common.cuh"
#include "vc_kernels.h"

namespace vc {

struct ConvArgs {
  const __nv_bfloat16* in;   // [S_in][RT][8]
  const __nv_bfloat16* w;    // [nsplit][ntaps][S_in][ncta][8]
  const float* scale;        // [nsplit*ncta]
  const float* bias;         // [nsplit*ncta]
  __nv_bfloat16* out;        // [S_out_total][RT][8]
  long long RT;
  int S_in, ncta, ntaps, P, n_patches, ntiles, out_slice_off, relu, nstages, kpb, debug_flags;
  // multi-plane input (shared stem of dense scene inference, abi.cu): tap t reads plane p iff bit t of
  // tapmask[p]; taps in no mask are dropped (they would leave the window).  Plain conv: one plane, all taps.
  const __nv_bfloat16* planes[9];
  unsigned int tapmask[9];
  int nplanes;
};

constexpr int kConvThreads = 192;  // warp0 producer, warp1 MMA issuer, warps 2..5 epilogue

__global__ void __launch_bounds__(kConvThreads) conv_sps_tc_kernel(ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HALO = sps_halo(a.P), ROWS = 128 + 2 * HALO, PP = sps_pp(a.P), PW = a.P + 1;
  const int KPB = a.kpb;  // K=16 steps per pipeline stage
  const int NPL = a.nplanes;  // input planes resident per K step (1 for a plain conv)
  const uint32_t slice_bytes = (uint32_t)ROWS * 16u, kstep_bytes = 2u * slice_bytes, stage_bytes = (uint32_t)(KPB * NPL) * kstep_bytes;
  const uint32_t wbytes = (uint32_t)a.ntaps * a.S_in * a.ncta * 16u;
  const int KS = a.S_in / 2;  // K=16 steps per tap
  const int half = blockIdx.y;

  uint8_t* w_s = smem;
  uint8_t* stage_s = smem + ((wbytes + 127u) & ~127u);
  uint8_t* tail = stage_s + (size_t)a.nstages * stage_bytes;
  float* sc_s = reinterpret_cast<float*>(tail);
  float* bi_s = sc_s + a.ncta;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bi_s + a.ncta) + 7) & ~uintptr_t(7));
  uint64_t* full = bars;
  uint64_t* empty = bars + a.nstages;
  uint64_t* wfull = bars + 2 * a.nstages;
  uint64_t* tfull = wfull + 1;   // [2]
  uint64_t* tempty = wfull + 3;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 5);

  const uint32_t tmem_cols = (2 * a.ncta <= 32) ? 32u : (2 * a.ncta <= 64) ? 64u : (2 * a.ncta <= 128) ? 128u : 256u;

  for (int i = threadIdx.x; i < a.ncta; i += blockDim.x) {
    sc_s[i] = a.scale[half * a.ncta + i];
    bi_s[i] = a.bias[half * a.ncta + i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(wfull, 1);
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 128);
    mbar_init(&tempty[1], 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long dbg_c0 = 0, dbg_t0 = 0;
  if ((a.debug_flags & 64) && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(dbg_t0));
  }

  if (warp == 0) {
    // ===== producer: weights once, then one stage (2 channel slices) per K=16 step =====
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) + (size_t)half * wbytes;
      mbar_arrive_expect_tx(wfull, wbytes);
      for (uint32_t off = 0; off < wbytes; off += 32768u) {
        uint32_t n = wbytes - off < 32768u ? wbytes - off : 32768u;
        bulk_g2s(w_s + off, wsrc + off, n, wfull);
      }
      // rows of a plane's slab that its taps read: [HALO + lo, HALO + hi + 128) with lo / hi the smallest / largest
      // tap shift -- a plane read by one tap only (corner variants of the shared stem: up to nine such planes per
      // tile) is staged as 128 rows instead of 128 + 2 HALO, which is what bounds those launches (L2 -> smem)
      uint32_t pl_off[9], pl_bytes[9], stage_tx = 0;
      for (int pl = 0; pl < NPL; ++pl) {
        int lo = 1 << 20, hi = -(1 << 20);
        for (int tap = 0; tap < a.ntaps; ++tap)
          if (a.tapmask[pl] & (1u << tap)) {
            const int shift = a.ntaps == 9 ? (tap / 3 - 1) * PW + (tap % 3 - 1) : 0;
            lo = shift < lo ? shift : lo;
            hi = shift > hi ? shift : hi;
          }
        // whole 128-byte lines (8 rows): HALO is a multiple of 8, so a plane read by every tap is staged as before
        const int r0 = (HALO + lo) & ~7, r1 = (HALO + hi + 128 + 7) & ~7;
        pl_off[pl] = (uint32_t)r0;
        pl_bytes[pl] = (uint32_t)(r1 - r0) * 16u;
        stage_tx += 2u * pl_bytes[pl];
      }
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < ((a.debug_flags & 8) ? 0 : a.ntiles); tile += gridDim.x) {
        const long long row0 = (long long)tile * 128;  // = R0 - HALO
        for (int ks0 = 0; ks0 < KS; ks0 += KPB) {
          const int nk = KS - ks0 < KPB ? KS - ks0 : KPB;
          mbar_wait(&empty[st], ph ^ 1u);
          if (a.debug_flags & 2) {   // timing probe: no operand traffic
            mbar_arrive(&full[st]);
          } else {
            mbar_arrive_expect_tx(&full[st], (uint32_t)nk * stage_tx);
            uint8_t* dst = stage_s + (size_t)st * stage_bytes;
            for (int kl = 0; kl < nk; ++kl)       // stage layout: [K step][plane][2 slices][ROWS][16 B]
              for (int pl = 0; pl < NPL; ++pl)
                for (int h2 = 0; h2 < 2; ++h2)
                  bulk_g2s(dst + (size_t)((kl * NPL + pl) * 2 + h2) * slice_bytes + pl_off[pl] * 16u,
                           a.planes[pl] + ((long long)(2 * (ks0 + kl) + h2) * a.RT + row0 + pl_off[pl]) * 8, pl_bytes[pl], &full[st]);
          }
          if (++st == a.nstages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform control flow, one elected lane drives the tensor core =====
    // Descriptors are built once; inside the loop a tap / K-step is one 32-bit add on the
    // start-address field (units of 16 B = one SPS row), so issue costs a few instructions.
    const uint32_t idesc = umma_idesc_bf16(128, a.ncta);
    const uint64_t a_hi = umma_desc(0, slice_bytes, 128) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = umma_desc(0, (uint32_t)a.ncta * 16u, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lbo_lo = (uint32_t)(umma_desc(0, slice_bytes, 128) & 0xFFFF0000u);
    const uint32_t b_lbo_lo = (uint32_t)(umma_desc(0, (uint32_t)a.ncta * 16u, 128) & 0xFFFF0000u);
    const uint32_t w_lo = b_lbo_lo | ((smem_u32(w_s) & 0x3FFFFu) >> 4);
    const uint32_t a_lo0 = a_lbo_lo | (((smem_u32(stage_s) & 0x3FFFFu) >> 4) + (uint32_t)HALO);
    const uint32_t a_stage = stage_bytes >> 4;              // 16 B units between stages
    const uint32_t b_tap = (uint32_t)(a.S_in * a.ncta);     // ... between taps
    const uint32_t b_ks = (uint32_t)(2 * a.ncta);           // ... between K=16 steps
    const bool nine = a.ntaps == 9;
    unsigned long long tap_plane = 0;      // 4 bits per tap: the plane it reads, 15 = dropped
    for (int tap = 0; tap < 9; ++tap) {
      unsigned long long pl = 15;
      for (int q = 0; q < NPL; ++q)
        if (a.tapmask[q] & (1u << tap)) pl = (unsigned long long)q;
      tap_plane |= pl << (4 * tap);
    }
    mbar_wait(wfull, 0);
    int st = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      if (!(a.debug_flags & 32)) mbar_wait(&tempty[acc], accph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.ncta);
      // taps are issued in the plain conv's order (K step, then tap 0..8) whatever plane they read, so a
      // multi-plane conv accumulates in exactly the order of the per-window conv it replaces
      for (int ks0 = 0; ks0 < KS; ks0 += KPB) {
        const int nk = KS - ks0 < KPB ? KS - ks0 : KPB;
        if (!(a.debug_flags & 8)) mbar_wait(&full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          uint32_t go = ks0 != 0 ? 1u : 0u;   // the first MMA of a tile overwrites the accumulator
          for (int kl = 0; kl < nk; ++kl) {
            const int ks = ks0 + kl;
            const uint32_t a_lo = a_lo0 + (uint32_t)st * a_stage + (uint32_t)(kl * NPL) * (kstep_bytes >> 4);
            const uint32_t b_lo = w_lo + (uint32_t)ks * b_ks;
            if (nine) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int pl = (int)((tap_plane >> (4 * tap)) & 15ull);
                if (pl != 15) {
                  const int shift = (tap / 3 - 1) * PW + (tap % 3 - 1);
                  umma_bf16(d_tmem, a_hi | (uint64_t)(a_lo + (uint32_t)pl * (kstep_bytes >> 4) + (uint32_t)shift),
                            b_hi | (uint64_t)(b_lo + (uint32_t)tap * b_tap), idesc, go);
                  go = 1u;
                }
              }
            } else {
              umma_bf16(d_tmem, a_hi | (uint64_t)a_lo, b_hi | (uint64_t)b_lo, idesc, go);
              go = 1u;
            }
          }
          umma_commit(&empty[st]);
          if (ks0 + nk == KS) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++st == a.nstages) { st = 0; ph ^= 1u; }
      }
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> affine (+ReLU) -> bf16 -> SPS rows =====
    const int quarter = warp & 3;  // TMEM lanes this warp may touch
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const long long r = (long long)tile * 128 + row_in_tile;  // row index without the lead halo
      const long long R = r + HALO;
      const long long b = r / PP;
      const int q = (int)(r - b * PP);
      const int i = q / PW, j = q - i * PW;
      const bool valid = (b < a.n_patches) && (i < a.P) && (j < a.P);
      if (a.debug_flags & 128) {       // probe: one polling lane per warp
        if (lane == 0) mbar_wait(&tfull[acc], accph);
        __syncwarp();
      } else {
        mbar_wait(&tfull[acc], accph);
      }
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * a.ncta);
      for (int c0 = 0; c0 < ((a.debug_flags & 16) ? 0 : a.ncta); c0 += 16) {
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
        const int slice = a.out_slice_off + (half * a.ncta + c0) / 8;
        uint4* o0 = reinterpret_cast<uint4*>(a.out + ((long long)slice * a.RT + R) * 8);
        uint4* o1 = reinterpret_cast<uint4*>(a.out + ((long long)(slice + 1) * a.RT + R) * 8);
        if (!(a.debug_flags & 4)) {
          *o0 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *o1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  if ((a.debug_flags & 64) && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    const long long c1 = clock64();
    printf("conv dbg: %lld cycles in %lld ns -> %.0f MHz\n", c1 - dbg_c0, t1 - dbg_t0,
           1e3 * (double)(c1 - dbg_c0) / (double)(t1 - dbg_t0));
  }
}

// ---- CTA-pair variant (cta_group::2) for 128-channel layers whose weights do not fit one CTA ----
// conv 1 of the HSI stem has 9 x 144 x 128 bf16 = 332 KB of weights.  The single-CTA kernel splits
// the output channels over two CTAs (N = 64 per instruction), which leaves the tensor pipe
// operand-fetch bound (A: 4 KB + B: 2 KB per 32-cycle instruction).  Here a cluster of two CTAs
// works as ONE M = 256, N = 128 tile: each CTA stages its own 128 rows (A) and holds HALF of the
// weights (64 of the 128 output channels, exactly the per-half packing of the split kernel); the
// hardware reads the B halves from both shared memories, every instruction does N = 128 per CTA,
// and each CTA's TMEM receives its own 128 rows x all 128 channels.  The leader CTA issues; the
// peer forwards "my stage is full" / "my accumulator is drained" with remote mbarrier arrives;
// tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads) conv_sps_tc2_kernel(ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HALO = sps_halo(a.P), ROWS = 128 + 2 * HALO, PP = sps_pp(a.P), PW = a.P + 1;
  const int KPB = a.kpb;
  const uint32_t slice_bytes = (uint32_t)ROWS * 16u, kstep_bytes = 2u * slice_bytes, stage_bytes = (uint32_t)KPB * kstep_bytes;
  const uint32_t wbytes = (uint32_t)a.ntaps * a.S_in * a.ncta * 16u;     // this CTA's half (ncta = 64 channels)
  const int KS = a.S_in / 2;
  const uint32_t rank = cluster_ctarank();
  const int NOUT = 2 * a.ncta;                                           // 128
  const int npairs = gridDim.x / 2, pair = blockIdx.x / 2;
  const int nsuper = (a.ntiles + 1) / 2;                                 // 256-row super tiles

  uint8_t* w_s = smem;
  uint8_t* stage_s = smem + ((wbytes + 127u) & ~127u);
  uint8_t* tail = stage_s + (size_t)a.nstages * stage_bytes;
  float* sc_s = reinterpret_cast<float*>(tail);
  float* bi_s = sc_s + NOUT;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bi_s + NOUT) + 7) & ~uintptr_t(7));
  uint64_t* full = bars;
  uint64_t* empty = bars + a.nstages;
  uint64_t* pfull = bars + 2 * a.nstages;      // leader: the peer's stage is full
  uint64_t* wfull = bars + 3 * a.nstages;
  uint64_t* pwfull = wfull + 1;                // leader: the peer's weights are resident
  uint64_t* tfull = wfull + 2;                 // [2]
  uint64_t* tempty = wfull + 4;                // [2] leader: both CTAs' epilogues are done with the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 6);

  for (int i = threadIdx.x; i < NOUT; i += blockDim.x) {
    sc_s[i] = a.scale[i];
    bi_s[i] = a.bias[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&pfull[s], 1);
    }
    mbar_init(wfull, 1);
    mbar_init(pwfull, 1);
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 256);
    mbar_init(&tempty[1], 256);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 256);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer (both CTAs): own weight half once, then own 128 rows per stage =====
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) + (size_t)rank * wbytes;
      mbar_arrive_expect_tx(wfull, wbytes);
      for (uint32_t off = 0; off < wbytes; off += 32768u) {
        uint32_t n = wbytes - off < 32768u ? wbytes - off : 32768u;
        bulk_g2s(w_s + off, wsrc + off, n, wfull);
      }
      int st = 0;
      uint32_t ph = 0;
      for (int u = pair; u < nsuper; u += npairs) {
        int tile = 2 * u + (int)rank;
        if (tile >= a.ntiles) tile = a.ntiles - 1;     // odd tile count: the peer re-reads the last tile, stores nothing
        const long long row0 = (long long)tile * 128;
        for (int ks0 = 0; ks0 < KS; ks0 += KPB) {
          const int nk = KS - ks0 < KPB ? KS - ks0 : KPB;
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&full[st], (uint32_t)nk * kstep_bytes);
          uint8_t* dst = stage_s + (size_t)st * stage_bytes;
          for (int sl = 0; sl < 2 * nk; ++sl)
            bulk_g2s(dst + (size_t)sl * slice_bytes, a.in + ((long long)(2 * ks0 + sl) * a.RT + row0) * 8, slice_bytes, &full[st]);
          if (++st == a.nstages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && rank == 1) {
    // ===== peer relay: tell the leader when this CTA's weights / stages have landed =====
    if (lane == 0) {
      mbar_wait(wfull, 0);
      mbar_arrive_remote(pwfull, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int u = pair; u < nsuper; u += npairs) {
        for (int ks0 = 0; ks0 < KS; ks0 += KPB) {
          mbar_wait(&full[st], ph);
          mbar_arrive_remote(&pfull[st], 0);
          if (++st == a.nstages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== leader MMA issuer: one instruction = 256 rows (both CTAs) x 128 channels x K=16 =====
    const uint32_t idesc = umma_idesc_bf16(256, NOUT);
    const uint64_t a_hi = umma_desc(0, slice_bytes, 128) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = umma_desc(0, (uint32_t)a.ncta * 16u, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lbo_lo = (uint32_t)(umma_desc(0, slice_bytes, 128) & 0xFFFF0000u);
    const uint32_t b_lbo_lo = (uint32_t)(umma_desc(0, (uint32_t)a.ncta * 16u, 128) & 0xFFFF0000u);
    const uint32_t w_lo = b_lbo_lo | ((smem_u32(w_s) & 0x3FFFFu) >> 4);
    const uint32_t a_lo0 = a_lbo_lo | (((smem_u32(stage_s) & 0x3FFFFu) >> 4) + (uint32_t)HALO);
    const uint32_t a_stage = stage_bytes >> 4;
    const uint32_t b_tap = (uint32_t)(a.S_in * a.ncta);
    const uint32_t b_ks = (uint32_t)(2 * a.ncta);
    const bool nine = a.ntaps == 9;
    mbar_wait(wfull, 0);
    mbar_wait(pwfull, 0);
    int st = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t accph = 0;
    for (int u = pair; u < nsuper; u += npairs) {
      mbar_wait(&tempty[acc], accph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NOUT);
      for (int ks0 = 0; ks0 < KS; ks0 += KPB) {
        const int nk = KS - ks0 < KPB ? KS - ks0 : KPB;
        mbar_wait(&full[st], ph);
        mbar_wait(&pfull[st], ph);
        tc_fence_after();
        if (elect_one()) {
          for (int kl = 0; kl < nk; ++kl) {
            const int ks = ks0 + kl;
            const uint32_t a_lo = a_lo0 + (uint32_t)st * a_stage + (uint32_t)kl * (kstep_bytes >> 4);
            const uint32_t b_lo = w_lo + (uint32_t)ks * b_ks;
            if (nine) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int shift = (tap / 3 - 1) * PW + (tap % 3 - 1);
                umma_bf16_2(d_tmem, a_hi | (uint64_t)(a_lo + (uint32_t)shift), b_hi | (uint64_t)(b_lo + (uint32_t)tap * b_tap),
                            idesc, (ks | tap) != 0 ? 1u : 0u);
              }
            } else {
              umma_bf16_2(d_tmem, a_hi | (uint64_t)a_lo, b_hi | (uint64_t)b_lo, idesc, ks != 0 ? 1u : 0u);
            }
          }
          umma_commit2(&empty[st], 3);
          if (ks0 + nk == KS) umma_commit2(&tfull[acc], 3);
        }
        __syncwarp();
        if (++st == a.nstages) { st = 0; ph ^= 1u; }
      }
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  } else {
    // ===== epilogue (both CTAs): own 128 rows x 128 channels =====
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t accph = 0;
    for (int u = pair; u < nsuper; u += npairs) {
      const int tile = 2 * u + (int)rank;
      const bool have = tile < a.ntiles;
      const long long r = (long long)tile * 128 + row_in_tile;
      const long long R = r + HALO;
      const long long b = r / PP;
      const int q = (int)(r - b * PP);
      const int i = q / PW, j = q - i * PW;
      const bool valid = (b < a.n_patches) && (i < a.P) && (j < a.P);
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * NOUT);
      for (int c0 = 0; c0 < NOUT; c0 += 16) {
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
        if (have) {
          const int slice = a.out_slice_off + c0 / 8;
          *reinterpret_cast<uint4*>(a.out + ((long long)slice * a.RT + R) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(a.out + ((long long)(slice + 1) * a.RT + R) * 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before();
      if (rank == 0) mbar_arrive(&tempty[acc]);
      else mbar_arrive_remote(&tempty[acc], 0);
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) tmem_dealloc2(tmem_base, 256);
}

// ---- plain SIMT twin (tests / bring-up cross-check only; same arguments, same layouts) ------
__global__ void conv_sps_simt_kernel(ConvArgs a, int nsplit) {
  const int HALO = sps_halo(a.P), PP = sps_pp(a.P), PW = a.P + 1;
  const int ntot = a.ncta * nsplit;
  const long long total = (long long)a.ntiles * 128 * ntot;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int oc = (int)(idx % ntot);
    const long long r = idx / ntot;
    const long long R = r + HALO;
    const long long b = r / PP;
    const int q = (int)(r - b * PP);
    const int i = q / PW, j = q - i * PW;
    const bool valid = (b < a.n_patches) && (i < a.P) && (j < a.P);
    const int half = oc / a.ncta, n = oc % a.ncta;
    float acc = 0.f;
    if (valid) {
      for (int tap = 0; tap < a.ntaps; ++tap) {
        const int shift = (a.ntaps == 9) ? ((tap / 3 - 1) * PW + (tap % 3 - 1)) : 0;
        for (int s = 0; s < a.S_in; ++s) {
          const __nv_bfloat16* x = a.in + ((long long)s * a.RT + R + shift) * 8;
          const __nv_bfloat16* w = a.w + ((((long long)half * a.ntaps + tap) * a.S_in + s) * a.ncta + n) * 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc += __bfloat162float(x[k]) * __bfloat162float(w[k]);
        }
      }
      acc = acc * a.scale[oc] + a.bias[oc];
      if (a.relu) acc = fmaxf(acc, 0.f);
    }
    a.out[((long long)(a.out_slice_off + oc / 8) * a.RT + R) * 8 + (oc & 7)] = __float2bfloat16_rn(acc);
  }
}

static size_t conv_smem_bytes(int S_in, int ncta, int ntaps, int P, int nstages, int kpb, int nplanes = 1) {
  const int ROWS = 128 + 2 * sps_halo(P);
  size_t wbytes = ((size_t)ntaps * S_in * ncta * 16 + 127) & ~size_t(127);
  return wbytes + (size_t)nstages * kpb * nplanes * 2 * ROWS * 16 + (size_t)ncta * 8 + 8 + (2 * nstages + 5) * 8 + 16;
}

int conv_sps_launch(const void* in, int S_in, const void* w, const float* scale, const float* bias, void* out,
                    int out_slice_off, int n_out, int nsplit, int n_patches, int P, int ntaps, int relu, int impl,
                    int debug_flags, cudaStream_t stream) {
  const unsigned int mask = ntaps == 9 ? 0x1FFu : 1u;
  return conv_sps_planes_launch(&in, &mask, 1, S_in, w, scale, bias, out, out_slice_off, n_out, nsplit, n_patches, P, ntaps,
                                relu, impl, debug_flags, stream);
}

int conv_sps_planes_launch(const void* const* planes, const unsigned int* tapmasks, int nplanes, int S_in, const void* w,
                           const float* scale, const float* bias, void* out, int out_slice_off, int n_out, int nsplit,
                           int n_patches, int P, int ntaps, int relu, int impl, int debug_flags, cudaStream_t stream) {
  if (S_in <= 0 || (S_in & 1) || nsplit <= 0 || n_out % nsplit || nplanes < 1 || nplanes > 9) return VC_ERR_ARG;
  const int ncta = n_out / nsplit;
  if (ncta % 16 || ncta < 16 || ncta > 128 || (ntaps != 9 && ntaps != 1) || P < 1 || n_patches <= 0) return VC_ERR_ARG;
  const unsigned int all_taps = ntaps == 9 ? 0x1FFu : 1u;
  const bool plain = nplanes == 1 && tapmasks[0] == all_taps;
  if (!plain && impl != 0) return VC_ERR_UNSUPPORTED;
  ConvArgs a;
  a.nplanes = nplanes;
  for (int i = 0; i < 9; ++i) {
    a.planes[i] = (const __nv_bfloat16*)planes[i < nplanes ? i : 0];
    a.tapmask[i] = i < nplanes ? (tapmasks[i] & all_taps) : 0u;
    if (i < nplanes && !a.tapmask[i]) return VC_ERR_ARG;
  }
  const void* in = planes[0];
  a.in = (const __nv_bfloat16*)in;
  a.w = (const __nv_bfloat16*)w;
  a.scale = scale;
  a.bias = bias;
  a.out = (__nv_bfloat16*)out;
  a.RT = sps_rows(n_patches, P);
  a.S_in = S_in;
  a.ncta = ncta;
  a.ntaps = ntaps;
  a.P = P;
  a.n_patches = n_patches;
  a.ntiles = sps_tiles(n_patches, P);
  a.out_slice_off = out_slice_off;
  a.relu = relu;
  a.debug_flags = debug_flags;
  a.nstages = 0;
  a.kpb = 1;
  if (impl == 1) {
    long long total = (long long)a.ntiles * 128 * n_out;
    int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    conv_sps_simt_kernel<<<blocks, 256, 0, stream>>>(a, nsplit);
    return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
  }
  static int max_smem = 0, num_sms = 0;
  if (!max_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // Stage = kpb K=16 steps: fewer barrier round trips per MMA (measured: 95 -> 69 cycles per
  // MMA going from 1 to 3..4 steps per stage).  Pick the kpb with the fewest stage iterations
  // per tile that still leaves >= 3 stages in flight; ties go to the smaller stage.
  const int KS = S_in / 2;
  int kpb = 1, nst = 0;
  const int force_kpb = (debug_flags >> 12) & 7, force_nst = (debug_flags >> 8) & 15;
  {
    int best_iters = 1 << 30;
    for (int cand = 1; cand <= 4; ++cand) {
      if (cand > KS || (force_kpb && cand != force_kpb)) continue;
      int n = force_nst ? force_nst : 6;
      while (n > 2 && conv_smem_bytes(S_in, ncta, ntaps, P, n, cand, nplanes) > (size_t)max_smem) --n;
      if (conv_smem_bytes(S_in, ncta, ntaps, P, n, cand, nplanes) > (size_t)max_smem) continue;
      if (n < 3 && cand > 1) continue;
      const int iters = (KS + cand - 1) / cand;
      if (iters < best_iters) { best_iters = iters; kpb = cand; nst = n; }
    }
  }
  if (nst == 0) return VC_ERR_UNSUPPORTED;
  // CTA-pair kernel: 128 output channels split over two CTAs (impl 3 forces the split kernel)
  static int pair_ok = -1;
  if (pair_ok < 0) {
    const char* e = getenv("VC_CONV_PAIR");
    pair_ok = (e && e[0] == '0') ? 0 : 1;
  }
  if (impl == 0 && plain && pair_ok && nsplit == 2 && ncta == 64 && a.ntiles >= 2) {
    const size_t smem2 = conv_smem_bytes(S_in, ncta, ntaps, P, nst, kpb) + (size_t)nst * 8 + 64 + 2 * 128 * 4;
    if (smem2 <= (size_t)max_smem) {
      a.nstages = nst;
      a.kpb = kpb;
      if (cudaFuncSetAttribute(conv_sps_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess)
        return VC_ERR_CUDA;
      int npairs = num_sms / 2;
      const int nsuper = (a.ntiles + 1) / 2;
      if (npairs > nsuper) npairs = nsuper;
      conv_sps_tc2_kernel<<<2 * npairs, kConvThreads, smem2, stream>>>(a);
      return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
    }
  }
  const size_t smem = conv_smem_bytes(S_in, ncta, ntaps, P, nst, kpb, nplanes);
  a.nstages = nst;
  a.kpb = kpb;
  if (cudaFuncSetAttribute(conv_sps_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VC_ERR_CUDA;
  // small layers are latency-bound per tile: let several CTAs share an SM (TMEM: 2*ncta columns each)
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_sps_tc_kernel, kConvThreads, smem);
  const int tmem_cols = 2 * ncta <= 32 ? 32 : 2 * ncta <= 64 ? 64 : 2 * ncta <= 128 ? 128 : 256;
  if (occ > 512 / tmem_cols) occ = 512 / tmem_cols;
  if (occ > 4) occ = 4;
  if (occ < 1) occ = 1;
  int gx = num_sms * occ / nsplit;
  if (gx < 1) gx = 1;
  if (gx > a.ntiles) gx = a.ntiles;
  dim3 grid(gx, nsplit);
  conv_sps_tc_kernel<<<grid, kConvThreads, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
