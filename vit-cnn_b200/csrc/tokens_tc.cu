// Token stage of the ViT-CNN hybrid on the 5th-gen tensor cores (eval mode, P*P + 1 <= 128 tokens).
//
// Same function as transformer_fwd_kernel (transformer.cu; spec SURVEY.md App. A, following
// model/compare_method/vit/timm/models/vision_transformer.py:57-105 Attention, :123-166 Block,
// :598-629 cls/pos, :680-701 norm/head and timm/layers/mlp.py:13-47), re-laid for tcgen05:
//
//  * one patch = one M = 128 tile: token row r lives in TMEM lane r and in thread r of a
//    128-thread group, so LayerNorm, the softmax row maximum and every bias / GELU / residual
//    epilogue are thread-local (no shuffles); the residual stream stays in 32 registers;
//  * every GEMM of the block (fusion 1x1 conv, qkv, QK^T per head, PV per head, proj, fc1, fc2)
//    is a tcgen05.mma with the accumulator in TMEM; A operands are written by the row threads
//    straight into the no-swizzle K-major canonical layout ([k/8][row][8] = 2 KB slabs), weights
//    are re-laid into that layout once per persistent CTA;
//  * head_dim = 8 is half of the K = 16 of one bf16 MMA: the second 8-element chunk of Q is
//    pointed (leading-byte-offset field) at a slab of (1,0,..,0) rows and that of K at a slab
//    holding (-30000,0,..,0) for padded keys and zeros for real ones, so S_h = Q_h K_h^T needs no
//    padded copies and padded keys arrive already masked; when a static bound on |q.k| (from the
//    weight norms, LayerNorm output has norm sqrt(32)) shows 2^s cannot overflow, the softmax
//    skips the row maximum altogether (softmax is shift invariant; bf16 keeps the fp32 exponent);
//    V is consumed MN-major ([key][8 dims] slabs) and its second N chunk is a
//    slab of (1,0,..,0) rows, so column 8 of the PV accumulator is the softmax denominator of
//    the bf16-rounded probabilities;
//  * a CTA runs three independent 128-thread groups ("slots": 56 KB of operand buffers with
//    non-overlapping lifetimes aliased, 160 TMEM columns each) so one slot's MMA round trips hide
//    behind the others' softmax, plus one MMA-issuer warp per slot:
//    row threads never wait for each other, they arrive on mbarriers the issuer waits on and
//    wait only for tensor-core completions; the fusion-conv input of the next patch is
//    prefetched with 16-byte cp.async (one token row per thread) over the K / V buffers once
//    the last PV of the current patch is done;
//  * the last block only feeds the head through the cls token (x[:, 0]): K / V of every token
//    come from one more MMA, the single-query attention is a thread-local dot product + a
//    butterfly reduction per warp, and the cls row (proj, MLP, final LN, head) is finished by
//    tokens_tail_kernel, one thread per patch, from a 704-byte record per patch.
#include "tokens_tc_common.cuh"

namespace vc {

// Warp roles: warps 0-3 = row threads of slot 0, 4-7 = slot 1 (token row r = thread & 127, TMEM lane r),
// warps 8 / 9 = MMA issuers of slot 0 / 1.  Row threads never wait for each other: they signal
// "operands written" / "S consumed" on two mbarriers (128 arrivals) that only the issuer waits on, and
// wait only for tensor-core completions (tcgen05.commit mbarriers).
__global__ void __launch_bounds__(tc::kThreads, 1) tokens_tc_kernel(TcArgs a) {   // 15 warps -> 128 registers (the file is allocated per 4 warps)
  using namespace tc;
  if (a.run_flag && *a.run_flag == 0.f) return;     // fallback launch behind tokens_tm_kernel: nothing to do
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp >= 4 * kSlots;
  const int slot = issuer ? warp - 4 * kSlots : warp >> 2;
  const int r = tid & 127, wq = warp & 3;
  const int T = a.T, P = a.P;
  const TLayout& L = a.L;
  const uint32_t sb = smem_u32(smem);
  float* vecf = reinterpret_cast<float*>(smem + VEC);
  float* q0_s = reinterpret_cast<float*>(smem + MISC + M_Q0) + slot * 32;          // [32]
  float* wmax_s = reinterpret_cast<float*>(smem + MISC + M_WMAX) + slot * 16;      // [4 warps][4 heads]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MISC + M_BARS) + slot * 5;
  uint64_t* b_rp = bars + 0;      // row threads: operands of the next GEMM are written (4 arrivals: one per row warp)
  uint64_t* b_rs = bars + 1;      // row threads: S_h has been read out of TMEM (4 arrivals)
  uint64_t* b_mma = bars + 2;     // tensor core: the GEMM just issued (fusion / qkv / proj / fc1 / fc2 / kv2) is done
  uint64_t* b_s = bars + 3;       // tensor core: S_h is in TMEM
  uint64_t* b_pv = bars + 4;      // tensor core: PV_h is done (P buffer free, O_h in TMEM)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + MISC + M_TMEM);
  const float qscale = 0.35355339059327376220f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)

  bool exact_softmax, exact_cls;   // block 1 (all queries) / last block (cls query only): the softmax needs the row maximum
  tc_setup(a, smem, tid, kThreads, VC_TC_WARP_ARRIVE ? 4 : 128, exact_softmax, exact_cls);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tb = tmem_base + (uint32_t)slot * C_SLOT;                // columns of this slot (issuer view)
  const uint32_t tl = tb + ((uint32_t)(wq * 32) << 16);                   // + the 32 lanes of this warp
  const uint32_t slot_s = sb + SLOT0 + (uint32_t)slot * SLOT_BYTES;
  const uint32_t fbuf = slot_s + S_FBUF, abuf = slot_s + S_ABUF, qbuf = slot_s + S_QBUF, kbuf = slot_s + S_KBUF,
                 vbuf = slot_s + S_VBUF, pbuf = slot_s + S_PBUF;
  const int NK = (T + 31) & ~31, NKS = NK >> 4;       // keys rounded to the 32-column chunks the row threads read
  const int nslots = kSlots * (int)gridDim.x;
  const int b0 = kSlots * (int)blockIdx.x + slot;

  if (issuer) {
    tc_issuer(a, sb, tb, slot_s, bars, b0, nslots);
  } else {
    // ============================ row threads ============================
    const uint32_t row16 = (uint32_t)r * 16u;
    const int ct = (T - 1) >> 5;                         // last 32-key chunk holding a real key
    const int PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P);
    const int bar_id = 1 + slot;
    const bool w0 = wq == 0;
    uint32_t ph_m = 0, ph_s = 0, ph_pv = 0;
    // stem outputs of token row r of patch b -> FBUF (8 slices x 16 B; cls row and padding rows stay zero)
    // planes mode (shared stem, depth D): the stem output of window pixel (i, j) is the variant plane of border_class(i) x
    // border_class(j) at the scene pixel's row in its scene block; (i, j) and with them the variant are fixed per thread
    const int tok_i = r >= 1 && r < T ? (r - 1) / P : 0, tok_j = r >= 1 && r < T ? (r - 1) - tok_i * P : 0;
    const bool planes = a.pl.h || a.pl.l;
    const long long voff0 = planes ? ((long long)(border_class(tok_i, P, a.pl.D) * (2 * a.pl.D + 1) + border_class(tok_j, P, a.pl.D)) * 4 * a.pl.RTb +
                                      sps_halo(a.pl.B)) : 0;
    auto fetch = [&](int b) {
      if (r >= 1 && r < T) {
        const __nv_bfloat16* src = a.f + (HALO + (long long)b * PP + tok_i * PW + tok_j) * 8;
        const __nv_bfloat16 *sh = src, *sl = src + 4 * a.RT * 8;
        long long ph = a.RT * 8, pls = a.RT * 8;          // slice pitch (elements) of the HSI / LiDAR source
        if (planes) {
          const int2 c = __ldg(reinterpret_cast<const int2*>(a.pl.xy) + b);
          const long long voff = (voff0 + __ldg(a.pl.rowterm + c.x + tok_i) + __ldg(a.pl.colterm + c.y + tok_j)) * 8;
          if (a.pl.h) { sh = a.pl.h + voff; ph = a.pl.RTb * 8; }
          if (a.pl.l) { sl = a.pl.l + voff; pls = a.pl.RTb * 8; }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) cp_async16(fbuf + s * SLAB + row16, sh + s * ph);
#pragma unroll
        for (int s = 0; s < 4; ++s) cp_async16(fbuf + (4 + s) * SLAB + row16, sl + s * pls);
      } else {          // the buffer doubles as P / hidden: the cls row and the padding rows are zeroed every time
#pragma unroll
        for (int s = 0; s < 8; ++s) sts128(fbuf + s * SLAB + row16, 0u, 0u, 0u, 0u);
      }
    };
    // this thread's operand rows are written -> visible to the tensor core; its TMEM reads are retired
    // (one arrival per warp: the lanes order their writes / TMEM reads before the elected lane's release-arrive
    // with __syncwarp, instead of 128 serialised arrivals on one shared-memory word per hand-off)
    auto row_arrive = [&](uint64_t* bar) {
#if VC_TC_WARP_ARRIVE
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
#else
      mbar_arrive(bar);
#endif
    };
    // MUFU-phase lock (VC_TC_MUFU_LOCK): the three slots' warps on one SM sub-partition share its MUFU pipe; left alone they
    // drift into the softmax phases together (processor sharing: all three finish late, then all three leave the pipe idle).
    // First-come-first-served per sub-partition lets one warp run its exponentials at the full pipe rate while the others
    // are still in (or return to) their latency-bound phases.
    uint32_t* mufu_lock = reinterpret_cast<uint32_t*>(smem + MISC + M_LOCK) + wq;
    auto mufu_acquire = [&]() {
#if VC_TC_MUFU_LOCK
      if (lane == 0)
        while (atomicCAS(mufu_lock, 0u, 1u) != 0u) __nanosleep(40);
      __syncwarp();
#endif
    };
    auto mufu_release = [&]() {
#if VC_TC_MUFU_LOCK
      __syncwarp();
      if (lane == 0) atomicExch(mufu_lock, 0u);
#endif
    };
    auto publish = [&]() {
      fence_proxy_async();
      tc_fence_before();
      row_arrive(b_rp);
    };
    auto wait_mma = [&]() {
      mbar_wait(b_mma, ph_m);
      ph_m ^= 1u;
      tc_fence_after();
    };
    if (b0 < a.n_patches) fetch(b0);
    // the slots run the same program: started together they would hit the SFU-bound softmax phases together and
    // idle together in the latency-bound ones, so slot k starts k thirds of a patch time later
    if (slot > 0 && a.stagger_ns > 0) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      } while (t1 - t0 < (unsigned long long)a.stagger_ns * slot);
    }

    for (int b = b0; b < a.n_patches; b += nslots) {
      float x[32];   // residual stream of token row r
      // ================= fusion 1x1 conv (64 -> 32) + folded BN + ReLU, + cls / pos =================
      cp_async_wait_all();
      publish();
      wait_mma();
      {
        uint32_t v[32];
        tmem_ld32(tl + C_SMALL, v);
        tc_wait_ld();
        const float rowmask = (r >= 1 && r < T) ? 1.f : 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 p = lds_f4(sb + POS + r * 128 + ((g ^ (r & 7)) << 4));
          const float4 sc = lds_f4(sb + VEC + (V_FSC + 4 * g) * 4), bi = lds_f4(sb + VEC + (V_FBI + 4 * g) * 4);
          x[4 * g + 0] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 0]), sc.x, bi.x), 0.f), rowmask, p.x);
          x[4 * g + 1] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 1]), sc.y, bi.y), 0.f), rowmask, p.y);
          x[4 * g + 2] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 2]), sc.z, bi.z), 0.f), rowmask, p.z);
          x[4 * g + 3] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 3]), sc.w, bi.w), 0.f), rowmask, p.w);
        }
      }

      // ================= block 1: LN1 -> qkv =================
      ln_store(x, sb + VEC + V_LN1G * 4, sb + VEC + V_LN1B * 4, abuf + row16);
      publish();
      wait_mma();
      {
        uint32_t v[2][32];
        tmem_ld32(tl + C_S, v[0]);
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          tc_wait_ld();
          if (part < 2) tmem_ld32(tl + C_S + 32 * (part + 1), v[(part + 1) & 1]);
          const uint32_t dst = (part == 0 ? qbuf : part == 1 ? kbuf : vbuf) + row16;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float4 b0_ = lds_f4(sb + VEC + (V_BQKV + 32 * part + 8 * h) * 4), b1_ = lds_f4(sb + VEC + (V_BQKV + 32 * part + 8 * h + 4) * 4);
            const float sc = part == 0 ? qscale : 1.f;
            const uint32_t* vv = v[part & 1] + 8 * h;
            sts128(dst + h * SLAB, pack_bf16(fmaf(__uint_as_float(vv[0]), sc, b0_.x), fmaf(__uint_as_float(vv[1]), sc, b0_.y)),
                   pack_bf16(fmaf(__uint_as_float(vv[2]), sc, b0_.z), fmaf(__uint_as_float(vv[3]), sc, b0_.w)),
                   pack_bf16(fmaf(__uint_as_float(vv[4]), sc, b1_.x), fmaf(__uint_as_float(vv[5]), sc, b1_.y)),
                   pack_bf16(fmaf(__uint_as_float(vv[6]), sc, b1_.z), fmaf(__uint_as_float(vv[7]), sc, b1_.w)));
          }
        }
      }
      publish();

      // ================= attention, one head at a time =================
      // p = 2^(s - m) -> bf16 -> P buffer (A operand of PV); S_h is released to the issuer as soon as its last
      // chunk is in registers.  `exact`: two passes over S_h in TMEM (row maximum first); otherwise m = 0.
      // O_h lands in one of two 16-column buffers and is read (normalised by column 8, the softmax denominator
      // of the bf16 probabilities) while the next head's probabilities are produced.
      // The normalised O_h goes straight into slab h of the Q buffer (= the A operand of proj): S_h is done, so
      // Q_h is dead, and later heads read their own slabs only.
      auto read_o = [&](int h) {
        uint32_t o[16];
        tmem_ld16(tl + C_O + 16 * (h & 1), o);
        tc_wait_ld();
        const float il = 1.f / __uint_as_float(o[8]);
        sts128(abuf + h * SLAB + row16, pack_bf16(__uint_as_float(o[0]) * il, __uint_as_float(o[1]) * il),
               pack_bf16(__uint_as_float(o[2]) * il, __uint_as_float(o[3]) * il),
               pack_bf16(__uint_as_float(o[4]) * il, __uint_as_float(o[5]) * il),
               pack_bf16(__uint_as_float(o[6]) * il, __uint_as_float(o[7]) * il));
      };
      auto softmax_head = [&](int h, auto exact_tag) {
        constexpr bool kExact = decltype(exact_tag)::value;
        uint32_t sc[32];
        mbar_wait(b_s, ph_s);
        ph_s ^= 1u;
        tc_fence_after();
        float m = 0.f;
        if constexpr (kExact) {
          m = -INFINITY;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c <= ct) {
              tmem_ld32(tl + C_S + 32 * c, sc);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; i += 2) m = fmaxf(m, fmaxf(__uint_as_float(sc[i]), __uint_as_float(sc[i + 1])));
            }
          }
        }
        if (h > 0) {                      // PV of the previous head is done: the P buffer is free, O_{h-1} is in TMEM
          mbar_wait(b_pv, ph_pv);
          ph_pv ^= 1u;
          tc_fence_after();
          read_o(h - 1);
        }
        mufu_acquire();
#if VC_TC_LD16
        // 16-key half chunks, double buffered: the TMEM load of half chunk c + 1 is in flight while c is exponentiated
        {
          uint32_t sb2[2][16];
          const int ct2 = 2 * ct + 1;      // last half chunk of the last 32-key chunk holding a real key
          tmem_ld16(tl + C_S, sb2[0]);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (c <= ct2) {
              tc_wait_ld();
              if (c < ct2) tmem_ld16(tl + C_S + 16 * (c + 1), sb2[(c + 1) & 1]);
              if (c == ct2 && h < 3) {       // S_h is out of TMEM: the next head's S may overwrite it
                tc_fence_before();
                row_arrive(b_rs);
              }
              const uint32_t(&sv)[16] = sb2[c & 1];
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float s0 = __uint_as_float(sv[8 * g + 2 * e]), s1 = __uint_as_float(sv[8 * g + 2 * e + 1]);
                  if constexpr (kExact) {
                    pk[e] = pack_bf16(ex2(s0 - m), ex2(s1 - m));
                  } else {
                    const float p0 = ((VC_TC_POLY_MASK >> (2 * e)) & 1) ? ex2_poly(s0) : ex2(s0);
                    const float p1 = ((VC_TC_POLY_MASK >> (2 * e + 1)) & 1) ? ex2_poly(s1) : ex2(s1);
                    pk[e] = pack_bf16(p0, p1);
                  }
                }
                sts128(pbuf + (2 * c + g) * SLAB + row16, pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
#else
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c <= ct) {
            tmem_ld32(tl + C_S + 32 * c, sc);
            tc_wait_ld();
            if (c == ct && h < 3) {       // S_h is out of TMEM: the next head's S may overwrite it
              tc_fence_before();
              row_arrive(b_rs);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float s0 = __uint_as_float(sc[8 * g + 2 * e]), s1 = __uint_as_float(sc[8 * g + 2 * e + 1]);
                if constexpr (kExact) {
                  pk[e] = pack_bf16(ex2(s0 - m), ex2(s1 - m));
                } else {
                  const float p0 = ((VC_TC_POLY_MASK >> (2 * e)) & 1) ? ex2_poly(s0) : ex2(s0);
                  const float p1 = ((VC_TC_POLY_MASK >> (2 * e + 1)) & 1) ? ex2_poly(s1) : ex2(s1);
                  pk[e] = pack_bf16(p0, p1);
                }
              }
              sts128(pbuf + (4 * c + g) * SLAB + row16, pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
#endif
        mufu_release();
        publish();
      };
      if (exact_softmax) {
#pragma unroll 1
        for (int h = 0; h < 4; ++h) softmax_head(h, std::true_type{});
      } else {
#pragma unroll 1
        for (int h = 0; h < 4; ++h) softmax_head(h, std::false_type{});
      }
      mbar_wait(b_pv, ph_pv);
      ph_pv ^= 1u;
      tc_fence_after();
      if (b + nslots < a.n_patches) fetch(b + nslots);   // K and V are dead: the next patch's fusion input lands over them
      read_o(3);
      publish();
      wait_mma();
      {
        uint32_t v[32];
        tmem_ld32(tl + C_SMALL, v);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bb = lds_f4(sb + VEC + (V_BPROJ + 4 * g) * 4);
          x[4 * g + 0] += __uint_as_float(v[4 * g + 0]) + bb.x;
          x[4 * g + 1] += __uint_as_float(v[4 * g + 1]) + bb.y;
          x[4 * g + 2] += __uint_as_float(v[4 * g + 2]) + bb.z;
          x[4 * g + 3] += __uint_as_float(v[4 * g + 3]) + bb.w;
        }
      }

      // ================= MLP: LN2 -> fc1 (+bias, GELU) -> fc2 (+bias, +residual) =================
      ln_store(x, sb + VEC + V_LN2G * 4, sb + VEC + V_LN2B * 4, abuf + row16);
      publish();
      wait_mma();
      {
        uint32_t v[2][32];
        tmem_ld32(tl + C_S, v[0]);
        mufu_acquire();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc_wait_ld();
          if (c < 3) tmem_ld32(tl + C_S + 32 * (c + 1), v[(c + 1) & 1]);
          const uint32_t(&vc_)[32] = v[c & 1];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b0_ = lds_f4(sb + VEC + (V_BFC1 + 32 * c + 8 * g) * 4), b1_ = lds_f4(sb + VEC + (V_BFC1 + 32 * c + 8 * g + 4) * 4);
            const uint32_t* vv = vc_ + 8 * g;
            sts128(pbuf + (4 * c + g) * SLAB + row16,
                   pack_bf16(gelu2(__uint_as_float(vv[0]) + b0_.x), gelu2(__uint_as_float(vv[1]) + b0_.y)),
                   pack_bf16(gelu2(__uint_as_float(vv[2]) + b0_.z), gelu2(__uint_as_float(vv[3]) + b0_.w)),
                   pack_bf16(gelu2(__uint_as_float(vv[4]) + b1_.x), gelu2(__uint_as_float(vv[5]) + b1_.y)),
                   pack_bf16(gelu2(__uint_as_float(vv[6]) + b1_.z), gelu2(__uint_as_float(vv[7]) + b1_.w)));
          }
        }
        mufu_release();
      }
      publish();
      wait_mma();
      {
        uint32_t v[32];
        tmem_ld32(tl + C_SMALL, v);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bb = lds_f4(sb + VEC + (V_BFC2 + 4 * g) * 4);
          x[4 * g + 0] += __uint_as_float(v[4 * g + 0]) + bb.x;
          x[4 * g + 1] += __uint_as_float(v[4 * g + 1]) + bb.y;
          x[4 * g + 2] += __uint_as_float(v[4 * g + 2]) + bb.z;
          x[4 * g + 3] += __uint_as_float(v[4 * g + 3]) + bb.w;
        }
      }

      // ================= last block: K / V of every token, attention of the cls query only =================
      ln_store(x, sb + VEC + V_L2G * 4, sb + VEC + V_L2B * 4, abuf + row16);
      publish();
      wait_mma();
      float* trec = a.tail + (long long)b * kTailFloats;
      {
        uint32_t kk[32], vv[32];
        tmem_ld32(tl + C_S + 32, kk);
        tmem_ld32(tl + C_S + 64, vv);
        if (w0) {     // the cls token is row 0: its (scaled) query and its residual stream
          uint32_t qq[32];
          tmem_ld32(tl + C_S, qq);
          tc_wait_ld();
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              q0_s[c] = fmaf(__uint_as_float(qq[c]), qscale, vecf[V_BQKV2 + c]);
              trec[144 + c] = x[c];
            }
          }
        }
        tc_wait_ld();
        bar_sync(bar_id, 128);
        float sc[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float4 q0 = lds_f4(smem_u32(q0_s) + 32 * h), q1 = lds_f4(smem_u32(q0_s) + 32 * h + 16);
          const float4 k0 = lds_f4(sb + VEC + (V_BQKV2 + 32 + 8 * h) * 4), k1 = lds_f4(sb + VEC + (V_BQKV2 + 36 + 8 * h) * 4);
          float d = q0.x * (__uint_as_float(kk[8 * h + 0]) + k0.x);
          d = fmaf(q0.y, __uint_as_float(kk[8 * h + 1]) + k0.y, d);
          d = fmaf(q0.z, __uint_as_float(kk[8 * h + 2]) + k0.z, d);
          d = fmaf(q0.w, __uint_as_float(kk[8 * h + 3]) + k0.w, d);
          d = fmaf(q1.x, __uint_as_float(kk[8 * h + 4]) + k1.x, d);
          d = fmaf(q1.y, __uint_as_float(kk[8 * h + 5]) + k1.y, d);
          d = fmaf(q1.z, __uint_as_float(kk[8 * h + 6]) + k1.z, d);
          d = fmaf(q1.w, __uint_as_float(kk[8 * h + 7]) + k1.w, d);
          sc[h] = r < T ? d : -INFINITY;
          if (exact_cls) {           // the maximum over all keys is only needed when 2^s could overflow
            float mw = sc[h];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
            if (lane == 0) wmax_s[wq * 4 + h] = mw;
          }
        }
        if (exact_cls) bar_sync(bar_id, 128);
        float val[32], pl[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float m = exact_cls ? fmaxf(fmaxf(wmax_s[h], wmax_s[4 + h]), fmaxf(wmax_s[8 + h], wmax_s[12 + h])) : 0.f;
          const float p = ex2(sc[h] - m);
          pl[h] = p;
          const float4 v0 = lds_f4(sb + VEC + (V_BQKV2 + 64 + 8 * h) * 4), v1 = lds_f4(sb + VEC + (V_BQKV2 + 68 + 8 * h) * 4);
          val[8 * h + 0] = p * (__uint_as_float(vv[8 * h + 0]) + v0.x);
          val[8 * h + 1] = p * (__uint_as_float(vv[8 * h + 1]) + v0.y);
          val[8 * h + 2] = p * (__uint_as_float(vv[8 * h + 2]) + v0.z);
          val[8 * h + 3] = p * (__uint_as_float(vv[8 * h + 3]) + v0.w);
          val[8 * h + 4] = p * (__uint_as_float(vv[8 * h + 4]) + v1.x);
          val[8 * h + 5] = p * (__uint_as_float(vv[8 * h + 5]) + v1.y);
          val[8 * h + 6] = p * (__uint_as_float(vv[8 * h + 6]) + v1.z);
          val[8 * h + 7] = p * (__uint_as_float(vv[8 * h + 7]) + v1.w);
        }
        // butterfly reduction over the 32 rows of this warp: lane i ends with sum over rows of val[i]
#pragma unroll
        for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float send = up ? val[i] : val[i + n / 2];
            const float keep = up ? val[i + n / 2] : val[i];
            val[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        {   // denominators: head = 2 * bit4 + bit3 of the lane after two halving steps, then a full reduce over bits 2..0
          const bool up4 = (lane & 16) != 0, up3 = (lane & 8) != 0;
          float a0 = (up4 ? pl[2] : pl[0]) + __shfl_xor_sync(0xffffffffu, up4 ? pl[0] : pl[2], 16);
          float a1 = (up4 ? pl[3] : pl[1]) + __shfl_xor_sync(0xffffffffu, up4 ? pl[1] : pl[3], 16);
          float l = (up3 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up3 ? a0 : a1, 8);
          l += __shfl_xor_sync(0xffffffffu, l, 4);
          l += __shfl_xor_sync(0xffffffffu, l, 2);
          l += __shfl_xor_sync(0xffffffffu, l, 1);
          trec[wq * 36 + lane] = val[0];
          if ((lane & 7) == 0) trec[wq * 36 + 32 + (lane >> 3)] = l;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

// ---- cls row of the last block: attention normalisation, proj, MLP, final LayerNorm, head ----
// One THREAD per patch (the tail of transformer_fwd_kernel as its own launch): the cls row is 32 channels, so
// the whole tail is ~9.5 k FMAs with 32 independent accumulators per thread and no shuffles; the weights are
// converted to fp32 once per CTA and read as warp-uniform (broadcast) 16-byte loads.
constexpr int kTailThreads = 64;
constexpr int T_WPROJ = 0, T_WFC1 = T_WPROJ + kD * kD, T_WFC2T = T_WFC1 + kHidden * kD, T_END = T_WFC2T + kHidden * kD;   // floats

__global__ void __launch_bounds__(kTailThreads) tokens_tail_kernel(TailArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TLayout& L = a.L;
  const TLayerOff& OL = L.layer[kLayers - 1];
  float* w = reinterpret_cast<float*>(smem);
  {
    const __nv_bfloat16* wproj = reinterpret_cast<const __nv_bfloat16*>(a.blob + OL.wproj);
    const __nv_bfloat16* wfc1 = reinterpret_cast<const __nv_bfloat16*>(a.blob + OL.wfc1);
    const __nv_bfloat16* wfc2 = reinterpret_cast<const __nv_bfloat16*>(a.blob + OL.wfc2);
    for (int i = threadIdx.x; i < kD * kD; i += blockDim.x) w[T_WPROJ + i] = __bfloat162float(wproj[(i >> 5) * kLdD + (i & 31)]);
    for (int i = threadIdx.x; i < kHidden * kD; i += blockDim.x) w[T_WFC1 + i] = __bfloat162float(wfc1[(i >> 5) * kLdD + (i & 31)]);
    for (int i = threadIdx.x; i < kHidden * kD; i += blockDim.x)     // transposed: [hidden j][out n]
      w[T_WFC2T + i] = __bfloat162float(wfc2[(i & 31) * kLdHid + (i >> 5)]);
  }
  __syncthreads();
  const float* fb = reinterpret_cast<const float*>(a.blob);      // fp32 vectors: warp-uniform read-only loads
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.n_patches) return;
  const float4* rec = reinterpret_cast<const float4*>(a.tail + (long long)b * tc::kTailFloats);
  float x0[kD], y[kD];
  {
    float o[kD], l[kHeads];
#pragma unroll
    for (int q4 = 0; q4 < 9; ++q4) {          // warp 0's partial: 32 o + 4 l
      const float4 v = __ldg(rec + q4);
      if (q4 < 8) { o[4 * q4] = v.x; o[4 * q4 + 1] = v.y; o[4 * q4 + 2] = v.z; o[4 * q4 + 3] = v.w; }
      else { l[0] = v.x; l[1] = v.y; l[2] = v.z; l[3] = v.w; }
    }
#pragma unroll
    for (int wq = 1; wq < 4; ++wq)
#pragma unroll
      for (int q4 = 0; q4 < 9; ++q4) {
        const float4 v = __ldg(rec + wq * 9 + q4);
        if (q4 < 8) { o[4 * q4] += v.x; o[4 * q4 + 1] += v.y; o[4 * q4 + 2] += v.z; o[4 * q4 + 3] += v.w; }
        else { l[0] += v.x; l[1] += v.y; l[2] += v.z; l[3] += v.w; }
      }
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const float4 v = __ldg(rec + 36 + q4);
      x0[4 * q4] = v.x; x0[4 * q4 + 1] = v.y; x0[4 * q4 + 2] = v.z; x0[4 * q4 + 3] = v.w;
    }
#pragma unroll
    for (int d = 0; d < kD; ++d) y[d] = o[d] / l[d >> 3];       // attention output of the cls query
  }
  // proj (+bias, +residual)
#pragma unroll 4
  for (int n = 0; n < kD; ++n) {
    float acc = __ldg(fb + OL.bproj / 4 + n);
    const float4* wr = reinterpret_cast<const float4*>(w + T_WPROJ + n * kD);
#pragma unroll
    for (int k4 = 0; k4 < kD / 4; ++k4) {
      const float4 wv = wr[k4];
      acc = fmaf(wv.x, y[4 * k4], fmaf(wv.y, y[4 * k4 + 1], fmaf(wv.z, y[4 * k4 + 2], fmaf(wv.w, y[4 * k4 + 3], acc))));
    }
    x0[n] += acc;
  }
  auto layer_norm = [&](const float (&x)[kD], float (&out)[kD], int g_off, int b_off) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < kD; ++d) s += x[d];
    const float mean = s * (1.f / kD);
    float v = 0.f;
#pragma unroll
    for (int d = 0; d < kD; ++d) { const float dd = x[d] - mean; v = fmaf(dd, dd, v); }
    const float rs = rsqrtf(v * (1.f / kD) + 1e-6f);
#pragma unroll
    for (int d = 0; d < kD; ++d) out[d] = (x[d] - mean) * rs * __ldg(fb + g_off / 4 + d) + __ldg(fb + b_off / 4 + d);
  };
  layer_norm(x0, y, OL.ln2_g, OL.ln2_b);
  // fc1 + GELU + fc2 (+bias, +residual), one hidden unit at a time: 32 FMAs in, 32 FMAs out
  float z[kD];
#pragma unroll
  for (int n = 0; n < kD; ++n) z[n] = __ldg(fb + OL.bfc2 / 4 + n);
#pragma unroll 2
  for (int j = 0; j < kHidden; ++j) {
    float h0 = __ldg(fb + OL.bfc1 / 4 + j), h1 = 0.f;
    const float4* w1 = reinterpret_cast<const float4*>(w + T_WFC1 + j * kD);
#pragma unroll
    for (int k4 = 0; k4 < kD / 4; k4 += 2) {
      const float4 wa = w1[k4], wb = w1[k4 + 1];
      h0 = fmaf(wa.x, y[4 * k4], fmaf(wa.y, y[4 * k4 + 1], fmaf(wa.z, y[4 * k4 + 2], fmaf(wa.w, y[4 * k4 + 3], h0))));
      h1 = fmaf(wb.x, y[4 * k4 + 4], fmaf(wb.y, y[4 * k4 + 5], fmaf(wb.z, y[4 * k4 + 6], fmaf(wb.w, y[4 * k4 + 7], h1))));
    }
    const float hv = gelu_tanh_approx(h0 + h1);
    const float4* w2 = reinterpret_cast<const float4*>(w + T_WFC2T + j * kD);
#pragma unroll
    for (int n4 = 0; n4 < kD / 4; ++n4) {
      const float4 wv = w2[n4];
      z[4 * n4] = fmaf(wv.x, hv, z[4 * n4]);
      z[4 * n4 + 1] = fmaf(wv.y, hv, z[4 * n4 + 1]);
      z[4 * n4 + 2] = fmaf(wv.z, hv, z[4 * n4 + 2]);
      z[4 * n4 + 3] = fmaf(wv.w, hv, z[4 * n4 + 3]);
    }
  }
#pragma unroll
  for (int n = 0; n < kD; ++n) x0[n] += z[n];
  layer_norm(x0, y, L.lnf_g, L.lnf_b);
  // head + first-maximum argmax (like np.argmax)
  const long long orow = a.out_index ? a.out_index[b] : (long long)b;
  int best = 0;
  float bv = -INFINITY;
  for (int k = 0; k < a.K; ++k) {
    float acc = __ldg(fb + L.bhead / 4 + k);
    const float4* wh = reinterpret_cast<const float4*>(fb + L.whead / 4 + k * kD);
#pragma unroll
    for (int d4 = 0; d4 < kD / 4; ++d4) {
      const float4 wv = __ldg(wh + d4);
      acc = fmaf(wv.x, y[4 * d4], fmaf(wv.y, y[4 * d4 + 1], fmaf(wv.z, y[4 * d4 + 2], fmaf(wv.w, y[4 * d4 + 3], acc))));
    }
    a.logits[orow * a.K + k] = acc;
    if (acc > bv) { bv = acc; best = k; }
  }
  if (a.argmax_map) a.argmax_map[orow] = (unsigned char)best;
}

// staging area of tokens_tm_main_launch, then one cls record per patch
size_t tokens_tc_scratch_bytes(int n_patches) { return tokens_tm_stage_bytes() + (size_t)n_patches * tc::kTailFloats * 4; }

bool tokens_tc_supported(int P, int K) { return P >= 1 && P * P + 1 <= 128 && K >= 1 && K <= 64; }

int tokens_tc_launch(const void* f_sps, const void* tparams, int n_patches, int P, int K, float* logits,
                     const long long* out_index, unsigned char* argmax_map, void* scratch, const TcPlanes* planes,
                     cudaStream_t stream) {
  if (n_patches <= 0 || !scratch || !tokens_tc_supported(P, K)) return VC_ERR_ARG;
  if (planes && (planes->h || planes->l) && (P < 2 * planes->D + 1 || !planes->xy || !planes->rowterm || !planes->colterm)) return VC_ERR_ARG;
  int dev = 0, max_smem = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if ((int)tc::SMEM_BYTES > max_smem) return VC_ERR_UNSUPPORTED;
  TcArgs a;
  a.f = (const __nv_bfloat16*)f_sps;
  a.blob = (const uint8_t*)tparams;
  float* stage = (float*)scratch;
  a.tail = (float*)((uint8_t*)scratch + tokens_tm_stage_bytes());
  a.run_flag = nullptr;
  a.RT = sps_rows(n_patches, P);
  a.n_patches = n_patches;
  a.P = P;
  a.T = P * P + 1;
  a.L = tlayout(P, K);
  if (planes) a.pl = *planes;
  else memset(&a.pl, 0, sizeof(a.pl));
  {
    static const int stagger = [] {
      const char* e = getenv("VITCNN_TC_STAGGER_NS");
      return e ? atoi(e) : 7000;
    }();
    a.stagger_ns = stagger;
  }
  // VITCNN_TC_SPLIT = 0 (default): one thread per token row (tokens_tc_kernel); 1 / 2 / 3: two threads per row with that
  // many patches in flight per CTA (tokens_tc2.cu).  Measured per 131 072 patches at P = 11 on one box
  // (profiles/r02_tokens_variants.txt): one thread per row 10.57 / 7.08 / 5.54 ms with 1 / 2 / 3 patches in flight, two
  // threads per row 9.25 / 6.94 / 5.57 ms -- with three patches in flight both forms end at the same throughput (the
  // shared MUFU, issue and shared-memory pipes, not the length of one thread's program, set the pace), so the simpler
  // kernel stays the default.
  static const int split = [] {
    const char* e = getenv("VITCNN_TC_SPLIT");
    return e ? atoi(e) : 0;
  }();
  // VITCNN_TC_KERNEL = tm4 (default) / tm3: tokens_tm_kernel with four / three patches in flight per CTA, and
  // tokens_tc_kernel behind it gated on the prep kernel's flag (weights whose attention logits need the row maximum);
  // tc: tokens_tc_kernel only.
  static const int tm_slots = [] {
    const char* e = getenv("VITCNN_TC_KERNEL");
    if (!e || !strcmp(e, "tm4")) return 4;
    if (!strcmp(e, "tm3")) return 3;
    return 0;
  }();
  if (split > 0) {
    const int rc = tokens_tc2_main_launch(a, n_patches, num_sms, split, stream);
    if (rc != VC_OK) return rc;
  } else {
    if (tm_slots > 0) {
      const int rc = tokens_tm_main_launch(a, stage, n_patches, num_sms, max_smem, tm_slots, stream);
      if (rc != VC_OK) return rc;
      a.run_flag = stage + kTmFlagIndex;
    }
    if (cudaFuncSetAttribute(tokens_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES) != cudaSuccess)
      return VC_ERR_CUDA;
    int blocks = (n_patches + tc::kSlots - 1) / tc::kSlots;
    if (blocks > num_sms) blocks = num_sms;
    tokens_tc_kernel<<<blocks, tc::kThreads, tc::SMEM_BYTES, stream>>>(a);
    if (cudaGetLastError() != cudaSuccess) return VC_ERR_CUDA;
  }

  TailArgs t;
  t.blob = a.blob;
  t.tail = a.tail;
  t.logits = logits;
  t.out_index = out_index;
  t.argmax_map = argmax_map;
  t.n_patches = n_patches;
  t.K = K;
  t.L = a.L;
  const size_t tail_smem = (size_t)T_END * 4;
  if (cudaFuncSetAttribute(tokens_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem) != cudaSuccess)
    return VC_ERR_CUDA;
  tokens_tail_kernel<<<(n_patches + kTailThreads - 1) / kTailThreads, kTailThreads, tail_smem, stream>>>(t);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

}  // namespace vc
