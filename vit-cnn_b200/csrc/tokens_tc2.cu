// Token stage on tcgen05, TWO threads per token row.
//
// Same function, shared-memory / TMEM maps, MMA-issuer protocol and cls-tail kernel as tokens_tc.cu (whose header
// describes the layout tricks); what changes is how the row work is spread over threads.  tokens_tc_kernel gives one
// thread a whole token row (32 residual columns, 128 scores per head, 128 hidden units): ~4 150 instructions per patch
// issued in order by ONE warp per 32 rows, so a patch is a ~23 000-cycle dependency chain even with the SM to itself
// (profiles/r02_tokens_stalls.txt: three slots give 1.90x of one, the softmax phases saturate the MUFU with all slots in
// them together, everything else waits on latencies), and TMEM (512 columns) / shared memory cap the slots at three.
// Here warps w and w + 4 of a slot own the same 32 TMEM lanes (a warp may touch lanes 32 (w % 4) .. +31 of every
// column) and split the COLUMNS of every accumulator: each thread carries 16 residual columns, exponentiates 64 of the
// 128 scores of a head, applies GELU to 64 hidden units, handles two of the four heads of the cls attention.  The
// per-patch chain halves and the SM holds twice the warps on the same buffers.  The only exchange between the two
// threads of a row is LayerNorm's statistics (local mean / M2 of 16 columns, merged Chan-style through 2 KB of the dead
// P buffer and a 64-thread named barrier); the softmax needs none (no row maximum when the static bound on |q.k| holds,
// denominators from the ones column of PV); with the exact softmax each thread scans all the scores for the maximum.
#include "tokens_tc_common.cuh"

namespace vc {

namespace tc2 {
constexpr int kRowWarps = 8;                                  // per slot: warps 0-3 columns' first half, 4-7 second half
__host__ __device__ constexpr int threads(int slots) { return (kRowWarps * slots + slots) * 32; }
}  // namespace tc2

// Development aid: -DVC_TC_TRACE makes thread (block 0, slot 0, row 0, first half) stamp clock64() at every phase
// boundary of its 6th patch and print the deltas (tools/build_variants.py + tools/time_tokens.py).
#ifdef VC_TC_TRACE
#define TC_STAMP(k) do { if (tracing) ts[k] = clock64(); } while (0)
#else
#define TC_STAMP(k) do { } while (0)
#endif

template <int SLOTS>
__global__ void __launch_bounds__(tc2::threads(SLOTS), 1) tokens_tc2_kernel(TcArgs a) {
  using namespace tc;
  static_assert(SLOTS >= 1 && SLOTS <= kSlots && SLOTS <= 3, "shared-memory map and named barriers are laid out for <= 3 slots");
  constexpr int kThreads2 = tc2::threads(SLOTS);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp >= tc2::kRowWarps * SLOTS;
  const int slot = issuer ? warp - tc2::kRowWarps * SLOTS : warp / tc2::kRowWarps;
  const int wq = warp & 3;                                   // TMEM lane quarter of this warp (8 row warps per slot: 8 s + k)
  const int hsel = (warp >> 2) & 1;                          // which half of the columns this thread owns
  const int r = wq * 32 + lane;                              // token row = TMEM lane
  const int T = a.T, P = a.P;
  const uint32_t sb = smem_u32(smem);
  float* vecf = reinterpret_cast<float*>(smem + VEC);
  float* q0_s = reinterpret_cast<float*>(smem + MISC + M_Q0) + slot * 32;          // [32]
  float* wmax_s = reinterpret_cast<float*>(smem + MISC + M_WMAX) + slot * 16;      // [4 quarters][4 heads]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MISC + M_BARS) + slot * 5;
  uint64_t* b_rp = bars + 0;      // row threads: operands of the next GEMM are written (256 arrivals)
  uint64_t* b_rs = bars + 1;      // row threads: S_h has been read out of TMEM (256 arrivals)
  uint64_t* b_mma = bars + 2;     // tensor core: the GEMM just issued is done
  uint64_t* b_s = bars + 3;       // tensor core: S_h is in TMEM
  uint64_t* b_pv = bars + 4;      // tensor core: PV_h is done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + MISC + M_TMEM);
  const float qscale = 0.35355339059327376220f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)

  bool exact_softmax, exact_cls;
  tc_setup(a, smem, tid, kThreads2, 256, exact_softmax, exact_cls);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tb = tmem_base + (uint32_t)slot * C_SLOT;
  const uint32_t tl = tb + ((uint32_t)(wq * 32) << 16);
  const uint32_t slot_s = sb + SLOT0 + (uint32_t)slot * SLOT_BYTES;
  const uint32_t fbuf = slot_s + S_FBUF, abuf = slot_s + S_ABUF, qbuf = slot_s + S_QBUF, kbuf = slot_s + S_KBUF,
                 vbuf = slot_s + S_VBUF, pbuf = slot_s + S_PBUF;
  const int nslots = SLOTS * (int)gridDim.x;
  const int b0 = SLOTS * (int)blockIdx.x + slot;

  if (issuer) {
    tc_issuer(a, sb, tb, slot_s, bars, b0, nslots);
  } else {
    // ============================ row threads ============================
    const uint32_t row16 = (uint32_t)r * 16u;
    const int ct = (T - 1) >> 5;                         // last 32-key chunk holding a real key
    const int PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P);
    const int pair_bar = 1 + slot * 4 + wq;              // the two warps that share these 32 rows (64 threads)
    const int slot_bar = 13 + slot;                      // all row threads of the slot (256)
    float2* xch = reinterpret_cast<float2*>(smem + SLOT0 + (size_t)slot * SLOT_BYTES + S_PBUF);   // [2][128] (mean, M2): P buffer is dead at every LayerNorm
    uint32_t ph_m = 0, ph_s = 0, ph_pv = 0;

    // stem outputs of token row r of patch b -> FBUF: this thread brings 4 of the 8 slices (hsel 0: HSI, 1: LiDAR)
    const int tok_i = r >= 1 && r < T ? (r - 1) / P : 0, tok_j = r >= 1 && r < T ? (r - 1) - tok_i * P : 0;
    const __nv_bfloat16* plane = hsel ? a.pl.l : a.pl.h;
    const long long voff0 = plane ? ((long long)(border_class(tok_i, P, a.pl.D) * (2 * a.pl.D + 1) + border_class(tok_j, P, a.pl.D)) * 4 * a.pl.RTb +
                                     sps_halo(a.pl.B)) : 0;
    auto fetch = [&](int b) {
      if (r >= 1 && r < T) {
        const __nv_bfloat16* src = a.f + ((long long)(4 * hsel) * a.RT + HALO + (long long)b * PP + tok_i * PW + tok_j) * 8;
        long long pitch = a.RT * 8;
        if (plane) {
          const int2 c = __ldg(reinterpret_cast<const int2*>(a.pl.xy) + b);
          src = plane + (voff0 + __ldg(a.pl.rowterm + c.x + tok_i) + __ldg(a.pl.colterm + c.y + tok_j)) * 8;
          pitch = a.pl.RTb * 8;
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) cp_async16(fbuf + (4 * hsel + s) * SLAB + row16, src + s * pitch);
      } else {          // the buffer doubles as K / V: the cls row and the padding rows are zeroed every time
#pragma unroll
        for (int s = 0; s < 4; ++s) sts128(fbuf + (4 * hsel + s) * SLAB + row16, 0u, 0u, 0u, 0u);
      }
    };
    auto publish = [&]() {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(b_rp);
    };
    auto wait_mma = [&]() {
      mbar_wait(b_mma, ph_m);
      ph_m ^= 1u;
      tc_fence_after();
    };
    // LayerNorm (eps 1e-6) of the row whose columns 16 hsel .. +15 this thread holds -> bf16 -> slabs 2 hsel, 2 hsel + 1
    // of the K-major A operand.  Each thread reduces its own 16 columns (two-pass, exact), the pair merges
    // (mean, M2) of the halves: mean = (m0 + m1) / 2, M2 = M2_0 + M2_1 + 8 (m0 - m1)^2.
    auto ln_store = [&](const float (&x)[16], int vg, int vb) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) s += x[c];
      const float m = s * (1.f / 16.f);
      float M2 = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) { const float d = x[c] - m; M2 = fmaf(d, d, M2); }
      xch[hsel * 128 + r] = make_float2(m, M2);
      bar_sync(pair_bar, 64);
      const float2 o = xch[(hsel ^ 1) * 128 + r];
      const float mean = 0.5f * (m + o.x), dm = m - o.x;
      const float rs = rsqrtf((M2 + o.y + 8.f * dm * dm) * (1.f / 32.f) + 1e-6f);
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const uint32_t go = sb + VEC + (uint32_t)(vg + 16 * hsel + 8 * sl) * 4, bo = sb + VEC + (uint32_t)(vb + 16 * hsel + 8 * sl) * 4;
        const float4 g0 = lds_f4(go), g1 = lds_f4(go + 16), c0 = lds_f4(bo), c1 = lds_f4(bo + 16);
        const float* xx = x + 8 * sl;
        sts128(abuf + (uint32_t)(2 * hsel + sl) * SLAB + row16,
               pack_bf16(fmaf((xx[0] - mean) * rs, g0.x, c0.x), fmaf((xx[1] - mean) * rs, g0.y, c0.y)),
               pack_bf16(fmaf((xx[2] - mean) * rs, g0.z, c0.z), fmaf((xx[3] - mean) * rs, g0.w, c0.w)),
               pack_bf16(fmaf((xx[4] - mean) * rs, g1.x, c1.x), fmaf((xx[5] - mean) * rs, g1.y, c1.y)),
               pack_bf16(fmaf((xx[6] - mean) * rs, g1.z, c1.z), fmaf((xx[7] - mean) * rs, g1.w, c1.w)));
      }
    };
    // x[0..15] += accumulator columns 16 hsel .. +15 of the 32-column buffer + bias
    auto residual_add = [&](float (&x)[16], int vbias) {
      uint32_t v[16];
      tmem_ld16(tl + C_SMALL + 16 * hsel, v);
      tc_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 bb = lds_f4(sb + VEC + (uint32_t)(vbias + 16 * hsel + 4 * g) * 4);
        x[4 * g + 0] += __uint_as_float(v[4 * g + 0]) + bb.x;
        x[4 * g + 1] += __uint_as_float(v[4 * g + 1]) + bb.y;
        x[4 * g + 2] += __uint_as_float(v[4 * g + 2]) + bb.z;
        x[4 * g + 3] += __uint_as_float(v[4 * g + 3]) + bb.w;
      }
    };

    if (b0 < a.n_patches) fetch(b0);
    if (slot > 0 && a.stagger_ns > 0) {     // see tokens_tc_kernel
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      } while (t1 - t0 < (unsigned long long)a.stagger_ns * slot);
    }

#ifdef VC_TC_TRACE
    long long ts[32];
    int iter = 0;
#endif
    for (int b = b0; b < a.n_patches; b += nslots) {
#ifdef VC_TC_TRACE
      const bool tracing = blockIdx.x == 0 && slot == 0 && r == 0 && hsel == 0 && iter == 5;
      ++iter;
#endif
      float x[16];   // residual stream of token row r, columns 16 hsel .. +15
      // ================= fusion 1x1 conv (64 -> 32) + folded BN + ReLU, + cls / pos =================
      TC_STAMP(0);
      cp_async_wait_all();
      publish();
      TC_STAMP(1);
      wait_mma();
      TC_STAMP(2);
      {
        uint32_t v[16];
        tmem_ld16(tl + C_SMALL + 16 * hsel, v);
        tc_wait_ld();
        const float rowmask = (r >= 1 && r < T) ? 1.f : 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int gg = 4 * hsel + g;
          const float4 p = lds_f4(sb + POS + r * 128 + ((gg ^ (r & 7)) << 4));
          const float4 sc = lds_f4(sb + VEC + (uint32_t)(V_FSC + 4 * gg) * 4), bi = lds_f4(sb + VEC + (uint32_t)(V_FBI + 4 * gg) * 4);
          x[4 * g + 0] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 0]), sc.x, bi.x), 0.f), rowmask, p.x);
          x[4 * g + 1] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 1]), sc.y, bi.y), 0.f), rowmask, p.y);
          x[4 * g + 2] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 2]), sc.z, bi.z), 0.f), rowmask, p.z);
          x[4 * g + 3] = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * g + 3]), sc.w, bi.w), 0.f), rowmask, p.w);
        }
      }

      // ================= block 1: LN1 -> qkv =================
      TC_STAMP(3);
      ln_store(x, V_LN1G, V_LN1B);
      publish();
      TC_STAMP(4);
      wait_mma();
      TC_STAMP(5);
      {
        // 96 accumulator columns = six 16-column groups (q q k k v v); this thread takes groups 3 hsel .. +2
        uint32_t v[2][16];
        tmem_ld16(tl + C_S + 16 * (3 * hsel), v[0]);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int cg = 3 * hsel + t;
          tc_wait_ld();
          if (t < 2) tmem_ld16(tl + C_S + 16 * (cg + 1), v[(t + 1) & 1]);
          const uint32_t dst = (cg < 2 ? qbuf : cg < 4 ? kbuf : vbuf) + (uint32_t)(2 * (cg & 1)) * SLAB + row16;
          const float sc = cg < 2 ? qscale : 1.f;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t bo = sb + VEC + (uint32_t)(V_BQKV + 16 * cg + 8 * hh) * 4;
            const float4 b0_ = lds_f4(bo), b1_ = lds_f4(bo + 16);
            const uint32_t* vv = v[t & 1] + 8 * hh;
            sts128(dst + hh * SLAB, pack_bf16(fmaf(__uint_as_float(vv[0]), sc, b0_.x), fmaf(__uint_as_float(vv[1]), sc, b0_.y)),
                   pack_bf16(fmaf(__uint_as_float(vv[2]), sc, b0_.z), fmaf(__uint_as_float(vv[3]), sc, b0_.w)),
                   pack_bf16(fmaf(__uint_as_float(vv[4]), sc, b1_.x), fmaf(__uint_as_float(vv[5]), sc, b1_.y)),
                   pack_bf16(fmaf(__uint_as_float(vv[6]), sc, b1_.z), fmaf(__uint_as_float(vv[7]), sc, b1_.w)));
          }
        }
      }
      publish();
      TC_STAMP(6);

      // ================= attention, one head at a time =================
      // this thread exponentiates the 32-key chunks 2 hsel, 2 hsel + 1 of S_h (those that hold real keys); O_{h-1} is
      // normalised and stored by the thread with hsel == (h - 1) & 1
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
        TC_STAMP(7 + 3 * h);
        float m = 0.f;
        if constexpr (kExact) {          // every thread scans the whole row: no exchange with its partner
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
          if (((h - 1) & 1) == hsel) read_o(h - 1);
        }
        TC_STAMP(8 + 3 * h);
        const int c_last = ct < 2 * hsel + 1 ? ct : 2 * hsel + 1;      // last chunk this thread reads (none if < 2 hsel)
        if (c_last < 2 * hsel && h < 3) {
          tc_fence_before();
          mbar_arrive(b_rs);
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * hsel + cc;
          if (c <= ct) {
            tmem_ld32(tl + C_S + 32 * c, sc);
            tc_wait_ld();
            if (c == c_last && h < 3) {       // this thread's part of S_h is out of TMEM
              tc_fence_before();
              mbar_arrive(b_rs);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float s0 = __uint_as_float(sc[8 * g + 2 * e]), s1 = __uint_as_float(sc[8 * g + 2 * e + 1]);
                if constexpr (kExact) pk[e] = pack_bf16(ex2(s0 - m), ex2(s1 - m));
                else pk[e] = pack_bf16(ex2(s0), ex2(s1));
              }
              sts128(pbuf + (uint32_t)(4 * c + g) * SLAB + row16, pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        publish();
        TC_STAMP(9 + 3 * h);
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
      TC_STAMP(19);
      if (b + nslots < a.n_patches) fetch(b + nslots);   // K and V are dead: the next patch's fusion input lands over them
      if (hsel == 1) read_o(3);
      publish();
      TC_STAMP(20);
      wait_mma();
      TC_STAMP(21);
      residual_add(x, V_BPROJ);

      // ================= MLP: LN2 -> fc1 (+bias, GELU) -> fc2 (+bias, +residual) =================
      ln_store(x, V_LN2G, V_LN2B);
      publish();
      TC_STAMP(22);
      wait_mma();
      TC_STAMP(23);
      {
        // 128 hidden units = eight 16-column groups; this thread takes groups 4 hsel .. +3
        uint32_t v[2][16];
        tmem_ld16(tl + C_S + 64 * hsel, v[0]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int cg = 4 * hsel + t;
          tc_wait_ld();
          if (t < 3) tmem_ld16(tl + C_S + 16 * (cg + 1), v[(t + 1) & 1]);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t bo = sb + VEC + (uint32_t)(V_BFC1 + 16 * cg + 8 * hh) * 4;
            const float4 b0_ = lds_f4(bo), b1_ = lds_f4(bo + 16);
            const uint32_t* vv = v[t & 1] + 8 * hh;
            sts128(pbuf + (uint32_t)(2 * cg + hh) * SLAB + row16,
                   pack_bf16(gelu2(__uint_as_float(vv[0]) + b0_.x), gelu2(__uint_as_float(vv[1]) + b0_.y)),
                   pack_bf16(gelu2(__uint_as_float(vv[2]) + b0_.z), gelu2(__uint_as_float(vv[3]) + b0_.w)),
                   pack_bf16(gelu2(__uint_as_float(vv[4]) + b1_.x), gelu2(__uint_as_float(vv[5]) + b1_.y)),
                   pack_bf16(gelu2(__uint_as_float(vv[6]) + b1_.z), gelu2(__uint_as_float(vv[7]) + b1_.w)));
          }
        }
      }
      publish();
      TC_STAMP(24);
      wait_mma();
      TC_STAMP(25);
      residual_add(x, V_BFC2);

      // ================= last block: K / V of every token, attention of the cls query only =================
      ln_store(x, V_L2G, V_L2B);
      publish();
      TC_STAMP(26);
      wait_mma();
      TC_STAMP(27);
      float* trec = a.tail + (long long)b * kTailFloats;
      {
        // this thread: heads 2 hsel, 2 hsel + 1 = columns 16 hsel .. +15 of q (cls row only), k and v
        uint32_t kk[16], vv[16];
        tmem_ld16(tl + C_S + 32 + 16 * hsel, kk);
        tmem_ld16(tl + C_S + 64 + 16 * hsel, vv);
        if (wq == 0) {     // the cls token is row 0: its (scaled) query and its residual stream
          uint32_t qq[16];
          tmem_ld16(tl + C_S + 16 * hsel, qq);
          tc_wait_ld();
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              q0_s[16 * hsel + c] = fmaf(__uint_as_float(qq[c]), qscale, vecf[V_BQKV2 + 16 * hsel + c]);
              trec[144 + 16 * hsel + c] = x[c];
            }
          }
        }
        tc_wait_ld();
        bar_sync(slot_bar, 256);
        float sc[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int h = 2 * hsel + hh;
          const float4 q0 = lds_f4(smem_u32(q0_s) + 32 * h), q1 = lds_f4(smem_u32(q0_s) + 32 * h + 16);
          const float4 k0 = lds_f4(sb + VEC + (uint32_t)(V_BQKV2 + 32 + 8 * h) * 4), k1 = lds_f4(sb + VEC + (uint32_t)(V_BQKV2 + 36 + 8 * h) * 4);
          float d = q0.x * (__uint_as_float(kk[8 * hh + 0]) + k0.x);
          d = fmaf(q0.y, __uint_as_float(kk[8 * hh + 1]) + k0.y, d);
          d = fmaf(q0.z, __uint_as_float(kk[8 * hh + 2]) + k0.z, d);
          d = fmaf(q0.w, __uint_as_float(kk[8 * hh + 3]) + k0.w, d);
          d = fmaf(q1.x, __uint_as_float(kk[8 * hh + 4]) + k1.x, d);
          d = fmaf(q1.y, __uint_as_float(kk[8 * hh + 5]) + k1.y, d);
          d = fmaf(q1.z, __uint_as_float(kk[8 * hh + 6]) + k1.z, d);
          d = fmaf(q1.w, __uint_as_float(kk[8 * hh + 7]) + k1.w, d);
          sc[hh] = r < T ? d : -INFINITY;
          if (exact_cls) {           // the maximum over all keys is only needed when 2^s could overflow
            float mw = sc[hh];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
            if (lane == 0) wmax_s[wq * 4 + h] = mw;
          }
        }
        if (exact_cls) bar_sync(slot_bar, 256);
        float val[16], pl[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int h = 2 * hsel + hh;
          const float m = exact_cls ? fmaxf(fmaxf(wmax_s[h], wmax_s[4 + h]), fmaxf(wmax_s[8 + h], wmax_s[12 + h])) : 0.f;
          const float p = ex2(sc[hh] - m);
          pl[hh] = p;
          const float4 v0 = lds_f4(sb + VEC + (uint32_t)(V_BQKV2 + 64 + 8 * h) * 4), v1 = lds_f4(sb + VEC + (uint32_t)(V_BQKV2 + 68 + 8 * h) * 4);
          val[8 * hh + 0] = p * (__uint_as_float(vv[8 * hh + 0]) + v0.x);
          val[8 * hh + 1] = p * (__uint_as_float(vv[8 * hh + 1]) + v0.y);
          val[8 * hh + 2] = p * (__uint_as_float(vv[8 * hh + 2]) + v0.z);
          val[8 * hh + 3] = p * (__uint_as_float(vv[8 * hh + 3]) + v0.w);
          val[8 * hh + 4] = p * (__uint_as_float(vv[8 * hh + 4]) + v1.x);
          val[8 * hh + 5] = p * (__uint_as_float(vv[8 * hh + 5]) + v1.y);
          val[8 * hh + 6] = p * (__uint_as_float(vv[8 * hh + 6]) + v1.z);
          val[8 * hh + 7] = p * (__uint_as_float(vv[8 * hh + 7]) + v1.w);
        }
        // butterfly reduction over the 32 rows of this warp: four halving steps leave lane l with element l >> 1
        // summed over 16 lanes, the last step adds the neighbour lane
#pragma unroll
        for (int off = 16, n = 16; off >= 2; off >>= 1, n >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float send = up ? val[i] : val[i + n / 2];
            const float keep = up ? val[i + n / 2] : val[i];
            val[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        val[0] += __shfl_xor_sync(0xffffffffu, val[0], 1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          pl[0] += __shfl_xor_sync(0xffffffffu, pl[0], o);
          pl[1] += __shfl_xor_sync(0xffffffffu, pl[1], o);
        }
        if ((lane & 1) == 0) trec[wq * 36 + 16 * hsel + (lane >> 1)] = val[0];
        if (lane < 2) trec[wq * 36 + 32 + 2 * hsel + lane] = lane ? pl[1] : pl[0];
      }
      TC_STAMP(28);
#ifdef VC_TC_TRACE
      if (tracing) {
        const char* names[28] = {"cp.async wait+publish", "wait fusion MMA", "fusion epilogue", "LN1+publish", "wait qkv MMA", "qkv epilogue+publish",
                                 "wait S0", "h0 pre", "h0 exp+publish", "wait S1", "h1 wait PV0+read_o", "h1 exp+publish", "wait S2", "h2 wait PV1+read_o",
                                 "h2 exp+publish", "wait S3", "h3 wait PV2+read_o", "h3 exp+publish", "wait PV3", "fetch+read_o+publish",
                                 "wait proj MMA", "proj epi+LN2+publish", "wait fc1 MMA", "GELU+publish", "wait fc2 MMA", "fc2 epi+LN3+publish",
                                 "wait kv2 MMA", "cls attention"};
        for (int k = 0; k < 28; ++k) printf("trace %2d %-28s %6lld\n", k, names[k], ts[k + 1] - ts[k]);
        printf("trace total %lld\n", ts[28] - ts[0]);
      }
#endif
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

template <int SLOTS>
static int launch_split(const TcArgs& a, int n_patches, int num_sms, cudaStream_t stream) {
  if (cudaFuncSetAttribute(tokens_tc2_kernel<SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES) != cudaSuccess)
    return VC_ERR_CUDA;
  int blocks = (n_patches + SLOTS - 1) / SLOTS;
  if (blocks > num_sms) blocks = num_sms;
  tokens_tc2_kernel<SLOTS><<<blocks, tc2::threads(SLOTS), tc::SMEM_BYTES, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// main kernel of the split variant; the caller (tokens_tc.cu) fills the arguments and runs the shared cls-tail kernel
int tokens_tc2_main_launch(const TcArgs& a, int n_patches, int num_sms, int slots, cudaStream_t stream) {
  if (slots == 3 && tc::kSlots >= 3) return launch_split<(tc::kSlots >= 3 ? 3 : 1)>(a, n_patches, num_sms, stream);
  if (slots == 2 && tc::kSlots >= 2) return launch_split<(tc::kSlots >= 2 ? 2 : 1)>(a, n_patches, num_sms, stream);
  return launch_split<1>(a, n_patches, num_sms, stream);
}

}  // namespace vc
