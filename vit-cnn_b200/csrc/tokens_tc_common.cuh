// Shared pieces of the tcgen05 token-stage kernels (tokens_tc.cu: one thread per token row; tokens_tc2.cu: two
// threads per token row): shared-memory / TMEM maps, argument blocks, small device helpers.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "vc_common.cuh"
#include "vc_kernels.h"
#include "vc_tparams.h"
#include "vc_tokens.cuh"

namespace vc {

namespace tc {
#ifndef VC_TC_SLOTS
#define VC_TC_SLOTS 3
#endif
constexpr int kSlots = VC_TC_SLOTS;       // patches in flight per CTA
constexpr int kThreads = (4 * kSlots + kSlots) * 32;   // 4 warps of row threads per slot + one MMA-issuer warp per slot
constexpr uint32_t SLAB = 2048;   // 128 rows x 16 B: one 8-element K chunk (or 8-dim V group) of a tile
// ---- shared-memory map (bytes) ----
constexpr uint32_t W_FUS = 0;                     // [64/8][32][8]
constexpr uint32_t W_QKV1 = W_FUS + 4096;         // [32/8][96][8]
constexpr uint32_t W_PROJ1 = W_QKV1 + 6144;       // [32/8][32][8]
constexpr uint32_t W_FC1 = W_PROJ1 + 2048;        // [32/8][128][8]
constexpr uint32_t W_FC2 = W_FC1 + 8192;          // [128/8][32][8]
constexpr uint32_t W_QKV2 = W_FC2 + 8192;         // [32/8][96][8]
constexpr uint32_t VEC = W_QKV2 + 6144;           // fp32 vectors, see V_* (floats)
constexpr int V_FSC = 0, V_FBI = 32, V_LN1G = 64, V_LN1B = 96, V_BQKV = 128, V_BPROJ = 224, V_LN2G = 256, V_LN2B = 288,
              V_BFC1 = 320, V_BFC2 = 448, V_L2G = 480, V_L2B = 512, V_BQKV2 = 544, V_TOTAL = 640;
constexpr uint32_t POS = VEC + V_TOTAL * 4;       // [128][32] fp32, 16-byte granules XOR-swizzled by (row & 7)
constexpr uint32_t SLOT0 = POS + 128 * 32 * 4;
// 56 KB per slot so that three fit: buffers whose lifetimes do not overlap share storage
constexpr uint32_t S_QBUF = 0;                    // [4 heads][128][8] scaled queries
constexpr uint32_t S_ABUF = S_QBUF;               // [4][128][8] LN output (dead once qkv is done) / attention output (after the last S)
constexpr uint32_t S_KBUF = S_QBUF + 4 * SLAB;
constexpr uint32_t S_VBUF = S_KBUF + 4 * SLAB;    // [4 heads][128 keys][8 dims]
constexpr uint32_t S_PBUF = S_VBUF + 4 * SLAB;    // [16][128][8] probabilities of one head / MLP hidden
constexpr uint32_t S_FBUF = S_KBUF;               // [8][128][8] fusion-conv input of the NEXT patch: over K and V once the last PV is done
constexpr uint32_t SLOT_BYTES = S_PBUF + 16 * SLAB;
constexpr uint32_t ONES = SLOT0 + kSlots * SLOT_BYTES;  // [128][8] = (1,0,0,0,0,0,0,0)
constexpr uint32_t MASK = ONES + SLAB;             // [128 keys][8] = (0 | -30000 for padded keys, 0, ..): K's second K chunk
constexpr uint32_t MISC = MASK + SLAB;            // q0 [slots][32] f32, wmax [slots][4][4] f32, barriers [slots][5], tmem slot
constexpr uint32_t M_Q0 = 0, M_WMAX = M_Q0 + kSlots * 128, M_BARS = M_WMAX + kSlots * 64, M_TMEM = M_BARS + kSlots * 40;
constexpr uint32_t M_LOCK = M_TMEM + 8;            // [4] one word per SM sub-partition: which slot's warp holds its MUFU pipe (VC_TC_MUFU_LOCK)
constexpr uint32_t SMEM_BYTES = MISC + ((M_LOCK + 16 + 127) & ~127u);
// ---- TMEM columns inside a slot's 160: S / qkv / fc1 accumulators at 0, two 16-column O_h buffers at 128 (the
// 32-column fusion / proj / fc2 accumulators reuse them) ----
constexpr uint32_t C_SLOT = 160, C_S = 0, C_O = 128, C_SMALL = 128;
constexpr int kTailFloats = 176;                  // per patch: [4 warps][32 o + 4 l], x0[32]
}  // namespace tc

struct TcArgs {
  const __nv_bfloat16* f;   // [8][RT][8]: slices 0-3 HSI stem, 4-7 LiDAR stem
  TcPlanes pl;              // stem outputs taken straight from the scene-level variant planes instead (h / l non-null)
  const uint8_t* blob;      // parameter blob (vc_tparams.h)
  float* tail;              // [n][kTailFloats]
  long long RT;
  int n_patches, P, T, stagger_ns;
  const float* run_flag;    // nullable: the kernel runs only if *run_flag != 0 (fallback behind tokens_tm_kernel)
  TLayout L;
};

struct TailArgs {
  const uint8_t* blob;
  const float* tail;
  float* logits;
  const long long* out_index;
  unsigned char* argmax_map;
  int n_patches, K;
  TLayout L;
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// instruction descriptor: bf16 x bf16 -> fp32, A K-major, B K-major (b_mn = 0) or MN-major (1)
__device__ __forceinline__ uint32_t idesc(int N, int b_mn) {
  return umma_idesc_bf16(128, N) | ((uint32_t)b_mn << 16);
}

// 2 * GELU(v) in the tanh form of vc_tokens.cuh: v + v tanh(u); the factor 0.5 is folded into W_fc2
__device__ __forceinline__ float gelu2(float v) {
  const float u = v * fmaf(0.0356774081f, v * v, 0.7978845608f);
  return fmaf(v, tanh_fast(u), v);
}

// 2^s on the FMA / ALU pipes (Cody-Waite split + cubic, max relative error 1.9e-4, far below the bf16 rounding of the
// result): the SFU does 16 ex2 per clock and SM and is the busiest unit of the softmax phases, so a fixed share of
// the probabilities (VC_TC_POLY_MASK: which of the 8 elements of a group) can be computed here instead.  Measured per
// 32 768 patches at P = 11: mask 0x00 1.533 ms, 0x88 (a quarter) 1.509 ms, 0xAA (half) 1.583 ms -- the phases are
// latency bound, not SFU-throughput bound, so the split stays off.
#ifndef VC_TC_POLY_MASK
#define VC_TC_POLY_MASK 0x00
#endif
// tuning switches (tools/build_variants.py builds one library per setting and times them on the same box):
// VC_TC_WARP_ARRIVE: one elected mbarrier arrival per row warp instead of one per row thread;
// VC_TC_LD16: the softmax reads S_h in 16-column halves, the load of the next half in flight behind the exponentials.
// Measured per 131 072 patches at P = 11 (same box, profiles/r02_tokens_variants.txt): base 5.535 ms, WARP_ARRIVE 5.610,
// LD16 5.562, both 5.713, POLY_MASK 0x88 5.749, 0xAA 5.876 -- none pays: the softmax phases are bound by MUFU
// throughput with all three slots in them together (stall sampling: profiles/r02_tokens_stalls.txt), the rest by the
// length of the per-patch dependency chain, not by these hand-offs.
#ifndef VC_TC_WARP_ARRIVE
#define VC_TC_WARP_ARRIVE 0
#endif
#ifndef VC_TC_LD16
#define VC_TC_LD16 0
#endif
__device__ __forceinline__ float ex2_poly(float s) {
  const float x = fmaxf(s, -126.f);                 // masked keys sit at -30000
  const float t = x + 12582912.f;                   // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.f);             // in [-0.5, 0.5]
  float p = fmaf(f, 0.05587554f, 0.24229463f);
  p = fmaf(p, f, 0.69312726f);
  p = fmaf(p, f, 0.99994823f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// LayerNorm (eps 1e-6) of the row held by this thread -> bf16 -> K-major A operand (4 slabs)
__device__ __forceinline__ void ln_store(const float (&x)[32], uint32_t vec_g, uint32_t vec_b, uint32_t dst_row) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) s += x[c];
  const float mean = s * (1.f / 32.f);
  float v = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) { const float d = x[c] - mean; v = fmaf(d, d, v); }
  const float rs = rsqrtf(v * (1.f / 32.f) + 1e-6f);
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    const float4 g0 = lds_f4(vec_g + sl * 32), g1 = lds_f4(vec_g + sl * 32 + 16);
    const float4 b0 = lds_f4(vec_b + sl * 32), b1 = lds_f4(vec_b + sl * 32 + 16);
    const float* xx = x + 8 * sl;
    const uint32_t p0 = pack_bf16(fmaf((xx[0] - mean) * rs, g0.x, b0.x), fmaf((xx[1] - mean) * rs, g0.y, b0.y));
    const uint32_t p1 = pack_bf16(fmaf((xx[2] - mean) * rs, g0.z, b0.z), fmaf((xx[3] - mean) * rs, g0.w, b0.w));
    const uint32_t p2 = pack_bf16(fmaf((xx[4] - mean) * rs, g1.x, b1.x), fmaf((xx[5] - mean) * rs, g1.y, b1.y));
    const uint32_t p3 = pack_bf16(fmaf((xx[6] - mean) * rs, g1.z, b1.z), fmaf((xx[7] - mean) * rs, g1.w, b1.w));
    sts128(dst_row + sl * tc::SLAB, p0, p1, p2, p3);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One-time image of the parameters in the layouts the tensor core reads, barriers, TMEM allocation, and the static
// bound on |q.k| of both blocks (exact_softmax / exact_cls: the row maximum is needed).  Called by every thread of the CTA.
__device__ __forceinline__ void tc_setup(const TcArgs& a, uint8_t* smem, int tid, int kThreads, int row_arrivals, bool& exact_softmax_out,
                                         bool& exact_cls_out) {
  using namespace tc;
  const int T = a.T;
  const TLayout& L = a.L;
  float* vecf = reinterpret_cast<float*>(smem + VEC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + MISC + M_TMEM);
  const float qscale = 0.35355339059327376220f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)
  // ---------------- one-time image of the parameters in the layouts the tensor core reads ----------------
  {
    auto copy_w = [&](uint32_t dst, int src, int N, int K, int ld, bool halve = false) {
      for (int i = tid; i < N * (K / 8); i += kThreads) {
        const int n = i % N, kc = i / N;
        uint4 g = __ldg(reinterpret_cast<const uint4*>(a.blob + src + (size_t)(n * ld + kc * 8) * 2));
        if (halve) {   // exact in bf16: one less in the exponent field (weights are far from subnormal)
          __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&g);
          for (int e = 0; e < 4; ++e) hp[e] = __hmul2(hp[e], __floats2bfloat162_rn(0.5f, 0.5f));
        }
        *reinterpret_cast<uint4*>(smem + dst + (size_t)kc * N * 16 + n * 16) = g;
      }
    };
    copy_w(W_FUS, L.wfus, 32, 64, kLdFus);
    copy_w(W_QKV1, L.layer[0].wqkv, 96, 32, kLdD);
    copy_w(W_PROJ1, L.layer[0].wproj, 32, 32, kLdD);
    copy_w(W_FC1, L.layer[0].wfc1, 128, 32, kLdD);
    copy_w(W_FC2, L.layer[0].wfc2, 32, 128, kLdHid, true);   // the 0.5 of GELU lives here: H = 2 gelu(.)
    copy_w(W_QKV2, L.layer[1].wqkv, 96, 32, kLdD);
    auto copy_v = [&](int dst, int src, int n, float scale_first32) {
      for (int i = tid; i < n; i += kThreads) {
        const float v = __ldg(reinterpret_cast<const float*>(a.blob + src) + i);
        vecf[dst + i] = i < 32 ? v * scale_first32 : v;
      }
    };
    copy_v(V_FSC, L.fus_scale, 32, 1.f);
    copy_v(V_FBI, L.fus_bias, 32, 1.f);
    copy_v(V_LN1G, L.layer[0].ln1_g, 32, 1.f);
    copy_v(V_LN1B, L.layer[0].ln1_b, 32, 1.f);
    copy_v(V_BQKV, L.layer[0].bqkv, 96, qscale);      // q bias pre-scaled: q = acc * qscale + b * qscale
    copy_v(V_BPROJ, L.layer[0].bproj, 32, 1.f);
    copy_v(V_LN2G, L.layer[0].ln2_g, 32, 1.f);
    copy_v(V_LN2B, L.layer[0].ln2_b, 32, 1.f);
    copy_v(V_BFC1, L.layer[0].bfc1, 128, 1.f);
    copy_v(V_BFC2, L.layer[0].bfc2, 32, 1.f);
    copy_v(V_L2G, L.layer[1].ln1_g, 32, 1.f);
    copy_v(V_L2B, L.layer[1].ln1_b, 32, 1.f);
    copy_v(V_BQKV2, L.layer[1].bqkv, 96, qscale);
    // pos-embed rows (row 0 = cls + pos[0], rows >= T zero), 16-byte granules swizzled by row
    const float* pos = reinterpret_cast<const float*>(a.blob + L.pos);
    const float* cls = reinterpret_cast<const float*>(a.blob + L.cls);
    for (int i = tid; i < 128 * 8; i += kThreads) {
      const int row = i >> 3, g = i & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < T) {
        v = __ldg(reinterpret_cast<const float4*>(pos + row * 32 + 4 * g));
        if (row == 0) {
          const float4 c = __ldg(reinterpret_cast<const float4*>(cls + 4 * g));
          v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
        }
      }
      *reinterpret_cast<float4*>(smem + POS + row * 128 + ((g ^ (row & 7)) << 4)) = v;
    }
    // slot buffers, ZERO slab: zeros (row 0 and rows >= T of FBUF are never written again); ONES slab
    for (uint32_t i = tid; i < (kSlots * SLOT_BYTES + 2 * SLAB) / 16; i += kThreads)
      *reinterpret_cast<uint4*>(smem + SLOT0 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // static bound on |q.k| of both blocks: row n of Wq / Wk contributes (||W_n diag(g)||^2, (W_n . beta + b_n)^2)
    if (tid < 128) {
      const int l = tid >> 6, n = tid & 63;
      const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(a.blob + L.layer[l].wqkv) + n * kLdD;
      const float* g = reinterpret_cast<const float*>(a.blob + L.layer[l].ln1_g);
      const float* be = reinterpret_cast<const float*>(a.blob + L.layer[l].ln1_b);
      float f2 = 0.f, bs = __ldg(reinterpret_cast<const float*>(a.blob + L.layer[l].bqkv) + n);
      for (int c = 0; c < 32; ++c) {
        const float wv = __bfloat162float(w[c]);
        f2 = fmaf(wv * __ldg(g + c), wv * __ldg(g + c), f2);
        bs = fmaf(wv, __ldg(be + c), bs);
      }
      float* scr = reinterpret_cast<float*>(smem + SLOT0 + S_PBUF);
      scr[2 * tid] = f2;           // [layer][64 rows][2]
      scr[2 * tid + 1] = bs * bs;
    }
    for (int i = tid; i < 128; i += kThreads) {
      *reinterpret_cast<uint32_t*>(smem + ONES + i * 16) = 0x00003F80u;                 // bf16 1.0
      *reinterpret_cast<uint32_t*>(smem + MASK + i * 16) = i >= T ? 0x0000C6EAu : 0u;   // bf16 -29952 for padded keys
    }
    if (tid == 0) {
      for (int s = 0; s < kSlots; ++s) {
        uint64_t* bb = reinterpret_cast<uint64_t*>(smem + MISC + M_BARS) + s * 5;
        mbar_init(bb + 0, row_arrivals);      // one elected arrival per row warp, or one per thread
        mbar_init(bb + 1, row_arrivals);
        mbar_init(bb + 2, 1);
        mbar_init(bb + 3, 1);
        mbar_init(bb + 4, 1);
      }
      fence_mbar_init();
    }
    if (tid < 4) reinterpret_cast<uint32_t*>(smem + MISC + M_LOCK)[tid] = 0u;
    if (tid < 32) {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // |s| <= qscale * (sqrt(32) ||Wq_h diag(g)||_F + ||Wq_h beta + bq_h||) * (same for k): LayerNorm output is
  // g * xhat + beta with ||xhat|| <= sqrt(32).  Below 2^100 neither 2^s nor its row sums leave fp32 / bf16
  // range, so the row maximum is not needed (bf16 rounding of the operands is far inside the margin).
  bool exact_softmax = false, exact_cls = false;   // block 1 (all queries) / last block (cls query only)
  {
    const float* scr0 = reinterpret_cast<const float*>(smem + SLOT0 + S_PBUF);
    for (int l = 0; l < 2; ++l) {
      const float* scr = scr0 + 128 * l;
      for (int h = 0; h < 4; ++h) {
        float qf = 0.f, qb = 0.f, kf = 0.f, kb = 0.f;
        for (int n = 0; n < 8; ++n) {
          qf += scr[2 * (8 * h + n)]; qb += scr[2 * (8 * h + n) + 1];
          kf += scr[2 * (32 + 8 * h + n)]; kb += scr[2 * (32 + 8 * h + n) + 1];
        }
        const float bound = qscale * (sqrtf(32.f * qf) + sqrtf(qb)) * (sqrtf(32.f * kf) + sqrtf(kb));
        if (!(bound < 100.f)) (l == 0 ? exact_softmax : exact_cls) = true;
      }
    }
    __syncthreads();   // the scratch is part of a P buffer
  }
  exact_softmax_out = exact_softmax;
  exact_cls_out = exact_cls;
}

// MMA issuer warp of one slot: waits for the row threads' "operands written" / "S consumed" arrivals, issues the GEMMs of
// the block in program order, commits to the barriers the row threads wait on.
__device__ __forceinline__ void tc_issuer(const TcArgs& a, uint32_t sb, uint32_t tb, uint32_t slot_s, uint64_t* bars, int b0, int nslots) {
  using namespace tc;
  uint64_t *b_rp = bars + 0, *b_rs = bars + 1, *b_mma = bars + 2, *b_s = bars + 3, *b_pv = bars + 4;
  const uint32_t fbuf = slot_s + S_FBUF, abuf = slot_s + S_ABUF, qbuf = slot_s + S_QBUF, kbuf = slot_s + S_KBUF,
                 vbuf = slot_s + S_VBUF, pbuf = slot_s + S_PBUF;
  const int NK = (a.T + 31) & ~31, NKS = NK >> 4;       // keys rounded to the 32-column chunks the row threads read
  {
    // ============================ MMA issuer of this slot ============================
    uint32_t ph_rp = 0, ph_rs = 0;
    auto ready = [&]() {       // the row threads have written the operands of the next GEMM
      mbar_wait(b_rp, ph_rp);
      ph_rp ^= 1u;
      tc_fence_after();
    };
    // S_h = Q_h K_h^T + mask: A chunks (Q_h, ones), B chunks (K_h, mask) -> the second half of K = 16 adds
    // 1 * (-30000) to the columns of padded keys and nothing to the others
    auto issue_s = [&](int h) {
      if (elect_one()) {
        umma_bf16(tb + C_S, umma_desc(qbuf + h * SLAB, (sb + ONES) - (qbuf + h * SLAB), 128),
                  umma_desc(kbuf + h * SLAB, (sb + MASK) - (kbuf + h * SLAB), 128), idesc(NK, 0), 0u);
        umma_commit(b_s);
      }
      __syncwarp();
    };
    // D[128 x N] (TMEM column `col`) = A[128 x 16 ksteps] (K-major slabs at `abase`) . W^T (weights [k/8][N][8] at `wbase`)
    auto issue_gemm = [&](uint32_t col, uint32_t abase, uint32_t wbase, int N, int ksteps) {
      if (elect_one()) {
        for (int k = 0; k < ksteps; ++k)
          umma_bf16(tb + col, umma_desc(abase + 2 * k * SLAB, SLAB, 128), umma_desc(wbase + 2 * k * N * 16, (uint32_t)N * 16u, 128),
                    idesc(N, 0), k ? 1u : 0u);
        umma_commit(b_mma);
      }
      __syncwarp();
    };
    for (int b = b0; b < a.n_patches; b += nslots) {
      ready(); issue_gemm(C_SMALL, fbuf, sb + W_FUS, 32, 4);       // fusion 1x1 conv
      ready(); issue_gemm(C_S, abuf, sb + W_QKV1, 96, 2);          // qkv
      ready(); issue_s(0);
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        if (h < 3) {
          mbar_wait(b_rs, ph_rs);
          ph_rs ^= 1u;
          tc_fence_after();
          issue_s(h + 1);
        }
        ready();
        if (elect_one()) {
          // O_h[128 x 16] = P_h[128 x keys] . [V_h | ones]: B is MN-major, N chunk 0 = V_h slab, chunk 1 = ONES slab
          const uint32_t vb = vbuf + h * SLAB;
          for (int k = 0; k < NKS; ++k)
            umma_bf16(tb + C_O + 16 * (h & 1), umma_desc(pbuf + 2 * k * SLAB, SLAB, 128), umma_desc(vb + k * 256, 128, (sb + ONES) - vb),
                      idesc(16, 1), k ? 1u : 0u);
          umma_commit(b_pv);
        }
        __syncwarp();
      }
      ready(); issue_gemm(C_SMALL, abuf, sb + W_PROJ1, 32, 2);     // proj
      ready(); issue_gemm(C_S, abuf, sb + W_FC1, 128, 2);          // fc1
      ready(); issue_gemm(C_SMALL, pbuf, sb + W_FC2, 32, 8);       // fc2
      ready(); issue_gemm(C_S, abuf, sb + W_QKV2, 96, 2);          // last block: q (cls row), k, v
    }
  }
}

// tokens_tm.cu: probabilities / hidden units through TMEM, per-channel vectors in the constant bank (no-row-maximum softmax
// only: `stage` receives the flag that gates the fallback launch of tokens_tc_kernel)
size_t tokens_tm_stage_bytes();
int tokens_tm_main_launch(const TcArgs& a, float* stage, int n_patches, int num_sms, int max_smem, int slots, cudaStream_t stream);
constexpr int kTmFlagIndex = tc::V_TOTAL;   // stage[kTmFlagIndex] != 0: the softmax needs its row maximum

// tokens_tc2.cu: the main kernel with two threads per token row (`slots` patches in flight per CTA)
int tokens_tc2_main_launch(const TcArgs& a, int n_patches, int num_sms, int slots, cudaStream_t stream);

}  // namespace vc
