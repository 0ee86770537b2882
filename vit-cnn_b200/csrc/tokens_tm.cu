// Token stage on tcgen05 with the attention probabilities and the MLP hidden units handed to the tensor core
// THROUGH TMEM (tcgen05.st + TMEM-sourced A operand) instead of shared memory, and the per-channel vectors
// (folded BN, LayerNorm gamma / beta, biases) read as constant-bank operands instead of shared-memory loads.
//
// Same function, parameter blob, stem-input addressing, cls record and tail kernel as tokens_tc.cu (whose header
// describes the layout tricks: one patch = one M = 128 tile, token row r = TMEM lane r = thread r of a 128-thread
// group, head_dim 8 padded to K = 16 by pointing the second K chunk at a ones / mask slab, softmax denominators
// from a ones column of V, last block pruned to the cls query).  What the ncu record of tokens_tc_kernel showed
// (profiles/r02_tokens_stalls.txt, r02_tokens_variants.txt): MUFU 46 % busy, issue 39 %, and every phase of a slot
// stretched ~1.8x as soon as three slots share an SM -- the softmax phases because they share the MUFU, all the
// others because their shared-memory loads / stores queue behind the other slots' MUFU instructions in the same
// MIO queue; plus one exposed tensor-core round trip per head (the next head's probabilities could not be written
// before PV of the previous head had released the single 32 KB P buffer).  This kernel removes those:
//
//  * S_h is produced in key chunks of CK columns into a ring of RD TMEM buffers (3 slots: 2 x 64 keys, 4 slots:
//    3 x 32 keys).  A row thread reads a chunk, exponentiates, packs to bf16 and stores the probabilities back
//    over the first half of the SAME columns (tcgen05.st); PV takes its A operand from there (K = CK keys per
//    step, accumulating into O_h) and the S chunk RD steps ahead is issued right behind it into the freed buffer,
//    so the row threads never wait for a PV and S chunks are ready before they are needed;
//  * fc1's accumulators are overwritten in place by the packed GELU outputs, fc2 reads them from TMEM;
//    the 32 KB P / hidden buffer of a slot is gone (24 KB of shared memory per slot instead of 56 KB, ~60 % less
//    shared-memory traffic per patch, 80 fewer 16-byte stores per row and patch);
//  * LayerNorm gamma / beta, the fusion conv's folded-BN scale, the query scale and every bias are folded into the
//    GEMMs (VC_TM_FOLD: scaled bf16 weight images; the bias is one more K step against the ones slab), so the epilogues
//    are pack / ReLU / GELU / residual only; the per-channel constants that remain (cls query bias, flags; all of them
//    with VC_TM_FOLD=0) live in __constant__ memory, filled per launch by a one-block prep kernel and a
//    device-to-device copy on the same stream, and enter the FFMAs as uniform / constant operands: no shared-memory
//    loads, nothing in the MIO queue;
//  * the next patch's fusion GEMM is issued with the K / V GEMM of the last block, the cls query is computed by the row
//    warps while that GEMM is in flight (no single-lane TMEM read, no barrier behind the GEMM);
//  * the MMA-issuer warps run warp-uniform control flow with the tcgen05 instructions in elect blocks (UTCHMMA takes
//    uniform-register operands: inside a divergent region every one of them costs an ELECT / R2UR.BROADCAST loop);
//  * the static bound that decides whether the softmax may skip the row maximum is evaluated once by the prep
//    kernel.  This kernel implements the no-maximum path only; when the bound fails it exits at once and
//    tokens_tc_kernel (launched right behind it, gated on the same flag) does the work.
#include <mutex>
#include "tokens_tc_common.cuh"

namespace vc {

namespace tm {
using tc::SLAB;
using tc::V_BFC1; using tc::V_BFC2; using tc::V_BPROJ; using tc::V_BQKV; using tc::V_BQKV2; using tc::V_FBI; using tc::V_FSC;
using tc::V_L2B; using tc::V_L2G; using tc::V_LN1B; using tc::V_LN1G; using tc::V_LN2B; using tc::V_LN2G; using tc::V_TOTAL;
constexpr int V_EXACT = V_TOTAL, V_EXACT_CLS = V_TOTAL + 1;   // flags (0 / 1) behind the vectors
constexpr int kVecFloats = V_TOTAL + 16;
constexpr size_t kStageBytes = 4096;                          // staging area at the head of the scratch buffer

__constant__ float c_vec[kVecFloats];

// VC_TM_FOLD: LayerNorm gamma / beta, the folded-BN scale of the fusion conv, the query scale and every bias are folded into
// the GEMMs.  Weights image: W'[n][k] = W[n][k] * gamma[k] * rowscale[n] (bf16); biases b'[n] = rowscale[n] (b[n] + W[n] . beta)
// ride as one more K step of every GEMM: A = the ones slab (rows (1,0,..,0) for both K chunks, leading-byte offset 0), B = two
// extra chunks behind the weights holding (hi(b'), 0, ..) and (lo(b'), 0, ..) -- a bf16 pair keeps 16 mantissa bits.  The
// epilogues lose their per-channel FFMA / FADD (416 instructions per token row and patch).
#ifndef VC_TM_FOLD
#define VC_TM_FOLD 1
#endif
constexpr bool kFold = VC_TM_FOLD != 0;
constexpr int kXC = kFold ? 2 : 0;         // extra 8-element K chunks per weight matrix (bias hi / lo)
// shared-memory image of the weights, [k / 8][N][8] bf16 each
constexpr uint32_t W_FUS = 0;
constexpr uint32_t W_QKV1 = W_FUS + (8 + kXC) * 32 * 16;
constexpr uint32_t W_PROJ1 = W_QKV1 + (4 + kXC) * 96 * 16;
constexpr uint32_t W_FC1 = W_PROJ1 + (4 + kXC) * 32 * 16;
constexpr uint32_t W_FC2 = W_FC1 + (4 + kXC) * 128 * 16;
constexpr uint32_t W_QKV2 = W_FC2 + (16 + kXC) * 32 * 16;
constexpr uint32_t W_END = W_QKV2 + (4 + kXC) * 96 * 16;
#ifndef VC_TM_PARKX
#define VC_TM_PARKX 0      // 1: tm4 parks the residual stream in shared memory during the attention (measured: +1 %, off)
#endif
template <int SLOTS>
struct Cfg {
  static constexpr int CK = SLOTS >= 4 ? 32 : 64;       // keys per attention step
  static constexpr int RD = SLOTS >= 4 ? 3 : 2;         // S chunk buffers per slot
  static constexpr uint32_t C_O = RD * CK;              // two 16-column O_h buffers / the 32-column accumulators
  static constexpr uint32_t C_SLOT = C_O + 32;
  static constexpr int kThreads = SLOTS * 5 * 32;       // 4 row warps + 1 MMA-issuer warp per slot
  static constexpr uint32_t POS = W_END;                // weights (W_FUS .. W_QKV2), then pos-embed rows
  static constexpr uint32_t SLOT0 = POS + 128 * 32 * 4;
  static constexpr uint32_t S_QBUF = 0, S_ABUF = 0, S_KBUF = 4 * SLAB, S_VBUF = 8 * SLAB, S_FBUF = S_KBUF;
  // four slots run at 96 registers per thread: the residual stream (32 fp32 per row) is parked in shared memory from
  // LayerNorm 1 to the proj epilogue, so the attention loop keeps its addresses in registers instead of recomputing them
  static constexpr bool kParkX = VC_TM_PARKX && SLOTS >= 4;
  static constexpr uint32_t S_XBUF = 12 * SLAB;          // [8][128 rows][4 fp32]
  static constexpr uint32_t SLOT_BYTES = (kParkX ? 20 : 12) * SLAB;
  static constexpr uint32_t ONES = SLOT0 + SLOTS * SLOT_BYTES;
  static constexpr uint32_t MASK = ONES + SLAB;
  static constexpr uint32_t MISC = MASK + SLAB;
  // q0: [slot][row warp][32] fp32 cls query (one copy per warp); y0: [slot][2][32] bf16 LayerNorm output of the cls row
  static constexpr uint32_t M_Q0 = 0, M_Y0 = M_Q0 + SLOTS * 512, M_WMAX = M_Y0 + SLOTS * 128, M_BARS = M_WMAX + SLOTS * 64,
                            M_TMEM = M_BARS + SLOTS * 128;
  static constexpr uint32_t SMEM_BYTES = MISC + ((M_TMEM + 16 + 127) & ~127u);
  static_assert(SLOTS * C_SLOT <= 512, "TMEM columns");
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem: lane = row, 2 bf16 per 32-bit column, K = 16 -> 8 columns] . B[smem descriptor]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier operations on 32-bit shared-memory addresses (computed once per thread: the generic-pointer forms re-derive
// the shared window base at every use).  The wait tries once without any bookkeeping (in the attention loop the S chunk
// is normally there already) and only then enters the bounded polling loop: a protocol bug must surface as a trapped
// kernel, never as a hung GPU.  VC_TM_WAIT_HINT: try_wait with a suspend-time hint (compiles to NANOSLEEP.SYNCS between
// two phase checks) instead of the plain form in the polling loop.
#ifndef VC_TM_WAIT_HINT
#define VC_TM_WAIT_HINT 1
#endif
__device__ __forceinline__ bool mbar_try_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_hint_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity), "r"(10000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  if (mbar_try_a(addr, parity)) return;
  uint32_t spins = 0;
  while (!(VC_TM_WAIT_HINT ? mbar_try_hint_a(addr, parity) : mbar_try_a(addr, parity)))
    if (++spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
// Make a value opaque to the compiler: it then lives in a register instead of being re-derived from tid / the shared
// window base inside the loops (S2R / S2UR + shifts at every use under register pressure).
__device__ __forceinline__ void pin(uint32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ float rcp_fast(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// LayerNorm (eps 1e-6) of the row held by this thread, gamma / beta from the constant bank -> bf16 -> K-major A operand
template <int G, int B>
__device__ __forceinline__ void ln_store_c(const float (&x)[32], uint32_t dst_row, uint32_t dst2 = 0u) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) s += x[c];
  const float mean = s * (1.f / 32.f);
  float v = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) { const float d = x[c] - mean; v = fmaf(d, d, v); }
  const float rs = rsqrtf(v * (1.f / 32.f) + 1e-6f);
  const float nm = -mean * rs;
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    uint32_t p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c0 = 8 * sl + 2 * e;
      p[e] = kFold ? pack_bf16(fmaf(x[c0], rs, nm), fmaf(x[c0 + 1], rs, nm))      // gamma / beta live in the next GEMM
                   : pack_bf16(fmaf(fmaf(x[c0], rs, nm), c_vec[G + c0], c_vec[B + c0]),
                               fmaf(fmaf(x[c0 + 1], rs, nm), c_vec[G + c0 + 1], c_vec[B + c0 + 1]));
    }
    sts128(dst_row + sl * SLAB, p[0], p[1], p[2], p[3]);
    if (dst2) sts128(dst2 + sl * 16, p[0], p[1], p[2], p[3]);
  }
}

// One block: the per-channel vectors in the order of tc::V_* (q biases pre-scaled), then the two flags of the static
// bound on |q.k| (see tc_setup in tokens_tc_common.cuh: same arithmetic, evaluated once per launch instead of per CTA)
__global__ void __launch_bounds__(128) tm_prep_kernel(const uint8_t* blob, TLayout L, float* out) {
  __shared__ float scr[256];
  const int tid = threadIdx.x;
  const float qscale = 0.35355339059327376220f * 1.44269504088896340736f;
  auto copy_v = [&](int dst, int src, int n, float scale_first32) {
    for (int i = tid; i < n; i += 128) {
      const float v = __ldg(reinterpret_cast<const float*>(blob + src) + i);
      out[dst + i] = i < 32 ? v * scale_first32 : v;
    }
  };
  copy_v(V_FSC, L.fus_scale, 32, 1.f);
  copy_v(V_FBI, L.fus_bias, 32, 1.f);
  copy_v(V_LN1G, L.layer[0].ln1_g, 32, 1.f);
  copy_v(V_LN1B, L.layer[0].ln1_b, 32, 1.f);
  copy_v(V_BPROJ, L.layer[0].bproj, 32, 1.f);
  copy_v(V_LN2G, L.layer[0].ln2_g, 32, 1.f);
  copy_v(V_LN2B, L.layer[0].ln2_b, 32, 1.f);
  copy_v(V_BFC2, L.layer[0].bfc2, 32, 1.f);
  copy_v(V_L2G, L.layer[1].ln1_g, 32, 1.f);
  copy_v(V_L2B, L.layer[1].ln1_b, 32, 1.f);
  if (!kFold) {
    copy_v(V_BQKV, L.layer[0].bqkv, 96, qscale);
    copy_v(V_BFC1, L.layer[0].bfc1, 128, 1.f);
    copy_v(V_BQKV2, L.layer[1].bqkv, 96, qscale);
  } else {
    // b'[n] = rowscale[n] (b[n] + W[n] . beta): the LayerNorm in front of the GEMM writes (x - mean) rstd only
    auto fold = [&](int dst, int wsrc, int bsrc, int betasrc, int N, bool qrows) {
      for (int n = tid; n < N; n += 128) {
        const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(blob + wsrc) + n * kLdD;
        const float* be = reinterpret_cast<const float*>(blob + betasrc);
        float acc = __ldg(reinterpret_cast<const float*>(blob + bsrc) + n);
        for (int c = 0; c < 32; ++c) acc = fmaf(__bfloat162float(w[c]), __ldg(be + c), acc);
        out[dst + n] = (qrows && n < 32) ? acc * qscale : acc;
      }
    };
    fold(V_BQKV, L.layer[0].wqkv, L.layer[0].bqkv, L.layer[0].ln1_b, 96, true);
    fold(V_BFC1, L.layer[0].wfc1, L.layer[0].bfc1, L.layer[0].ln2_b, 128, false);
    fold(V_BQKV2, L.layer[1].wqkv, L.layer[1].bqkv, L.layer[1].ln1_b, 96, true);
  }
  {
    const int l = tid >> 6, n = tid & 63;
    const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(blob + L.layer[l].wqkv) + n * kLdD;
    const float* g = reinterpret_cast<const float*>(blob + L.layer[l].ln1_g);
    const float* be = reinterpret_cast<const float*>(blob + L.layer[l].ln1_b);
    float f2 = 0.f, bs = __ldg(reinterpret_cast<const float*>(blob + L.layer[l].bqkv) + n);
    for (int c = 0; c < 32; ++c) {
      const float wv = __bfloat162float(w[c]);
      f2 = fmaf(wv * __ldg(g + c), wv * __ldg(g + c), f2);
      bs = fmaf(wv, __ldg(be + c), bs);
    }
    scr[2 * tid] = f2;
    scr[2 * tid + 1] = bs * bs;
  }
  __syncthreads();
  if (tid < 2) {
    const float* s = scr + 128 * tid;
    bool exact = false;
    for (int h = 0; h < 4; ++h) {
      float qf = 0.f, qb = 0.f, kf = 0.f, kb = 0.f;
      for (int n = 0; n < 8; ++n) {
        qf += s[2 * (8 * h + n)]; qb += s[2 * (8 * h + n) + 1];
        kf += s[2 * (32 + 8 * h + n)]; kb += s[2 * (32 + 8 * h + n) + 1];
      }
      const float bound = qscale * (sqrtf(32.f * qf) + sqrtf(qb)) * (sqrtf(32.f * kf) + sqrtf(kb));
      if (!(bound < 100.f)) exact = true;
    }
    out[V_EXACT + tid] = exact ? 1.f : 0.f;
  }
  for (int i = V_TOTAL + 2 + tid; i < kVecFloats; i += 128) out[i] = 0.f;
}

}  // namespace tm

template <int SLOTS>
__global__ void __launch_bounds__(tm::Cfg<SLOTS>::kThreads, 1) tokens_tm_kernel(TcArgs a) {
  using C = tm::Cfg<SLOTS>;
  using namespace tm;
  constexpr int CK = C::CK, RD = C::RD;
  if (c_vec[V_EXACT] != 0.f) return;      // the softmax needs its row maximum: tokens_tc_kernel runs instead
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp >= 4 * SLOTS;
  const int slot = issuer ? warp - 4 * SLOTS : warp >> 2;
  const int r = tid & 127, wq = warp & 3;
  const int T = a.T, P = a.P;
  const TLayout& L = a.L;
  uint32_t sb = smem_u32(smem);
  pin(sb);
  float* q0_s = reinterpret_cast<float*>(smem + C::MISC + C::M_Q0) + (issuer ? 0 : warp * 32);   // [32] per row warp
  const uint32_t y0_s = sb + C::MISC + C::M_Y0 + (uint32_t)slot * 128u;                  // [2][4 x 16 B]
  float* wmax_s = reinterpret_cast<float*>(smem + C::MISC + C::M_WMAX) + slot * 16;      // [4 warps][4 heads]
  uint32_t bars = sb + C::MISC + C::M_BARS + (uint32_t)slot * 128u;   // 16 mbarriers per slot (shared addresses)
  pin(bars);
  const uint32_t b_rp = bars + 0;      // row threads: operands of the next GEMM are written (128 arrivals)
  const uint32_t b_mma = bars + 8;     // tensor core: the GEMM just issued (fusion / qkv / proj / fc1 / fc2 / kv2) is done
  const uint32_t b_pv = bars + 16;     // tensor core: O_h is complete
  const uint32_t b_s = bars + 24;      // [RD] tensor core: the S chunk in ring buffer k is in TMEM (and the PV that read the buffer is done)
  // [RD] row threads: the probabilities in ring buffer k are written (128 arrivals).  One barrier per buffer: S chunks are
  // issued ahead, so a row thread may finish step i + 1 before another has finished step i -- on a shared barrier its
  // second arrival would complete the phase of step i; on buffer k it cannot arrive again before PV of step i has run.
  const uint32_t b_p = bars + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::MISC + C::M_TMEM);
  const float qscale = 0.35355339059327376220f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)
  const bool exact_cls = c_vec[V_EXACT_CLS] != 0.f;

  // ---------------- one-time image of the parameters in the layouts the tensor core reads ----------------
  {
    // W'[n][k] = W[n][k] * colscale[k] * rowscale[n] (* qscale for the first 32 rows when `qrows`) -> [k / 8][N][8] bf16;
    // fold mode: two more chunks with (hi, lo) of the folded bias c_vec[bias + n] in their first element
    auto copy_w = [&](uint32_t dst, int src, int N, int K, int ld, int colscale, int rowscale, bool qrows, float all, int bias) {
      for (int i = tid; i < N * (K / 8); i += C::kThreads) {
        const int n = i % N, kc = i / N;
        uint4 g = __ldg(reinterpret_cast<const uint4*>(a.blob + src + (size_t)(n * ld + kc * 8) * 2));
        float rs = all;
        if (kFold && rowscale >= 0) rs *= __ldg(reinterpret_cast<const float*>(a.blob + rowscale) + n);
        if (kFold && qrows && n < 32) rs *= qscale;
        if (rs != 1.f || (kFold && colscale >= 0)) {
          __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(&g);
          for (int e = 0; e < 8; ++e) {
            const float cs = (kFold && colscale >= 0) ? __ldg(reinterpret_cast<const float*>(a.blob + colscale) + kc * 8 + e) : 1.f;
            hp[e] = __float2bfloat16_rn(__bfloat162float(hp[e]) * cs * rs);
          }
        }
        *reinterpret_cast<uint4*>(smem + dst + (size_t)kc * N * 16 + n * 16) = g;
      }
      if (kFold) {
        for (int n = tid; n < N; n += C::kThreads) {
          const float b = c_vec[bias + n];
          const __nv_bfloat16 hi = __float2bfloat16_rn(b), lo = __float2bfloat16_rn(b - __bfloat162float(hi));
          uint4 z = make_uint4(0u, 0u, 0u, 0u);
          z.x = (uint32_t)__bfloat16_as_ushort(hi);
          *reinterpret_cast<uint4*>(smem + dst + (size_t)(K / 8) * N * 16 + n * 16) = z;
          z.x = (uint32_t)__bfloat16_as_ushort(lo);
          *reinterpret_cast<uint4*>(smem + dst + (size_t)(K / 8 + 1) * N * 16 + n * 16) = z;
        }
      }
    };
    copy_w(W_FUS, L.wfus, 32, 64, kLdFus, -1, L.fus_scale, false, 1.f, V_FBI);
    copy_w(W_QKV1, L.layer[0].wqkv, 96, 32, kLdD, L.layer[0].ln1_g, -1, true, 1.f, V_BQKV);
    copy_w(W_PROJ1, L.layer[0].wproj, 32, 32, kLdD, -1, -1, false, 1.f, V_BPROJ);
    copy_w(W_FC1, L.layer[0].wfc1, 128, 32, kLdD, L.layer[0].ln2_g, -1, false, 1.f, V_BFC1);
    copy_w(W_FC2, L.layer[0].wfc2, 32, 128, kLdHid, -1, -1, false, 0.5f, V_BFC2);   // the 0.5 of GELU lives here: H = 2 gelu(.)
    copy_w(W_QKV2, L.layer[1].wqkv, 96, 32, kLdD, L.layer[1].ln1_g, -1, true, 1.f, V_BQKV2);
    // pos-embed rows (row 0 = cls + pos[0], rows >= T zero), 16-byte granules swizzled by row
    const float* pos = reinterpret_cast<const float*>(a.blob + L.pos);
    const float* cls = reinterpret_cast<const float*>(a.blob + L.cls);
    for (int i = tid; i < 128 * 8; i += C::kThreads) {
      const int row = i >> 3, g = i & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < T) {
        v = __ldg(reinterpret_cast<const float4*>(pos + row * 32 + 4 * g));
        if (row == 0) {
          const float4 c = __ldg(reinterpret_cast<const float4*>(cls + 4 * g));
          v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
        }
      }
      *reinterpret_cast<float4*>(smem + C::POS + row * 128 + ((g ^ (row & 7)) << 4)) = v;
    }
    // slot buffers, ONES / MASK slabs: zeros first
    for (uint32_t i = tid; i < (SLOTS * C::SLOT_BYTES + 2 * SLAB) / 16; i += C::kThreads)
      *reinterpret_cast<uint4*>(smem + C::SLOT0 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int i = tid; i < 128; i += C::kThreads) {
      *reinterpret_cast<uint32_t*>(smem + C::ONES + i * 16) = 0x00003F80u;                 // bf16 1.0
      *reinterpret_cast<uint32_t*>(smem + C::MASK + i * 16) = i >= T ? 0x0000C6EAu : 0u;   // bf16 -29952 for padded keys
    }
    if (tid == 0) {
      for (int s = 0; s < SLOTS; ++s) {
        uint64_t* bb = reinterpret_cast<uint64_t*>(smem + C::MISC + C::M_BARS) + s * 16;   // order: see b_rp .. b_p
        mbar_init(bb + 0, 128);
        for (int k = 1; k < 8; ++k) mbar_init(bb + k, 1);
        for (int k = 8; k < 16; ++k) mbar_init(bb + k, 128);
      }
      fence_mbar_init();
    }
    if (tid < 32) {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tb = tmem_base + (uint32_t)slot * C::C_SLOT;             // columns of this slot (issuer view)
  uint32_t tl = tb + ((uint32_t)(wq * 32) << 16);                   // + the 32 lanes of this warp
  uint32_t slot_s = sb + C::SLOT0 + (uint32_t)slot * C::SLOT_BYTES;
  pin(tl);
  pin(slot_s);
  const uint32_t fbuf = slot_s + C::S_FBUF, abuf = slot_s + C::S_ABUF, qbuf = slot_s + C::S_QBUF, kbuf = slot_s + C::S_KBUF,
                 vbuf = slot_s + C::S_VBUF;
  const int NK = (T + 31) & ~31;                    // keys rounded to the 32-column chunks the row threads read
  const int SPH = (NK + CK - 1) / CK;             // attention steps per head
  const int nslots = SLOTS * (int)gridDim.x;
  const int b0 = SLOTS * (int)blockIdx.x + slot;

  if (issuer) {
    // ============================ MMA issuer warp of this slot ============================
    // Warp-uniform control flow (every lane waits and steps the descriptors; they live in uniform registers), the
    // tcgen05 instructions themselves sit in `if (elect_one())` blocks: inside a divergent `lane == 0` region the
    // compiler wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~14 instructions each).
    // Descriptors are kept as 64-bit values and stepped by constant increments: the start-address field counts 16-byte
    // units from bit 0, the leading-byte-offset field from bit 16 (no carries: every address is below 256 KB).
    {
      uint32_t ph_rp = 0, ph_p = 0;
      int rb = 0;               // ring buffer of the next attention step (ph_p: parity of its "probabilities written" barrier)
      auto ready = [&]() {      // the row threads have written the operands of the next GEMM
        mbar_wait_a(b_rp, ph_rp);
        ph_rp ^= 1u;
        tc_fence_after();
      };
      const uint64_t d_ones = umma_desc(sb + C::ONES, 0, 128);      // A of the bias K step: both K chunks = the ones slab
      // D[128 x N] (TMEM column `col`) = A[128 x 16 ksteps] (K-major slabs at `abase`) . W^T (weights [k/8][N][8] at `wbase`)
      auto issue_gemm = [&](uint32_t col, uint32_t abase, uint32_t wbase, int N, int ksteps) {
        if (elect_one()) {
          uint64_t da = umma_desc(abase, SLAB, 128), dw = umma_desc(wbase, (uint32_t)N * 16u, 128);
          const uint32_t id = idesc(N, 0);
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16(tb + col, da, dw, id, k ? 1u : 0u);
            da += (2 * SLAB) >> 4;
            dw += (uint64_t)(2 * N);      // 2 N 16-byte rows per K step
          }
          if (kFold) umma_bf16(tb + col, d_ones, dw, id, 1u);     // + bias: ones . (hi, lo) chunks behind the weights
          umma_commit_a(b_mma);
        }
        __syncwarp();
      };
      const uint32_t ones = sb + C::ONES, mask = sb + C::MASK;
      // head 0, chunk 0: Q / K descriptors (A chunks (Q_h, ones), B chunks (keys of K_h, their mask rows)) and the V descriptor
      const uint64_t dq0 = umma_desc(qbuf, ones - qbuf, 128), dk0 = umma_desc(kbuf, mask - kbuf, 128);
      const uint64_t dv0 = umma_desc(vbuf, 128, ones - vbuf);
      const uint64_t d_head = (uint64_t)(SLAB >> 4) - ((uint64_t)(SLAB >> 4) << 16);   // next head: start + SLAB, LBO - SLAB
      const uint64_t dv_head = (uint64_t)(SLAB >> 4) - ((uint64_t)(SLAB >> 4) << 32);  // V: the ones slab sits in the SBO field
      const uint32_t id_pv = idesc(16, 1);
      for (int b = b0; b < a.n_patches; b += nslots) {
        if (b == b0) { ready(); issue_gemm(C::C_O, fbuf, sb + W_FUS, 32, 4); }   // fusion 1x1 conv of the first patch
        ready(); issue_gemm(0, abuf, sb + W_QKV1, 96, 2);           // qkv
        ready();
        // S chunk of step (hs, cs) -> ring buffer: issued RD steps ahead of the PV that frees the buffer
        int hs = 0, cs = 0;
        uint64_t dq = dq0, dk = dk0;
        auto issue_s = [&](int buf) {        // called by the elected lane; the step counters advance on every lane (s_next)
          const int keys = CK == 32 ? 32 : min(CK, NK - cs * CK);
          umma_bf16(tb + buf * CK, dq, dk + (uint64_t)(cs * CK), idesc(keys, 0), 0u);
          umma_commit_a(b_s + 8 * buf);
        };
        auto s_next = [&]() {
          if (++cs == SPH) { cs = 0; ++hs; dq += d_head; dk += d_head; }
        };
        {
          int buf = rb;
          for (int j = 0; j < RD && hs < 4; ++j) {
            if (elect_one()) issue_s(buf);
            __syncwarp();
            s_next();
            if (++buf == RD) buf = 0;
          }
        }
        uint64_t dv = dv0;
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
#pragma unroll 1
          for (int c = 0; c < SPH; ++c) {
            // O_h[128 x 16] (+)= P[128 x keys] . [V_h | ones]: B is MN-major, N chunk 0 = V_h slab, chunk 1 = ONES slab
            const int ksteps = (CK == 32 ? 32 : min(CK, NK - c * CK)) / 16;
            const uint32_t d_o = tb + C::C_O + 16 * (h & 1), a_p = tb + rb * CK;
            const uint64_t dvk0 = dv + (uint64_t)(c * CK);    // 16 keys = 256 B = 16 units per K step
            mbar_wait_a(b_p + 8 * rb, ph_p);                  // the probabilities of step (h, c) are in the first half of buffer rb
            tc_fence_after();
            if (elect_one()) {
              uint64_t dvk = dvk0;
              for (int kk = 0; kk < ksteps; ++kk) {
                umma_bf16_ts(d_o, a_p + 8 * kk, dvk, id_pv, (c | kk) ? 1u : 0u);
                dvk += 16;
              }
              if (c == SPH - 1) umma_commit_a(b_pv);
              if (hs < 4) issue_s(rb);                        // the buffer is free as soon as this PV has read it (in-order pipe)
            }
            __syncwarp();
            if (hs < 4) s_next();
            if (++rb == RD) { rb = 0; ph_p ^= 1u; }
          }
          dv += dv_head;
        }
        ready(); issue_gemm(C::C_O, abuf, sb + W_PROJ1, 32, 2);     // proj
        ready(); issue_gemm(0, abuf, sb + W_FC1, 128, 2);           // fc1
        ready();                                                        // fc2: A = packed hidden units in TMEM columns 0..63
        if (elect_one()) {
          uint64_t dw = umma_desc(sb + W_FC2, 32 * 16, 128);
          const uint32_t id = idesc(32, 0);
          for (int k = 0; k < 8; ++k) {
            umma_bf16_ts(tb + C::C_O, tb + 8 * k, dw, id, k ? 1u : 0u);
            dw += 64;
          }
          if (kFold) umma_bf16(tb + C::C_O, d_ones, dw, id, 1u);
          umma_commit_a(b_mma);
        }
        __syncwarp();
        ready();
        if (elect_one()) {
          // last block: k, v of every token (the cls query is computed by the row threads), columns 0..63 -- and, in the
          // same hand-off, the fusion 1x1 conv of the slot's NEXT patch (its input has landed over the dead K / V buffers)
          uint64_t da = umma_desc(abuf, SLAB, 128), dw = umma_desc(sb + W_QKV2 + 32 * 16, 96 * 16, 128);
          const uint32_t id = idesc(64, 0);
          for (int k = 0; k < 2; ++k) {
            umma_bf16(tb, da, dw, id, k ? 1u : 0u);
            da += (2 * SLAB) >> 4;
            dw += 2 * 96;
          }
          if (kFold) umma_bf16(tb, d_ones, dw, id, 1u);
          if (b + nslots < a.n_patches) {
            uint64_t df = umma_desc(fbuf, SLAB, 128), dwf = umma_desc(sb + W_FUS, 32 * 16, 128);
            const uint32_t idf = idesc(32, 0);
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tb + C::C_O, df, dwf, idf, k ? 1u : 0u);
              df += (2 * SLAB) >> 4;
              dwf += 64;
            }
            if (kFold) umma_bf16(tb + C::C_O, d_ones, dwf, idf, 1u);
          }
          umma_commit_a(b_mma);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ row threads ============================
    uint32_t row16 = (uint32_t)r * 16u;
    pin(row16);
    const int PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P);
    const int bar_id = 1 + slot;
    uint32_t ph_m = 0, ph_pv = 0, rpar = 0;
    int rbuf = 0;
    // stem outputs of token row r of patch b -> FBUF (8 slices x 16 B; cls row and padding rows stay zero); see tokens_tc.cu
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
      } else {          // the buffer doubles as K / V: the cls row and the padding rows are zeroed every time
#pragma unroll
        for (int s = 0; s < 8; ++s) sts128(fbuf + s * SLAB + row16, 0u, 0u, 0u, 0u);
      }
    };
    auto publish = [&]() {        // shared-memory operands written by this thread -> visible to the tensor core
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive_a(b_rp);
    };
    auto publish_tmem = [&](uint32_t bar) {   // TMEM operands stored by this thread (and its TMEM reads) are complete
      tc_wait_st();
      tc_fence_before();
      mbar_arrive_a(bar);
    };
    auto wait_mma = [&]() {
      mbar_wait_a(b_mma, ph_m);
      ph_m ^= 1u;
      tc_fence_after();
    };
    if (b0 < a.n_patches) fetch(b0);

    for (int b = b0; b < a.n_patches; b += nslots) {
      float x[32];   // residual stream of token row r
      // ================= fusion 1x1 conv (64 -> 32) + folded BN + ReLU, + cls / pos =================
      // (first patch of the slot: its own hand-off; later ones were issued with kv2 of the previous patch)
      if (b == b0) {
        cp_async_wait_all();
        publish();
        wait_mma();
      }
      {
        uint32_t v[32];
        tmem_ld32(tl + C::C_O, v);
        tc_wait_ld();
        const float rowmask = (r >= 1 && r < T) ? 1.f : 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 p = lds_f4(sb + C::POS + r * 128 + ((g ^ (r & 7)) << 4));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pe = e == 0 ? p.x : e == 1 ? p.y : e == 2 ? p.z : p.w;
            const float cv = __uint_as_float(v[4 * g + e]);
            x[4 * g + e] = kFold ? fmaf(fmaxf(cv, 0.f), rowmask, pe)
                                 : fmaf(fmaxf(fmaf(cv, c_vec[V_FSC + 4 * g + e], c_vec[V_FBI + 4 * g + e]), 0.f), rowmask, pe);
          }
        }
      }

      // ================= block 1: LN1 -> qkv =================
      ln_store_c<V_LN1G, V_LN1B>(x, abuf + row16);
      publish();
      if (C::kParkX) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          sts128(slot_s + C::S_XBUF + g * SLAB + row16, __float_as_uint(x[4 * g]), __float_as_uint(x[4 * g + 1]), __float_as_uint(x[4 * g + 2]),
                 __float_as_uint(x[4 * g + 3]));
      }
      wait_mma();
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        uint32_t v[32];
        tmem_ld32(tl + 32 * part, v);
        tc_wait_ld();
        const uint32_t dst = (part == 0 ? qbuf : part == 1 ? kbuf : vbuf) + row16;
        const float sc = part == 0 ? qscale : 1.f;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t p[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            p[e] = kFold ? pack_bf16(__uint_as_float(v[8 * h + 2 * e]), __uint_as_float(v[8 * h + 2 * e + 1]))
                         : pack_bf16(fmaf(__uint_as_float(v[8 * h + 2 * e]), sc, c_vec[V_BQKV + 32 * part + 8 * h + 2 * e]),
                                     fmaf(__uint_as_float(v[8 * h + 2 * e + 1]), sc, c_vec[V_BQKV + 32 * part + 8 * h + 2 * e + 1]));
          sts128(dst + h * SLAB, p[0], p[1], p[2], p[3]);
        }
      }
      publish();

      // ================= attention: 4 heads x SPH key chunks =================
      // p = 2^s (the static bound holds: no row maximum; padded keys arrive at -29952 -> 0), packed to bf16 over the
      // first half of the chunk's own TMEM columns.  O_h (two 16-column buffers) is read one step after its last PV
      // was issued, normalised by column 8 (the denominator of the bf16 probabilities) and written to slab h of the
      // Q buffer (= A operand of proj; every S of head h is complete by then).
      auto read_o = [&](int h) {
        uint32_t o[16];
        mbar_wait_a(b_pv, ph_pv);
        ph_pv ^= 1u;
        tc_fence_after();
        tmem_ld16(tl + C::C_O + 16 * (h & 1), o);
        tc_wait_ld();
        const float il = rcp_fast(__uint_as_float(o[8]));
        sts128(abuf + h * SLAB + row16, pack_bf16(__uint_as_float(o[0]) * il, __uint_as_float(o[1]) * il),
               pack_bf16(__uint_as_float(o[2]) * il, __uint_as_float(o[3]) * il),
               pack_bf16(__uint_as_float(o[4]) * il, __uint_as_float(o[5]) * il),
               pack_bf16(__uint_as_float(o[6]) * il, __uint_as_float(o[7]) * il));
      };
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
#pragma unroll 1
        for (int c = 0; c < SPH; ++c) {
          mbar_wait_a(b_s + 8 * rbuf, rpar);
          tc_fence_after();
          const uint32_t scol = tl + (uint32_t)(rbuf * CK);
          const int keys = CK == 32 ? 32 : min(CK, NK - c * CK);
#pragma unroll
          for (int sub = 0; sub < CK / 32; ++sub) {
            if (CK == 32 || 32 * sub < keys) {
              uint32_t sc[32], pk[16];
              tmem_ld32(scol + 32 * sub, sc);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(ex2(__uint_as_float(sc[2 * e])), ex2(__uint_as_float(sc[2 * e + 1])));
              tmem_st16(scol + 16 * sub, pk);
            }
          }
          // O_{h-1}: with one step per head it must be taken BEFORE this arrival (else PV of head h could complete b_pv a
          // second time before this thread has seen the first), otherwise one step later, when its PV is surely done
          if (c == 0 && h > 0 && SPH == 1) read_o(h - 1);
          publish_tmem(b_p + 8 * rbuf);
          if (++rbuf == RD) { rbuf = 0; rpar ^= 1u; }
          if (c == 0 && h > 0 && SPH > 1) read_o(h - 1);
        }
      }
      {   // O_3: wait first, then start the next patch's input over the dead K / V buffers
        uint32_t o[16];
        mbar_wait_a(b_pv, ph_pv);
        ph_pv ^= 1u;
        tc_fence_after();
        if (b + nslots < a.n_patches) fetch(b + nslots);
        tmem_ld16(tl + C::C_O + 16, o);
        tc_wait_ld();
        const float il = rcp_fast(__uint_as_float(o[8]));
        sts128(abuf + 3 * SLAB + row16, pack_bf16(__uint_as_float(o[0]) * il, __uint_as_float(o[1]) * il),
               pack_bf16(__uint_as_float(o[2]) * il, __uint_as_float(o[3]) * il),
               pack_bf16(__uint_as_float(o[4]) * il, __uint_as_float(o[5]) * il),
               pack_bf16(__uint_as_float(o[6]) * il, __uint_as_float(o[7]) * il));
      }
      publish();
      wait_mma();
      {
        uint32_t v[32];
        tmem_ld32(tl + C::C_O, v);
        tc_wait_ld();
        if (C::kParkX) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 xv = lds_f4(slot_s + C::S_XBUF + g * SLAB + row16);
            x[4 * g] = xv.x; x[4 * g + 1] = xv.y; x[4 * g + 2] = xv.z; x[4 * g + 3] = xv.w;
          }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] += kFold ? __uint_as_float(v[c]) : __uint_as_float(v[c]) + c_vec[V_BPROJ + c];
      }

      // ================= MLP: LN2 -> fc1 (+bias, GELU, in place) -> fc2 (+bias, +residual) =================
      ln_store_c<V_LN2G, V_LN2B>(x, abuf + row16);
      publish();
      wait_mma();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32], pk[16];
        tmem_ld32(tl + 32 * c, v);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          pk[e] = kFold ? pack_bf16(gelu2(__uint_as_float(v[2 * e])), gelu2(__uint_as_float(v[2 * e + 1])))
                        : pack_bf16(gelu2(__uint_as_float(v[2 * e]) + c_vec[V_BFC1 + 32 * c + 2 * e]),
                                    gelu2(__uint_as_float(v[2 * e + 1]) + c_vec[V_BFC1 + 32 * c + 2 * e + 1]));
        tmem_st16(tl + 16 * c, pk);       // columns [16 c, 16 c + 16) were read with chunk c / 2
      }
      publish_tmem(b_rp);
      wait_mma();
      {
        uint32_t v[32];
        tmem_ld32(tl + C::C_O, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] += kFold ? __uint_as_float(v[c]) : __uint_as_float(v[c]) + c_vec[V_BFC2 + c];
      }

      // ================= last block: K / V of every token, attention of the cls query only =================
      // The cls query q0 = qscale (W_q y_0 + b_q) is computed by the CUDA cores of every row warp (lane c: channel c) from
      // the bf16 LayerNorm output of row 0 while the K / V GEMM is in flight: no TMEM read of a single lane, no barrier
      // between the GEMM and the scores (the one barrier here sits where every thread waits for the hand-off anyway).
      const uint32_t y0buf = y0_s + 64u * ((uint32_t)(b - b0) / (uint32_t)nslots & 1u);
      ln_store_c<V_L2G, V_L2B>(x, abuf + row16, r == 0 ? y0buf : 0u);
      cp_async_wait_all();            // the next patch's fusion input (issued after the last PV) has landed
      publish();
      float* trec = a.tail + (long long)b * tc::kTailFloats;
      if (r == 0) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(trec + 144 + c) = make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]);
      }
      bar_sync(bar_id, 128);          // y_0 is in shared memory (its buffer is reused two patches later: another barrier in between)
      {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint4 yv, wv;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(yv.x), "=r"(yv.y), "=r"(yv.z), "=r"(yv.w) : "r"(y0buf + kc * 16));
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(wv.x), "=r"(wv.y), "=r"(wv.z), "=r"(wv.w)
                       : "r"(sb + W_QKV2 + kc * 96 * 16 + lane * 16));
          acc0 = fmaf(bf_lo(yv.x), bf_lo(wv.x), acc0); acc1 = fmaf(bf_hi(yv.x), bf_hi(wv.x), acc1);
          acc0 = fmaf(bf_lo(yv.y), bf_lo(wv.y), acc0); acc1 = fmaf(bf_hi(yv.y), bf_hi(wv.y), acc1);
          acc0 = fmaf(bf_lo(yv.z), bf_lo(wv.z), acc0); acc1 = fmaf(bf_hi(yv.z), bf_hi(wv.z), acc1);
          acc0 = fmaf(bf_lo(yv.w), bf_lo(wv.w), acc0); acc1 = fmaf(bf_hi(yv.w), bf_hi(wv.w), acc1);
        }
        q0_s[lane] = kFold ? acc0 + acc1 + c_vec[V_BQKV2 + lane] : fmaf(acc0 + acc1, qscale, c_vec[V_BQKV2 + lane]);
        __syncwarp();
      }
      wait_mma();
      {
        uint32_t kk[32], vv[32];
        tmem_ld32(tl, kk);
        tmem_ld32(tl + 32, vv);
        tc_wait_ld();
        float sc[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float4 q0 = lds_f4(smem_u32(q0_s) + 32 * h), q1 = lds_f4(smem_u32(q0_s) + 32 * h + 16);
          float kb[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) kb[e] = kFold ? __uint_as_float(kk[8 * h + e]) : __uint_as_float(kk[8 * h + e]) + c_vec[V_BQKV2 + 32 + 8 * h + e];
          float d = q0.x * kb[0];
          d = fmaf(q0.y, kb[1], d);
          d = fmaf(q0.z, kb[2], d);
          d = fmaf(q0.w, kb[3], d);
          d = fmaf(q1.x, kb[4], d);
          d = fmaf(q1.y, kb[5], d);
          d = fmaf(q1.z, kb[6], d);
          d = fmaf(q1.w, kb[7], d);
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
#pragma unroll
          for (int e = 0; e < 8; ++e)
            val[8 * h + e] = kFold ? p * __uint_as_float(vv[8 * h + e]) : p * (__uint_as_float(vv[8 * h + e]) + c_vec[V_BQKV2 + 64 + 8 * h + e]);
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

template <int SLOTS>
static int tm_launch_slots(const TcArgs& a, int n_patches, int num_sms, int max_smem, cudaStream_t stream) {
  using C = tm::Cfg<SLOTS>;
  if ((int)C::SMEM_BYTES > max_smem) return VC_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(tokens_tm_kernel<SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES) != cudaSuccess)
    return VC_ERR_CUDA;
  int blocks = (n_patches + SLOTS - 1) / SLOTS;
  if (blocks > num_sms) blocks = num_sms;
  tokens_tm_kernel<SLOTS><<<blocks, C::kThreads, C::SMEM_BYTES, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

size_t tokens_tm_stage_bytes() { return tm::kStageBytes; }

// Prep kernel -> constant bank -> main kernel on `stream`.  `stage` = tm::kStageBytes of device memory owned by the
// caller's scratch buffer; stage[V_EXACT] stays readable for the gated fallback launch (tokens_tc.cu).
int tokens_tm_main_launch(const TcArgs& a, float* stage, int n_patches, int num_sms, int max_smem, int slots, cudaStream_t stream) {
  // The constant bank is one per device: a launch on another stream must not overwrite it under a kernel still running.
  static std::mutex mu;
  static cudaEvent_t ev[64] = {};
  static cudaStream_t last[64] = {};
  static bool have[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return VC_ERR_UNSUPPORTED;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cap);
  const bool track = cap == cudaStreamCaptureStatusNone;
  std::lock_guard<std::mutex> lock(mu);
  if (track && have[dev] && last[dev] != stream && cudaStreamWaitEvent(stream, ev[dev], 0) != cudaSuccess) return VC_ERR_CUDA;
  tm::tm_prep_kernel<<<1, 128, 0, stream>>>(a.blob, a.L, stage);
  if (cudaGetLastError() != cudaSuccess) return VC_ERR_CUDA;
  if (cudaMemcpyToSymbolAsync(tm::c_vec, stage, tm::kVecFloats * sizeof(float), 0, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return VC_ERR_CUDA;
  const int rc = slots >= 4 ? tm_launch_slots<4>(a, n_patches, num_sms, max_smem, stream) : tm_launch_slots<3>(a, n_patches, num_sms, max_smem, stream);
  if (rc != VC_OK) return rc;
  if (track) {
    if (!ev[dev] && cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming) != cudaSuccess) return VC_ERR_CUDA;
    if (cudaEventRecord(ev[dev], stream) != cudaSuccess) return VC_ERR_CUDA;
    last[dev] = stream;
    have[dev] = true;
  }
  return VC_OK;
}

}  // namespace vc
