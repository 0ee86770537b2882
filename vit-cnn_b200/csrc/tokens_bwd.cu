// Backward pass of the token stage (cls/pos -> 2 pre-norm transformer blocks -> final LayerNorm
// -> head on the cls token) as ONE kernel per batch: what autograd derives for
// model/compare_method/vit/timm/models/vision_transformer.py:57-105 (Attention), :123-166 (Block),
// :598-629 (cls/pos), :680-701 (norm/head) and timm/layers/mlp.py:13-47 when the reference's
// train() calls loss.backward() (model_utils.py:936).
//
// Same work split as the forward kernel (transformer.cu): one CTA owns one patch at a time, warp
// w owns token rows 16w..16w+15 and keeps residual stream and its gradient in registers in the
// mma.sync accumulator layout.  Nothing of the forward pass is read back from HBM except the
// token inputs: the block is recomputed, and only Q, K, V, O and the softmax statistics of the
// current patch live in shared memory.  Attention backward is the two-sweep scheme (query-owned
// sweep for dQ, key-owned sweep for dK / dV), so no cross-warp reduction is needed; transposed
// operands come from ldmatrix.trans instead of transposed copies.  The last block only feeds the
// head through the cls token, so its backward is a single-query attention plus vector-matrix
// products (one warp), and K / V gradients for every token.
//
// Weight gradients (sums over ALL token rows of the batch) are not reduced here: the kernel
// writes the bf16 operand pairs (layer input X, output gradient dY) of every linear layer in
// the [slice][row][8] layout and the tcgen05 weight-gradient kernel (wgrad_tc.cu) contracts
// them over rows; a constant-one channel in X makes the bias gradient one more column.
#include "vc_common.cuh"
#include "vc_kernels.h"
#include "vc_tparams.h"
#include "vc_tokens.cuh"

namespace vc {

struct TBArgs {
  const __nv_bfloat16* zf;   // [4][RT][8]  tokens before "+ pos" (relu(bn(fusion conv)))
  const uint8_t* blob;       // parameter blob (vc_tparams.h)
  const float* dlogits;      // [n][K]
  __nv_bfloat16* dzf;        // [4][RT][8]  gradient w.r.t. zf (pad cells untouched: pre-zeroed)
  // operand dumps, token-row space [slice][RTt][8] (row = patch * TP + token)
  __nv_bfloat16 *xln1[2], *dqkv[2];
  __nv_bfloat16 *xo0, *dxa0, *xln2_0, *dh0, *xh0, *dxb0;
  // operand dumps of the cls-only block / head, compact space [slice][RTc][8] (row = patch)
  __nv_bfloat16 *c_xo, *c_dxa, *c_xln2, *c_dh, *c_xh, *c_dxb, *c_xc, *c_dlog;
  float* g_ln[5][2];         // fp32 [32] gamma / beta grads: ln1_0, ln2_0, ln1_1, ln2_1, final norm (atomics)
  float *g_cls, *g_pos;      // fp32 [32], [T][32] (atomics; caller zeroes all of these)
  long long RT, RTt, RTc;
  int n, P, K, T;
  uint32_t drop_thr;            // dropout of the forward pass being differentiated (0 = off)
  const uint32_t* drop_seed;    // device: seed of this step
  TLayout L;
};


__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void acc2_to_afrag(const float (&t0)[4], const float (&t1)[4], uint32_t (&A)[4]) {
  A[0] = pack_bf16(t0[0], t0[1]);
  A[1] = pack_bf16(t0[2], t0[3]);
  A[2] = pack_bf16(t1[0], t1[1]);
  A[3] = pack_bf16(t1[2], t1[3]);
}
// dump a bf16 pair (row, cols c, c+1) into an operand buffer [slice][rows][8]
__device__ __forceinline__ void dump2(__nv_bfloat16* buf, long long rows, long long row, int c, uint32_t v) {
  *reinterpret_cast<uint32_t*>(buf + ((long long)(c >> 3) * rows + row) * 8 + (c & 7)) = v;
}
__device__ __forceinline__ uint32_t load2(const __nv_bfloat16* buf, long long rows, long long row, int c) {
  return *reinterpret_cast<const volatile uint32_t*>(buf + ((long long)(c >> 3) * rows + row) * 8 + (c & 7));
}
// dump the K=32 A fragments of this thread (rows r0 / r1 of the patch)
__device__ __forceinline__ void dump_afrag32(__nv_bfloat16* buf, long long rows, long long row0, int q,
                                             const uint32_t (&A)[2][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    dump2(buf, rows, row0, 16 * kk + 2 * q, A[kk][0]);
    dump2(buf, rows, row0 + 8, 16 * kk + 2 * q, A[kk][1]);
    dump2(buf, rows, row0, 16 * kk + 8 + 2 * q, A[kk][2]);
    dump2(buf, rows, row0 + 8, 16 * kk + 8 + 2 * q, A[kk][3]);
  }
}
// LayerNorm statistics of the two rows this thread shares with its quad
__device__ __forceinline__ void ln_stats(const float (&x)[4][4], float& m0, float& rs0, float& m1, float& rs1) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3]; }
  m0 = quad_sum(s0) * (1.f / kD);
  m1 = quad_sum(s1) * (1.f / kD);
  float v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float d;
    d = x[j][0] - m0; v0 += d * d;
    d = x[j][1] - m0; v0 += d * d;
    d = x[j][2] - m1; v1 += d * d;
    d = x[j][3] - m1; v1 += d * d;
  }
  rs0 = rsqrtf(quad_sum(v0) * (1.f / kD) + 1e-6f);
  rs1 = rsqrtf(quad_sum(v1) * (1.f / kD) + 1e-6f);
}

// LayerNorm backward for the thread's two rows: dx (+)= d LN(x) given dy (gradient w.r.t. the
// LN output), accumulating dgamma / dbeta partials (gacc[0][*] gamma, gacc[1][*] beta; 8 columns
// per thread: 8*jn + 2q + {0,1}).
__device__ __forceinline__ void ln_backward(const float (&x)[4][4], const float (&dy)[4][4], const float* gam, int q,
                                            float (&dx)[4][4], bool add, float (&gacc)[2][8]) {
  float m0, rs0, m1, rs1;
  ln_stats(x, m0, rs0, m1, rs1);
  float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
  float xh[4][4], dg[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 gg = *reinterpret_cast<const float2*>(gam + 8 * j + 2 * q);
    xh[j][0] = (x[j][0] - m0) * rs0; xh[j][1] = (x[j][1] - m0) * rs0;
    xh[j][2] = (x[j][2] - m1) * rs1; xh[j][3] = (x[j][3] - m1) * rs1;
    dg[j][0] = dy[j][0] * gg.x; dg[j][1] = dy[j][1] * gg.y;
    dg[j][2] = dy[j][2] * gg.x; dg[j][3] = dy[j][3] * gg.y;
    a0 += dg[j][0] + dg[j][1]; a1 += dg[j][2] + dg[j][3];
    b0 += dg[j][0] * xh[j][0] + dg[j][1] * xh[j][1];
    b1 += dg[j][2] * xh[j][2] + dg[j][3] * xh[j][3];
    gacc[0][2 * j] += dy[j][0] * xh[j][0] + dy[j][2] * xh[j][2];
    gacc[0][2 * j + 1] += dy[j][1] * xh[j][1] + dy[j][3] * xh[j][3];
    gacc[1][2 * j] += dy[j][0] + dy[j][2];
    gacc[1][2 * j + 1] += dy[j][1] + dy[j][3];
  }
  a0 = quad_sum(a0) * (1.f / kD); a1 = quad_sum(a1) * (1.f / kD);
  b0 = quad_sum(b0) * (1.f / kD); b1 = quad_sum(b1) * (1.f / kD);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float v0 = rs0 * (dg[j][0] - a0 - xh[j][0] * b0), v1 = rs0 * (dg[j][1] - a0 - xh[j][1] * b0);
    const float v2 = rs1 * (dg[j][2] - a1 - xh[j][2] * b1), v3 = rs1 * (dg[j][3] - a1 - xh[j][3] * b1);
    if (add) { dx[j][0] += v0; dx[j][1] += v1; dx[j][2] += v2; dx[j][3] += v3; }
    else { dx[j][0] = v0; dx[j][1] = v1; dx[j][2] = v2; dx[j][3] = v3; }
  }
}

// d_in[16 x 32] = d_out[16 x 16*KS] * W  with W stored [out][ld] (k = out index, n = in index):
// transposed operand through ldmatrix.trans.  A fragments given per K=16 step.
template <int KS>
__device__ __forceinline__ void gemm_dgrad32(const uint32_t (*A)[4], const __nv_bfloat16* W, int ld, int lane,
                                             float (&out)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) out[j][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      uint32_t b0, b1;
      ldsm_x2_trans(b0, b1, W + (16 * kk + (lane & 15)) * ld + 8 * jn);
      mma16816(out[jn], A[kk], b0, b1);
    }
  }
}

// scratch of the cls-only path (one warp)
struct ClsScratch {
  float q[kD], doo[kD], dq[kD], dxres[kD], dlt[kHeads], h[kHidden], du[kHidden];
  float part[16][kD];
  float p[kHeads][256], ds[kHeads][256];
};

template <int NW, bool DROP>
__global__ void __launch_bounds__(NW * 32, 1) transformer_bwd_kernel(TBArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int TP = 16 * NW;
  constexpr int LD = kLdD;                 // 40: pitch of the [TP][32] bf16 arrays
  constexpr int NT = 2 * NW;               // 8-wide key / query tiles
  const TLayout& L = a.L;
  const int PB = L.pos;                    // parameter bytes kept in shared memory (pos stays in L2)
  // five [TP][40] bf16 arrays; the block-1 arrays reuse the block-0 ones (Q/K/V of block 0 are
  // recomputed at the start of its backward: 24 MMAs per warp for 3 arrays less shared memory)
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem + PB);
  __nv_bfloat16* Ks = Qs + TP * LD;
  __nv_bfloat16* Vs = Ks + TP * LD;
  __nv_bfloat16* Os = Vs + TP * LD;
  __nv_bfloat16* dOs = Os + TP * LD;
  __nv_bfloat16* K1s = Qs;
  __nv_bfloat16* V1s = Ks;
  __nv_bfloat16* dK1s = Vs;
  __nv_bfloat16* dV1s = dOs;
  float* st_m = reinterpret_cast<float*>(dOs + TP * LD);   // [4][TP]
  float* st_il = st_m + kHeads * TP;
  float* st_dl = st_il + kHeads * TP;
  ClsScratch* cs = reinterpret_cast<ClsScratch*>(st_dl + kHeads * TP);

  {
    const uint4* src = reinterpret_cast<const uint4*>(a.blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < PB / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int T = a.T, P = a.P, K = a.K;
  const int r0 = 16 * warp + g, r1 = r0 + 8;
  const int PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P);
  const float* f32 = reinterpret_cast<const float*>(smem);
  const float* posg = reinterpret_cast<const float*>(a.blob + L.pos);
  const float qscale = 0.35355339059327376220f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)
  const float kLn2 = 0.69314718055994530942f, kScale = 0.35355339059327376220f;
  DropCfg dcfg;
  dcfg.thr = DROP ? a.drop_thr : 0u;
  dcfg.seed = DROP ? __ldg(a.drop_seed) : 0u;
  dcfg.inv_keep = 65536.f / (65536.f - (float)dcfg.thr);
  const TLayerOff& O0 = L.layer[0];
  const TLayerOff& O1 = L.layer[1];

  // persistent gradient partials
  float gpos[4][4];
  float gln[3][2][8];            // ln1 block0, ln2 block0, ln1 block1
  float gcls_ln2[2] = {0.f, 0.f}, gcls_lnf[2] = {0.f, 0.f};   // warp 0: lane = channel
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) gpos[j][e] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) { gln[i][0][k] = 0.f; gln[i][1][k] = 0.f; }

  for (int b = blockIdx.x; b < a.n; b += gridDim.x) {
    const long long trow = (long long)b * TP + 16 * warp + g;   // token-row-space row of r0
    long long R0 = -1, R1 = -1;
    if (r0 >= 1 && r0 < T) { const int p = r0 - 1; R0 = HALO + (long long)b * PP + (p / P) * PW + (p % P); }
    if (r1 >= 1 && r1 < T) { const int p = r1 - 1; R1 = HALO + (long long)b * PP + (p / P) * PW + (p % P); }

    // ------------------------------ tokens: x0 = [cls | zf] + pos ------------------------------
    auto load_x0 = [&](float (&x)[4][4]) {
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        const __nv_bfloat16* sl = a.zf + (long long)jn * a.RT * 8 + 2 * q;
        const uint32_t u0 = R0 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(sl + R0 * 8)) : 0u;
        const uint32_t u1 = R1 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(sl + R1 * 8)) : 0u;
        const int col = 8 * jn + 2 * q;
        float2 p0 = make_float2(0.f, 0.f), p1 = make_float2(0.f, 0.f);
        if (r0 < T) p0 = __ldg(reinterpret_cast<const float2*>(posg + r0 * kD + col));
        if (r1 < T) p1 = __ldg(reinterpret_cast<const float2*>(posg + r1 * kD + col));
        x[jn][0] = bf_lo(u0) + p0.x; x[jn][1] = bf_hi(u0) + p0.y;
        x[jn][2] = bf_lo(u1) + p1.x; x[jn][3] = bf_hi(u1) + p1.y;
        if (r0 == 0) { x[jn][0] = f32[L.cls / 4 + col] + p0.x; x[jn][1] = f32[L.cls / 4 + col + 1] + p0.y; }
        if (r0 >= T) { x[jn][0] = 0.f; x[jn][1] = 0.f; }
        if (r1 >= T) { x[jn][2] = 0.f; x[jn][3] = 0.f; }
        if (DROP) {   // pos_drop
          drop2(x[jn][0], x[jn][1], drop_key(b, 0, r0, 4 * jn + q), dcfg);
          drop2(x[jn][2], x[jn][3], drop_key(b, 0, r1, 4 * jn + q), dcfg);
        }
      }
    };
    float x[4][4];
    load_x0(x);
    // LN1 -> qkv of block 0: Q (scaled, log2 domain), K, V of this warp's rows into shared memory
    auto qkv0 = [&](const float (&xin)[4][4], bool dump) {
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wqkv);
      uint32_t A1[2][4];
      ln_to_afrag(xin, f32 + O0.ln1_g / 4, f32 + O0.ln1_b / 4, q, A1);
      if (dump) dump_afrag32(a.xln1[0], a.RTt, trow, q, A1);
#pragma unroll
      for (int jn = 0; jn < 12; ++jn) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const __nv_bfloat16* w = wqkv + (8 * jn + g) * LD + 16 * kk + 2 * q;
          mma16816(c, A1[kk], lds32(w), lds32(w + 8));
        }
        const float2 bb = *reinterpret_cast<const float2*>(f32 + O0.bqkv / 4 + 8 * jn + 2 * q);
        c[0] += bb.x; c[1] += bb.y; c[2] += bb.x; c[3] += bb.y;
        const int col = 8 * (jn & 3) + 2 * q;
        if (jn < 4) {
          *reinterpret_cast<uint32_t*>(Qs + r0 * LD + col) = pack_bf16(c[0] * qscale, c[1] * qscale);
          *reinterpret_cast<uint32_t*>(Qs + r1 * LD + col) = pack_bf16(c[2] * qscale, c[3] * qscale);
        } else {
          __nv_bfloat16* dst = jn < 8 ? Ks : Vs;
          *reinterpret_cast<uint32_t*>(dst + r0 * LD + col) = pack_bf16(c[0], c[1]);
          *reinterpret_cast<uint32_t*>(dst + r1 * LD + col) = pack_bf16(c[2], c[3]);
        }
      }
    };

    // =============================== F1: block 0 forward (with saves) ===========================
    {
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wqkv);
      const __nv_bfloat16* wproj = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wproj);
      const __nv_bfloat16* wfc1 = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wfc1);
      const __nv_bfloat16* wfc2 = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wfc2);
      qkv0(x, true);
      __syncthreads();
      // attention forward: all keys at once per head (scores in registers), stats saved
      uint32_t oa[2][4];
#pragma unroll 1
      for (int h = 0; h < kHeads; ++h) {
        float s[NT][4];
        float mx0 = -INFINITY, mx1 = -INFINITY;
        const uint32_t qa0 = lds32(Qs + r0 * LD + 8 * h + 2 * q), qa1 = lds32(Qs + r1 * LD + 8 * h + 2 * q);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
          mma1688(s[t], qa0, qa1, lds32(Ks + (8 * t + g) * LD + 8 * h + 2 * q));
          if (t >= NT - 2) {
            const int kc = 8 * t + 2 * q;
            if (kc >= T) { s[t][0] = -INFINITY; s[t][2] = -INFINITY; }
            if (kc + 1 >= T) { s[t][1] = -INFINITY; s[t][3] = -INFINITY; }
          }
          mx0 = fmaxf(mx0, fmaxf(s[t][0], s[t][1]));
          mx1 = fmaxf(mx1, fmaxf(s[t][2], s[t][3]));
        }
        mx0 = quad_max(mx0); mx1 = quad_max(mx1);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          s[t][0] = ex2(s[t][0] - mx0); s[t][1] = ex2(s[t][1] - mx0);
          s[t][2] = ex2(s[t][2] - mx1); s[t][3] = ex2(s[t][3] - mx1);
        }
        float oh[4] = {0.f, 0.f, 0.f, 0.f}, lacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < NW; ++kk) {
          uint32_t Pa[4];
          acc2_to_afrag(s[2 * kk], s[2 * kk + 1], Pa);
          uint32_t b0, b1;
          ldsm_x2_trans(b0, b1, Vs + (16 * kk + (lane & 15)) * LD + 8 * h);
          mma16816(oh, Pa, b0, b1);
          mma16816(lacc, Pa, 0x3F803F80u, 0x3F803F80u);   // row sums of (bf16) P on the tensor core
        }
        const float il0 = 1.f / lacc[0], il1 = 1.f / lacc[2];
        if (q == 0) {
          st_m[h * TP + r0] = mx0; st_m[h * TP + r1] = mx1;
          st_il[h * TP + r0] = il0; st_il[h * TP + r1] = il1;
        }
        // save O (bf16) for the backward of this block
        *reinterpret_cast<uint32_t*>(Os + r0 * LD + 8 * h + 2 * q) = pack_bf16(oh[0] * il0, oh[1] * il0);
        *reinterpret_cast<uint32_t*>(Os + r1 * LD + 8 * h + 2 * q) = pack_bf16(oh[2] * il1, oh[3] * il1);
      }
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        oa[kk][0] = lds32(Os + r0 * LD + 16 * kk + 2 * q);
        oa[kk][1] = lds32(Os + r1 * LD + 16 * kk + 2 * q);
        oa[kk][2] = lds32(Os + r0 * LD + 16 * kk + 8 + 2 * q);
        oa[kk][3] = lds32(Os + r1 * LD + 16 * kk + 8 + 2 * q);
      }
      dump_afrag32(a.xo0, a.RTt, trow, q, oa);
      // proj (+bias, +residual)
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const __nv_bfloat16* w = wproj + (8 * jn + g) * LD + 16 * kk + 2 * q;
          mma16816(c, oa[kk], lds32(w), lds32(w + 8));
        }
        const float2 bb = *reinterpret_cast<const float2*>(f32 + O0.bproj / 4 + 8 * jn + 2 * q);
        c[0] += bb.x; c[1] += bb.y; c[2] += bb.x; c[3] += bb.y;
        if (DROP) {
          drop2(c[0], c[1], drop_key(b, 1, r0, 4 * jn + q), dcfg);
          drop2(c[2], c[3], drop_key(b, 1, r1, 4 * jn + q), dcfg);
        }
        x[jn][0] += c[0]; x[jn][1] += c[1]; x[jn][2] += c[2]; x[jn][3] += c[3];
      }
      // MLP
      uint32_t A2[2][4];
      ln_to_afrag(x, f32 + O0.ln2_g / 4, f32 + O0.ln2_b / 4, q, A2);
      float acc2[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[j][e] = 0.f;
#pragma unroll 1
      for (int hk = 0; hk < kHidden / 16; ++hk) {
        float h0[4] = {0.f, 0.f, 0.f, 0.f}, h1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const __nv_bfloat16* w0 = wfc1 + (16 * hk + g) * LD + 16 * kk + 2 * q;
          const __nv_bfloat16* w1 = w0 + 8 * LD;
          mma16816(h0, A2[kk], lds32(w0), lds32(w0 + 8));
          mma16816(h1, A2[kk], lds32(w1), lds32(w1 + 8));
        }
        const float2 b0 = *reinterpret_cast<const float2*>(f32 + O0.bfc1 / 4 + 16 * hk + 2 * q);
        const float2 b1 = *reinterpret_cast<const float2*>(f32 + O0.bfc1 / 4 + 16 * hk + 8 + 2 * q);
        h0[0] = gelu_tanh_approx(h0[0] + b0.x); h0[1] = gelu_tanh_approx(h0[1] + b0.y); h0[2] = gelu_tanh_approx(h0[2] + b0.x); h0[3] = gelu_tanh_approx(h0[3] + b0.y);
        h1[0] = gelu_tanh_approx(h1[0] + b1.x); h1[1] = gelu_tanh_approx(h1[1] + b1.y); h1[2] = gelu_tanh_approx(h1[2] + b1.x); h1[3] = gelu_tanh_approx(h1[3] + b1.y);
        if (DROP) {
          drop2(h0[0], h0[1], drop_key(b, 2, r0, 8 * hk + q), dcfg);
          drop2(h0[2], h0[3], drop_key(b, 2, r1, 8 * hk + q), dcfg);
          drop2(h1[0], h1[1], drop_key(b, 2, r0, 8 * hk + 4 + q), dcfg);
          drop2(h1[2], h1[3], drop_key(b, 2, r1, 8 * hk + 4 + q), dcfg);
        }
        uint32_t Ha[4];
        Ha[0] = pack_bf16(h0[0], h0[1]);
        Ha[1] = pack_bf16(h0[2], h0[3]);
        Ha[2] = pack_bf16(h1[0], h1[1]);
        Ha[3] = pack_bf16(h1[2], h1[3]);
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          const __nv_bfloat16* w = wfc2 + (8 * jn + g) * kLdHid + 16 * hk + 2 * q;
          mma16816(acc2[jn], Ha, lds32(w), lds32(w + 8));
        }
      }
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        const float2 bb = *reinterpret_cast<const float2*>(f32 + O0.bfc2 / 4 + 8 * jn + 2 * q);
        acc2[jn][0] += bb.x; acc2[jn][1] += bb.y; acc2[jn][2] += bb.x; acc2[jn][3] += bb.y;
        if (DROP) {
          drop2(acc2[jn][0], acc2[jn][1], drop_key(b, 3, r0, 4 * jn + q), dcfg);
          drop2(acc2[jn][2], acc2[jn][3], drop_key(b, 3, r1, 4 * jn + q), dcfg);
        }
        x[jn][0] += acc2[jn][0]; x[jn][1] += acc2[jn][1]; x[jn][2] += acc2[jn][2]; x[jn][3] += acc2[jn][3];
      }
    }
    // x = x1: output of block 0 (rows >= T hold finite garbage; they never reach a valid row)

    // =============================== F2: block 1, K / V for every token =========================
    __syncthreads();   // K1s / V1s alias Qs / Ks: every warp must be done with block 0's attention
    {
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + O1.wqkv);
      uint32_t A1[2][4];
      ln_to_afrag(x, f32 + O1.ln1_g / 4, f32 + O1.ln1_b / 4, q, A1);
      dump_afrag32(a.xln1[1], a.RTt, trow, q, A1);
#pragma unroll
      for (int jn = 4; jn < 12; ++jn) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const __nv_bfloat16* w = wqkv + (8 * jn + g) * LD + 16 * kk + 2 * q;
          mma16816(c, A1[kk], lds32(w), lds32(w + 8));
        }
        const float2 bb = *reinterpret_cast<const float2*>(f32 + O1.bqkv / 4 + 8 * jn + 2 * q);
        const int col = 8 * (jn & 3) + 2 * q;
        __nv_bfloat16* dst = jn < 8 ? K1s : V1s;
        *reinterpret_cast<uint32_t*>(dst + r0 * LD + col) = pack_bf16(c[0] + bb.x, c[1] + bb.y);
        *reinterpret_cast<uint32_t*>(dst + r1 * LD + col) = pack_bf16(c[2] + bb.x, c[3] + bb.y);
      }
      if (warp == 0 && g == 0) {   // the cls row of x1 for the single-warp path
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          cs->dxres[8 * jn + 2 * q] = x[jn][0];
          cs->dxres[8 * jn + 2 * q + 1] = x[jn][1];
        }
      }
    }
    __syncthreads();

    // ====================== cls path of block 1 + head: forward and backward =======================
    // Vector-matrix steps run on warp 0 (lane = channel); everything that is "per key" (scores,
    // P.V, dP, dK / dV, dQ) is spread over all warps between CTA barriers.
    {
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + O1.wqkv);
      const __nv_bfloat16* wproj = reinterpret_cast<const __nv_bfloat16*>(smem + O1.wproj);
      const __nv_bfloat16* wfc1 = reinterpret_cast<const __nv_bfloat16*>(smem + O1.wfc1);
      const __nv_bfloat16* wfc2 = reinterpret_cast<const __nv_bfloat16*>(smem + O1.wfc2);
      const int hd = lane >> 3;                       // head of channel `lane`
      const int nthr = NW * 32;
      float x10 = 0.f;
      // ---- C1 (warp 0): LN1 -> q ----
      if (warp == 0) {
        x10 = cs->dxres[lane];
        const float mean = warp_sum(x10) * (1.f / kD);
        const float d = x10 - mean;
        const float rstd = rsqrtf(warp_sum(d * d) * (1.f / kD) + 1e-6f);
        const float y1 = __bfloat162float(__float2bfloat16_rn(d * rstd * f32[O1.ln1_g / 4 + lane] + f32[O1.ln1_b / 4 + lane]));
        float qv = f32[O1.bqkv / 4 + lane];
#pragma unroll
        for (int k2 = 0; k2 < kD / 2; ++k2) {
          const uint32_t wv = lds32(wqkv + lane * LD + 2 * k2);
          qv = fmaf(bf_lo(wv), __shfl_sync(0xffffffffu, y1, 2 * k2), qv);
          qv = fmaf(bf_hi(wv), __shfl_sync(0xffffffffu, y1, 2 * k2 + 1), qv);
        }
        cs->q[lane] = qv * qscale;
      }
      __syncthreads();
      // ---- C2 (all): scores of the cls query against every key, log2 domain ----
      for (int i = threadIdx.x; i < kHeads * TP; i += nthr) {
        const int h = i / TP, key = i - h * TP;
        float sv = -INFINITY;
        if (key < T) {
          const uint4 kv = *reinterpret_cast<const uint4*>(K1s + key * LD + 8 * h);
          const float* qq = cs->q + 8 * h;
          sv = qq[0] * bf_lo(kv.x) + qq[1] * bf_hi(kv.x) + qq[2] * bf_lo(kv.y) + qq[3] * bf_hi(kv.y) +
               qq[4] * bf_lo(kv.z) + qq[5] * bf_hi(kv.z) + qq[6] * bf_lo(kv.w) + qq[7] * bf_hi(kv.w);
        }
        cs->p[h][key] = sv;
      }
      __syncthreads();
      // ---- C3 (one warp per head): softmax ----
      for (int h = warp; h < kHeads; h += NW) {
        float mx = -INFINITY;
        for (int key = lane; key < TP; key += 32) mx = fmaxf(mx, cs->p[h][key]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float l = 0.f;
        for (int key = lane; key < TP; key += 32) {
          const float e = ex2(cs->p[h][key] - mx);
          cs->p[h][key] = e;
          l += e;
        }
        const float il = 1.f / warp_sum(l);
        for (int key = lane; key < TP; key += 32) cs->p[h][key] *= il;
      }
      __syncthreads();
      // ---- C4 (all): o = sum_key p * V, keys split over the warps ----
      {
        float part = 0.f;
        for (int key = warp; key < T; key += NW) part = fmaf(cs->p[hd][key], __bfloat162float(V1s[key * LD + lane]), part);
        cs->part[warp][lane] = part;
      }
      __syncthreads();
      // ---- C5 (warp 0): rest of the block for the cls row, head, and their backward ----
      if (warp == 0) {
        float ov = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) ov += cs->part[w][lane];
        float y = f32[O1.bproj / 4 + lane];
#pragma unroll
        for (int k2 = 0; k2 < kD / 2; ++k2) {
          const uint32_t wv = lds32(wproj + lane * LD + 2 * k2);
          y = fmaf(bf_lo(wv), __shfl_sync(0xffffffffu, ov, 2 * k2), y);
          y = fmaf(bf_hi(wv), __shfl_sync(0xffffffffu, ov, 2 * k2 + 1), y);
        }
        if (DROP) y = drop1(y, drop_key(b, 4, 0, lane >> 1), lane & 1, dcfg);
        const float xm = x10 + y;
        const float mean2 = warp_sum(xm) * (1.f / kD);
        const float d2 = xm - mean2;
        const float rstd2 = rsqrtf(warp_sum(d2 * d2) * (1.f / kD) + 1e-6f);
        const float xh2 = d2 * rstd2;
        const float y2 = xh2 * f32[O1.ln2_g / 4 + lane] + f32[O1.ln2_b / 4 + lane];
        float u[4], hv[4], hder[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = f32[O1.bfc1 / 4 + lane + 32 * i];
#pragma unroll
        for (int k2 = 0; k2 < kD / 2; ++k2) {
          const float ya = __shfl_sync(0xffffffffu, y2, 2 * k2), yb = __shfl_sync(0xffffffffu, y2, 2 * k2 + 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t wv = lds32(wfc1 + (lane + 32 * i) * LD + 2 * k2);
            u[i] = fmaf(bf_lo(wv), ya, fmaf(bf_hi(wv), yb, u[i]));
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          gelu_tanh_approx_grad(u[i], hv[i], hder[i]);
          if (DROP) {   // post-dropout activation; the mask (0 or 1/keep) also multiplies the derivative
            hv[i] = drop1(hv[i], drop_key(b, 5, 0, (lane + 32 * i) >> 1), lane & 1, dcfg);
            hder[i] = drop1(hder[i], drop_key(b, 5, 0, (lane + 32 * i) >> 1), lane & 1, dcfg);
          }
          cs->h[lane + 32 * i] = hv[i];
        }
        __syncwarp();
        float z0 = f32[O1.bfc2 / 4 + lane], z1 = 0.f;
#pragma unroll 8
        for (int k2 = 0; k2 < kHidden / 2; k2 += 2) {
          const uint2 wv = *reinterpret_cast<const uint2*>(wfc2 + lane * kLdHid + 2 * k2);
          const float4 hh = *reinterpret_cast<const float4*>(cs->h + 2 * k2);
          z0 = fmaf(bf_lo(wv.x), hh.x, fmaf(bf_hi(wv.x), hh.y, z0));
          z1 = fmaf(bf_lo(wv.y), hh.z, fmaf(bf_hi(wv.y), hh.w, z1));
        }
        float zz = z0 + z1;
        if (DROP) zz = drop1(zz, drop_key(b, 6, 0, lane >> 1), lane & 1, dcfg);
        const float x2 = xm + zz;
        const float meanf = warp_sum(x2) * (1.f / kD);
        const float df = x2 - meanf;
        const float rstdf = rsqrtf(warp_sum(df * df) * (1.f / kD) + 1e-6f);
        const float xhf = df * rstdf;
        const float cfin = xhf * f32[L.lnf_g / 4 + lane] + f32[L.lnf_b / 4 + lane];
        // ---------------- backward ----------------
        const float* dl = a.dlogits + (long long)b * K;
        float dc = 0.f;
        for (int k = 0; k < K; ++k) dc = fmaf(__ldg(dl + k), f32[L.whead / 4 + k * kD + lane], dc);
        a.c_xc[((long long)(lane >> 3) * a.RTc + b) * 8 + (lane & 7)] = __float2bfloat16_rn(cfin);
        for (int k = lane; k < ((K + 15) / 16) * 16; k += 32)
          a.c_dlog[((long long)(k >> 3) * a.RTc + b) * 8 + (k & 7)] = __float2bfloat16_rn(k < K ? __ldg(dl + k) : 0.f);
        gcls_lnf[0] += dc * xhf;
        gcls_lnf[1] += dc;
        float dgm = dc * f32[L.lnf_g / 4 + lane];
        float c1 = warp_sum(dgm) * (1.f / kD), c2 = warp_sum(dgm * xhf) * (1.f / kD);
        const float dx2 = rstdf * (dgm - c1 - xhf * c2);
        const float dz2 = DROP ? drop1(dx2, drop_key(b, 6, 0, lane >> 1), lane & 1, dcfg) : dx2;   // d(fc2 output)
        // fc2 backward: dh[j] = sum_o W2[o][j] dx2[o]; du = dh * gelu'(u)
        float dh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int o = 0; o < kD; ++o) {
          const float dv = __shfl_sync(0xffffffffu, dz2, o);
#pragma unroll
          for (int i = 0; i < 4; ++i) dh[i] = fmaf(__bfloat162float(wfc2[o * kLdHid + lane + 32 * i]), dv, dh[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float du = dh[i] * hder[i];
          cs->du[lane + 32 * i] = du;
          const int j = lane + 32 * i;
          a.c_xh[((long long)(j >> 3) * a.RTc + b) * 8 + (j & 7)] = __float2bfloat16_rn(hv[i]);
          a.c_dh[((long long)(j >> 3) * a.RTc + b) * 8 + (j & 7)] = __float2bfloat16_rn(du);
        }
        a.c_dxb[((long long)(lane >> 3) * a.RTc + b) * 8 + (lane & 7)] = __float2bfloat16_rn(dz2);
        a.c_xln2[((long long)(lane >> 3) * a.RTc + b) * 8 + (lane & 7)] = __float2bfloat16_rn(y2);
        __syncwarp();
        // fc1 backward: dy2[i] = sum_j W1[j][i] du[j]  (four independent chains)
        float dy2a = 0.f, dy2b = 0.f, dy2c = 0.f, dy2d = 0.f;
#pragma unroll 8
        for (int j = 0; j < kHidden; j += 4) {
          dy2a = fmaf(__bfloat162float(wfc1[j * LD + lane]), cs->du[j], dy2a);
          dy2b = fmaf(__bfloat162float(wfc1[(j + 1) * LD + lane]), cs->du[j + 1], dy2b);
          dy2c = fmaf(__bfloat162float(wfc1[(j + 2) * LD + lane]), cs->du[j + 2], dy2c);
          dy2d = fmaf(__bfloat162float(wfc1[(j + 3) * LD + lane]), cs->du[j + 3], dy2d);
        }
        const float dy2 = (dy2a + dy2b) + (dy2c + dy2d);
        gcls_ln2[0] += dy2 * xh2;
        gcls_ln2[1] += dy2;
        dgm = dy2 * f32[O1.ln2_g / 4 + lane];
        c1 = warp_sum(dgm) * (1.f / kD);
        c2 = warp_sum(dgm * xh2) * (1.f / kD);
        const float dxm = dx2 + rstd2 * (dgm - c1 - xh2 * c2);
        const float dpo = DROP ? drop1(dxm, drop_key(b, 4, 0, lane >> 1), lane & 1, dcfg) : dxm;   // d(proj output)
        a.c_dxa[((long long)(lane >> 3) * a.RTc + b) * 8 + (lane & 7)] = __float2bfloat16_rn(dpo);
        a.c_xo[((long long)(lane >> 3) * a.RTc + b) * 8 + (lane & 7)] = __float2bfloat16_rn(ov);
        // proj backward: do[i] = sum_o Wp[o][i] dxm[o]
        float dov = 0.f;
#pragma unroll 8
        for (int o = 0; o < kD; ++o)
          dov = fmaf(__bfloat162float(wproj[o * LD + lane]), __shfl_sync(0xffffffffu, dpo, o), dov);
        cs->doo[lane] = dov;
        float delta = dov * ov;   // per head: sum over its 8 channels
        delta += __shfl_xor_sync(0xffffffffu, delta, 1);
        delta += __shfl_xor_sync(0xffffffffu, delta, 2);
        delta += __shfl_xor_sync(0xffffffffu, delta, 4);
        if ((lane & 7) == 0) cs->dlt[hd] = delta;
        cs->dxres[lane] = dxm;      // residual gradient reaching x1[cls]
      }
      __syncthreads();
      // ---- C6 (all): single-query attention backward per (head, key): ds, dK, dV ----
      for (int i = threadIdx.x; i < kHeads * TP; i += nthr) {
        const int h = i / TP, key = i - h * TP;
        const uint4 vv = *reinterpret_cast<const uint4*>(V1s + key * LD + 8 * h);
        const float* dd = cs->doo + 8 * h;
        const float* qq = cs->q + 8 * h;
        const float dp = dd[0] * bf_lo(vv.x) + dd[1] * bf_hi(vv.x) + dd[2] * bf_lo(vv.y) + dd[3] * bf_hi(vv.y) +
                         dd[4] * bf_lo(vv.z) + dd[5] * bf_hi(vv.z) + dd[6] * bf_lo(vv.w) + dd[7] * bf_hi(vv.w);
        const float pk = key < T ? cs->p[h][key] : 0.f;
        const float dsv = pk * (dp - cs->dlt[h]);
        cs->ds[h][key] = dsv;
        const float dsk = dsv * kLn2;   // dK = ds * scale * q = ds * qhat / log2(e)
        *reinterpret_cast<uint4*>(dK1s + key * LD + 8 * h) =
            make_uint4(pack_bf16(dsk * qq[0], dsk * qq[1]), pack_bf16(dsk * qq[2], dsk * qq[3]),
                       pack_bf16(dsk * qq[4], dsk * qq[5]), pack_bf16(dsk * qq[6], dsk * qq[7]));
        *reinterpret_cast<uint4*>(dV1s + key * LD + 8 * h) =
            make_uint4(pack_bf16(pk * dd[0], pk * dd[1]), pack_bf16(pk * dd[2], pk * dd[3]),
                       pack_bf16(pk * dd[4], pk * dd[5]), pack_bf16(pk * dd[6], pk * dd[7]));
      }
      __syncthreads();
      // ---- C7 (all): dq[d] = scale * sum_key ds[h(d)][key] * K[key][d], keys split over the warps ----
      {
        float part = 0.f;
        for (int key = warp; key < T; key += NW) part = fmaf(cs->ds[hd][key], __bfloat162float(K1s[key * LD + lane]), part);
        cs->part[warp][lane] = part;
      }
      __syncthreads();
      if (warp == 0) {
        float dqv = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) dqv += cs->part[w][lane];
        cs->dq[lane] = dqv * kScale;
      }
    }
    __syncthreads();

    // ====================== B2: block 1, all rows: dqkv -> dLN1 -> dx1 ===========================
    float dx[4][4];
    {
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + O1.wqkv);
      uint32_t Aq[6][4];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {   // dQ: only the cls row
        Aq[kk][0] = Aq[kk][1] = Aq[kk][2] = Aq[kk][3] = 0u;
        if (r0 == 0) {
          Aq[kk][0] = pack_bf16(cs->dq[16 * kk + 2 * q], cs->dq[16 * kk + 2 * q + 1]);
          Aq[kk][2] = pack_bf16(cs->dq[16 * kk + 8 + 2 * q], cs->dq[16 * kk + 8 + 2 * q + 1]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const __nv_bfloat16* pk = dK1s + 16 * kk + 2 * q;
        const __nv_bfloat16* pv = dV1s + 16 * kk + 2 * q;
        Aq[2 + kk][0] = lds32(pk + r0 * LD); Aq[2 + kk][1] = lds32(pk + r1 * LD);
        Aq[2 + kk][2] = lds32(pk + r0 * LD + 8); Aq[2 + kk][3] = lds32(pk + r1 * LD + 8);
        Aq[4 + kk][0] = lds32(pv + r0 * LD); Aq[4 + kk][1] = lds32(pv + r1 * LD);
        Aq[4 + kk][2] = lds32(pv + r0 * LD + 8); Aq[4 + kk][3] = lds32(pv + r1 * LD + 8);
      }
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) {
        dump2(a.dqkv[1], a.RTt, trow, 16 * kk + 2 * q, Aq[kk][0]);
        dump2(a.dqkv[1], a.RTt, trow + 8, 16 * kk + 2 * q, Aq[kk][1]);
        dump2(a.dqkv[1], a.RTt, trow, 16 * kk + 8 + 2 * q, Aq[kk][2]);
        dump2(a.dqkv[1], a.RTt, trow + 8, 16 * kk + 8 + 2 * q, Aq[kk][3]);
      }
      float dy[4][4];
      gemm_dgrad32<6>(Aq, wqkv, LD, lane, dy);
      ln_backward(x, dy, f32 + O1.ln1_g / 4, q, dx, false, gln[2]);
      if (r0 == 0) {
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          dx[jn][0] += cs->dxres[8 * jn + 2 * q];
          dx[jn][1] += cs->dxres[8 * jn + 2 * q + 1];
        }
      }
    }
    __syncthreads();   // dK1s is reused as dOs below

    // ====================== B1: block 0 backward ===================================================
    {
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wqkv);
      const __nv_bfloat16* wproj = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wproj);
      const __nv_bfloat16* wfc1 = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wfc1);
      const __nv_bfloat16* wfc2 = reinterpret_cast<const __nv_bfloat16*>(smem + O0.wfc2);
      // recompute Q / K / V of block 0 (their arrays were reused by block 1) and x_mid = x0 + proj(O) + b
      load_x0(x);
      qkv0(x, false);
      uint32_t oa[2][4];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        oa[kk][0] = lds32(Os + r0 * LD + 16 * kk + 2 * q);
        oa[kk][1] = lds32(Os + r1 * LD + 16 * kk + 2 * q);
        oa[kk][2] = lds32(Os + r0 * LD + 16 * kk + 8 + 2 * q);
        oa[kk][3] = lds32(Os + r1 * LD + 16 * kk + 8 + 2 * q);
      }
      float xm[4][4];
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const __nv_bfloat16* w = wproj + (8 * jn + g) * LD + 16 * kk + 2 * q;
          mma16816(c, oa[kk], lds32(w), lds32(w + 8));
        }
        const float2 bb = *reinterpret_cast<const float2*>(f32 + O0.bproj / 4 + 8 * jn + 2 * q);
        c[0] += bb.x; c[1] += bb.y; c[2] += bb.x; c[3] += bb.y;
        if (DROP) {
          drop2(c[0], c[1], drop_key(b, 1, r0, 4 * jn + q), dcfg);
          drop2(c[2], c[3], drop_key(b, 1, r1, 4 * jn + q), dcfg);
        }
        xm[jn][0] = x[jn][0] + c[0]; xm[jn][1] = x[jn][1] + c[1];
        xm[jn][2] = x[jn][2] + c[2]; xm[jn][3] = x[jn][3] + c[3];
      }
      // ---- MLP backward ----
      uint32_t A2[2][4], Adx[2][4];
      ln_to_afrag(xm, f32 + O0.ln2_g / 4, f32 + O0.ln2_b / 4, q, A2);
      dump_afrag32(a.xln2_0, a.RTt, trow, q, A2);
      {
        float d3[4][4];    // d(fc2 output) = dropout mask (site 3) applied to d x1
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
#pragma unroll
          for (int e = 0; e < 4; ++e) d3[jn][e] = dx[jn][e];
          if (DROP) {
            drop2(d3[jn][0], d3[jn][1], drop_key(b, 3, r0, 4 * jn + q), dcfg);
            drop2(d3[jn][2], d3[jn][3], drop_key(b, 3, r1, 4 * jn + q), dcfg);
          }
        }
        acc2_to_afrag(d3[0], d3[1], Adx[0]);
        acc2_to_afrag(d3[2], d3[3], Adx[1]);
      }
      dump_afrag32(a.dxb0, a.RTt, trow, q, Adx);
      float dln2[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dln2[j][e] = 0.f;
#pragma unroll 1
      for (int hk = 0; hk < kHidden / 16; ++hk) {
        float h0[4] = {0.f, 0.f, 0.f, 0.f}, h1[4] = {0.f, 0.f, 0.f, 0.f};
        float e0[4] = {0.f, 0.f, 0.f, 0.f}, e1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const __nv_bfloat16* w0 = wfc1 + (16 * hk + g) * LD + 16 * kk + 2 * q;
          const __nv_bfloat16* w1 = w0 + 8 * LD;
          mma16816(h0, A2[kk], lds32(w0), lds32(w0 + 8));
          mma16816(h1, A2[kk], lds32(w1), lds32(w1 + 8));
          uint32_t b0, b1;
          ldsm_x2_trans(b0, b1, wfc2 + (16 * kk + (lane & 15)) * kLdHid + 16 * hk);
          mma16816(e0, Adx[kk], b0, b1);
          ldsm_x2_trans(b0, b1, wfc2 + (16 * kk + (lane & 15)) * kLdHid + 16 * hk + 8);
          mma16816(e1, Adx[kk], b0, b1);
        }
        const float2 b0 = *reinterpret_cast<const float2*>(f32 + O0.bfc1 / 4 + 16 * hk + 2 * q);
        const float2 b1 = *reinterpret_cast<const float2*>(f32 + O0.bfc1 / 4 + 16 * hk + 8 + 2 * q);
        float hv0[4], hv1[4], du0[4], du1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float der;
          gelu_tanh_approx_grad(h0[e] + ((e & 1) ? b0.y : b0.x), hv0[e], der);
          du0[e] = e0[e] * der;
          gelu_tanh_approx_grad(h1[e] + ((e & 1) ? b1.y : b1.x), hv1[e], der);
          du1[e] = e1[e] * der;
        }
        if (DROP) {
          const uint32_t k00 = drop_key(b, 2, r0, 8 * hk + q), k01 = drop_key(b, 2, r1, 8 * hk + q);
          const uint32_t k10 = drop_key(b, 2, r0, 8 * hk + 4 + q), k11 = drop_key(b, 2, r1, 8 * hk + 4 + q);
          drop2(hv0[0], hv0[1], k00, dcfg); drop2(hv0[2], hv0[3], k01, dcfg);
          drop2(hv1[0], hv1[1], k10, dcfg); drop2(hv1[2], hv1[3], k11, dcfg);
          drop2(du0[0], du0[1], k00, dcfg); drop2(du0[2], du0[3], k01, dcfg);
          drop2(du1[0], du1[1], k10, dcfg); drop2(du1[2], du1[3], k11, dcfg);
        }
        uint32_t Ah[4], Ad[4];
        acc2_to_afrag(hv0, hv1, Ah);
        acc2_to_afrag(du0, du1, Ad);
        dump2(a.xh0, a.RTt, trow, 16 * hk + 2 * q, Ah[0]);
        dump2(a.xh0, a.RTt, trow + 8, 16 * hk + 2 * q, Ah[1]);
        dump2(a.xh0, a.RTt, trow, 16 * hk + 8 + 2 * q, Ah[2]);
        dump2(a.xh0, a.RTt, trow + 8, 16 * hk + 8 + 2 * q, Ah[3]);
        dump2(a.dh0, a.RTt, trow, 16 * hk + 2 * q, Ad[0]);
        dump2(a.dh0, a.RTt, trow + 8, 16 * hk + 2 * q, Ad[1]);
        dump2(a.dh0, a.RTt, trow, 16 * hk + 8 + 2 * q, Ad[2]);
        dump2(a.dh0, a.RTt, trow + 8, 16 * hk + 8 + 2 * q, Ad[3]);
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          uint32_t c0, c1;
          ldsm_x2_trans(c0, c1, wfc1 + (16 * hk + (lane & 15)) * LD + 8 * jn);
          mma16816(dln2[jn], Ad, c0, c1);
        }
      }
      ln_backward(xm, dln2, f32 + O0.ln2_g / 4, q, dx, true, gln[1]);   // dx = d x_mid
      // ---- attention block backward ----
      uint32_t Adm[2][4];
      {
        float d1[4][4];    // d(proj output) = dropout mask (site 1) applied to d x_mid
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
#pragma unroll
          for (int e = 0; e < 4; ++e) d1[jn][e] = dx[jn][e];
          if (DROP) {
            drop2(d1[jn][0], d1[jn][1], drop_key(b, 1, r0, 4 * jn + q), dcfg);
            drop2(d1[jn][2], d1[jn][3], drop_key(b, 1, r1, 4 * jn + q), dcfg);
          }
        }
        acc2_to_afrag(d1[0], d1[1], Adm[0]);
        acc2_to_afrag(d1[2], d1[3], Adm[1]);
      }
      dump_afrag32(a.dxa0, a.RTt, trow, q, Adm);
      float dO[4][4];
      gemm_dgrad32<2>(Adm, wproj, LD, lane, dO);
#pragma unroll
      for (int h = 0; h < kHeads; ++h) {
        const uint32_t d0 = pack_bf16(dO[h][0], dO[h][1]), d1 = pack_bf16(dO[h][2], dO[h][3]);
        *reinterpret_cast<uint32_t*>(dOs + r0 * LD + 8 * h + 2 * q) = d0;
        *reinterpret_cast<uint32_t*>(dOs + r1 * LD + 8 * h + 2 * q) = d1;
        const uint32_t o0 = oa[h >> 1][(h & 1) * 2 + 0], o1 = oa[h >> 1][(h & 1) * 2 + 1];
        const float dl0 = quad_sum(dO[h][0] * bf_lo(o0) + dO[h][1] * bf_hi(o0));
        const float dl1 = quad_sum(dO[h][2] * bf_lo(o1) + dO[h][3] * bf_hi(o1));
        if (q == 0) { st_dl[h * TP + r0] = dl0; st_dl[h * TP + r1] = dl1; }
      }
      __syncthreads();
      // per head: dQ / dK / dV tiles of this warp's rows (nothing indexed dynamically in registers)
#pragma unroll 1
      for (int h = 0; h < kHeads; ++h) {
        float dq[4] = {0.f, 0.f, 0.f, 0.f}, dk[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
        // ---- sweep A: this warp's rows as queries -> dQ ----
        {
          const uint32_t qa0 = lds32(Qs + r0 * LD + 8 * h + 2 * q), qa1 = lds32(Qs + r1 * LD + 8 * h + 2 * q);
          const uint32_t da0 = lds32(dOs + r0 * LD + 8 * h + 2 * q), da1 = lds32(dOs + r1 * LD + 8 * h + 2 * q);
          const float m0 = st_m[h * TP + r0], m1 = st_m[h * TP + r1];
          const float il0 = st_il[h * TP + r0], il1 = st_il[h * TP + r1];
          const float dl0 = st_dl[h * TP + r0], dl1 = st_dl[h * TP + r1];
#pragma unroll 2
          for (int kk = 0; kk < NW; ++kk) {
            float ds[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int t = 2 * kk + u;
              float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
              mma1688(s, qa0, qa1, lds32(Ks + (8 * t + g) * LD + 8 * h + 2 * q));
              mma1688(dp, da0, da1, lds32(Vs + (8 * t + g) * LD + 8 * h + 2 * q));
              const int kc = 8 * t + 2 * q;
              const float p0 = kc < T ? ex2(s[0] - m0) * il0 : 0.f, p1 = kc + 1 < T ? ex2(s[1] - m0) * il0 : 0.f;
              const float p2 = kc < T ? ex2(s[2] - m1) * il1 : 0.f, p3 = kc + 1 < T ? ex2(s[3] - m1) * il1 : 0.f;
              ds[u][0] = p0 * (dp[0] - dl0); ds[u][1] = p1 * (dp[1] - dl0);
              ds[u][2] = p2 * (dp[2] - dl1); ds[u][3] = p3 * (dp[3] - dl1);
            }
            uint32_t Ads[4], b0, b1;
            acc2_to_afrag(ds[0], ds[1], Ads);
            ldsm_x2_trans(b0, b1, Ks + (16 * kk + (lane & 15)) * LD + 8 * h);
            mma16816(dq, Ads, b0, b1);
          }
        }
        // ---- sweep B: this warp's rows as keys -> dK, dV ----
        {
          const uint32_t ka0 = lds32(Ks + r0 * LD + 8 * h + 2 * q), ka1 = lds32(Ks + r1 * LD + 8 * h + 2 * q);
          const uint32_t va0 = lds32(Vs + r0 * LD + 8 * h + 2 * q), va1 = lds32(Vs + r1 * LD + 8 * h + 2 * q);
          const bool kv0 = r0 < T, kv1 = r1 < T;
#pragma unroll 2
          for (int kk = 0; kk < NW; ++kk) {
            float ds[2][4], pt[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int t = 2 * kk + u;
              float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
              mma1688(s, ka0, ka1, lds32(Qs + (8 * t + g) * LD + 8 * h + 2 * q));
              mma1688(dp, va0, va1, lds32(dOs + (8 * t + g) * LD + 8 * h + 2 * q));
              const int qc = 8 * t + 2 * q;       // query columns qc, qc + 1
              const float2 mm = *reinterpret_cast<const float2*>(st_m + h * TP + qc);
              const float2 ii = *reinterpret_cast<const float2*>(st_il + h * TP + qc);
              const float2 dd = *reinterpret_cast<const float2*>(st_dl + h * TP + qc);
              pt[u][0] = kv0 ? ex2(s[0] - mm.x) * ii.x : 0.f; pt[u][1] = kv0 ? ex2(s[1] - mm.y) * ii.y : 0.f;
              pt[u][2] = kv1 ? ex2(s[2] - mm.x) * ii.x : 0.f; pt[u][3] = kv1 ? ex2(s[3] - mm.y) * ii.y : 0.f;
              ds[u][0] = pt[u][0] * (dp[0] - dd.x); ds[u][1] = pt[u][1] * (dp[1] - dd.y);
              ds[u][2] = pt[u][2] * (dp[2] - dd.x); ds[u][3] = pt[u][3] * (dp[3] - dd.y);
            }
            uint32_t Ads[4], Apt[4], b0, b1;
            acc2_to_afrag(ds[0], ds[1], Ads);
            acc2_to_afrag(pt[0], pt[1], Apt);
            ldsm_x2_trans(b0, b1, Qs + (16 * kk + (lane & 15)) * LD + 8 * h);
            mma16816(dk, Ads, b0, b1);
            ldsm_x2_trans(b0, b1, dOs + (16 * kk + (lane & 15)) * LD + 8 * h);
            mma16816(dv, Apt, b0, b1);
          }
        }
        // scale: dq = scale * dS K ; dk = dS^T qhat / log2(e)
        const int col = 8 * h + 2 * q;
        // straight into the operand dump of the qkv weight gradient; each thread re-reads its own
        // writes below (same-thread global RAW), so no staging array is needed
        dump2(a.dqkv[0], a.RTt, trow, col, pack_bf16(dq[0] * kScale, dq[1] * kScale));
        dump2(a.dqkv[0], a.RTt, trow + 8, col, pack_bf16(dq[2] * kScale, dq[3] * kScale));
        dump2(a.dqkv[0], a.RTt, trow, 32 + col, pack_bf16(dk[0] * kLn2, dk[1] * kLn2));
        dump2(a.dqkv[0], a.RTt, trow + 8, 32 + col, pack_bf16(dk[2] * kLn2, dk[3] * kLn2));
        dump2(a.dqkv[0], a.RTt, trow, 64 + col, pack_bf16(dv[0], dv[1]));
        dump2(a.dqkv[0], a.RTt, trow + 8, 64 + col, pack_bf16(dv[2], dv[3]));
      }
      uint32_t Aq[6][4];
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) {
        Aq[kk][0] = load2(a.dqkv[0], a.RTt, trow, 16 * kk + 2 * q);
        Aq[kk][1] = load2(a.dqkv[0], a.RTt, trow + 8, 16 * kk + 2 * q);
        Aq[kk][2] = load2(a.dqkv[0], a.RTt, trow, 16 * kk + 8 + 2 * q);
        Aq[kk][3] = load2(a.dqkv[0], a.RTt, trow + 8, 16 * kk + 8 + 2 * q);
      }
      float dy[4][4];
      gemm_dgrad32<6>(Aq, wqkv, LD, lane, dy);
      ln_backward(x, dy, f32 + O0.ln1_g / 4, q, dx, true, gln[0]);      // dx = d x0
    }

    // ------------------------------ outputs: d zf, d pos / d cls ------------------------------
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      if (r0 >= T) { dx[jn][0] = 0.f; dx[jn][1] = 0.f; }
      if (r1 >= T) { dx[jn][2] = 0.f; dx[jn][3] = 0.f; }
      if (DROP) {   // back through pos_drop: d(tokens + pos)
        drop2(dx[jn][0], dx[jn][1], drop_key(b, 0, r0, 4 * jn + q), dcfg);
        drop2(dx[jn][2], dx[jn][3], drop_key(b, 0, r1, 4 * jn + q), dcfg);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) gpos[jn][e] += dx[jn][e];
      __nv_bfloat16* sl = a.dzf + (long long)jn * a.RT * 8 + 2 * q;
      if (R0 >= 0) *reinterpret_cast<uint32_t*>(sl + R0 * 8) = pack_bf16(dx[jn][0], dx[jn][1]);
      if (R1 >= 0) *reinterpret_cast<uint32_t*>(sl + R1 * 8) = pack_bf16(dx[jn][2], dx[jn][3]);
    }
    __syncthreads();   // shared arrays are rewritten by the next patch
  }

  // ---------------------------------- flush the small gradients ----------------------------------
#pragma unroll
  for (int jn = 0; jn < 4; ++jn) {
    const int col = 8 * jn + 2 * q;
    if (r0 < T) { atomicAdd(a.g_pos + r0 * kD + col, gpos[jn][0]); atomicAdd(a.g_pos + r0 * kD + col + 1, gpos[jn][1]); }
    if (r1 < T) { atomicAdd(a.g_pos + r1 * kD + col, gpos[jn][2]); atomicAdd(a.g_pos + r1 * kD + col + 1, gpos[jn][3]); }
    if (r0 == 0) { atomicAdd(a.g_cls + col, gpos[jn][0]); atomicAdd(a.g_cls + col + 1, gpos[jn][1]); }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = g_sum(gln[i][w][k]);
        if (g == 0) atomicAdd(a.g_ln[i][w] + 8 * (k >> 1) + 2 * q + (k & 1), v);
      }
  if (warp == 0) {
    atomicAdd(a.g_ln[3][0] + lane, gcls_ln2[0]);
    atomicAdd(a.g_ln[3][1] + lane, gcls_ln2[1]);
    atomicAdd(a.g_ln[4][0] + lane, gcls_lnf[0]);
    atomicAdd(a.g_ln[4][1] + lane, gcls_lnf[1]);
  }
}

template <int NW, bool DROP>
static int launch_bwd(const TBArgs& a, cudaStream_t stream) {
  constexpr int TP = 16 * NW;
  const size_t smem = (size_t)a.L.pos + 5 * (size_t)TP * kLdD * 2 + 3 * (size_t)kHeads * TP * 4 + sizeof(ClsScratch) + 16;
  int dev = 0, max_smem = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (smem > (size_t)max_smem || TP > 256) return VC_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(transformer_bwd_kernel<NW, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VC_ERR_CUDA;
  int blocks = num_sms < a.n ? num_sms : a.n;
  transformer_bwd_kernel<NW, DROP><<<blocks, NW * 32, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

// tok_dumps[10]: xln1_0, dqkv_0, xo0, dxa0, xln2_0, dh0, xh0, dxb0, xln1_1, dqkv_1 (token-row space);
// cls_dumps[8]: c_xo, c_dxa, c_xln2, c_dh, c_xh, c_dxb, c_xc, c_dlog (compact space);
// small[12]: gamma / beta grads of ln1_0, ln2_0, ln1_1, ln2_1, final norm, then cls, pos.
int transformer_bwd_launch(const void* zf, const void* tparams, const float* dlogits, void* dzf, void* const* tok_dumps,
                           void* const* cls_dumps, float* const* small, int n_patches, int P, int K, unsigned int drop_thr,
                           const unsigned int* drop_seed, cudaStream_t stream) {
  if (n_patches <= 0 || P < 1 || K < 1 || K > 64 || (drop_thr && !drop_seed) || drop_thr >= 65536u) return VC_ERR_ARG;
  const int T = P * P + 1, NW = (T + 15) / 16, TP = 16 * NW;
  TBArgs a;
  a.zf = (const __nv_bfloat16*)zf;
  a.blob = (const uint8_t*)tparams;
  a.dlogits = dlogits;
  a.dzf = (__nv_bfloat16*)dzf;
  __nv_bfloat16* const* td = (__nv_bfloat16* const*)tok_dumps;
  a.xln1[0] = td[0]; a.dqkv[0] = td[1]; a.xo0 = td[2]; a.dxa0 = td[3]; a.xln2_0 = td[4]; a.dh0 = td[5];
  a.xh0 = td[6]; a.dxb0 = td[7]; a.xln1[1] = td[8]; a.dqkv[1] = td[9];
  __nv_bfloat16* const* cd = (__nv_bfloat16* const*)cls_dumps;
  a.c_xo = cd[0]; a.c_dxa = cd[1]; a.c_xln2 = cd[2]; a.c_dh = cd[3]; a.c_xh = cd[4]; a.c_dxb = cd[5]; a.c_xc = cd[6];
  a.c_dlog = cd[7];
  for (int i = 0; i < 5; ++i) { a.g_ln[i][0] = small[2 * i]; a.g_ln[i][1] = small[2 * i + 1]; }
  a.g_cls = small[10];
  a.g_pos = small[11];
  a.RT = sps_rows(n_patches, P);
  a.RTt = ((long long)n_patches * TP + 127) / 128 * 128;
  a.RTc = ((long long)n_patches + 127) / 128 * 128;
  a.n = n_patches;
  a.P = P;
  a.K = K;
  a.T = T;
  a.drop_thr = drop_thr;
  a.drop_seed = drop_seed;
  a.L = tlayout(P, K);
  if (drop_thr) {
    switch (NW) {
#define VC_CASE(N) case N: return launch_bwd<N, true>(a, stream);
      VC_CASE(1) VC_CASE(2) VC_CASE(3) VC_CASE(4) VC_CASE(5) VC_CASE(6) VC_CASE(7) VC_CASE(8) VC_CASE(11) VC_CASE(15)
#undef VC_CASE
      default: return VC_ERR_UNSUPPORTED;
    }
  }
  switch (NW) {
#define VC_CASE(N) case N: return launch_bwd<N, false>(a, stream);
    VC_CASE(1) VC_CASE(2) VC_CASE(3) VC_CASE(4) VC_CASE(5) VC_CASE(6) VC_CASE(7) VC_CASE(8)
    VC_CASE(11) VC_CASE(15)   // P = 13, 15 (odd patch sizes: the centre pixel is classified)
#undef VC_CASE
    default: return VC_ERR_UNSUPPORTED;   // even P > 11
  }
}

}  // namespace vc
