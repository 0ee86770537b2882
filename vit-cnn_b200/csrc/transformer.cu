// Token stage of the ViT-CNN hybrid as ONE kernel per chunk of patches: 1x1 fusion conv
// (+folded BN, ReLU) -> cls token + pos-embed -> 2 x [LN -> MHSA -> +res -> LN -> MLP(GELU)
// -> +res] -> LN -> head on the cls token.  Spec: SURVEY.md App. A (R0), following
// model/compare_method/vit/timm/models/vision_transformer.py:57-105 (Attention), :123-166
// (Block), :598-629 (cls/pos), :680-701 (norm/head) and timm/layers/mlp.py:13-47.
//
// One CTA owns one patch at a time; warp w owns token rows 16w..16w+15 and keeps their
// residual stream in registers in the mma.sync accumulator layout for the whole stage, so
// LayerNorm is a GEMM prologue (quad shuffles), bias/GELU/residual are GEMM epilogues and
// QK^T -> online softmax (exp2, warp-shuffle row reductions) -> PV never leaves registers;
// only K and V^T of the current layer go through shared memory (whole token set resident).
#include <math.h>
#include "vc_common.cuh"
#include "vc_kernels.h"
#include "vc_tparams.h"
#include "vc_tokens.cuh"

namespace vc {

struct TArgs {
  const __nv_bfloat16* f;  // [8][RT][8] : slices 0-3 HSI stem, 4-7 LiDAR stem
                           // (prefused: [4][RT][8] = relu(bn(fusion conv)), the tokens themselves)
  int prefused;            // training path: BatchNorm needs batch statistics, so the fusion conv
                           // runs as its own launch and this kernel starts at "+ pos"
  const uint8_t* blob;
  float* logits;               // [n][K] or scattered through out_index
  const long long* out_index;  // nullable: logits row of patch b
  unsigned char* argmax_map;   // nullable: argmax written at out_index[b]
  long long RT;
  int n_patches, P, K, T;
  uint32_t drop_thr;            // training-mode dropout (0 = off): see DropCfg
  const uint32_t* drop_seed;    // device: seed of this step
  TLayout L;  // parameter-blob offsets (kernel-argument space: constant-bank reads)
};

// Hand-off buffer between the token warps and the cls-tail warp (double buffered).
template <int NW>
struct TailBuf {
  float o[NW][kD];       // per-warp partial sum_j p_j * v_j of the cls query, all heads
  float l[NW][kHeads];   // per-warp partial sum_j p_j
  float x0[kD];          // residual stream of the cls token before the last attention
};

// NW token warps (rows 16w..16w+15) + 1 tail warp.  The last block only needs the cls
// token's output (the head reads x[:, 0], vision_transformer.py:692-701), so there the token
// warps only produce K / V for every token and the cls query's attention partials; the tail
// warp finishes the cls row (proj, MLP, final LN, head) while the token warps already work
// on the next patch.
template <int NW, bool DROP>
__global__ void __launch_bounds__((NW + 1) * 32, (NW <= 8 ? 2 : 1)) transformer_fwd_kernel(TArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int TP = 16 * NW;    // padded token count
  constexpr int LDV = TP + 8;    // pitch of V^T rows
  constexpr int NKB = (2 * NW + 7) / 8;  // key blocks of 64
  constexpr int NMAIN = NW * 32, NALL = (NW + 1) * 32;
  enum { BAR_MAIN = 1, BAR_READY0 = 2, BAR_READY1 = 3, BAR_FREE0 = 4, BAR_FREE1 = 5 };
  const TLayout& L = a.L;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem + L.total);  // [TP][40]
  __nv_bfloat16* Vt = Ks + TP * kLdD;                                    // [32][LDV]
  float* q0_s = reinterpret_cast<float*>(Vt + kD * LDV);                 // [32] cls query (scaled)
  float* max_s = q0_s + kD;                                              // [NW][4]
  TailBuf<NW>* tb = reinterpret_cast<TailBuf<NW>*>(max_s + NW * kHeads); // [2]
  float* h_s = reinterpret_cast<float*>(tb + 2);                         // [128] tail: MLP hidden
  float* logit_s = h_s + kHidden;                                        // [K]

  {  // parameters -> shared memory (once per persistent CTA)
    const uint4* src = reinterpret_cast<const uint4*>(a.blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < L.total / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int T = a.T, P = a.P;
  DropCfg dc;
  dc.thr = DROP ? a.drop_thr : 0u;
  dc.seed = DROP ? __ldg(a.drop_seed) : 0u;
  dc.inv_keep = 65536.f / (65536.f - (float)dc.thr);
  const int my_patches = (a.n_patches - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const TLayerOff& OL = L.layer[kLayers - 1];   // the cls-only block

  if (warp == NW) {
    // ======================= tail warp: finish the cls token, one channel per lane ==============
    const __nv_bfloat16* wproj = reinterpret_cast<const __nv_bfloat16*>(smem + OL.wproj);
    const __nv_bfloat16* wfc1 = reinterpret_cast<const __nv_bfloat16*>(smem + OL.wfc1);
    const __nv_bfloat16* wfc2 = reinterpret_cast<const __nv_bfloat16*>(smem + OL.wfc2);
    const float* f32 = reinterpret_cast<const float*>(smem);
    for (int it = 0; it < my_patches; ++it) {
      const int b = blockIdx.x + it * gridDim.x, par = it & 1;
      bar_sync(BAR_READY0 + par, NALL);
      const TailBuf<NW>& B = tb[par];
      float o = 0.f, l = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) { o += B.o[w][lane]; l += B.l[w][lane >> 3]; }
      const float att = o / l;
      float x0 = B.x0[lane];
      if (it + 2 < my_patches) bar_arrive(BAR_FREE0 + par, NALL);   // buffer may be refilled
      // proj (+bias, +residual)
      float y = f32[OL.bproj / 4 + lane];
#pragma unroll
      for (int k2 = 0; k2 < kD / 2; ++k2) {
        const uint32_t wv = lds32(wproj + lane * kLdD + 2 * k2);
        y = fmaf(bf_lo(wv), __shfl_sync(0xffffffffu, att, 2 * k2), y);
        y = fmaf(bf_hi(wv), __shfl_sync(0xffffffffu, att, 2 * k2 + 1), y);
      }
      if (DROP) y = drop1(y, drop_key(b, 4, 0, lane >> 1), lane & 1, dc);
      x0 += y;
      // LN2
      float mean = warp_sum(x0) * (1.f / kD);
      float d = x0 - mean;
      float rstd = rsqrtf(warp_sum(d * d) * (1.f / kD) + 1e-6f);
      const float y2 = d * rstd * f32[OL.ln2_g / 4 + lane] + f32[OL.ln2_b / 4 + lane];
      // fc1 + GELU: hidden unit j = lane + 32 i
      float hacc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) hacc[i] = f32[OL.bfc1 / 4 + lane + 32 * i];
#pragma unroll
      for (int k2 = 0; k2 < kD / 2; ++k2) {
        const float ya = __shfl_sync(0xffffffffu, y2, 2 * k2), yb = __shfl_sync(0xffffffffu, y2, 2 * k2 + 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t wv = lds32(wfc1 + (lane + 32 * i) * kLdD + 2 * k2);
          hacc[i] = fmaf(bf_lo(wv), ya, fmaf(bf_hi(wv), yb, hacc[i]));
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float hv = gelu_tanh_approx(hacc[i]);
        if (DROP) hv = drop1(hv, drop_key(b, 5, 0, (lane + 32 * i) >> 1), lane & 1, dc);
        h_s[lane + 32 * i] = hv;
      }
      __syncwarp();
      // fc2 (+bias, +residual)
      float z = f32[OL.bfc2 / 4 + lane];
#pragma unroll 8
      for (int k2 = 0; k2 < kHidden / 2; ++k2) {
        const uint32_t wv = lds32(wfc2 + lane * kLdHid + 2 * k2);
        const float2 hh = *reinterpret_cast<const float2*>(h_s + 2 * k2);
        z = fmaf(bf_lo(wv), hh.x, fmaf(bf_hi(wv), hh.y, z));
      }
      __syncwarp();
      if (DROP) z = drop1(z, drop_key(b, 6, 0, lane >> 1), lane & 1, dc);
      x0 += z;
      // final LayerNorm + head
      mean = warp_sum(x0) * (1.f / kD);
      d = x0 - mean;
      rstd = rsqrtf(warp_sum(d * d) * (1.f / kD) + 1e-6f);
      const float c = d * rstd * f32[L.lnf_g / 4 + lane] + f32[L.lnf_b / 4 + lane];
      const long long orow = a.out_index ? a.out_index[b] : (long long)b;
      for (int k0 = 0; k0 < a.K; k0 += 32) {
        const int k = k0 + lane;
        float acc = k < a.K ? f32[L.bhead / 4 + k] : 0.f;
#pragma unroll
        for (int dd = 0; dd < kD; ++dd) {
          const float cd = __shfl_sync(0xffffffffu, c, dd);
          if (k < a.K) acc = fmaf(f32[L.whead / 4 + k * kD + dd], cd, acc);
        }
        if (k < a.K) {
          a.logits[orow * a.K + k] = acc;
          logit_s[k] = acc;
        }
      }
      if (a.argmax_map) {
        __syncwarp();
        if (lane == 0) {
          int best = 0;
          float bv = logit_s[0];
          for (int k = 1; k < a.K; ++k)
            if (logit_s[k] > bv) { bv = logit_s[k]; best = k; }  // first maximum, like np.argmax
          a.argmax_map[orow] = (unsigned char)best;
        }
        __syncwarp();
      }
    }
    return;
  }

  // ============================== token warps ====================================================
  const int r0 = 16 * warp + g, r1 = r0 + 8;
  const int PW = P + 1, PP = sps_pp(P), HALO = sps_halo(P);
  const __nv_bfloat16* wfus = reinterpret_cast<const __nv_bfloat16*>(smem + L.wfus);
  const float* fus_scale = reinterpret_cast<const float*>(smem + L.fus_scale);
  const float* fus_bias = reinterpret_cast<const float*>(smem + L.fus_bias);
  const float* cls = reinterpret_cast<const float*>(smem + L.cls);
  const float* pos = reinterpret_cast<const float*>(smem + L.pos);
  const float qscale = 0.35355339059327376220f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)

  for (int it = 0; it < my_patches; ++it) {
    const int b = blockIdx.x + it * gridDim.x, par = it & 1;
    // ---------------- tokens: fusion 1x1 conv over the concatenated stems ----------------
    float x[4][4];
    {
      long long R0 = -1, R1 = -1;
      if (r0 >= 1 && r0 < T) { const int p = r0 - 1; R0 = HALO + (long long)b * PP + (p / P) * PW + (p % P); }
      if (r1 >= 1 && r1 < T) { const int p = r1 - 1; R1 = HALO + (long long)b * PP + (p / P) * PW + (p % P); }
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
      if (a.prefused) {
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          const __nv_bfloat16* sl = a.f + (long long)jn * a.RT * 8 + 2 * q;
          const uint32_t u0 = R0 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(sl + R0 * 8)) : 0u;
          const uint32_t u1 = R1 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(sl + R1 * 8)) : 0u;
          acc[jn][0] = bf_lo(u0); acc[jn][1] = bf_hi(u0); acc[jn][2] = bf_lo(u1); acc[jn][3] = bf_hi(u1);
        }
      }
#pragma unroll
      for (int kk = 0; kk < (a.prefused ? 0 : kFusK / 16); ++kk) {
        uint32_t A[4];
        const __nv_bfloat16* s0 = a.f + (long long)(2 * kk) * a.RT * 8 + 2 * q;
        const __nv_bfloat16* s1 = a.f + (long long)(2 * kk + 1) * a.RT * 8 + 2 * q;
        A[0] = R0 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(s0 + R0 * 8)) : 0u;
        A[1] = R1 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(s0 + R1 * 8)) : 0u;
        A[2] = R0 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(s1 + R0 * 8)) : 0u;
        A[3] = R1 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(s1 + R1 * 8)) : 0u;
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          const __nv_bfloat16* w = wfus + (8 * jn + g) * kLdFus + 16 * kk + 2 * q;
          mma16816(acc[jn], A, lds32(w), lds32(w + 8));
        }
      }
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        const int col = 8 * jn + 2 * q;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = (e < 2) ? r0 : r1, c = col + (e & 1);
          float v = 0.f;
          if (r == 0) v = cls[c] + pos[c];
          else if (r < T) v = (a.prefused ? acc[jn][e] : fmaxf(acc[jn][e] * fus_scale[c] + fus_bias[c], 0.f)) + pos[r * kD + c];
          x[jn][e] = v;
        }
        if (DROP) {   // pos_drop
          drop2(x[jn][0], x[jn][1], drop_key(b, 0, r0, 4 * jn + q), dc);
          drop2(x[jn][2], x[jn][3], drop_key(b, 0, r1, 4 * jn + q), dc);
        }
      }
    }

    // ---------------- full transformer blocks (all but the last) ----------------
#pragma unroll 1
    for (int l = 0; l < kLayers - 1; ++l) {
      const TLayerOff& O = L.layer[l];
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + O.wqkv);
      const __nv_bfloat16* wproj = reinterpret_cast<const __nv_bfloat16*>(smem + O.wproj);
      const __nv_bfloat16* wfc1 = reinterpret_cast<const __nv_bfloat16*>(smem + O.wfc1);
      const __nv_bfloat16* wfc2 = reinterpret_cast<const __nv_bfloat16*>(smem + O.wfc2);
      const float* bqkv = reinterpret_cast<const float*>(smem + O.bqkv);
      const float* bproj = reinterpret_cast<const float*>(smem + O.bproj);
      const float* bfc1 = reinterpret_cast<const float*>(smem + O.bfc1);
      const float* bfc2 = reinterpret_cast<const float*>(smem + O.bfc2);

      // ---- LN1 -> qkv GEMM; Q stays in registers, K and V^T go to shared memory ----
      uint32_t A1[2][4];
      ln_to_afrag(x, reinterpret_cast<const float*>(smem + O.ln1_g), reinterpret_cast<const float*>(smem + O.ln1_b), q, A1);
      uint32_t qa[kHeads][2];
#pragma unroll
      for (int jn = 0; jn < 12; ++jn) {
        const float2 bb = *reinterpret_cast<const float2*>(bqkv + 8 * jn + 2 * q);
        float c[4] = {bb.x, bb.y, bb.x, bb.y};      // bias = accumulator init
        uint32_t B[4];   // B fragments of both K=16 steps in one ldmatrix
        ldsm4(B, wqkv + (8 * jn + (lane & 7)) * kLdD + 8 * (lane >> 3));
        mma16816(c, A1[0], B[0], B[1]);
        mma16816(c, A1[1], B[2], B[3]);
        if (jn < 4) {
          qa[jn][0] = pack_bf16(c[0] * qscale, c[1] * qscale);
          qa[jn][1] = pack_bf16(c[2] * qscale, c[3] * qscale);
        } else if (jn < 8) {
          const int col = 8 * (jn - 4) + 2 * q;
          *reinterpret_cast<uint32_t*>(Ks + r0 * kLdD + col) = pack_bf16(c[0], c[1]);
          *reinterpret_cast<uint32_t*>(Ks + r1 * kLdD + col) = pack_bf16(c[2], c[3]);
        } else {
          const int d = 8 * (jn - 8) + 2 * q;
          Vt[d * LDV + r0] = __float2bfloat16_rn(c[0]);
          Vt[(d + 1) * LDV + r0] = __float2bfloat16_rn(c[1]);
          Vt[d * LDV + r1] = __float2bfloat16_rn(c[2]);
          Vt[(d + 1) * LDV + r1] = __float2bfloat16_rn(c[3]);
        }
      }
      bar_sync(BAR_MAIN, NMAIN);

      // ---- attention: per head QK^T -> online softmax -> PV, all in registers ----
      uint32_t oa[2][4];  // attention output as the two K=16 A fragments of the proj GEMM
#pragma unroll
      for (int h = 0; h < kHeads; ++h) {
        float m0 = -INFINITY, m1 = -INFINITY;
        float lacc[4] = {0.f, 0.f, 0.f, 0.f};   // softmax denominators from the tensor core: P x ones
        float oh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          constexpr int kFull = 8;
          const int ntile = (2 * NW - 8 * kb) < kFull ? (2 * NW - 8 * kb) : kFull;
          float s[8][4];
          uint32_t kfr[4];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            if (t < ntile) {
              s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
              const int key0 = (8 * kb + t) * 8;
              if ((t & 3) == 0 && t + 3 < ntile)     // K fragments of four key tiles in one ldmatrix
                ldsm4(kfr, Ks + (key0 + 8 * (lane >> 3) + (lane & 7)) * kLdD + 8 * h);
              const uint32_t kb0 = ((t & ~3) + 3 < ntile) ? kfr[t & 3] : lds32(Ks + (key0 + g) * kLdD + 8 * h + 2 * q);
              mma1688(s[t], qa[h][0], qa[h][1], kb0);
              if (8 * kb + t >= 2 * NW - 2) {   // only the last 16 keys can be padding (T > 16(NW-1))
                const int kc = key0 + 2 * q;
                if (kc >= T) { s[t][0] = -INFINITY; s[t][2] = -INFINITY; }
                if (kc + 1 >= T) { s[t][1] = -INFINITY; s[t][3] = -INFINITY; }
              }
            }
          }
          float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t < ntile) {
              bm0 = fmaxf(bm0, fmaxf(s[t][0], s[t][1]));
              bm1 = fmaxf(bm1, fmaxf(s[t][2], s[t][3]));
            }
          const float mn0 = fmaxf(m0, quad_max(bm0)), mn1 = fmaxf(m1, quad_max(bm1));
          if (kb > 0) {
            const float al0 = ex2(m0 - mn0), al1 = ex2(m1 - mn1);
            lacc[0] *= al0; lacc[2] *= al1;
            oh[0] *= al0; oh[1] *= al0; oh[2] *= al1; oh[3] *= al1;
          }
          m0 = mn0; m1 = mn1;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t < ntile) {
              s[t][0] = ex2(s[t][0] - mn0); s[t][1] = ex2(s[t][1] - mn0);
              s[t][2] = ex2(s[t][2] - mn1); s[t][3] = ex2(s[t][3] - mn1);
            }
          uint32_t vfr[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (2 * kk < ntile) {
              uint32_t Pa[4];
              Pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
              Pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
              Pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
              Pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
              if ((kk & 1) == 0 && 2 * kk + 2 < ntile)
                ldsm4(vfr, Vt + (8 * h + (lane & 7)) * LDV + (8 * kb + 2 * kk) * 8 + 8 * (lane >> 3));
              const __nv_bfloat16* v = Vt + (8 * h + g) * LDV + (8 * kb + 2 * kk) * 8 + 2 * q;
              const bool pair = 2 * (kk & ~1) + 2 < ntile;
              const uint32_t vb0 = pair ? vfr[2 * (kk & 1)] : lds32(v), vb1 = pair ? vfr[2 * (kk & 1) + 1] : lds32(v + 8);
              mma16816(oh, Pa, vb0, vb1);
              mma16816(lacc, Pa, 0x3F803F80u, 0x3F803F80u);   // every column = row sum of (bf16) P
            }
        }
        const float il0 = 1.f / lacc[0], il1 = 1.f / lacc[2];
        oa[h >> 1][(h & 1) * 2 + 0] = pack_bf16(oh[0] * il0, oh[1] * il0);
        oa[h >> 1][(h & 1) * 2 + 1] = pack_bf16(oh[2] * il1, oh[3] * il1);
      }

      // ---- proj GEMM, bias + residual epilogue ----
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {   // the residual stream is the accumulator
        const float2 bb = *reinterpret_cast<const float2*>(bproj + 8 * jn + 2 * q);
        uint32_t B[4];
        ldsm4(B, wproj + (8 * jn + (lane & 7)) * kLdD + 8 * (lane >> 3));
        if (!DROP) {
          x[jn][0] += bb.x; x[jn][1] += bb.y; x[jn][2] += bb.x; x[jn][3] += bb.y;
          mma16816(x[jn], oa[0], B[0], B[1]);
          mma16816(x[jn], oa[1], B[2], B[3]);
        } else {     // proj_drop acts on the branch output before the residual add
          float c[4] = {bb.x, bb.y, bb.x, bb.y};
          mma16816(c, oa[0], B[0], B[1]);
          mma16816(c, oa[1], B[2], B[3]);
          drop2(c[0], c[1], drop_key(b, 1 + 3 * l, r0, 4 * jn + q), dc);
          drop2(c[2], c[3], drop_key(b, 1 + 3 * l, r1, 4 * jn + q), dc);
          x[jn][0] += c[0]; x[jn][1] += c[1]; x[jn][2] += c[2]; x[jn][3] += c[3];
        }
      }

      // ---- LN2 -> fc1 (+bias, GELU) -> fc2 (+bias, +residual), 16 hidden units at a time ----
      uint32_t A2[2][4];
      ln_to_afrag(x, reinterpret_cast<const float*>(smem + O.ln2_g), reinterpret_cast<const float*>(smem + O.ln2_b), q, A2);
      float acc2[DROP ? 4 : 1][4];        // dropout needs the branch output separate from the residual
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {   // fc2 accumulates straight into the residual stream
        const float2 bb = *reinterpret_cast<const float2*>(bfc2 + 8 * jn + 2 * q);
        if (!DROP) { x[jn][0] += bb.x; x[jn][1] += bb.y; x[jn][2] += bb.x; x[jn][3] += bb.y; }
        else { acc2[DROP ? jn : 0][0] = bb.x; acc2[DROP ? jn : 0][1] = bb.y; acc2[DROP ? jn : 0][2] = bb.x; acc2[DROP ? jn : 0][3] = bb.y; }
      }
#pragma unroll
      for (int hk = 0; hk < kHidden / 16; ++hk) {
        const float2 b0 = *reinterpret_cast<const float2*>(bfc1 + 16 * hk + 2 * q);
        const float2 b1 = *reinterpret_cast<const float2*>(bfc1 + 16 * hk + 8 + 2 * q);
        float h0[4] = {b0.x, b0.y, b0.x, b0.y}, h1[4] = {b1.x, b1.y, b1.x, b1.y};
        {
          uint32_t B0[4], B1[4];
          ldsm4(B0, wfc1 + (16 * hk + (lane & 7)) * kLdD + 8 * (lane >> 3));
          ldsm4(B1, wfc1 + (16 * hk + 8 + (lane & 7)) * kLdD + 8 * (lane >> 3));
          mma16816(h0, A2[0], B0[0], B0[1]);
          mma16816(h0, A2[1], B0[2], B0[3]);
          mma16816(h1, A2[0], B1[0], B1[1]);
          mma16816(h1, A2[1], B1[2], B1[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { h0[e] = gelu_tanh_approx(h0[e]); h1[e] = gelu_tanh_approx(h1[e]); }
        if (DROP) {
          drop2(h0[0], h0[1], drop_key(b, 2 + 3 * l, r0, 8 * hk + q), dc);
          drop2(h0[2], h0[3], drop_key(b, 2 + 3 * l, r1, 8 * hk + q), dc);
          drop2(h1[0], h1[1], drop_key(b, 2 + 3 * l, r0, 8 * hk + 4 + q), dc);
          drop2(h1[2], h1[3], drop_key(b, 2 + 3 * l, r1, 8 * hk + 4 + q), dc);
        }
        uint32_t Ha[4];
        Ha[0] = pack_bf16(h0[0], h0[1]);
        Ha[1] = pack_bf16(h0[2], h0[3]);
        Ha[2] = pack_bf16(h1[0], h1[1]);
        Ha[3] = pack_bf16(h1[2], h1[3]);
#pragma unroll
        for (int jn = 0; jn < 4; jn += 2) {   // one ldmatrix = B fragments of two output tiles
          uint32_t B[4];
          ldsm4(B, wfc2 + (8 * (jn + (lane >> 4)) + (lane & 7)) * kLdHid + 16 * hk + 8 * ((lane >> 3) & 1));
          if (!DROP) {
            mma16816(x[jn], Ha, B[0], B[1]);
            mma16816(x[jn + 1], Ha, B[2], B[3]);
          } else {
            mma16816(acc2[DROP ? jn : 0], Ha, B[0], B[1]);
            mma16816(acc2[DROP ? jn + 1 : 0], Ha, B[2], B[3]);
          }
        }
      }
      if (DROP) {
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          float* c = acc2[DROP ? jn : 0];
          drop2(c[0], c[1], drop_key(b, 3 + 3 * l, r0, 4 * jn + q), dc);
          drop2(c[2], c[3], drop_key(b, 3 + 3 * l, r1, 4 * jn + q), dc);
          x[jn][0] += c[0]; x[jn][1] += c[1]; x[jn][2] += c[2]; x[jn][3] += c[3];
        }
      }
      bar_sync(BAR_MAIN, NMAIN);  // every warp is done with this layer's K / V^T
    }

    // ---------------- last block: K, V for every token, attention of the cls query only -------
    {
      const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(smem + OL.wqkv);
      const float* bqkv = reinterpret_cast<const float*>(smem + OL.bqkv);
      uint32_t A1[2][4];
      ln_to_afrag(x, reinterpret_cast<const float*>(smem + OL.ln1_g), reinterpret_cast<const float*>(smem + OL.ln1_b), q, A1);
      float kc[kHeads][4], vv[kHeads][4];
#pragma unroll
      for (int jn = 4; jn < 12; ++jn) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const __nv_bfloat16* w = wqkv + (8 * jn + g) * kLdD + 16 * kk + 2 * q;
          mma16816(c, A1[kk], lds32(w), lds32(w + 8));
        }
        const float2 bb = *reinterpret_cast<const float2*>(bqkv + 8 * jn + 2 * q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = c[e] + ((e & 1) ? bb.y : bb.x);
          if (jn < 8) kc[jn - 4][e] = v; else vv[jn - 8][e] = v;
        }
      }
      if (it >= 2) bar_sync(BAR_FREE0 + par, NALL);   // tail warp has consumed this buffer
      TailBuf<NW>& B = tb[par];
      if (warp == 0) {   // the cls token is row 0: its query and its residual stream
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const __nv_bfloat16* w = wqkv + (8 * jn + g) * kLdD + 16 * kk + 2 * q;
            mma16816(c, A1[kk], lds32(w), lds32(w + 8));
          }
          if (g == 0) {
            const float2 bb = *reinterpret_cast<const float2*>(bqkv + 8 * jn + 2 * q);
            q0_s[8 * jn + 2 * q] = (c[0] + bb.x) * qscale;
            q0_s[8 * jn + 2 * q + 1] = (c[1] + bb.y) * qscale;
            B.x0[8 * jn + 2 * q] = x[jn][0];
            B.x0[8 * jn + 2 * q + 1] = x[jn][1];
          }
        }
      }
      bar_sync(BAR_MAIN, NMAIN);
      float s0[kHeads], s1[kHeads];
#pragma unroll
      for (int h = 0; h < kHeads; ++h) {
        const float2 qq = *reinterpret_cast<const float2*>(q0_s + 8 * h + 2 * q);
        s0[h] = quad_sum(qq.x * kc[h][0] + qq.y * kc[h][1]);
        s1[h] = quad_sum(qq.x * kc[h][2] + qq.y * kc[h][3]);
        if (r0 >= T) s0[h] = -INFINITY;
        if (r1 >= T) s1[h] = -INFINITY;
        float m = fmaxf(s0[h], s1[h]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        if (lane == 0) max_s[warp * kHeads + h] = m;
      }
      bar_sync(BAR_MAIN, NMAIN);
#pragma unroll
      for (int h = 0; h < kHeads; ++h) {
        float m = max_s[h];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = fmaxf(m, max_s[w * kHeads + h]);
        const float p0 = ex2(s0[h] - m), p1 = ex2(s1[h] - m);
        const float ls = g_sum(p0 + p1);
        const float oa_ = g_sum(p0 * vv[h][0] + p1 * vv[h][2]);
        const float ob_ = g_sum(p0 * vv[h][1] + p1 * vv[h][3]);
        if (g == 0) {
          B.o[warp][8 * h + 2 * q] = oa_;
          B.o[warp][8 * h + 2 * q + 1] = ob_;
          if (q == 0) B.l[warp][h] = ls;
        }
      }
      bar_arrive(BAR_READY0 + par, NALL);
      // q0_s / max_s are rewritten only after the next patch's BAR_MAIN syncs
    }
  }
}

size_t tparams_bytes(int P, int K) { return (size_t)tlayout(P, K).total; }

template <int NW, bool DROP>
static int launch_nw(const TArgs& a, cudaStream_t stream) {
  const TLayout L = tlayout(a.P, a.K);
  const size_t smem = (size_t)L.total + (size_t)(16 * NW) * kLdD * 2 + (size_t)kD * (16 * NW + 8) * 2 + kD * 4 +
                      (size_t)NW * kHeads * 4 + 2 * sizeof(TailBuf<NW>) + kHidden * 4 + (size_t)a.K * 4 + 16;
  int dev = 0, max_smem = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (smem > (size_t)max_smem) return VC_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(transformer_fwd_kernel<NW, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return VC_ERR_CUDA;
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, transformer_fwd_kernel<NW, DROP>, (NW + 1) * 32, smem);
  if (occ < 1) occ = 1;
  long long blocks = (long long)num_sms * occ;
  if (blocks > a.n_patches) blocks = a.n_patches;
  transformer_fwd_kernel<NW, DROP><<<(int)blocks, (NW + 1) * 32, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? VC_OK : VC_ERR_CUDA;
}

int transformer_fwd_launch(const void* f_sps, const void* tparams, int n_patches, int P, int K, float* logits,
                           const long long* out_index, unsigned char* argmax_map, int prefused, unsigned int drop_thr,
                           const unsigned int* drop_seed, cudaStream_t stream) {
  if (n_patches <= 0 || P < 1 || K < 1 || K > 64 || (drop_thr && !drop_seed) || drop_thr >= 65536u) return VC_ERR_ARG;
  TArgs a;
  a.f = (const __nv_bfloat16*)f_sps;
  a.blob = (const uint8_t*)tparams;
  a.prefused = prefused;
  a.logits = logits;
  a.out_index = out_index;
  a.argmax_map = argmax_map;
  a.RT = sps_rows(n_patches, P);
  a.n_patches = n_patches;
  a.P = P;
  a.K = K;
  a.T = P * P + 1;
  a.drop_thr = drop_thr;
  a.drop_seed = drop_seed;
  a.L = tlayout(P, K);
  const int NW = (a.T + 15) / 16;
  if (drop_thr) {   // training-mode dropout: the patch sizes the training kernels cover
    switch (NW) {
#define VC_CASE(N) case N: return launch_nw<N, true>(a, stream);
      VC_CASE(1) VC_CASE(2) VC_CASE(3) VC_CASE(4) VC_CASE(5) VC_CASE(6) VC_CASE(7) VC_CASE(8) VC_CASE(11) VC_CASE(15)
#undef VC_CASE
      default: return VC_ERR_UNSUPPORTED;
    }
  }
  switch (NW) {
#define VC_CASE(N) case N: return launch_nw<N, false>(a, stream);
    VC_CASE(1) VC_CASE(2) VC_CASE(3) VC_CASE(4) VC_CASE(5) VC_CASE(6) VC_CASE(7) VC_CASE(8)
    VC_CASE(9) VC_CASE(10) VC_CASE(11) VC_CASE(12) VC_CASE(13) VC_CASE(14) VC_CASE(15)
#undef VC_CASE
    default: return VC_ERR_UNSUPPORTED;  // P > 15
  }
}

}  // namespace vc
