// Layout of the packed parameter blob of the token stage (fusion 1x1 conv, cls/pos, two
// pre-norm transformer blocks, final norm, head).  bf16 GEMM weights are stored [out][in+8]
// (row pitch padded by 8 elements = 16 B so the mma.sync B-fragment loads are bank-conflict
// free), everything else fp32.  Offsets are in BYTES from the start of the blob; every array
// is 16-byte aligned.  Python fills the blob through vc_tparams_layout().
#pragma once
#include <stdint.h>

namespace vc {

constexpr int kD = 32, kHeads = 4, kHd = 8, kHidden = 128, kLayers = 2, kFusK = 64;
constexpr int kLdD = kD + 8;        // 40: pitch of [.][32] weights
constexpr int kLdFus = kFusK + 8;   // 72
constexpr int kLdHid = kHidden + 8; // 136

struct TLayerOff {
  int wqkv, wproj, wfc1, wfc2;                               // bf16
  int ln1_g, ln1_b, bqkv, bproj, ln2_g, ln2_b, bfc1, bfc2;   // fp32
};
struct TLayout {
  int wfus;                         // bf16 [32][72]
  int fus_scale, fus_bias, cls;     // fp32 [32]
  TLayerOff layer[kLayers];
  int lnf_g, lnf_b;                 // fp32 [32]
  int whead, bhead;                 // fp32 [K][32], [K]
  int pos;                          // fp32 [T][32]
  int total;                        // bytes
};

#ifdef __CUDACC__
__host__ __device__
#endif
inline int tl_take(int& off, int bytes) { int o = off; off += (bytes + 15) & ~15; return o; }
#ifdef __CUDACC__
__host__ __device__
#endif
inline TLayout tlayout(int P, int K) {
  TLayout L;
  int off = 0;
  L.wfus = tl_take(off, kD * kLdFus * 2);
  for (int l = 0; l < kLayers; ++l) {
    L.layer[l].wqkv = tl_take(off, 3 * kD * kLdD * 2);
    L.layer[l].wproj = tl_take(off, kD * kLdD * 2);
    L.layer[l].wfc1 = tl_take(off, kHidden * kLdD * 2);
    L.layer[l].wfc2 = tl_take(off, kD * kLdHid * 2);
  }
  L.fus_scale = tl_take(off, kD * 4);
  L.fus_bias = tl_take(off, kD * 4);
  L.cls = tl_take(off, kD * 4);
  for (int l = 0; l < kLayers; ++l) {
    L.layer[l].ln1_g = tl_take(off, kD * 4);
    L.layer[l].ln1_b = tl_take(off, kD * 4);
    L.layer[l].bqkv = tl_take(off, 3 * kD * 4);
    L.layer[l].bproj = tl_take(off, kD * 4);
    L.layer[l].ln2_g = tl_take(off, kD * 4);
    L.layer[l].ln2_b = tl_take(off, kD * 4);
    L.layer[l].bfc1 = tl_take(off, kHidden * 4);
    L.layer[l].bfc2 = tl_take(off, kD * 4);
  }
  L.lnf_g = tl_take(off, kD * 4);
  L.lnf_b = tl_take(off, kD * 4);
  L.whead = tl_take(off, K * kD * 4);
  L.bhead = tl_take(off, K * 4);
  L.pos = tl_take(off, (P * P + 1) * kD * 4);
  L.total = off;
  return L;
}

}  // namespace vc
