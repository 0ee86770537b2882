// extern "C" surface of libvitcnn.so (include/vitcnn.h) and the host-side orchestration of
// one forward pass: pack -> HSI stem (3 convs) -> LiDAR stem (3 convs) -> token stage.
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <vector>
#include "../../include/vitcnn.h"
#include "vc_common.cuh"
#include "vc_kernels.h"
#include "vc_tparams.h"

namespace {

thread_local char g_err[256] = "";

int fail(int code, const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  snprintf(g_err, sizeof(g_err), "%s (code %d, cuda: %s)", what, code, cudaGetErrorString(e));
  return code;
}

#define VC_TRY(expr)                       \
  do {                                     \
    int _rc = (expr);                      \
    if (_rc != VC_OK) return fail(_rc, #expr); \
  } while (0)

// ---- optional per-kernel-class timing (bench.py's roofline leg) ---------------------------------
enum { KC_INDEX = 0, KC_PACK, KC_CONV_H1, KC_CONV_H2, KC_CONV_H3, KC_CONV_L, KC_TOKENS, KC_HALO, KC_BN, KC_WGRAD, KC_DGRAD,
       KC_TOKENS_BWD, KC_MISC, KC_COUNT };
struct ProfRec { int cls; cudaEvent_t e0, e1; };
thread_local bool g_prof_on = false;
thread_local std::vector<ProfRec>* g_prof = nullptr;
std::atomic<long long> g_launches{0};

struct Scope {  // brackets one kernel launch with events when profiling is on
  cudaStream_t st;
  ProfRec rec;
  bool on;
  Scope(int cls, cudaStream_t s) : st(s), on(g_prof_on) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (on) {
      rec.cls = cls;
      cudaEventCreate(&rec.e0);
      cudaEventCreate(&rec.e1);
      cudaEventRecord(rec.e0, st);
    }
  }
  ~Scope() {
    if (on) {
      cudaEventRecord(rec.e1, st);
      g_prof->push_back(rec);
    }
  }
};
#define VC_LAUNCH(cls, st, expr) \
  do {                           \
    Scope _sc(cls, st);          \
    VC_TRY(expr);                \
  } while (0)

struct Workspace {
  uint8_t *a0, *l0, *a1, *a2, *f, *l1, *l2, *tscr;
  long long *off1, *off2, *oidx;
  long long bytes;
};

Workspace carve(void* base, int n, int P, int S1, int S2) {
  Workspace w;
  const long long RT = vc::sps_rows(n, P);
  const long long sl = RT * 16;  // bytes per slice
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](long long bytes) { uint8_t* r = p; p += (bytes + 255) & ~255LL; return r; };
  w.a0 = take(sl * S1);
  w.l0 = take(sl * S2);
  w.a1 = take(sl * 16);
  w.a2 = take(sl * 8);
  w.f = take(sl * 8);
  w.l1 = take(sl * 2);
  w.l2 = take(sl * 2);
  w.off1 = reinterpret_cast<long long*>(take(8LL * n));
  w.off2 = reinterpret_cast<long long*>(take(8LL * n));
  w.oidx = reinterpret_cast<long long*>(take(8LL * n));
  w.tscr = take((long long)vc::tokens_tc_scratch_bytes(n));     // cls records of the tcgen05 token kernel
  w.bytes = p - reinterpret_cast<uint8_t*>(base);
  return w;
}

inline int slices_for(int C) { return (C + 15) / 16 * 2; }

// One 3-conv stem as the shared-stem code sees it (HSI: 128 / 64 / 32 channels, LiDAR: 8 / 16 / 32 padded to 16 / 16 / 32).
struct StemDesc {
  int C, S0;                   // raster channels, slices of the packed blocks
  const void* w[3];
  const float *scale[3], *bias[3];
  int nsplit[3], n_out[3];
  const void* w1_border;       // conv 1: nine weight copies with the taps that leave the window zeroed (the CTA-pair
  long long w1_bytes;          //   kernel takes no tap masks), or null: one copy + tap masks
  int kc[3];                   // profiling class of each layer
};

StemDesc hsi_stem(const vc_model* m) {
  StemDesc d;
  d.C = m->C1; d.S0 = m->S1;
  const int n_out[3] = {128, 64, 32};
  for (int i = 0; i < 3; ++i) {
    d.w[i] = m->w_h[i]; d.scale[i] = m->scale_h[i]; d.bias[i] = m->bias_h[i]; d.nsplit[i] = m->nsplit_h[i]; d.n_out[i] = n_out[i];
  }
  d.w1_border = m->w_h1_border;
  d.w1_bytes = 9LL * m->S1 * 128 * 16;
  d.kc[0] = KC_CONV_H1; d.kc[1] = KC_CONV_H2; d.kc[2] = KC_CONV_H3;
  return d;
}

StemDesc lidar_stem(const vc_model* m) {
  StemDesc d;
  d.C = m->C2; d.S0 = m->S2;
  const int n_out[3] = {16, 16, 32};
  for (int i = 0; i < 3; ++i) {
    d.w[i] = m->w_l[i]; d.scale[i] = m->scale_l[i]; d.bias[i] = m->bias_l[i]; d.nsplit[i] = m->nsplit_l[i]; d.n_out[i] = n_out[i];
  }
  d.w1_border = nullptr;
  d.w1_bytes = 0;
  d.kc[0] = d.kc[1] = d.kc[2] = KC_CONV_L;
  return d;
}

// Scene-level buffers of the shared stem (pack.cu), carved behind the chunk workspace: raster offsets of the
// blocks, the blocks as SPS patches of B x B pixels, the border-class variants of the first D stem convs.
// depth D: 3 / 2: B = 31, the first three / two convs shared (P >= 2D + 1); 1: B = 15, conv 1 only (P >= 2).
struct SceneWs {
  long long* boff;
  int *rowterm, *colterm;           // block-row offset tables of the raster rows / columns (token kernel reading the planes)
  uint8_t *blocks, *var[3];
  long long RTb, plane[3], bytes;   // bytes == 0: this depth does not apply to the model / raster; else the END offset
  int nb, B, D;
};

// Edge of the scene blocks.  Depth 1 (conv 1 only, small patches): 15.  Depth >= 2: a block's exact region is B - 2D wide, so
// small blocks recompute a lot ((31 / 25)^2 = 1.54x at depth 3) while large ones overshoot the raster; pick, among the edges
// with B + 1 a multiple of 32 (whole tiles per block, aligned slabs in conv_var.cu), the one with the fewest block rows in
// total.  Houston scene at depth 3: 1 064 blocks of 31 (1.09 M rows), 238 of 63 (0.97 M), 88 of 95 (0.81 M).  95 is the
// largest: the CTA-pair conv-1 kernel stages 128 + 2 (B + 9) rows per tile.  VITCNN_SCENE_BLOCK forces an edge.
int scene_block(int H, int W, int depth) {
  if (depth < 2) return 15;
  static const int forced = [] {
    const char* e = getenv("VITCNN_SCENE_BLOCK");
    return e ? atoi(e) : 0;
  }();
  int best = 31;
  long long best_rows = -1;
  for (int B = 31; B <= 95; B += 32) {
    if (H < B || W < B) break;
    if (forced && B != forced) continue;
    const long long rows = (long long)vc::blk_count(H, B, depth) * vc::blk_count(W, B, depth) * (B + 1) * (B + 1);
    if (best_rows < 0 || rows < best_rows) { best = B; best_rows = rows; }
  }
  return best;
}

SceneWs carve_scene(void* base, long long off, const StemDesc& sd, int H, int W, int P, int depth) {
  SceneWs s;
  memset(&s, 0, sizeof(s));
  const int B = scene_block(H, W, depth), D = depth;
  if (depth < 1 || depth > 3 || H < B || W < B || P < (depth == 1 ? 2 : 2 * depth + 1)) return s;
  s.B = B;
  s.D = D;
  s.nb = vc::blk_count(H, B, D) * vc::blk_count(W, B, D);
  s.RTb = vc::sps_rows(s.nb, B);
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](long long bytes) { uint8_t* r = p + off; off += (bytes + 255) & ~255LL; return r; };
  s.boff = reinterpret_cast<long long*>(take(8LL * s.nb));
  s.rowterm = reinterpret_cast<int*>(take(4LL * H));
  s.colterm = reinterpret_cast<int*>(take(4LL * W));
  s.blocks = take(s.RTb * 16 * sd.S0);
  for (int l = 0; l < D; ++l) {
    s.plane[l] = (long long)(sd.n_out[l] / 8) * s.RTb * 16;
    s.var[l] = take((long long)(2 * l + 3) * (2 * l + 3) * s.plane[l]);
  }
  s.bytes = off;
  return s;
}

long long chunk_ws_end(const vc_model* m, int chunk) {
  return (carve(nullptr, chunk, m->P, m->S1, m->S2).bytes + 256 + 255) & ~255LL;
}

// The token kernel that reads its stem inputs straight from the depth-3 variant planes (tokens_tc.cu): see use_tokens_tc
bool use_tokens_tc(int P, int K);

// The LiDAR stem is shared at depth 3 whenever the geometry allows: its planes are small (C2 -> 8 -> 16 -> 32
// channels) and the per-window alternative recomputes every LiDAR pixel up to P^2 times.
bool lidar_shared_ok(const vc_model* m, int H, int W) {
  static const int off = [] {
    const char* e = getenv("VITCNN_LIDAR_SHARED");
    return e && e[0] == '0';
  }();
  return !off && m->w_h1_border && m->P >= 7 && H >= 31 && W >= 31;
}

// Which sharing depth of the HSI stem pays for n windows and fits the workspace the caller gave.  Costs in units of
// one per-window conv-3 pass over a 128-row tile, measured on B200 at the Houston shape (profiles/r02_SUMMARY.md):
// per-window conv 1 / 2 / 3 = 2.8 / 1.8 / 1 and patch gather 0.9; a block tile of the shared conv 1 (9 variants) /
// conv 2 (25) / conv 3 (49) costs 27 / 53 / 38 (conv 2 / 3: conv_var.cu, all variants of a row class per work unit);
// the variant gather costs 1.65 / 0.6 / 0.5 (16 / 8 / 4 slices per row) and nothing at depth 3 when the tcgen05 token
// kernel runs (it reads the planes itself).  VITCNN_SCENE_DEPTH forces a depth (0 = per-window path).
int scene_mode(const vc_model* m, int H, int W, int chunk, long long n_windows, long long workspace_bytes) {
  static const int forced = [] {
    const char* e = getenv("VITCNN_SCENE_DEPTH");
    return e ? atoi(e) : -1;
  }();
  if (!m->w_h1_border || forced == 0) return 0;
  const StemDesc sd = hsi_stem(m);
  const double win_tiles = (double)n_windows * vc::sps_pp(m->P) / 128.0;
  const double conv[3] = {2.8, 1.8, 1.0}, shared[3] = {27.0, 53.0, 38.0};
  double gather[3] = {1.65, 0.6, 0.5};
  if (use_tokens_tc(m->P, m->K)) gather[2] = 0.0;
  const long long lid = lidar_shared_ok(m, H, W) ? carve_scene(nullptr, 0, lidar_stem(m), H, W, m->P, 3).bytes : 0;
  int best = 0;
  double best_cost = win_tiles * (conv[0] + conv[1] + conv[2] + 0.9);      // per-window path incl. its patch gather
  for (int mode = 1; mode <= 3; ++mode) {
    const SceneWs s = carve_scene(nullptr, chunk_ws_end(m, chunk), sd, H, W, m->P, mode);
    if (s.bytes <= 0 || workspace_bytes < s.bytes + lid) continue;
    if (forced >= 0 && mode != forced) continue;
    double cost = win_tiles * gather[mode - 1];
    for (int l = 0; l < 3; ++l) cost += l < mode ? shared[l] * s.nb * vc::sps_pp(s.B) / 128.0 : win_tiles * conv[l];
    if (cost < best_cost || forced == mode) { best = mode; best_cost = cost; }
  }
  return best;
}

// class (at depth L-1) of the neighbour d = -1 / 0 / +1 of a pixel whose class at depth L is c; -1: outside the window
int neighbour_class(int L, int c, int d, int P) {
  const int i = c < L ? c : (c == L ? L : P - 1 - (2 * L - c));
  const int ip = i + d;
  if (ip < 0 || ip > P - 1) return -1;
  return vc::border_class(ip, P, L - 1);
}

// The shared stem over the scene blocks: conv 1 as 9 launches (border-class weight copies, or one copy with the taps
// that leave the window masked out), conv L >= 2 as (2L+1)^2 launches that read, per tap, the conv L-1 variant of
// the neighbour's class (<= 9 planes per launch).
int shared_stem(const StemDesc& sd, const SceneWs& sw, const float* img, int H, int W, int P, cudaStream_t st) {
  VC_LAUNCH(KC_INDEX, st, vc::block_offsets_launch(H, W, sd.C, sw.B, sw.D, sw.boff, st));
  VC_LAUNCH(KC_INDEX, st, vc::scene_tables_launch(H, W, sw.B, sw.D, sw.rowterm, sw.colterm, st));
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(img, 0, 1, (long long)W * sd.C, sd.C, sw.boff, nullptr, sw.nb, sd.C, sw.B, sw.blocks,
                                             sd.S0, st));
  for (int v = 0; v < 9; ++v) {
    Scope sc(sd.kc[0], st);
    uint8_t* out = sw.var[0] + (size_t)v * sw.plane[0];
    if (sd.w1_border) {
      VC_TRY(vc::conv_sps_launch(sw.blocks, sd.S0, (const uint8_t*)sd.w1_border + (size_t)v * sd.w1_bytes, sd.scale[0], sd.bias[0], out, 0,
                                 sd.n_out[0], sd.nsplit[0], sw.nb, sw.B, 9, 1, 0, 0, st));
    } else {
      const int cy = v / 3, cx = v % 3;
      unsigned int mask = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx)
          if (!((cy == 0 && dy < 0) || (cy == 2 && dy > 0) || (cx == 0 && dx < 0) || (cx == 2 && dx > 0)))
            mask |= 1u << ((dy + 1) * 3 + (dx + 1));
      const void* plane = sw.blocks;
      VC_TRY(vc::conv_sps_planes_launch(&plane, &mask, 1, sd.S0, sd.w[0], sd.scale[0], sd.bias[0], out, 0, sd.n_out[0], sd.nsplit[0], sw.nb,
                                        sw.B, 9, 1, 0, 0, st));
    }
  }
  static const bool per_variant = [] {       // VITCNN_STEM_IMPL=planes: one launch per output variant (the round-1 path)
    const char* e = getenv("VITCNN_STEM_IMPL");
    return e && e[0] == 'p';
  }();
  for (int L = 2; L <= sw.D; ++L) {
    const int NC = 2 * L + 1, NP = 2 * L - 1;     // classes per axis at depth L / L-1
    if (!per_variant && sd.nsplit[L - 1] == 1) {
      // all variants of a row class per work unit: input slabs staged once for 2L+1 accumulators (conv_var.cu)
      signed char cls[7 * 3];
      for (int c = 0; c < NC; ++c)
        for (int d = 0; d < 3; ++d) cls[c * 3 + d] = (signed char)neighbour_class(L, c, d - 1, P);
      Scope sc(sd.kc[L - 1], st);
      const int rc = vc::conv_var_launch(sw.var[L - 2], sd.n_out[L - 2] / 8, sd.w[L - 1], sd.scale[L - 1], sd.bias[L - 1], sw.var[L - 1],
                                         sd.n_out[L - 1], L, cls, cls, sw.nb, sw.B, 1, st);
      if (rc == VC_OK) continue;
      if (rc != VC_ERR_UNSUPPORTED) return fail(rc, "conv_var_launch");
    }
    for (int cy = 0; cy < NC; ++cy)
      for (int cx = 0; cx < NC; ++cx) {
        const void* planes[9];
        unsigned int masks[9];
        int ids[9], np = 0;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            const int ry = neighbour_class(L, cy, dy, P), rx = neighbour_class(L, cx, dx, P);
            if (ry < 0 || rx < 0) continue;           // the tap leaves the window: zero padding
            const int id = ry * NP + rx;
            int k = 0;
            while (k < np && ids[k] != id) ++k;
            if (k == np) {
              ids[np] = id;
              planes[np] = sw.var[L - 2] + (size_t)id * sw.plane[L - 2];
              masks[np++] = 0u;
            }
            masks[k] |= 1u << ((dy + 1) * 3 + (dx + 1));
          }
        Scope sc(sd.kc[L - 1], st);
        VC_TRY(vc::conv_sps_planes_launch(planes, masks, np, sd.n_out[L - 2] / 8, sd.w[L - 1], sd.scale[L - 1], sd.bias[L - 1],
                                          sw.var[L - 1] + (size_t)(cy * NC + cx) * sw.plane[L - 1], 0, sd.n_out[L - 1],
                                          sd.nsplit[L - 1], sw.nb, sw.B, 9, 1, 0, 0, st));
      }
  }
  return VC_OK;
}

// lead / trailing halo rows of the intermediates are read by the next conv and written by no
// kernel: zero them once per workspace geometry (they stay zero across chunks of equal size)
int zero_halos(const Workspace& w, int n, int P, cudaStream_t st) {
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.a1, 16, n, P, st));
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.a2, 8, n, P, st));
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.l1, 2, n, P, st));
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.l2, 2, n, P, st));
  return VC_OK;
}

// Which token-stage kernel: the tcgen05 kernels (tokens_tm.cu, tokens_tc.cu) treat one patch as one M = 128 tile.  Measured
// per 131 072 patches, tokens_tm_kernel<4> vs the mma.sync kernel (transformer.cu): P = 11 4.63 vs 7.0 ms, P = 8 4.08 vs 4.59,
// P = 7 3.55 vs 3.84, P = 5 3.02 vs 3.22 (round 1, tokens_tc_kernel: the mma.sync kernel still won below P = 9): default =
// tcgen05 for 26 <= P*P + 1 <= 128 (P = 5 .. 11), mma.sync otherwise (larger token sets do not fit the tile; smaller ones
// are not measured).  VITCNN_TOKENS_IMPL=0 / 1 forces mma.sync / tcgen05 (where it applies).
bool use_tokens_tc(int P, int K) {
  static const int forced = [] {
    const char* e = getenv("VITCNN_TOKENS_IMPL");
    return e ? atoi(e) : -1;
  }();
  if (!vc::tokens_tc_supported(P, K) || forced == 0) return false;
  return forced == 1 || P * P + 1 >= 26;
}

// stems + token stage on packed inputs already in w.a0 / w.l0.
// stem_done: HSI stem convs already taken from the scene-level variants (1: w.a1 holds conv 1, 2: w.a2 conv 2, 3: w.f
//            slices 0-3 hold conv 3 -- or nothing does and `planes->h` is set: the token kernel reads the planes);
// lidar_done: w.f slices 4-7 already hold the LiDAR stem, or `planes->l` is set.
int forward_sps(const vc_model* m, const Workspace& w, int n, float* logits, const long long* out_index,
                uint8_t* argmax_map, cudaStream_t st, int stem_done = 0, bool lidar_done = false,
                const vc::TcPlanes* planes = nullptr) {
  const int P = m->P;
  if (stem_done < 1)
    VC_LAUNCH(KC_CONV_H1, st, vc::conv_sps_launch(w.a0, m->S1, m->w_h[0], m->scale_h[0], m->bias_h[0], w.a1, 0, 128, m->nsplit_h[0], n, P, 9,
                               1, 0, 0, st));
  if (stem_done < 2)
    VC_LAUNCH(KC_CONV_H2, st, vc::conv_sps_launch(w.a1, 16, m->w_h[1], m->scale_h[1], m->bias_h[1], w.a2, 0, 64, m->nsplit_h[1], n, P, 9, 1,
                               0, 0, st));
  if (stem_done < 3)
    VC_LAUNCH(KC_CONV_H3, st, vc::conv_sps_launch(w.a2, 8, m->w_h[2], m->scale_h[2], m->bias_h[2], w.f, 0, 32, m->nsplit_h[2], n, P, 9, 1, 0,
                               0, st));
  // Per-window LiDAR stem: three tcgen05 convs -- the same kernel and accumulation order as the shared LiDAR stem of
  // dense scenes, so forward() and the scene path agree bit for bit.  VITCNN_LIDAR_IMPL=fused selects the single-launch
  // mma.sync kernel (lidar_stem.cu: 7.3 instead of 9.6 ms per 642 k windows, results differ in the last bf16 bit).
  static const bool fused_lidar = [] {
    const char* e = getenv("VITCNN_LIDAR_IMPL");
    return e && e[0] == 'f';
  }();
  if (!lidar_done && fused_lidar && m->lidar_blob && m->C2 <= 8) {
    Scope sc(KC_CONV_L, st);
    const int rc = vc::lidar_stem_launch(w.l0, m->lidar_blob, w.f, 4, n, P, st);
    if (rc == VC_OK) lidar_done = true;
    else if (rc != VC_ERR_UNSUPPORTED) return fail(rc, "lidar_stem_launch");
  }
  if (!lidar_done) {
    VC_LAUNCH(KC_CONV_L, st, vc::conv_sps_launch(w.l0, m->S2, m->w_l[0], m->scale_l[0], m->bias_l[0], w.l1, 0, 16, m->nsplit_l[0], n, P,
                                                 9, 1, 0, 0, st));
    VC_LAUNCH(KC_CONV_L, st, vc::conv_sps_launch(w.l1, 2, m->w_l[1], m->scale_l[1], m->bias_l[1], w.l2, 0, 16, m->nsplit_l[1], n, P, 9, 1,
                                                 0, 0, st));
    VC_LAUNCH(KC_CONV_L, st, vc::conv_sps_launch(w.l2, 2, m->w_l[2], m->scale_l[2], m->bias_l[2], w.f, 4, 32, m->nsplit_l[2], n, P, 9, 1,
                                                 0, 0, st));
  }
  // token stage (see use_tokens_tc)
  if (use_tokens_tc(P, m->K))
    VC_LAUNCH(KC_TOKENS, st, vc::tokens_tc_launch(w.f, m->tparams, n, P, m->K, logits, out_index, argmax_map, w.tscr, planes, st));
  else
    VC_LAUNCH(KC_TOKENS, st, vc::transformer_fwd_launch(w.f, m->tparams, n, P, m->K, logits, out_index, argmax_map, 0, 0, nullptr, st));
  return VC_OK;
}

int check_model(const vc_model* m) {
  if (!m) return VC_ERR_ARG;
  if (m->P < 1 || m->P > 15 || m->K < 1 || m->K > 64 || m->C1 < 1 || m->C2 < 1) return VC_ERR_ARG;
  if (m->S1 != slices_for(m->C1) || m->S2 != slices_for(m->C2)) return VC_ERR_ARG;
  for (int i = 0; i < 3; ++i)
    if (!m->w_h[i] || !m->w_l[i] || !m->scale_h[i] || !m->scale_l[i] || !m->bias_h[i] || !m->bias_l[i]) return VC_ERR_ARG;
  return m->tparams ? VC_OK : VC_ERR_ARG;
}

}  // namespace

extern "C" {

int vc_abi_version(void) { return VC_ABI_VERSION; }
int64_t vc_lidar_blob_bytes(void) { return (int64_t)vc::lidar_blob_bytes(); }
const char* vc_last_error(void) { return g_err; }

int64_t vc_sps_rows(int32_t n_patches, int32_t P) { return vc::sps_rows(n_patches, P); }

int64_t vc_launch_count(void) { return g_launches.load(); }

int vc_profile_begin(void) {
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  g_prof->clear();
  g_prof_on = true;
  return VC_OK;
}

int vc_profile_end(double* ms_per_class, int64_t* launches_per_class, int32_t n_classes) {
  g_prof_on = false;
  if (!g_prof) return fail(VC_ERR_ARG, "vc_profile_end without vc_profile_begin");
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(VC_ERR_CUDA, "vc_profile_end: sync");
  for (int i = 0; i < n_classes; ++i) {
    if (ms_per_class) ms_per_class[i] = 0.0;
    if (launches_per_class) launches_per_class[i] = 0;
  }
  for (const ProfRec& r : *g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (r.cls < n_classes) {
      if (ms_per_class) ms_per_class[r.cls] += ms;
      if (launches_per_class) launches_per_class[r.cls] += 1;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof->clear();
  return VC_OK;
}

int64_t vc_workspace_bytes(int32_t n_patches, int32_t P, int32_t C1, int32_t C2) {
  if (n_patches <= 0 || P < 1) return -1;
  return carve(nullptr, n_patches, P, slices_for(C1), slices_for(C2)).bytes + 256;
}

int32_t vc_tparams_layout(int32_t P, int32_t K, int64_t* out, int32_t n) {
  const vc::TLayout L = vc::tlayout(P, K);
  int64_t v[10 + 12 * vc::kLayers];
  int k = 0;
  v[k++] = L.total; v[k++] = L.wfus; v[k++] = L.fus_scale; v[k++] = L.fus_bias; v[k++] = L.cls;
  v[k++] = L.lnf_g; v[k++] = L.lnf_b; v[k++] = L.whead; v[k++] = L.bhead; v[k++] = L.pos;
  for (int l = 0; l < vc::kLayers; ++l) {
    const vc::TLayerOff& o = L.layer[l];
    v[k++] = o.wqkv; v[k++] = o.wproj; v[k++] = o.wfc1; v[k++] = o.wfc2; v[k++] = o.ln1_g; v[k++] = o.ln1_b;
    v[k++] = o.bqkv; v[k++] = o.bproj; v[k++] = o.ln2_g; v[k++] = o.ln2_b; v[k++] = o.bfc1; v[k++] = o.bfc2;
  }
  if (out)
    for (int i = 0; i < k && i < n; ++i) out[i] = v[i];
  return k;
}

int vc_gather_patches_f32(const float* img1, const float* img2, const void* gt, int32_t gt_elem_bytes, int32_t H,
                          int32_t W, int32_t C1, int32_t C2, const int32_t* xy, const uint8_t* ops, int32_t n, int32_t P,
                          int32_t center_mode, float* hsi, float* lidar, int64_t* labels, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return VC_OK;
  if (!xy || n < 0) return fail(VC_ERR_ARG, "vc_gather_patches_f32: bad arguments");
  if (img1 && hsi) {      // tensor-map TMA form where the raster allows it (C1 % 4 == 0, aligned), else the generic kernel
    int rc = vc::gather_tma_launch(img1, H, W, C1, xy, ops, n, P, center_mode, hsi, st);
    if (rc == VC_ERR_UNSUPPORTED) rc = vc::gather_f32_launch(img1, H, W, C1, xy, ops, n, P, center_mode, hsi, st);
    if (rc != VC_OK) return fail(rc, "patch gather (hsi)");
  }
  if (img2 && lidar) VC_TRY(vc::gather_f32_launch(img2, H, W, C2, xy, ops, n, P, center_mode, lidar, st));
  if (gt && labels)
    VC_TRY(vc::gather_labels_launch(gt, gt_elem_bytes, H, W, xy, ops, n, P, center_mode, (long long*)labels, st));
  return VC_OK;
}

int vc_scene_index(const int32_t* xs, const int32_t* ys, int32_t nx, int32_t ny, int32_t first, int32_t count,
                   int32_t W, int32_t C1, int32_t C2, int32_t P, int64_t* off1, int64_t* off2, int64_t* out_idx,
                   int32_t* xy, void* stream) {
  if (count == 0) return VC_OK;
  VC_TRY(vc::scene_index_launch(xs, ys, nx, ny, first, count, W, C1, C2, P, 0, (long long*)off1, (long long*)off2,
                                (long long*)out_idx, xy, (cudaStream_t)stream));
  return VC_OK;
}

int vc_confusion_matrix(const void* prediction, int32_t pred_elem_bytes, const void* target, int32_t target_elem_bytes, int64_t n,
                        int32_t n_classes, uint64_t ignored_mask, int64_t* cm, void* stream) {
  VC_LAUNCH(KC_MISC, (cudaStream_t)stream, vc::confusion_launch(prediction, pred_elem_bytes, target, target_elem_bytes, n, n_classes,
                                                               ignored_mask, (long long*)cm, (cudaStream_t)stream));
  return VC_OK;
}

int vc_minmax_normalise(float* img, int64_t n_pixels, int32_t C, int32_t per_band, float* scratch, void* stream) {
  VC_LAUNCH(KC_MISC, (cudaStream_t)stream, vc::minmax_normalise_launch(img, n_pixels, C, per_band, scratch, (cudaStream_t)stream));
  return VC_OK;
}

int vc_pack_sps(const float* src, int64_t sb, int64_t sc, int64_t si, int64_t sj, const int64_t* patch_off,
                int32_t n_patches, int32_t C, int32_t P, void* sps, int32_t S, void* stream) {
  VC_TRY(vc::pack_sps_launch(src, sb, sc, si, sj, (const long long*)patch_off, nullptr, n_patches, C, P, sps, S,
                             (cudaStream_t)stream));
  return VC_OK;
}

int vc_conv_sps(const void* in_sps, int32_t S_in, const void* w_packed, const float* scale, const float* bias,
                void* out_sps, int32_t out_slice_off, int32_t n_out, int32_t nsplit, int32_t n_patches, int32_t P,
                int32_t taps, int32_t relu, int32_t impl, int32_t debug_flags, void* stream) {
  VC_TRY(vc::conv_sps_launch(in_sps, S_in, w_packed, scale, bias, out_sps, out_slice_off, n_out, nsplit, n_patches, P,
                             taps, relu, impl, debug_flags, (cudaStream_t)stream));
  return VC_OK;
}

int vc_tokens_forward(const void* f_sps, const void* tparams, int32_t n_patches, int32_t P, int32_t K, float* logits,
                      const int64_t* out_index, uint8_t* argmax_map, void* stream) {
  VC_TRY(vc::transformer_fwd_launch(f_sps, tparams, n_patches, P, K, logits, (const long long*)out_index, argmax_map,
                                    0, 0, nullptr, (cudaStream_t)stream));
  return VC_OK;
}

int64_t vc_tokens_tc_scratch_bytes(int32_t n_patches) { return (int64_t)vc::tokens_tc_scratch_bytes(n_patches); }

int vc_tokens_forward_tc(const void* f_sps, const void* tparams, int32_t n_patches, int32_t P, int32_t K, float* logits,
                         const int64_t* out_index, uint8_t* argmax_map, void* scratch, int64_t scratch_bytes, void* stream) {
  if (!f_sps || !tparams || !logits || !scratch || scratch_bytes < vc_tokens_tc_scratch_bytes(n_patches))
    return fail(VC_ERR_ARG, "vc_tokens_forward_tc: bad argument / scratch too small");
  if (!vc::tokens_tc_supported(P, K)) return fail(VC_ERR_UNSUPPORTED, "vc_tokens_forward_tc: needs P*P + 1 <= 128");
  VC_TRY(vc::tokens_tc_launch(f_sps, tparams, n_patches, P, K, logits, (const long long*)out_index, argmax_map, scratch,
                              nullptr, (cudaStream_t)stream));
  return VC_OK;
}

int64_t vc_wgrad_workspace_bytes(int32_t SB, int32_t taps) { return (int64_t)vc::wgrad_workspace_bytes(SB, taps); }

int vc_wgrad_sps(const void* a_sps, int32_t SA, const void* b_sps, int32_t SB, int32_t n_patches, int32_t P, int32_t taps,
                 int32_t shift_on_a, void* workspace, int64_t workspace_bytes, float* out, int32_t M, int32_t N,
                 int64_t sm, int64_t sn, int64_t st, int32_t bias_col, float* out_bias, int32_t accumulate, void* stream) {
  if (!a_sps || !b_sps || !out || workspace_bytes < vc_wgrad_workspace_bytes(SB, taps))
    return fail(VC_ERR_ARG, "vc_wgrad_sps: bad arguments");
  VC_TRY(vc::wgrad_sps_launch(a_sps, SA, b_sps, SB, n_patches, P, taps, shift_on_a, workspace, out, M, N, sm, sn, st,
                              bias_col, out_bias, accumulate, (cudaStream_t)stream));
  return VC_OK;
}

int vc_forward_patches(const vc_model* m, const float* hsi, const int64_t hs[4], const float* lidar,
                       const int64_t ls[4], int32_t n, void* workspace, int64_t workspace_bytes, float* logits,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return VC_OK;
  if (check_model(m) != VC_OK || !hsi || !lidar || !hs || !ls || n < 0 || !workspace || !logits)
    return fail(VC_ERR_ARG, "vc_forward_patches: bad arguments");
  if (workspace_bytes < vc_workspace_bytes(n, m->P, m->C1, m->C2)) return fail(VC_ERR_ARG, "workspace too small");
  const Workspace w = carve(workspace, n, m->P, m->S1, m->S2);
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(hsi, hs[0], hs[1], hs[2], hs[3], nullptr, nullptr, n, m->C1, m->P, w.a0, m->S1, st));
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(lidar, ls[0], ls[1], ls[2], ls[3], nullptr, nullptr, n, m->C2, m->P, w.l0, m->S2, st));
  VC_TRY(zero_halos(w, n, m->P, st));
  return forward_sps(m, w, n, logits, nullptr, nullptr, st);
}

int vc_scene_infer(const vc_model* m, const float* img1, const float* img2, int32_t H, int32_t W, const int32_t* xs,
                   const int32_t* ys, int32_t nx, int32_t ny, int64_t first_window, int64_t n_windows, int32_t chunk,
                   void* workspace, int64_t workspace_bytes, float* logits_map, uint8_t* argmax_map, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_windows == 0) return VC_OK;
  if (check_model(m) != VC_OK || !img1 || !img2 || !xs || !ys || chunk <= 0 || !workspace || !logits_map ||
      first_window < 0 || n_windows < 0 || first_window + n_windows > (int64_t)nx * ny)
    return fail(VC_ERR_ARG, "vc_scene_infer: bad arguments");
  if (workspace_bytes < vc_workspace_bytes(chunk, m->P, m->C1, m->C2)) return fail(VC_ERR_ARG, "workspace too small");
  const long long s1 = (long long)W * m->C1, s2 = (long long)W * m->C2;
  // ---- shared stem: border-class variants of the stem convs over the scene blocks, once per call ----
  const int mode = scene_mode(m, H, W, chunk, n_windows, workspace_bytes);
  const StemDesc hd = hsi_stem(m), ld = lidar_stem(m);
  const SceneWs sw = carve_scene(workspace, chunk_ws_end(m, chunk), hd, H, W, m->P, mode);
  const long long hsi_end = mode ? sw.bytes : chunk_ws_end(m, chunk);
  SceneWs lw;
  memset(&lw, 0, sizeof(lw));
  if (lidar_shared_ok(m, H, W)) {
    lw = carve_scene(workspace, hsi_end, ld, H, W, m->P, 3);
    if (lw.bytes <= 0 || lw.bytes > workspace_bytes) memset(&lw, 0, sizeof(lw));     // the caller's workspace is the smaller kind
  }
  const bool lid_shared = lw.bytes > 0;
  if (mode) VC_TRY(shared_stem(hd, sw, img1, H, W, m->P, st));
  if (lid_shared) VC_TRY(shared_stem(ld, lw, img2, H, W, m->P, st));
  // the tcgen05 token kernel takes depth-3 stem outputs straight from the variant planes: no per-window copy of them
  const bool direct = use_tokens_tc(m->P, m->K);
  for (int64_t done = 0; done < n_windows; done += chunk) {
    const int n = (int)((n_windows - done) < chunk ? (n_windows - done) : chunk);
    const int first = (int)(first_window + done);
    // the SPS geometry depends on n: carve per chunk (only the last chunk differs)
    const Workspace w = carve(workspace, n, m->P, m->S1, m->S2);
    const long long sl = vc::sps_rows(n, m->P) * 16;
    if ((done == 0 || n != chunk) && (mode < 3 || !lid_shared)) VC_TRY(zero_halos(w, n, m->P, st));
    // window corners of the chunk: raster offsets for the per-window gathers, (row, column) pairs for the token kernel when it
    // reads the variant planes itself (the two uses are exclusive: the pairs live in the off1 array)
    const bool want_xy = direct && ((mode == 3) || lid_shared);
    VC_LAUNCH(KC_INDEX, st, vc::scene_index_launch(xs, ys, nx, ny, first, n, W, m->C1, m->C2, m->P, m->K, want_xy ? nullptr : w.off1, w.off2,
                                                   w.oidx, want_xy ? reinterpret_cast<int*>(w.off1) : nullptr, st));
    vc::TcPlanes tp;
    memset(&tp, 0, sizeof(tp));
    tp.xy = reinterpret_cast<const int*>(w.off1);
    if (mode == 3 && direct) {
      tp.h = (const __nv_bfloat16*)sw.var[2];
      tp.RTb = sw.RTb; tp.B = sw.B; tp.D = sw.D; tp.rowterm = sw.rowterm; tp.colterm = sw.colterm;
    } else if (mode) {
      VC_LAUNCH(KC_PACK, st, vc::border_gather_launch(sw.var[sw.D - 1], mode == 3 ? 4 : mode == 2 ? 8 : 16, sw.B, sw.D, H, W, xs, ys, ny,
                                                      first, n, m->P, mode == 3 ? w.f : mode == 2 ? w.a2 : w.a1, st));
    } else {
      // strip-staged gather (stride-1 runs reuse the overlap of consecutive windows); the generic
      // per-row gather is the fallback for geometries the strip kernel does not take
      Scope sc(KC_PACK, st);
      int rc = vc::pack_scene_launch(img1, W, m->C1, xs, ys, nx, ny, first, n, m->P, w.a0, m->S1, st);
      if (rc == VC_ERR_UNSUPPORTED) rc = vc::pack_sps_launch(img1, 0, 1, s1, m->C1, w.off1, nullptr, n, m->C1, m->P, w.a0, m->S1, st);
      if (rc != VC_OK) return fail(rc, "scene gather (hsi)");
    }
    if (lid_shared && direct) {
      tp.l = (const __nv_bfloat16*)lw.var[2];      // same block geometry as the depth-3 HSI planes (same B, D = 3)
      tp.RTb = lw.RTb; tp.B = lw.B; tp.D = lw.D; tp.rowterm = lw.rowterm; tp.colterm = lw.colterm;
    } else if (lid_shared) {
      VC_LAUNCH(KC_PACK, st, vc::border_gather_launch(lw.var[2], 4, lw.B, lw.D, H, W, xs, ys, ny, first, n, m->P, w.f + 4 * sl, st));
    } else {
      Scope sc(KC_PACK, st);
      int rc = vc::pack_scene_launch(img2, W, m->C2, xs, ys, nx, ny, first, n, m->P, w.l0, m->S2, st);
      if (rc == VC_ERR_UNSUPPORTED) rc = vc::pack_sps_launch(img2, 0, 1, s2, m->C2, w.off2, nullptr, n, m->C2, m->P, w.l0, m->S2, st);
      if (rc != VC_OK) return fail(rc, "scene gather (lidar)");
    }
    VC_TRY(forward_sps(m, w, n, logits_map, w.oidx, argmax_map, st, mode, lid_shared, (tp.h || tp.l) ? &tp : nullptr));
  }
  return VC_OK;
}

int32_t vc_scene_block(int32_t H, int32_t W, int32_t depth) { return scene_block(H, W, depth); }

int32_t vc_scene_shared_depth(const vc_model* m, int32_t H, int32_t W, int32_t chunk, int64_t n_windows, int64_t workspace_bytes) {
  if (check_model(m) != VC_OK || chunk <= 0) return -1;
  return scene_mode(m, H, W, chunk, n_windows, workspace_bytes);
}

int64_t vc_scene_workspace_bytes(const vc_model* m, int32_t H, int32_t W, int32_t chunk) {
  if (check_model(m) != VC_OK || chunk <= 0) return -1;
  long long need = vc_workspace_bytes(chunk, m->P, m->C1, m->C2);
  if (!m->w_h1_border) return need;
  const StemDesc hd = hsi_stem(m), ld = lidar_stem(m);
  const bool lid = lidar_shared_ok(m, H, W);
  for (int mode = 0; mode <= 3; ++mode) {
    long long end = chunk_ws_end(m, chunk);
    if (mode) {
      const SceneWs sw = carve_scene(nullptr, end, hd, H, W, m->P, mode);
      if (sw.bytes <= 0) continue;
      end = sw.bytes;
    }
    if (lid) {
      const SceneWs lw = carve_scene(nullptr, end, ld, H, W, m->P, 3);
      if (lw.bytes > 0) end = lw.bytes;
    }
    if (end > need) need = end;
  }
  return need;
}

}  // extern "C"

// =================================================================================================
// Training: forward with batch statistics, backward of the whole model (see include/vitcnn.h)
// =================================================================================================
namespace {

struct ConvPlan {
  int cin, cout, taps, S_in, n_out, nsplit;     // forward operand
  int has_dgrad, d_S_in, d_n_out, d_nsplit;     // data-gradient conv: n_out channels back to S_in*8
  int shift_on_a;                               // weight-gradient orientation (A = layer input)
};

int choose_nsplit(int s_in, int n_out, int taps) {
  for (int ns = 1; ns <= 8; ns *= 2) {
    const int ncta = n_out / ns;
    if (n_out % ns == 0 && ncta % 16 == 0 && ncta <= 128 && (long long)taps * s_in * ncta * 16 <= 180000) return ns;
  }
  return -1;
}

bool make_plans(const vc_train* t, ConvPlan pl[7]) {
  const int hp[3] = {128, 64, 32}, lp[3] = {8, 16, 32};
  int cin = t->C1;
  for (int i = 0; i < 3; ++i) { pl[i].cin = cin; pl[i].cout = hp[i]; pl[i].taps = 9; cin = hp[i]; }
  cin = t->C2;
  for (int i = 0; i < 3; ++i) { pl[3 + i].cin = cin; pl[3 + i].cout = lp[i]; pl[3 + i].taps = 9; cin = lp[i]; }
  pl[6].cin = 64; pl[6].cout = 32; pl[6].taps = 1;
  for (int i = 0; i < 7; ++i) {
    ConvPlan& p = pl[i];
    p.S_in = slices_for(p.cin);
    p.n_out = (p.cout + 15) / 16 * 16;
    p.nsplit = choose_nsplit(p.S_in, p.n_out, p.taps);
    p.has_dgrad = (i != 0 && i != 3);
    p.d_S_in = p.n_out / 8;
    p.d_n_out = p.S_in * 8;
    p.d_nsplit = p.has_dgrad ? choose_nsplit(p.d_S_in, p.d_n_out, p.taps) : 1;
    p.shift_on_a = (p.S_in * 8 <= 128 && p.S_in * 8 >= p.n_out) ? 1 : 0;
    if (p.nsplit < 0 || p.d_nsplit < 0 || p.S_in > 32) return false;
  }
  return true;
}

const int kTokSlices[10] = {6, 12, 6, 4, 6, 16, 18, 4, 6, 12};   // xln1_0 dqkv_0 xo0 dxa0 xln2_0 dh0 xh0 dxb0 xln1_1 dqkv_1
const int kClsSlices[8] = {6, 4, 6, 16, 18, 4, 6, 8};            // c_xo c_dxa c_xln2 c_dh c_xh c_dxb c_xc c_dlog
const int kLinSB[9] = {6, 6, 6, 18, 6, 6, 6, 18, 6};   // slices of the N-side operand of the nine linear weight-gradient GEMMs
const int kTokOnes[10] = {4, -1, 4, -1, 4, -1, 16, -1, 4, -1};    // slice holding the constant-one channel
const int kClsOnes[8] = {4, -1, 4, -1, 16, -1, 4, -1};

struct TrainWs {
  uint8_t *a0, *l0, *f;
  uint8_t *y[7], *z[7];
  uint8_t *dzf, *df, *dzh2, *dzh1, *dzl2, *dzl1;
  uint8_t *tok[10], *cls[8];
  uint8_t *wf[7], *wd[7], *blob;
  float *bias_pad[7], *ones, *zeros;
  float *bn_scale[7], *bn_shift[7], *bn_mean[7], *bn_rstd[7];
  double* sums;                 // [14][256]: batch sums of layer i forward at 256 i, backward at 256 (7 + i)
  uint8_t *wgl[9], *wgc[7];     // split-K partials of the nine linear / seven conv weight-gradient GEMMs (reduced in one launch)
  long long *off1, *off2;
  long long RT, RTt, RTc;
  long long bytes;
};

TrainWs carve_train(void* base, const vc_train* t, const ConvPlan pl[7], int n) {
  TrainWs w;
  const int P = t->P, TP = (P * P + 1 + 15) / 16 * 16;
  w.RT = vc::sps_rows(n, P);
  w.RTt = ((long long)n * TP + 127) / 128 * 128;
  w.RTc = ((long long)n + 127) / 128 * 128;
  const long long sl = w.RT * 16;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](long long bytes) { uint8_t* r = p; p += (bytes + 255) & ~255LL; return r; };
  w.a0 = take(sl * pl[0].S_in);
  w.l0 = take(sl * pl[3].S_in);
  w.f = take(sl * 8);
  for (int i = 0; i < 7; ++i) w.y[i] = take(sl * (pl[i].n_out / 8));
  w.z[0] = take(sl * 16); w.z[1] = take(sl * 8); w.z[2] = w.f;
  w.z[3] = take(sl * 2); w.z[4] = take(sl * 2); w.z[5] = w.f + sl * 4;
  w.z[6] = take(sl * 4);
  w.dzf = take(sl * 4); w.df = take(sl * 8); w.dzh2 = take(sl * 8); w.dzh1 = take(sl * 16);
  w.dzl2 = take(sl * 2); w.dzl1 = take(sl * 2);
  for (int i = 0; i < 10; ++i) w.tok[i] = take(w.RTt * 16 * kTokSlices[i]);
  for (int i = 0; i < 8; ++i) w.cls[i] = take(w.RTc * 16 * kClsSlices[i]);
  for (int i = 0; i < 7; ++i) {
    w.wf[i] = take((long long)pl[i].taps * pl[i].S_in * pl[i].n_out * 16);
    w.wd[i] = take(pl[i].has_dgrad ? (long long)pl[i].taps * pl[i].d_S_in * pl[i].d_n_out * 16 : 0);
  }
  w.blob = take(vc::tlayout(P, t->K).total);
  for (int i = 0; i < 7; ++i) w.bias_pad[i] = reinterpret_cast<float*>(take(128 * 4));
  w.ones = reinterpret_cast<float*>(take(256 * 4));
  w.zeros = reinterpret_cast<float*>(take(256 * 4));
  for (int i = 0; i < 7; ++i) {
    w.bn_scale[i] = reinterpret_cast<float*>(take(128 * 4));
    w.bn_shift[i] = reinterpret_cast<float*>(take(128 * 4));
    w.bn_mean[i] = reinterpret_cast<float*>(take(128 * 4));
    w.bn_rstd[i] = reinterpret_cast<float*>(take(128 * 4));
  }
  w.sums = reinterpret_cast<double*>(take(14 * 256 * 8));
  for (int j = 0; j < 9; ++j) w.wgl[j] = take((long long)vc::wgrad_workspace_bytes(kLinSB[j], 1));
  for (int i = 0; i < 7; ++i) {
    const int SB = pl[i].shift_on_a ? pl[i].n_out / 8 : pl[i].S_in;
    long long b = (long long)vc::wgrad_workspace_bytes(SB, pl[i].taps);
    if ((long long)vc::wgrad_small_workspace_bytes() > b) b = (long long)vc::wgrad_small_workspace_bytes();
    w.wgc[i] = take(b);
  }
  w.off1 = reinterpret_cast<long long*>(take(8LL * n));
  w.off2 = reinterpret_cast<long long*>(take(8LL * n));
  w.bytes = p - reinterpret_cast<uint8_t*>(base);
  return w;
}

__global__ void bump_seed_kernel(uint32_t* seed) { *seed = *seed * 747796405u + 2891336453u; }
inline unsigned int drop_threshold(const vc_train* t) {
  if (!(t->dropout > 0.f) || !t->drop_seed) return 0u;
  const float v = t->dropout * 65536.f + 0.5f;
  return v >= 65535.f ? 65535u : (unsigned int)v;
}

__global__ void fill_ones_slice_kernel(__nv_bfloat16* slice, long long rows) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
    slice[r * 8] = __float2bfloat16_rn(1.f);
}
__global__ void fill_f32_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int check_train(const vc_train* t, ConvPlan pl[7]) {
  if (!t || !t->params || !t->grads || t->P < 1 || (t->P > 11 && t->P != 13 && t->P != 15) || t->K < 1 || t->K > 64 || t->C1 < 1 || t->C2 < 1 ||
      !t->blob_segments || t->n_blob_segments <= 0 || t->dropout < 0.f || t->dropout >= 1.f || (t->dropout > 0.f && !t->drop_seed))
    return VC_ERR_ARG;
  return make_plans(t, pl) ? VC_OK : VC_ERR_UNSUPPORTED;
}

// Side streams of the training step: the weight-gradient GEMMs of a layer depend only on that layer's dz and on forward
// activations, never on each other or on the data-gradient chain, so they run on a second stream beside the
// BatchNorm-backward -> data-gradient chain (fork / join with events; under CUDA-graph capture they become parallel
// branches of the graph).  At 512 samples per GPU the chain and the GEMMs are ~0.3 ms each: the step is bound by
// launch latencies, not by SM time.  VITCNN_TRAIN_STREAMS=0 keeps everything on the caller's stream.
struct SideStreams {
  cudaStream_t s[2];
  cudaEvent_t ev[32];
  int next;
  bool ok;
};
SideStreams* side_streams() {
  static SideStreams pool[16];
  static bool tried[16] = {false};
  static const bool off = [] {
    const char* e = getenv("VITCNN_TRAIN_STREAMS");
    return e && e[0] == '0';
  }();
  if (off) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStreams& p = pool[dev];
  if (!tried[dev]) {
    tried[dev] = true;
    p.ok = true;
    for (int i = 0; i < 2; ++i) p.ok = p.ok && cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 32; ++i) p.ok = p.ok && cudaEventCreateWithFlags(&p.ev[i], cudaEventDisableTiming) == cudaSuccess;
    p.next = 0;
  }
  return p.ok ? &p : nullptr;
}
// `to` waits for everything queued on `from` so far
int stream_dep(SideStreams* ss, cudaStream_t from, cudaStream_t to) {
  cudaEvent_t e = ss->ev[ss->next];
  ss->next = (ss->next + 1) & 31;
  if (cudaEventRecord(e, from) != cudaSuccess || cudaStreamWaitEvent(to, e, 0) != cudaSuccess) return VC_ERR_CUDA;
  return VC_OK;
}

int train_forward_core(const vc_train* t, const ConvPlan pl[7], const TrainWs& w, int n, float* logits, cudaStream_t st) {
  const int P = t->P;
  // bf16 operand forms of the current fp32 master weights, conv biases, cleared BatchNorm sums: ONE launch
  vc::PrepTable pt;
  pt.n = 0;
  auto add = [&](int type, const float* src, void* dst, long long n_) -> vc::PrepJob& {
    vc::PrepJob& j = pt.job[pt.n++];
    memset(&j, 0, sizeof(j));
    j.type = type; j.src = src; j.dst = dst; j.n = n_;
    return j;
  };
  for (int i = 0; i < 7; ++i) {
    const ConvPlan& c = pl[i];
    vc::PrepJob& f = add(0, t->params + t->off[4 * i], w.wf[i], (long long)c.taps * c.S_in * c.n_out * 8);
    f.cout = c.cout; f.cin = c.cin; f.taps = c.taps; f.transpose = 0; f.S_in = c.S_in; f.n_out = c.n_out; f.nsplit = c.nsplit;
    if (c.has_dgrad) {
      vc::PrepJob& d = add(0, t->params + t->off[4 * i], w.wd[i], (long long)c.taps * c.d_S_in * c.d_n_out * 8);
      d.cout = c.cout; d.cin = c.cin; d.taps = c.taps; d.transpose = 1; d.S_in = c.d_S_in; d.n_out = c.d_n_out; d.nsplit = c.d_nsplit;
    }
    add(1, t->params + t->off[4 * i + 1], w.bias_pad[i], c.cout);
  }
  add(3, nullptr, w.sums, 14 * 256);
  VC_LAUNCH(KC_MISC, st, vc::train_prep_launch(&pt, st));
  VC_LAUNCH(KC_MISC, st, vc::pack_segments_launch(t->params, w.blob, (const long long*)t->blob_segments, t->n_blob_segments, st));
  // stems: conv (+bias) -> raw y -> BatchNorm with batch statistics -> ReLU -> z.  The LiDAR chain (layers 3-5) is independent
  // of the HSI chain (0-2) until the fusion conv (6): it runs on a side stream (a parallel branch of the captured graph)
  SideStreams* ss = side_streams();
  cudaStream_t main_st = st, lidar_st = ss ? ss->s[1] : st;
  if (ss) VC_TRY(stream_dep(ss, main_st, lidar_st));
  for (int i = 0; i < 7; ++i) {
    const ConvPlan& c = pl[i];
    if (i == 6 && ss) VC_TRY(stream_dep(ss, lidar_st, main_st));
    st = (i >= 3 && i < 6) ? lidar_st : main_st;
    const uint8_t* in = (i == 0) ? w.a0 : (i == 3) ? w.l0 : (i == 6) ? w.f : w.z[i - 1];
    const int cls = i == 0 ? KC_CONV_H1 : i == 1 ? KC_CONV_H2 : i == 2 ? KC_CONV_H3 : i < 6 ? KC_CONV_L : KC_TOKENS;
    VC_LAUNCH(cls, st, vc::conv_sps_launch(in, c.S_in, w.wf[i], w.ones, w.bias_pad[i], w.y[i], 0, c.n_out, c.nsplit, n, P, c.taps,
                                           0, 0, 0, st));
    VC_LAUNCH(KC_BN, st, vc::bn_forward_fused_launch(w.y[i], w.z[i], c.n_out / 8, c.cout, n, P, t->params + t->off[4 * i + 2],
                                                     t->params + t->off[4 * i + 3], t->bn_eps, t->bn_momentum, t->bn_running_mean[i],
                                                     t->bn_running_var[i], (long long*)t->bn_num_batches[i], w.sums + 256 * i,
                                                     w.bn_scale[i], w.bn_shift[i], w.bn_mean[i], w.bn_rstd[i], 1, st));
  }
  st = main_st;
  const unsigned int thr = drop_threshold(t);
  if (thr) bump_seed_kernel<<<1, 1, 0, st>>>(t->drop_seed);     // a fresh mask set for this step
  VC_LAUNCH(KC_TOKENS, st, vc::transformer_fwd_launch(w.z[6], w.blob, n, P, t->K, logits, nullptr, nullptr, 1, thr, t->drop_seed, st));
  return VC_OK;
}

int conv_wgrad(const vc_train* t, const ConvPlan& c, int i, const void* x, const void* dy, const TrainWs& w, int n,
               cudaStream_t st, vc::WgradReduceTable* defer) {
  float* out = t->grads + t->off[4 * i];
  const int P = t->P;
  if (c.taps == 9 && c.n_out <= 32 && c.S_in == 2) {   // thin LiDAR layers: mma.sync kernel
    const int rc = vc::wgrad_small_launch(dy, c.n_out / 8, x, c.S_in, n, P, w.wgc[i], out, c.cout, c.cin, st);
    if (rc != VC_ERR_UNSUPPORTED) return rc;
  }
  if (c.shift_on_a)
    return vc::wgrad_sps_launch(x, c.S_in, dy, c.n_out / 8, n, P, c.taps, 1, w.wgc[i], out, c.cin, c.cout, c.taps,
                                (long long)c.cin * c.taps, 1, -1, nullptr, 0, st, defer);
  return vc::wgrad_sps_launch(dy, c.n_out / 8, x, c.S_in, n, P, c.taps, 0, w.wgc[i], out, c.cout, c.cin, (long long)c.cin * c.taps,
                              c.taps, 1, -1, nullptr, 0, st, defer);
}

// j: which of the nine linear GEMMs (its own split-K workspace: the reductions run in one launch at the end of the pass)
int linear_wgrad(const vc_train* t, int j, const void* dy, int SA, const void* x, int SB, long long rows, int M, int N, int iw,
                 const TrainWs& w, cudaStream_t st, vc::WgradReduceTable* defer) {
  if (SB != kLinSB[j]) return VC_ERR_ARG;
  return vc::wgrad_sps_launch(dy, SA, x, SB, (int)(rows / 128), 0, 1, 0, w.wgl[j], t->grads + t->off[iw], M, N, N, 1, 0, N,
                              t->grads + t->off[iw + 1], 0, st, defer);
}

int train_backward_core(const vc_train* t, const ConvPlan pl[7], const TrainWs& w, int n, const float* dlogits,
                        cudaStream_t st) {
  const int P = t->P, K = t->K, T = P * P + 1;
  float* G = t->grads;
  // ---- token stage ----
  const int small_idx[12] = {30, 31, 36, 37, 42, 43, 48, 49, 54, 55, 28, 29};
  float* small[12];
  vc::PrepTable pt;             // the small gradients the backward kernel accumulates with atomics start from zero: one launch
  pt.n = 0;
  for (int i = 0; i < 12; ++i) {
    small[i] = G + t->off[small_idx[i]];
    vc::PrepJob& j = pt.job[pt.n++];
    memset(&j, 0, sizeof(j));
    j.type = 2; j.dst = small[i]; j.n = small_idx[i] == 29 ? (long long)T * 32 : 32;
  }
  VC_LAUNCH(KC_MISC, st, vc::train_prep_launch(&pt, st));
  vc::WgradReduceTable rt;      // split-K reductions of every weight-gradient GEMM of the pass: one launch at the end
  rt.n = 0;
  VC_LAUNCH(KC_TOKENS_BWD, st, vc::transformer_bwd_launch(w.z[6], w.blob, dlogits, w.dzf, (void* const*)w.tok, (void* const*)w.cls,
                                                         small, n, P, K, drop_threshold(t), t->drop_seed, st));
  SideStreams* ss = side_streams();
  cudaStream_t main_st = st;
  if (ss) {                     // the weight-gradient GEMMs run beside the BatchNorm-backward / data-gradient chain
    VC_TRY(stream_dep(ss, main_st, ss->s[0]));
    st = ss->s[0];
  }
  // linear-layer weight / bias gradients: contraction over all token rows on the tensor cores
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 0, w.tok[1], 12, w.tok[0], 6, w.RTt, 96, 32, 32, w, st, &rt));    // block 0 qkv
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 1, w.tok[3], 4, w.tok[2], 6, w.RTt, 32, 32, 34, w, st, &rt));     //         proj
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 2, w.tok[5], 16, w.tok[4], 6, w.RTt, 128, 32, 38, w, st, &rt));   //         fc1
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 3, w.tok[7], 4, w.tok[6], 18, w.RTt, 32, 128, 40, w, st, &rt));   //         fc2
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 4, w.tok[9], 12, w.tok[8], 6, w.RTt, 96, 32, 44, w, st, &rt));    // block 1 qkv
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 5, w.cls[1], 4, w.cls[0], 6, w.RTc, 32, 32, 46, w, st, &rt));     //         proj (cls row)
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 6, w.cls[3], 16, w.cls[2], 6, w.RTc, 128, 32, 50, w, st, &rt));   //         fc1
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 7, w.cls[5], 4, w.cls[4], 18, w.RTc, 32, 128, 52, w, st, &rt));   //         fc2
  VC_LAUNCH(KC_WGRAD, st, linear_wgrad(t, 8, w.cls[7], (K + 15) / 16 * 2, w.cls[6], 6, w.RTc, K, 32, 56, w, st, &rt));   // head
  cudaStream_t wst = st;        // stream of the weight-gradient GEMMs
  st = main_st;
  // ---- convolutions, last to first ----
  const int order[7] = {6, 2, 1, 0, 5, 4, 3};
  const long long sl = w.RT * 16;
  cudaStream_t lidar_st = ss ? ss->s[1] : st;      // layers 5, 4, 3 beside layers 2, 1, 0 once the fusion layer's data gradient exists
  for (int oi = 0; oi < 7; ++oi) {
    const int i = order[oi];
    const ConvPlan& c = pl[i];
    if (oi == 1 && ss) VC_TRY(stream_dep(ss, main_st, lidar_st));
    st = (i >= 3 && i < 6) ? lidar_st : main_st;
    uint8_t* dz = i == 6 ? w.dzf : i == 2 ? w.df : i == 1 ? w.dzh2 : i == 0 ? w.dzh1 : i == 5 ? w.df + 4 * sl : i == 4 ? w.dzl2 : w.dzl1;
    const uint8_t* x = (i == 0) ? w.a0 : (i == 3) ? w.l0 : (i == 6) ? w.f : w.z[i - 1];
    VC_LAUNCH(KC_BN, st, vc::bn_backward_fused_launch(dz, w.y[i], dz, c.n_out / 8, c.cout, n, P, w.bn_scale[i], w.bn_shift[i],
                                                      w.bn_mean[i], w.bn_rstd[i], 1, w.sums + 256 * (7 + i), G + t->off[4 * i + 2],
                                                      G + t->off[4 * i + 3], G + t->off[4 * i + 1], st));
    if (ss) VC_TRY(stream_dep(ss, st, wst));       // dz of this layer is final
    VC_LAUNCH(KC_WGRAD, wst, conv_wgrad(t, c, i, x, dz, w, n, wst, &rt));
    if (c.has_dgrad) {
      uint8_t* dprev = i == 6 ? w.df : i == 2 ? w.dzh2 : i == 1 ? w.dzh1 : i == 5 ? w.dzl2 : w.dzl1;
      VC_LAUNCH(KC_DGRAD, st, vc::conv_sps_launch(dz, c.d_S_in, w.wd[i], w.ones, w.zeros, dprev, 0, c.d_n_out, c.d_nsplit, n, P,
                                                  c.taps, 0, 0, 0, st));
    }
  }
  st = main_st;
  VC_LAUNCH(KC_WGRAD, wst, vc::wgrad_reduce_batched_launch(&rt, wst));
  if (ss) {                                        // join: every gradient is in place when the caller's stream continues
    VC_TRY(stream_dep(ss, lidar_st, st));
    VC_TRY(stream_dep(ss, wst, st));
  }
  return VC_OK;
}

}  // namespace

extern "C" {

int vc_bn_forward(const void* y, void* z, int32_t S, int32_t C, int32_t n_patches, int32_t P, const float* gamma,
                  const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                  int64_t* num_batches_tracked, double* sums, float* scale, float* shift, float* mean, float* rstd,
                  int32_t relu, void* stream) {
  VC_TRY(vc::bn_forward_launch(y, z, S, C, n_patches, P, gamma, beta, eps, momentum, running_mean, running_var,
                               (long long*)num_batches_tracked, sums, scale, shift, mean, rstd, relu, (cudaStream_t)stream));
  return VC_OK;
}

int vc_bn_backward(const void* dz, const void* y, void* dy, int32_t S, int32_t C, int32_t n_patches, int32_t P,
                   const float* scale, const float* shift, const float* mean, const float* rstd, int32_t relu,
                   double* sums, float* dgamma, float* dbeta, float* dbias, void* stream) {
  VC_TRY(vc::bn_backward_launch(dz, y, dy, S, C, n_patches, P, scale, shift, mean, rstd, relu, sums, dgamma, dbeta, dbias, 0,
                                (cudaStream_t)stream));
  return VC_OK;
}

int vc_pack_conv_weight(const float* w, int32_t cout, int32_t cin, int32_t taps, int32_t transpose, int32_t S_in,
                        int32_t n_out, int32_t nsplit, void* dst, void* stream) {
  VC_TRY(vc::pack_conv_w_launch(w, cout, cin, taps, transpose, S_in, n_out, nsplit, dst, (cudaStream_t)stream));
  return VC_OK;
}

int vc_pack_segments(const float* flat, void* blob, const int64_t* segs, int32_t nsegs, void* stream) {
  VC_TRY(vc::pack_segments_launch(flat, blob, (const long long*)segs, nsegs, (cudaStream_t)stream));
  return VC_OK;
}

int vc_ce_loss(const float* logits, const int64_t* labels, const float* weight, int32_t n, int32_t K, float grad_scale,
               float* loss_out, float* dlogits, double* scratch, void* stream) {
  if (n == 0) return VC_OK;
  VC_LAUNCH(KC_MISC, (cudaStream_t)stream, vc::ce_loss_launch(logits, (const long long*)labels, weight, n, K, grad_scale, loss_out,
                                                              dlogits, scratch, (cudaStream_t)stream));
  return VC_OK;
}

int vc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int32_t step, float grad_scale, void* stream) {
  VC_LAUNCH(KC_MISC, (cudaStream_t)stream,
            vc::adam_launch(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream));
  return VC_OK;
}

int vc_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* hyper, int32_t* step, float grad_scale,
                     void* stream) {
  VC_LAUNCH(KC_MISC, (cudaStream_t)stream, vc::adam_dev_launch(p, g, m, v, n, hyper, step, grad_scale, (cudaStream_t)stream));
  return VC_OK;
}

int64_t vc_train_workspace_bytes(const vc_train* t, int32_t n) {
  ConvPlan pl[7];
  if (n <= 0 || check_train(t, pl) != VC_OK) return -1;
  return carve_train(nullptr, t, pl, n).bytes + 256;
}

int vc_train_workspace_init(const vc_train* t, int32_t n, void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ConvPlan pl[7];
  VC_TRY(check_train(t, pl));
  if (!workspace || workspace_bytes < vc_train_workspace_bytes(t, n)) return fail(VC_ERR_ARG, "workspace too small");
  const TrainWs w = carve_train(workspace, t, pl, n);
  if (cudaMemsetAsync(workspace, 0, (size_t)w.bytes, st) != cudaSuccess) return fail(VC_ERR_CUDA, "memset");
  fill_f32_kernel<<<1, 256, 0, st>>>(w.ones, 256, 1.f);
  for (int i = 0; i < 10; ++i)
    if (kTokOnes[i] >= 0)
      fill_ones_slice_kernel<<<148, 256, 0, st>>>((__nv_bfloat16*)w.tok[i] + (long long)kTokOnes[i] * w.RTt * 8, w.RTt);
  for (int i = 0; i < 8; ++i)
    if (kClsOnes[i] >= 0)
      fill_ones_slice_kernel<<<16, 256, 0, st>>>((__nv_bfloat16*)w.cls[i] + (long long)kClsOnes[i] * w.RTc * 8, w.RTc);
  return cudaGetLastError() == cudaSuccess ? VC_OK : fail(VC_ERR_CUDA, "workspace init");
}

int vc_train_forward(const vc_train* t, const float* hsi, const int64_t hs[4], const float* lidar, const int64_t ls[4],
                     int32_t n, void* workspace, int64_t workspace_bytes, float* logits, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ConvPlan pl[7];
  VC_TRY(check_train(t, pl));
  if (n <= 0 || !hsi || !lidar || !hs || !ls || !workspace || !logits) return fail(VC_ERR_ARG, "vc_train_forward: bad arguments");
  if (workspace_bytes < vc_train_workspace_bytes(t, n)) return fail(VC_ERR_ARG, "workspace too small");
  const TrainWs w = carve_train(workspace, t, pl, n);
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(hsi, hs[0], hs[1], hs[2], hs[3], nullptr, nullptr, n, t->C1, t->P, w.a0, pl[0].S_in, st));
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(lidar, ls[0], ls[1], ls[2], ls[3], nullptr, nullptr, n, t->C2, t->P, w.l0, pl[3].S_in, st));
  return train_forward_core(t, pl, w, n, logits, st);
}

int vc_train_forward_gather(const vc_train* t, const float* img1, const float* img2, const void* gt, int32_t gt_elem_bytes,
                            int32_t H, int32_t W, const int32_t* xy, const uint8_t* ops, int32_t n, void* workspace,
                            int64_t workspace_bytes, float* logits, int64_t* labels, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ConvPlan pl[7];
  VC_TRY(check_train(t, pl));
  if (n <= 0 || !img1 || !img2 || !xy || !workspace || !logits || H < t->P || W < t->P)
    return fail(VC_ERR_ARG, "vc_train_forward_gather: bad arguments");
  if (workspace_bytes < vc_train_workspace_bytes(t, n)) return fail(VC_ERR_ARG, "workspace too small");
  const TrainWs w = carve_train(workspace, t, pl, n);
  VC_LAUNCH(KC_INDEX, st, vc::center_offsets_launch(xy, n, H, W, t->C1, t->C2, t->P, w.off1, w.off2, st));
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(img1, 0, 1, (long long)W * t->C1, t->C1, w.off1, ops, n, t->C1, t->P, w.a0, pl[0].S_in, st));
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(img2, 0, 1, (long long)W * t->C2, t->C2, w.off2, ops, n, t->C2, t->P, w.l0, pl[3].S_in, st));
  if (gt && labels)
    VC_LAUNCH(KC_INDEX, st, vc::gather_labels_launch(gt, gt_elem_bytes, H, W, xy, ops, n, t->P, 1, (long long*)labels, st));
  return train_forward_core(t, pl, w, n, logits, st);
}

int vc_train_backward(const vc_train* t, const float* dlogits, int32_t n, void* workspace, int64_t workspace_bytes,
                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ConvPlan pl[7];
  VC_TRY(check_train(t, pl));
  if (n <= 0 || !dlogits || !workspace) return fail(VC_ERR_ARG, "vc_train_backward: bad arguments");
  if (workspace_bytes < vc_train_workspace_bytes(t, n)) return fail(VC_ERR_ARG, "workspace too small");
  const TrainWs w = carve_train(workspace, t, pl, n);
  return train_backward_core(t, pl, w, n, dlogits, st);
}

}  // extern "C"
