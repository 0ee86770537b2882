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
enum { KC_INDEX = 0, KC_PACK, KC_CONV_H1, KC_CONV_H2, KC_CONV_H3, KC_CONV_L, KC_TOKENS, KC_HALO, KC_COUNT };
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
  uint8_t *a0, *l0, *a1, *a2, *f, *l1, *l2;
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
  w.bytes = p - reinterpret_cast<uint8_t*>(base);
  return w;
}

inline int slices_for(int C) { return (C + 15) / 16 * 2; }

// stems + token stage on packed inputs already in w.a0 / w.l0
// lead / trailing halo rows of the intermediates are read by the next conv and written by no
// kernel: zero them once per workspace geometry (they stay zero across chunks of equal size)
int zero_halos(const Workspace& w, int n, int P, cudaStream_t st) {
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.a1, 16, n, P, st));
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.a2, 8, n, P, st));
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.l1, 2, n, P, st));
  VC_LAUNCH(KC_HALO, st, vc::zero_halo_launch(w.l2, 2, n, P, st));
  return VC_OK;
}

int forward_sps(const vc_model* m, const Workspace& w, int n, float* logits, const long long* out_index,
                uint8_t* argmax_map, cudaStream_t st) {
  const int P = m->P;
  VC_LAUNCH(KC_CONV_H1, st, vc::conv_sps_launch(w.a0, m->S1, m->w_h[0], m->scale_h[0], m->bias_h[0], w.a1, 0, 128, m->nsplit_h[0], n, P, 9,
                             1, 0, 0, st));
  VC_LAUNCH(KC_CONV_H2, st, vc::conv_sps_launch(w.a1, 16, m->w_h[1], m->scale_h[1], m->bias_h[1], w.a2, 0, 64, m->nsplit_h[1], n, P, 9, 1,
                             0, 0, st));
  VC_LAUNCH(KC_CONV_H3, st, vc::conv_sps_launch(w.a2, 8, m->w_h[2], m->scale_h[2], m->bias_h[2], w.f, 0, 32, m->nsplit_h[2], n, P, 9, 1, 0,
                             0, st));
  VC_LAUNCH(KC_CONV_L, st, vc::conv_sps_launch(w.l0, m->S2, m->w_l[0], m->scale_l[0], m->bias_l[0], w.l1, 0, 16, m->nsplit_l[0], n, P, 9,
                             1, 0, 0, st));
  VC_LAUNCH(KC_CONV_L, st, vc::conv_sps_launch(w.l1, 2, m->w_l[1], m->scale_l[1], m->bias_l[1], w.l2, 0, 16, m->nsplit_l[1], n, P, 9, 1, 0,
                             0, st));
  VC_LAUNCH(KC_CONV_L, st, vc::conv_sps_launch(w.l2, 2, m->w_l[2], m->scale_l[2], m->bias_l[2], w.f, 4, 32, m->nsplit_l[2], n, P, 9, 1, 0,
                             0, st));
  VC_LAUNCH(KC_TOKENS, st, vc::transformer_fwd_launch(w.f, m->tparams, n, P, m->K, logits, out_index, argmax_map, st));
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
                          int32_t W, int32_t C1, int32_t C2, const int32_t* xy, int32_t n, int32_t P,
                          int32_t center_mode, float* hsi, float* lidar, int64_t* labels, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return VC_OK;
  if (!xy || n < 0) return fail(VC_ERR_ARG, "vc_gather_patches_f32: bad arguments");
  if (img1 && hsi) VC_TRY(vc::gather_f32_launch(img1, H, W, C1, xy, n, P, center_mode, hsi, st));
  if (img2 && lidar) VC_TRY(vc::gather_f32_launch(img2, H, W, C2, xy, n, P, center_mode, lidar, st));
  if (gt && labels)
    VC_TRY(vc::gather_labels_launch(gt, gt_elem_bytes, H, W, xy, n, P, center_mode, (long long*)labels, st));
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

int vc_pack_sps(const float* src, int64_t sb, int64_t sc, int64_t si, int64_t sj, const int64_t* patch_off,
                int32_t n_patches, int32_t C, int32_t P, void* sps, int32_t S, void* stream) {
  VC_TRY(vc::pack_sps_launch(src, sb, sc, si, sj, (const long long*)patch_off, n_patches, C, P, sps, S,
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
                                    (cudaStream_t)stream));
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
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(hsi, hs[0], hs[1], hs[2], hs[3], nullptr, n, m->C1, m->P, w.a0, m->S1, st));
  VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(lidar, ls[0], ls[1], ls[2], ls[3], nullptr, n, m->C2, m->P, w.l0, m->S2, st));
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
  for (int64_t done = 0; done < n_windows; done += chunk) {
    const int n = (int)((n_windows - done) < chunk ? (n_windows - done) : chunk);
    // the SPS geometry depends on n: carve per chunk (only the last chunk differs)
    const Workspace w = carve(workspace, n, m->P, m->S1, m->S2);
    if (done == 0 || n != chunk) VC_TRY(zero_halos(w, n, m->P, st));
    VC_LAUNCH(KC_INDEX, st, vc::scene_index_launch(xs, ys, nx, ny, (int)(first_window + done), n, W, m->C1, m->C2, m->P,
                                                   m->K, w.off1, w.off2, w.oidx, nullptr, st));
    VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(img1, 0, 1, s1, m->C1, w.off1, n, m->C1, m->P, w.a0, m->S1, st));
    VC_LAUNCH(KC_PACK, st, vc::pack_sps_launch(img2, 0, 1, s2, m->C2, w.off2, n, m->C2, m->P, w.l0, m->S2, st));
    VC_TRY(forward_sps(m, w, n, logits_map, w.oidx, argmax_map, st));
  }
  (void)H;
  return VC_OK;
}

}  // extern "C"
