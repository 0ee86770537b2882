// Internal launch prototypes shared by the translation units of libvitcnn.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vc {

// conv_tc.cu -- impl 0: tcgen05 tensor-core kernel, impl 1: SIMT twin (bring-up / tests)
int conv_sps_launch(const void* in, int S_in, const void* w, const float* scale, const float* bias, void* out,
                    int out_slice_off, int n_out, int nsplit, int n_patches, int P, int ntaps, int relu, int impl,
                    int debug_flags, cudaStream_t stream);

// the same conv reading up to 9 input planes (same geometry): tap t takes its rows from plane p iff bit t of
// tapmasks[p]; taps in no mask are dropped.  tcgen05 single-CTA kernel only (impl 0)
int conv_sps_planes_launch(const void* const* planes, const unsigned int* tapmasks, int nplanes, int S_in, const void* w,
                           const float* scale, const float* bias, void* out, int out_slice_off, int n_out, int nsplit,
                           int n_patches, int P, int ntaps, int relu, int impl, int debug_flags, cudaStream_t stream);

// conv_var.cu -- shared stem, conv L >= 2: all (2L+1)^2 border-class variants from the (2L-1)^2 variant planes of
// conv L-1 in ONE launch (work unit = tile x output row class; input slabs staged once per unit and K step).
// ry / rx: [2L+1][3] input class of tap row / column d - 1 for each output class (-1: the tap leaves the window).
// VC_ERR_UNSUPPORTED: not the B = 31 block geometry / does not fit -- fall back to conv_sps_planes_launch.
int conv_var_launch(const void* in, int S_in, const void* w, const float* scale, const float* bias, void* out, int n_out, int L,
                    const signed char* ry, const signed char* rx, int n_blocks, int B, int relu, cudaStream_t stream);

// pack.cu
int pack_sps_launch(const float* src, long long sb, long long sc, long long si, long long sj, const long long* patch_off,
                    const unsigned char* ops, int n_patches, int C, int P, void* sps, int S, cudaStream_t stream);
int pack_scene_launch(const float* img, int W, int C, const int* xs, const int* ys, int nx, int ny, int first, int count,
                      int P, void* sps, int S, cudaStream_t stream);
int zero_halo_launch(void* sps, int S, int n_patches, int P, cudaStream_t stream);
int gather_f32_launch(const float* img, int H, int W, int C, const int* xy, const unsigned char* ops, int n, int P,
                      int center_mode, float* out, cudaStream_t stream);
// gather_tma.cu -- the same gather through a 3-D tensor map (cp.async.bulk.tensor) + bulk stores; VC_ERR_UNSUPPORTED when the
// raster cannot be described by a tensor map (C % 4 != 0, misaligned) -> fall back to gather_f32_launch
int gather_tma_launch(const float* img, int H, int W, int C, const int* xy, const unsigned char* ops, int n, int P, int center_mode,
                      float* out, cudaStream_t stream);
int gather_labels_launch(const void* gt, int gt_elem_bytes, int H, int W, const int* xy, const unsigned char* ops, int n, int P,
                         int center_mode, long long* labels, cudaStream_t stream);
int scene_index_launch(const int* xs, const int* ys, int nx, int ny, int first, int count, int W, int C1, int C2, int P,
                       int K, long long* off1, long long* off2, long long* out_idx, int* xy, cudaStream_t stream);

// shared stem of dense sliding windows (B x B scene blocks, sharing depth D): raster offsets of the blocks, and
// the per-window gather of the (2D+1)^2 border-class variants [v][S][sps_rows(blocks, B)][8] into slices 0..S-1
// of the chunk's stem output [.][sps_rows(count, P)][8]
int block_offsets_launch(int H, int W, int C, int B, int D, long long* off, cudaStream_t stream);
int border_gather_launch(const void* variants, int S, int B, int D, int H, int W, const int* xs, const int* ys, int ny, int first,
                         int count, int P, void* out, cudaStream_t stream);

int center_offsets_launch(const int* xy, int n, int H, int W, int C1, int C2, int P, long long* off1, long long* off2,
                          cudaStream_t stream);

// lidar_stem.cu -- fused eval-mode LiDAR stem (C2 <= 8)
size_t lidar_blob_bytes();
int lidar_stem_launch(const void* in_sps, const void* blob, void* out_sps, int out_slice_off, int n_patches, int P,
                      cudaStream_t stream);

// metrics.cu
int confusion_launch(const void* pred, int peb, const void* target, int teb, long long n, int K, unsigned long long ignored_mask,
                     long long* cm, cudaStream_t stream);

int minmax_normalise_launch(float* img, long long npix, int C, int per_band, float* scratch, cudaStream_t stream);

// transformer.cu
size_t tparams_bytes(int P, int K);
int transformer_fwd_launch(const void* f_sps, const void* tparams, int n_patches, int P, int K, float* logits,
                           const long long* out_index, unsigned char* argmax_map, int prefused, unsigned int drop_thr,
                           const unsigned int* drop_seed, cudaStream_t stream);

// tokens_tc.cu -- the same token stage on tcgen05 (eval mode, P*P + 1 <= 128 tokens); scratch holds
// tokens_tc_scratch_bytes(n) bytes (one 704-byte cls record per patch)
size_t tokens_tc_scratch_bytes(int n_patches);
bool tokens_tc_supported(int P, int K);
// Stem outputs read straight from the variant planes of the shared stem (dense scenes, sharing depth D on B x B
// scene blocks): h / l = [(2D+1)^2 variants][4 slices][RTb][8] conv-3 planes of the HSI / LiDAR stem (null: that
// half comes from slices 0-3 / 4-7 of f_sps); patch b of the launch is window first + b of the (xs, ys) grid.
// xy: int32 [n][2] top-left corner (row, column) of every window of the launch (scene_index_launch); rowterm [H] / colterm [W]:
// block-row offset of raster row y / column x inside the block that holds it (scene_tables_launch), so that the plane row of
// a scene pixel is sps_halo(B) + rowterm[y] + colterm[x] -- two table reads instead of divisions per token and patch.
struct TcPlanes {
  const __nv_bfloat16* h;
  const __nv_bfloat16* l;
  const int* xy;
  const int* rowterm;
  const int* colterm;
  long long RTb;
  int B, D;
};
int scene_tables_launch(int H, int W, int B, int D, int* rowterm, int* colterm, cudaStream_t stream);
int tokens_tc_launch(const void* f_sps, const void* tparams, int n_patches, int P, int K, float* logits,
                     const long long* out_index, unsigned char* argmax_map, void* scratch, const TcPlanes* planes,
                     cudaStream_t stream);

// wgrad_tc.cu -- weight gradients (rows are the reduction axis; both operands MN-major)
size_t wgrad_workspace_bytes(int SB, int ntaps);
// `defer` (nullable): instead of launching the split-K reduction, append it to the table; the caller then runs
// wgrad_reduce_batched_launch once for the whole backward pass (every deferred job needs its own workspace)
constexpr int kMaxWgradJobs = 20;
struct WgradReduceJob {
  const float* part;
  float* out;
  float* out_bias;
  long long sm, sn, st;
  int nparts, ntaps, N, M, Nr, bias_col, accumulate;
};
struct WgradReduceTable {
  WgradReduceJob job[kMaxWgradJobs];
  int n;
};
int wgrad_sps_launch(const void* A, int SA, const void* B, int SB, int n_patches, int P, int ntaps, int shift_on_a,
                     void* workspace, float* out, int M, int Nr, long long sm, long long sn, long long st,
                     int bias_col, float* out_bias, int accumulate, cudaStream_t stream, WgradReduceTable* defer = nullptr);
int wgrad_reduce_batched_launch(const WgradReduceTable* t, cudaStream_t stream);


int wgrad_reduce_launch(const float* part, int nparts, int ntaps, int N, int M, int Nr, float* out, long long sm, long long sn,
                        long long st, int bias_col, float* out_bias, int accumulate, cudaStream_t stream);
// wgrad_small.cu -- thin 3x3 convs (<= 32 x <= 16 channels) on mma.sync
size_t wgrad_small_workspace_bytes();
int wgrad_small_launch(const void* dy, int SA, const void* x, int SB, int n_patches, int P, void* workspace, float* out,
                       int cout, int cin, cudaStream_t stream);

// train.cu -- BatchNorm (training mode) forward / backward over SPS, loss, optimiser, weight packing
int bn_forward_launch(const void* y, void* z, int S, int C, int n_patches, int P, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, long long* nbt, double* sums,
                      float* scale, float* shift, float* mean, float* rstd, int relu, cudaStream_t st);
int bn_backward_launch(const void* dz, const void* y, void* dy, int S, int C, int n_patches, int P, const float* scale,
                       const float* shift, const float* mean, const float* rstd, int relu, double* sums, float* dgamma,
                       float* dbeta, float* dbias, int accumulate, cudaStream_t st);
// The same two passes in two launches each (training step: launch-count bound at small per-GPU batches): the
// statistics are finalised / the parameter gradients written by the apply kernel itself; `sums` (2 * S * 8 doubles,
// this call's own) must be zero on entry and is NOT cleared (the step's prep launch clears every layer's at once).
int bn_forward_fused_launch(const void* y, void* z, int S, int C, int n_patches, int P, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, long long* nbt, double* sums,
                            float* scale, float* shift, float* mean, float* rstd, int relu, cudaStream_t st);
int bn_backward_fused_launch(const void* dz, const void* y, void* dy, int S, int C, int n_patches, int P, const float* scale,
                             const float* shift, const float* mean, const float* rstd, int relu, double* sums, float* dgamma,
                             float* dbeta, float* dbias, cudaStream_t st);
// One launch for the small per-step chores: type 0 pack_conv_w (src fp32 weight -> bf16 operand), 1 copy fp32,
// 2 zero fp32, 3 zero fp64
constexpr int kMaxPrepJobs = 40;
struct PrepJob {
  const float* src;
  void* dst;
  long long n;
  int type, cout, cin, taps, transpose, S_in, n_out, nsplit;
};
struct PrepTable {
  PrepJob job[kMaxPrepJobs];
  int n;
};
int train_prep_launch(const PrepTable* t, cudaStream_t st);
int ce_loss_launch(const float* logits, const long long* labels, const float* weight, int n, int K, float grad_scale,
                   float* loss_out, float* dlogits, double* acc, cudaStream_t st);
int adam_launch(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                float wd, int step, float grad_scale, cudaStream_t st);
int adam_dev_launch(float* p, const float* g, float* m, float* v, long long n, float* hyper, int* step, float grad_scale,
                    cudaStream_t st);
int pack_conv_w_launch(const float* w, int cout, int cin, int taps, int transpose, int S_in, int n_out, int nsplit,
                       void* dst, cudaStream_t st);
int pack_segments_launch(const float* flat, void* blob, const long long* segs, int nsegs, cudaStream_t st);

// tokens_bwd.cu
int transformer_bwd_launch(const void* zf, const void* tparams, const float* dlogits, void* dzf, void* const* tok_dumps,
                           void* const* cls_dumps, float* const* small, int n_patches, int P, int K, unsigned int drop_thr,
                           const unsigned int* drop_seed, cudaStream_t stream);

}  // namespace vc
