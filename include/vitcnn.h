/* libvitcnn.so -- C ABI of the B200-native ViT-CNN hot path.
 *
 * Drop-in boundary: the reference toolkit is pure Python on stock
 * PyTorch ops; its hot path is entered through
 *   datasets.py:550-593      MultiModalX.__getitem__  (per-pixel patch slice, HWC->CHW)
 *   utils.py:357-415         sliding_window / count_sliding_window
 *   model_utils.py:1067-1132 test()  (batch assembly, forward, logits scatter)
 *   model_utils.py:921,1118  net(data, data2)  (model forward)
 * Each entry point below names the reference interface it replaces.  The Python binding a
 * maintainer adds is a ctypes stub; torch only supplies device buffers and
 * the stream.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; the caller owns all buffers
 * (inputs, outputs, workspaces); nothing is allocated, freed or retained by the library; all
 * work is enqueued on `stream` (a cudaStream_t passed as void*); functions never throw and
 * return 0 (VC_OK) or a negative code, vc_last_error() giving a thread-local description.
 */
#ifndef VITCNN_H_
#define VITCNN_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC_ABI_VERSION 4

/* Parameters of one model instance in kernel-ready form (built by vc-side packing, see
 * vitcnn_b200/model.py::pack_for_inference).  Conv weights: bf16 [nsplit][taps][S_in][N/nsplit][8]
 * (tap = ky*3+kx, S_in = padded Cin / 8); scale/bias: fp32 [N] = BatchNorm folded with the
 * conv bias.  tparams: blob described by vc_tparams_layout(). */
typedef struct vc_model {
  int32_t C1, C2, P, K;        /* HSI bands, LiDAR bands, patch size, classes            */
  int32_t S1, S2;              /* input slices: ceil(C/16)*2                              */
  int32_t nsplit_h[3];         /* CTA split of the output channels, HSI stem convs        */
  int32_t nsplit_l[3];         /* ... LiDAR stem convs                                    */
  const void* w_h[3];          /* HSI stem: C1->128->64->32                               */
  const float* scale_h[3];
  const float* bias_h[3];
  const void* w_l[3];          /* LiDAR stem: C2->8->16->32 (8 padded to 16)              */
  const float* scale_l[3];
  const float* bias_l[3];
  const void* tparams;         /* fusion 1x1 + cls/pos + 2 blocks + norm + head           */
  const void* lidar_blob;      /* nullable: operands of the fused LiDAR-stem kernel (C2 <= 8),
                                  vc_lidar_blob_bytes() bytes: bf16 rows of 24 elements --
                                  [5 tap pairs][8][k: tap 2p ch 0-7 | tap 2p+1 ch 0-7],
                                  [5][16][same], [9 taps][32][16 ch] -- then fp32 scale/bias
                                  of the three layers (8+8, 16+16, 32+32); NULL = three
                                  tensor-core conv launches through w_l / scale_l / bias_l   */
  const void* w_h1_border;     /* nullable: 9 packed copies of w_h[0], copy cy*3+cx with the taps that leave a
                                  window dropped for a pixel in row class cy / column class cx of the window
                                  (0: first row / column -> ky or kx == 0 zeroed, 1: interior, 2: last ->
                                  ky or kx == 2 zeroed).  Enables the shared first conv of vc_scene_infer  */
} vc_model;

int vc_abi_version(void);
int64_t vc_lidar_blob_bytes(void);
const char* vc_last_error(void);

/* ---- instrumentation (bench.py): kernels launched by this library so far in the process, and
 * per-kernel-class device time between vc_profile_begin/vc_profile_end on the calling thread
 * (CUDA events around each launch).  Classes: 0 window index, 1 pack (patch gather to SPS),
 * 2/3/4 HSI stem conv 1/2/3, 5 LiDAR stem convs, 6 token stage, 7 halo zeroing, 8 BatchNorm passes,
 * 9 weight-gradient GEMMs, 10 data-gradient convs, 11 token-stage backward, 12 packing / loss / Adam. */
#define VC_KERNEL_CLASSES 13
int64_t vc_launch_count(void);
int vc_profile_begin(void);
int vc_profile_end(double* ms_per_class, int64_t* launches_per_class, int32_t n_classes);

/* ---- geometry of the SPS activation layout (rows per chunk of n patches) ------------------- */
int64_t vc_sps_rows(int32_t n_patches, int32_t P);
/* bytes of workspace vc_forward_* needs for a chunk of n patches */
int64_t vc_workspace_bytes(int32_t n_patches, int32_t P, int32_t C1, int32_t C2);
/* byte offsets of the token-stage parameter blob: fills out[0..n) and returns n (or the
 * needed n if out is NULL).  Order: total, wfus, fus_scale, fus_bias, cls, lnf_g, lnf_b,
 * whead, bhead, pos, then per layer (2x): wqkv, wproj, wfc1, wfc2, ln1_g, ln1_b, bqkv,
 * bproj, ln2_g, ln2_b, bfc1, bfc2. */
int32_t vc_tparams_layout(int32_t P, int32_t K, int64_t* out, int32_t n);

/* ---- exact patch extraction ----------------------------------------------------------------
 * Replaces MultiModalX.__getitem__ + default_collate (datasets.py:550-593, main.py:434-447)
 * when center_mode=1 (xy = patch centres) and test()'s batch assembly (model_utils.py:1103-
 * 1112, windows from utils.sliding_window) when center_mode=0 (xy = top-left corners).
 * img1/img2: f32 [H][W][C] rasters; xy: int32 [n][2]; hsi: f32 [n][C1][P][P]; lidar: f32
 * [n][C2][P][P]; labels (nullable with gt): int64 [n] = gt[centre]; gt element size 1/4/8 B.
 * ops (nullable): uint8 [n], the spatial augmentation of each sample as the host drew it with the
 * reference's RNG calls (datasets.py:510-526, 559-564): 0 identity, 1 fliplr, 2 flipud, 3 both,
 * 4/5/6 np.rot90 k = 1/2/3 -- applied as an index remap to data, LiDAR and label alike.
 * Bit-exact copies. */
int vc_gather_patches_f32(const float* img1, const float* img2, const void* gt, int32_t gt_elem_bytes, int32_t H,
                          int32_t W, int32_t C1, int32_t C2, const int32_t* xy, const uint8_t* ops, int32_t n, int32_t P,
                          int32_t center_mode, float* hsi, float* lidar, int64_t* labels, void* stream);

/* Window enumeration of utils.sliding_window (utils.py:374-397) for windows [first, first+count)
 * given the per-axis start lists xs[nx], ys[ny] (int32, device): xy int32 [count][2] (nullable),
 * off1/off2 = element offsets of the window corner in img1/img2 (nullable), out_idx = pixel
 * index (x+P/2)*W + (y+P/2) that test() scatters to (model_utils.py:1127-1129). */
int vc_scene_index(const int32_t* xs, const int32_t* ys, int32_t nx, int32_t ny, int32_t first, int32_t count,
                   int32_t W, int32_t C1, int32_t C2, int32_t P, int64_t* off1, int64_t* off2, int64_t* out_idx,
                   int32_t* xy, void* stream);

/* ---- prediction post-processing: confusion matrix of metrics() (utils.py:585-663) ------------
 * cm int64 [K][K] (row = target, column = prediction) += counts over the n elements whose target is
 * not in the ignored set (bit l of ignored_mask = label l ignored, utils.py:595-601) and whose
 * labels lie in range(K) (sklearn's labels=range(n_classes)).  Element sizes 1 / 4 / 8 bytes. */
int vc_confusion_matrix(const void* prediction, int32_t pred_elem_bytes, const void* target, int32_t target_elem_bytes, int64_t n,
                        int32_t n_classes, uint64_t ignored_mask, int64_t* cm, void* stream);

/* ---- raster ingest: min-max normalisation to [0,1] in place (datasets.py:124-133) ----------------
 * img f32 [n_pixels][C]; per_band = 1: each band by its own min / max (the HSI cube), 0: one min / max
 * for the whole array (the LiDAR raster); scratch: 2*C floats.  Bit-exact with numpy's float32 form. */
int vc_minmax_normalise(float* img, int64_t n_pixels, int32_t C, int32_t per_band, float* scratch, void* stream);

/* ---- building blocks (exposed for tests and profiling) --------------------------------------- */
/* fp32 patches (any strides, in elements) or raster windows (patch_off != NULL: per-patch
 * element offset, sb ignored) -> bf16 SPS buffer [S][rows][8]. */
int vc_pack_sps(const float* src, int64_t sb, int64_t sc, int64_t si, int64_t sj, const int64_t* patch_off,
                int32_t n_patches, int32_t C, int32_t P, void* sps, int32_t S, void* stream);
/* 3x3 pad-1 (taps=9) or 1x1 (taps=1) conv + per-channel affine (+ReLU) over SPS buffers.
 * impl 0 = tcgen05 tensor-core kernel, impl 1 = SIMT twin for cross-checks. */
int vc_conv_sps(const void* in_sps, int32_t S_in, const void* w_packed, const float* scale, const float* bias,
                void* out_sps, int32_t out_slice_off, int32_t n_out, int32_t nsplit, int32_t n_patches, int32_t P,
                int32_t taps, int32_t relu, int32_t impl, int32_t debug_flags, void* stream);
/* token stage on the fused 64-channel SPS feature buffer -> logits f32 [n][K] (row b, or row
 * out_index[b] when given); argmax_map (nullable, uint8) receives argmax at out_index[b]. */
int vc_tokens_forward(const void* f_sps, const void* tparams, int32_t n_patches, int32_t P, int32_t K, float* logits,
                      const int64_t* out_index, uint8_t* argmax_map, void* stream);
/* The same token stage on the tcgen05 tensor cores (eval mode; the token set of a patch must fit one
 * M = 128 tile: P*P + 1 <= 128, else VC_ERR_UNSUPPORTED).  scratch: caller-owned device buffer of
 * vc_tokens_tc_scratch_bytes(n_patches) bytes (a 4 KB staging area for the per-launch constants, then one
 * 704-byte cls record per patch).  vc_forward_patches / vc_scene_infer pick this kernel themselves when it
 * applies (P = 5 .. 11; scratch = a dead part of their workspace).  Kernels: tokens_tm_kernel (probabilities and
 * MLP hidden units handed to the tensor core through TMEM, per-channel vectors in the constant bank; launches on
 * different streams of one device are ordered by an event because that bank is per device) with
 * tokens_tc_kernel behind it for weights whose attention logits need the softmax row maximum. */
int64_t vc_tokens_tc_scratch_bytes(int32_t n_patches);
int vc_tokens_forward_tc(const void* f_sps, const void* tparams, int32_t n_patches, int32_t P, int32_t K, float* logits,
                         const int64_t* out_index, uint8_t* argmax_map, void* scratch, int64_t scratch_bytes, void* stream);

/* ---- training building blocks ------------------------------------------------------------------
 * Weight gradient of a 3x3 pad-1 conv (taps=9) or of a linear / 1x1 conv (taps=1) over SPS
 * buffers, i.e. what autograd computes for Conv2d / Linear weights when loss.backward() runs
 * (model_utils.py:936; the layers are those of SURVEY.md App. A):
 *   dW[tap][m][n] = sum_rows A[row + sa(tap)][m] * B[row + sb(tap)][n],
 * the tap's row shift applied to A (shift_on_a) or to B.  A: [SA][rows][8] (M <= SA*8 <= 128),
 * B: [SB][rows][8] (SB even, N <= SB*8 <= 256).  Result scattered as out[m*sm + n*sn + tap*st]
 * (= or +=); column bias_col (>= 0) of B goes to out_bias[m] (a constant-one channel in B makes
 * it the bias gradient).  Deterministic (fixed-order split-K reduction through `workspace`). */
int64_t vc_wgrad_workspace_bytes(int32_t SB, int32_t taps);
int vc_wgrad_sps(const void* a_sps, int32_t SA, const void* b_sps, int32_t SB, int32_t n_patches, int32_t P, int32_t taps,
                 int32_t shift_on_a, void* workspace, int64_t workspace_bytes, float* out, int32_t M, int32_t N,
                 int64_t sm, int64_t sn, int64_t st, int32_t bias_col, float* out_bias, int32_t accumulate, void* stream);

/* BatchNorm2d (training mode) + ReLU over an SPS buffer, forward and backward (the conv_bn_relu
 * idiom of the stems as autograd differentiates it).  y: raw conv output bf16 [S][rows][8];
 * sums: 2*S*8 doubles of scratch, zero on entry, left zero on exit; scale/shift/mean/rstd: fp32
 * [S*8] saved for the backward.  Running statistics follow nn.BatchNorm2d (momentum, unbiased
 * variance); pointers may be NULL. */
int vc_bn_forward(const void* y, void* z, int32_t S, int32_t C, int32_t n_patches, int32_t P, const float* gamma,
                  const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                  int64_t* num_batches_tracked, double* sums, float* scale, float* shift, float* mean, float* rstd,
                  int32_t relu, void* stream);
int vc_bn_backward(const void* dz, const void* y, void* dy, int32_t S, int32_t C, int32_t n_patches, int32_t P,
                   const float* scale, const float* shift, const float* mean, const float* rstd, int32_t relu,
                   double* sums, float* dgamma, float* dbeta, float* dbias, void* stream);
/* torch conv weight fp32 [cout][cin][taps] -> bf16 operand of vc_conv_sps; transpose=1 gives the
 * data-gradient operand (conv from cout back to cin channels with flipped taps). */
int vc_pack_conv_weight(const float* w, int32_t cout, int32_t cin, int32_t taps, int32_t transpose, int32_t S_in,
                        int32_t n_out, int32_t nsplit, void* dst, void* stream);
/* table-driven fp32 -> packed blob copy: segs int64 [n][6] = src element offset, dst byte offset,
 * rows, cols, dst pitch (elements), is_bf16 (device memory). */
int vc_pack_segments(const float* flat, void* blob, const int64_t* segs, int32_t nsegs, void* stream);

/* nn.CrossEntropyLoss(weight=w) (model_utils.py:63-66,216; applied at :929-936): loss_out[0] =
 * weighted mean NLL, loss_out[1] = sum of weights; dlogits (nullable) = grad_scale * dloss/dlogits.
 * labels int64 [n]; entries outside [0,K) are ignored (ignore_index semantics); scratch: 2 doubles. */
int vc_ce_loss(const float* logits, const int64_t* labels, const float* weight, int32_t n, int32_t K, float grad_scale,
               float* loss_out, float* dlogits, double* scratch, void* stream);
/* optim.Adam(lr) step (model_utils.py:214-215) on flat fp32 buffers; g is multiplied by
 * grad_scale first (1/world_size after a summing all-reduce). */
int vc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int32_t step, float grad_scale, void* stream);

/* Same step with the hyper-parameters and the step counter in DEVICE memory, so that a captured
 * CUDA graph of the training step replays correctly: hyper f32 [8] = lr, beta1, beta2, eps,
 * weight_decay, then 3 scratch floats; step int32 (incremented by the call). */
int vc_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* hyper, int32_t* step, float grad_scale,
                     void* stream);

/* ---- training: forward (batch statistics) / backward of the whole model -------------------------
 * Replaces net(data, data2) in train mode and loss.backward() (model_utils.py:921-936).
 * Parameters and gradients live in flat fp32 buffers; `off` gives the element offset of each
 * tensor in the canonical order: conv layer i (0..6 = hsi_stem.0-2, lidar_stem.0-2, fusion):
 * 4i+{0 conv.weight, 1 conv.bias, 2 bn.weight, 3 bn.bias}; 28 cls_token; 29 pos_embed; block l
 * (0,1): 30+12l+{0 norm1.w, 1 norm1.b, 2 qkv.w, 3 qkv.b, 4 proj.w, 5 proj.b, 6 norm2.w, 7 norm2.b,
 * 8 fc1.w, 9 fc1.b, 10 fc2.w, 11 fc2.b}; 54 norm.w; 55 norm.b; 56 head.w; 57 head.b. */
#define VC_NPARAMS 58
typedef struct vc_train {
  int32_t C1, C2, P, K;
  float* params;
  float* grads;
  int64_t off[VC_NPARAMS];
  float* bn_running_mean[7];
  float* bn_running_var[7];
  int64_t* bn_num_batches[7];
  float bn_eps, bn_momentum;
  const int64_t* blob_segments;   /* device table for vc_pack_segments: token-stage blob from params */
  int32_t n_blob_segments;
  float dropout;                  /* p of pos_drop / proj_drop / MLP drops (vision_transformer.py:598-629,
                                     mlp.py:41-47); masks are a stateless hash of (seed, sample, site, element) */
  uint32_t* drop_seed;            /* device: seed word, advanced by every vc_train_forward* call (graph-safe);
                                     required when dropout > 0 */
} vc_train;

int64_t vc_train_workspace_bytes(const vc_train* t, int32_t n);
/* zero the workspace and write its constant parts; call once per (workspace, n) */
int vc_train_workspace_init(const vc_train* t, int32_t n, void* workspace, int64_t workspace_bytes, void* stream);
/* hsi/lidar: f32 [n][C][P][P] with element strides, as in vc_forward_patches */
int vc_train_forward(const vc_train* t, const float* hsi, const int64_t hsi_strides[4], const float* lidar,
                     const int64_t lidar_strides[4], int32_t n, void* workspace, int64_t workspace_bytes, float* logits,
                     void* stream);
/* same, patches gathered on the device from the rasters: xy int32 [n][2] patch centres
 * (MultiModalX.__getitem__, datasets.py:550-556); ops (nullable) uint8 [n] flip / rot90 codes as in
 * vc_gather_patches_f32; labels (nullable, with gt) int64 [n] */
int vc_train_forward_gather(const vc_train* t, const float* img1, const float* img2, const void* gt, int32_t gt_elem_bytes,
                            int32_t H, int32_t W, const int32_t* xy, const uint8_t* ops, int32_t n, void* workspace,
                            int64_t workspace_bytes, float* logits, int64_t* labels, void* stream);
/* gradients of every parameter into t->grads (overwritten) for the batch of the last forward
 * on this workspace */
int vc_train_backward(const vc_train* t, const float* dlogits, int32_t n, void* workspace, int64_t workspace_bytes,
                      void* stream);

/* ---- model forward: replaces net(data, data2) at model_utils.py:921/1118/1144 (eval) ---------
 * hsi/lidar: f32 [n][C][P][P] with arbitrary element strides (sb, sc, si, sj) -- contiguous
 * NCHW from the DataLoader or the NHWC-memory view test() builds.  logits: f32 [n][K]. */
int vc_forward_patches(const vc_model* m, const float* hsi, const int64_t hsi_strides[4], const float* lidar,
                       const int64_t lidar_strides[4], int32_t n, void* workspace, int64_t workspace_bytes,
                       float* logits, void* stream);

/* ---- full-scene inference: replaces the loop of test() (model_utils.py:1086-1129) ------------
 * Windows are enumerated on the device from xs/ys (see vc_scene_index), processed in chunks of
 * `chunk` windows [first_window, first_window+n_windows); logits land in logits_map f32
 * [H][W][K] at the window centre (untouched pixels are left as they are: zero-fill first),
 * argmax_map (nullable) uint8 [H][W].  Row-band sharding = disjoint window ranges per GPU.
 *
 * Shared stem: the output of the first d stem convs at a window pixel depends on the window only through
 * the zero padding at the window border ((2d+1) row classes x (2d+1) column classes).  When m->w_h1_border is
 * set, the raster is large enough, the windows are dense enough for it to pay and the workspace holds
 * vc_scene_workspace_bytes(), the variants are computed once per call on overlapping scene blocks: the HSI stem to
 * the depth vc_scene_shared_depth() reports (all three convs on 31 / 63 / 95-pixel blocks when P >= 7, conv 1 only on
 * 15 x 15 blocks otherwise, P >= 2), the LiDAR stem to depth 3 whenever P >= 7 and H, W >= 31.  Every window's stem output is
 * then taken from the variant planes (same bits as the per-window convs): read in place by the token kernel when it is
 * the tcgen05 one (82 <= P*P + 1 <= 128) and the depth is 3, gathered into the chunk's buffers otherwise.  With only
 * vc_workspace_bytes(chunk, ...) of workspace, or without w_h1_border, the per-window path runs (three tcgen05 convs per
 * stem: the same kernels and accumulation order, so vc_forward_patches and vc_scene_infer agree bit for bit). */
int64_t vc_scene_workspace_bytes(const vc_model* m, int32_t H, int32_t W, int32_t chunk);
/* edge of the scene blocks the shared stem would use for an H x W raster at this sharing depth (15 / 31 / 63 / 95) and
 * how many HSI stem convs vc_scene_infer would share for this call (0 = per-window path); instrumentation, and what a host
 * pipeline needs to cut a scene into sub-bands at whole block rows */
int32_t vc_scene_block(int32_t H, int32_t W, int32_t depth);
int32_t vc_scene_shared_depth(const vc_model* m, int32_t H, int32_t W, int32_t chunk, int64_t n_windows, int64_t workspace_bytes);
int vc_scene_infer(const vc_model* m, const float* img1, const float* img2, int32_t H, int32_t W, const int32_t* xs,
                   const int32_t* ys, int32_t nx, int32_t ny, int64_t first_window, int64_t n_windows, int32_t chunk,
                   void* workspace, int64_t workspace_bytes, float* logits_map, uint8_t* argmax_map, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITCNN_H_ */
