// Internal launch prototypes shared by the translation units of libvitcnn.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vc {

// conv_tc.cu -- impl 0: tcgen05 tensor-core kernel, impl 1: SIMT twin (bring-up / tests)
int conv_sps_launch(const void* in, int S_in, const void* w, const float* scale, const float* bias, void* out,
                    int out_slice_off, int n_out, int nsplit, int n_patches, int P, int ntaps, int relu, int impl,
                    int debug_flags, cudaStream_t stream);

// pack.cu
int pack_sps_launch(const float* src, long long sb, long long sc, long long si, long long sj, const long long* patch_off,
                    int n_patches, int C, int P, void* sps, int S, cudaStream_t stream);
int zero_halo_launch(void* sps, int S, int n_patches, int P, cudaStream_t stream);
int gather_f32_launch(const float* img, int H, int W, int C, const int* xy, int n, int P, int center_mode, float* out,
                      cudaStream_t stream);
int gather_labels_launch(const void* gt, int gt_elem_bytes, int H, int W, const int* xy, int n, int P, int center_mode,
                         long long* labels, cudaStream_t stream);
int scene_index_launch(const int* xs, const int* ys, int nx, int ny, int first, int count, int W, int C1, int C2, int P,
                       int K, long long* off1, long long* off2, long long* out_idx, int* xy, cudaStream_t stream);

// transformer.cu
size_t tparams_bytes(int P, int K);
int transformer_fwd_launch(const void* f_sps, const void* tparams, int n_patches, int P, int K, float* logits,
                           const long long* out_index, unsigned char* argmax_map, cudaStream_t stream);

// wgrad_tc.cu -- weight gradients (rows are the reduction axis; both operands MN-major)
size_t wgrad_workspace_bytes(int SB, int ntaps);
int wgrad_sps_launch(const void* A, int SA, const void* B, int SB, int n_patches, int P, int ntaps, int shift_on_a,
                     void* workspace, float* out, int M, int Nr, long long sm, long long sn, long long st,
                     int bias_col, float* out_bias, int accumulate, cudaStream_t stream);

}  // namespace vc
