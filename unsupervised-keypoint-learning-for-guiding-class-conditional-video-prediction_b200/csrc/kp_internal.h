// Internal launcher prototypes shared between the kernel translation units and the C-ABI file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/kp_b200.h"

namespace kp {

// k1_keypoints.cu
int k1_softargmax_render_fwd(const float* logits, int B, int H, int W, int K, float* mu, float* prob_x, float* prob_y,
                             float* maps, int hm, int wm, float inv_std, cudaStream_t st);
int k1_softargmax_render_bwd(const float* d_maps, const float* d_mu_extra, const float* mu, const float* prob_x,
                             const float* prob_y, int B, int H, int W, int K, int hm, int wm, float inv_std,
                             float* d_logits, float* d_mu_scratch, cudaStream_t st);
int k1_render_fwd(const float* mu, int B, int K, int hm, int wm, float inv_std, float* maps, cudaStream_t st);
int k1_render_bwd(const float* d_maps, const float* d_mu_extra, const float* mu, int B, int K, int hm, int wm,
                  float inv_std, float* d_mu, cudaStream_t st);
int k1_render_colorize(const float* mu, const float* colors, int B, int K, int hm, int wm, float inv_std, float* out,
                       cudaStream_t st);
int k1_colorize(const float* maps, const float* colors, long long P, int K, float* out, cudaStream_t st);

// conv_tc.cu
int tapconv_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                   float* ssum, float* ssq, cudaStream_t st);

// conv_wgrad.cu
int wgrad_launch(const kp_wgrad_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st);

}  // namespace kp
