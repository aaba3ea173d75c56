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
// Launch with the PDL attribute when KP_PDL=1 (see kp_common.cuh; the kernel must call pdl_wait() before its first global
// access).  Inside a stream capture this becomes a programmatic dependency edge of the CUDA graph.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

bool tapconv2_eligible(const kp_tapconv_desc* d);
int tapconv2_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                    float* ssum, float* ssq, cudaStream_t st);
bool halo2_eligible(const kp_tapconv_desc* d, const float* ssum);
int halo2_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                 float* ssum, float* ssq, cudaStream_t st);
bool haloconv_eligible(const kp_tapconv_desc* d, const float* ssum);
int haloconv_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                    float* ssum, float* ssq, cudaStream_t st);
int tapconv_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                   float* ssum, float* ssq, cudaStream_t st);

// conv_wgrad.cu
int wgrad_launch(const kp_wgrad_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st);
// conv_wgrad2.cu (halo-tile weight gradient of the stride-1 layers)
bool wgrad2_eligible(const kp_wgrad_desc* d);
int wgrad2_launch(const kp_wgrad_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st);

// head1x1.cu
int head1x1_launch(const void* x, const float* w, const float* bias, long long P, int Cin, int Cout, float* out, cudaStream_t st);

// elementwise.cu
int ew_image_prep(const float* x, long long P, const float* a, const float* b, const int* perm, void* out, cudaStream_t st);
int ew_image_prep_bwd(const void* g, long long P, const float* a, const int* perm, int accumulate, float* dx, cudaStream_t st);
int ew_bn_finalize(const float* ssum, const float* ssq, const float* bias, const float* gamma, const float* beta, int C,
                   double count, float eps, float decay, float* mm, float* mv, float* scale, float* shift, float* smean,
                   float* srstd, cudaStream_t st);
int ew_bn_stats_apply(const float* ssum, const float* ssq, const float* bias, const float* gamma, const float* beta, double count,
                      float eps, float decay, float* mm, float* mv, float* scale, float* shift, float* smean, float* srstd,
                      const void* x, int relu, int upsample, int N, int H, int W, int C, void* out, int groups, cudaStream_t st);
int ew_bn_act_apply(const void* x, const float* scale, const float* shift, int relu, int upsample, int N, int H, int W, int C,
                    void* out, cudaStream_t st);
int ew_bn_act_bwd(const void* dout, const void* x, const float* scale, const float* shift, const float* mean,
                  const float* rstd, int relu, int upsample, int N, int H, int W, int C, float* dbeta, float* dgamma,
                  void* dx, float* gbeta_acc, float* ggamma_acc, int prezeroed, int groups, cudaStream_t st);
int ew_upsample2x_bwd(const void* dout, int N, int H, int W, int C, void* dact, cudaStream_t st);
int ew_act_mask_bwd(const void* dy, const void* y, float alpha, long long n_elems, void* g, cudaStream_t st);
int ew_maxpool_fwd(const void* x, int N, int H, int W, int C, void* out, cudaStream_t st);
int ew_maxpool_bwd(const void* dy, const void* x, int relu_mask, int N, int H, int W, int C, void* dx, cudaStream_t st);
int ew_compose_fwd(const float* heads, const float* im, long long P, int clip, float* final_out, float* crude_out,
                   float* mask_out, cudaStream_t st);
int ew_compose_bwd(const float* d_final, const float* heads, const float* im, long long P, void* d_heads, cudaStream_t st);
int ew_pack_channels(const void* const* src, const int* C, const int* is_f32, int n, long long P, int Ctot, void* out,
                     cudaStream_t st);
int ew_unpack_channels(const void* g, long long P, int Ctot, void* const* dst, const int* C, const int* is_f32, int n,
                       cudaStream_t st);
int ew_l1_pair(const void* feat_gt, const void* feat_pred, long long half_elems, float weight, float* loss, void* d_pred,
               cudaStream_t st);
int ew_bce_logits(const float* x, int n, float z, float weight, float* loss, void* d_logits, cudaStream_t st);
int ew_adam_tf(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, int t,
               float grad_scale, const float* lr_t_dev, cudaStream_t st);
int ew_channel_sum(const void* g, long long P, int C, float* out, int squares, cudaStream_t st);
int ew_pack_job_blocks(const kp_pack_desc* d);
int ew_pack_weights_batch(const void* jobs_dev, int n_jobs, int total_blocks, cudaStream_t st);
int ew_pack_weights(const float* w, const kp_pack_desc* d, const float* row_scale, void* dst, cudaStream_t st);
int ew_image_prep_unrolled(const float* x, int N, int H, int W, int KW, int pl, int cpad, const float* a, const float* b,
                           const int* perm, void* out, cudaStream_t st);
int ew_image_prep_unrolled_bwd(const void* g, int N, int H, int W, int KW, int pl, int cpad, const float* a, const int* perm,
                               int accumulate, float* dx, cudaStream_t st);

}  // namespace kp
