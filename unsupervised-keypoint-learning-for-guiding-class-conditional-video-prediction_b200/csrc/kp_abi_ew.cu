// extern "C" wrappers of the memory-bound kernels (elementwise.cu): argument validation + dispatch.
#include "kp_common.cuh"
#include "kp_internal.h"

using namespace kp;

#define KP_NONNULL(p) KP_REQUIRE((p) != nullptr, "%s: argument '%s' must not be NULL", __func__, #p)
#define KP_NONNEG(v) KP_REQUIRE((v) >= 0, "%s: argument '%s' must be non-negative", __func__, #v)
#define ST static_cast<cudaStream_t>(stream)

extern "C" {

int kp_image_prep(const float* x, long long P, const float* a, const float* b, const int* perm, void* out, void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(x); KP_NONNULL(a); KP_NONNULL(b); KP_NONNULL(perm); KP_NONNULL(out);
    for (int c = 0; c < 3; ++c) KP_REQUIRE(perm[c] >= 0 && perm[c] < 3, "%s: perm out of range", __func__);
    return ew_image_prep(x, P, a, b, perm, out, ST);
}
int kp_image_prep_bwd(const void* g, long long P, const float* a, const int* perm, int accumulate, float* dx, void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(g); KP_NONNULL(a); KP_NONNULL(perm); KP_NONNULL(dx);
    for (int c = 0; c < 3; ++c) KP_REQUIRE(perm[c] >= 0 && perm[c] < 3, "%s: perm out of range", __func__);
    return ew_image_prep_bwd(g, P, a, perm, accumulate, dx, ST);
}
int kp_bn_finalize(const float* stats_sum, const float* stats_sq, const float* conv_bias, const float* gamma,
                   const float* beta, int C, double count, float eps, float decay, float* moving_mean, float* moving_var,
                   float* scale, float* shift, float* save_mean, float* save_rstd, void* stream) {
    KP_REQUIRE(C > 0 && count > 0, "%s: C and count must be positive", __func__);
    KP_NONNULL(stats_sum); KP_NONNULL(stats_sq); KP_NONNULL(gamma); KP_NONNULL(beta); KP_NONNULL(scale); KP_NONNULL(shift);
    KP_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "%s: moving_mean and moving_var go together", __func__);
    return ew_bn_finalize(stats_sum, stats_sq, conv_bias, gamma, beta, C, count, eps, decay, moving_mean, moving_var, scale,
                          shift, save_mean, save_rstd, ST);
}
int kp_bn_act_apply(const void* x, const float* scale, const float* shift, int relu, int upsample, int N, int H, int W,
                    int C, void* out, void* stream) {
    KP_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", __func__);
    if (N == 0) return KP_OK;
    KP_NONNULL(x); KP_NONNULL(out);
    KP_REQUIRE((scale == nullptr) == (shift == nullptr), "%s: scale and shift go together", __func__);
    return ew_bn_act_apply(x, scale, shift, relu, upsample, N, H, W, C, out, ST);
}
int kp_bn_stats_apply(const float* stats_sum, const float* stats_sq, const float* conv_bias, const float* gamma,
                      const float* beta, double count, float eps, float decay, float* moving_mean, float* moving_var,
                      float* scale, float* shift, float* save_mean, float* save_rstd, const void* x, int relu, int upsample,
                      int N, int H, int W, int C, void* out, int segments, void* stream) {
    KP_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && count > 0, "%s: bad shape", __func__);
    KP_REQUIRE(segments >= 1 && N % segments == 0, "%s: %d images do not split into %d segments", __func__, N, segments);
    KP_NONNULL(stats_sum); KP_NONNULL(stats_sq); KP_NONNULL(gamma); KP_NONNULL(beta); KP_NONNULL(scale); KP_NONNULL(shift);
    KP_NONNULL(x); KP_NONNULL(out);
    KP_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "%s: moving_mean and moving_var go together", __func__);
    return ew_bn_stats_apply(stats_sum, stats_sq, conv_bias, gamma, beta, count, eps, decay, moving_mean, moving_var, scale,
                             shift, save_mean, save_rstd, x, relu, upsample, N, H, W, C, out, segments, ST);
}
int kp_bn_act_bwd(const void* dout, const void* x, const float* scale, const float* shift, const float* save_mean,
                  const float* save_rstd, int relu, int upsample, int N, int H, int W, int C, float* dbeta, float* dgamma,
                  void* dx, float* gbeta_acc, float* ggamma_acc, int prezeroed, int segments, void* stream) {
    KP_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", __func__);
    KP_REQUIRE(segments >= 1 && N % segments == 0, "%s: %d images do not split into %d segments", __func__, N, segments);
    KP_NONNULL(dout); KP_NONNULL(x); KP_NONNULL(scale); KP_NONNULL(shift); KP_NONNULL(save_mean); KP_NONNULL(save_rstd);
    KP_NONNULL(dbeta); KP_NONNULL(dgamma); KP_NONNULL(dx);
    KP_REQUIRE((gbeta_acc == nullptr) == (ggamma_acc == nullptr), "%s: gbeta_acc and ggamma_acc go together", __func__);
    return ew_bn_act_bwd(dout, x, scale, shift, save_mean, save_rstd, relu, upsample, N, H, W, C, dbeta, dgamma, dx, gbeta_acc,
                         ggamma_acc, prezeroed, segments, ST);
}
int kp_upsample2x_bwd(const void* dout, int N, int H, int W, int C, void* dact, void* stream) {
    KP_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", __func__);
    KP_NONNULL(dout); KP_NONNULL(dact);
    return ew_upsample2x_bwd(dout, N, H, W, C, dact, ST);
}
int kp_act_mask_bwd(const void* dy, const void* y, float alpha, long long n_elems, void* g, void* stream) {
    KP_NONNEG(n_elems);
    if (n_elems == 0) return KP_OK;
    KP_NONNULL(dy); KP_NONNULL(y); KP_NONNULL(g);
    return ew_act_mask_bwd(dy, y, alpha, n_elems, g, ST);
}
int kp_maxpool2x2_fwd(const void* x, int N, int H, int W, int C, void* out, void* stream) {
    KP_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", __func__);
    if (N == 0) return KP_OK;
    KP_NONNULL(x); KP_NONNULL(out);
    return ew_maxpool_fwd(x, N, H, W, C, out, ST);
}
int kp_maxpool2x2_bwd(const void* dy, const void* x, int relu_mask, int N, int H, int W, int C, void* dx, void* stream) {
    KP_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", __func__);
    if (N == 0) return KP_OK;
    KP_NONNULL(dy); KP_NONNULL(x); KP_NONNULL(dx);
    return ew_maxpool_bwd(dy, x, relu_mask, N, H, W, C, dx, ST);
}
int kp_mask_compose_fwd(const float* heads, const float* im, long long P, int clip, float* final_out, float* crude_out,
                        float* mask_out, void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(heads); KP_NONNULL(im); KP_NONNULL(final_out);
    return ew_compose_fwd(heads, im, P, clip, final_out, crude_out, mask_out, ST);
}
int kp_mask_compose_bwd(const float* d_final, const float* heads, const float* im, long long P, void* d_heads, void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(d_final); KP_NONNULL(heads); KP_NONNULL(im); KP_NONNULL(d_heads);
    return ew_compose_bwd(d_final, heads, im, P, d_heads, ST);
}
int kp_pack_channels(const void* const* src, const int* C, const int* is_f32, int n, long long P, int Ctot, void* out,
                     void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(src); KP_NONNULL(C); KP_NONNULL(is_f32); KP_NONNULL(out);
    return ew_pack_channels(src, C, is_f32, n, P, Ctot, out, ST);
}
int kp_unpack_channels(const void* g, long long P, int Ctot, void* const* dst, const int* C, const int* is_f32, int n,
                       void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(g); KP_NONNULL(dst); KP_NONNULL(C); KP_NONNULL(is_f32);
    return ew_unpack_channels(g, P, Ctot, dst, C, is_f32, n, ST);
}
int kp_l1_pair_fwd_bwd(const void* feat_gt, const void* feat_pred, long long n_elems, float weight, float* loss,
                       void* d_pred, void* stream) {
    KP_REQUIRE(n_elems > 0, "%s: empty feature", __func__);
    KP_NONNULL(feat_gt); KP_NONNULL(feat_pred); KP_NONNULL(loss);
    return ew_l1_pair(feat_gt, feat_pred, n_elems, weight, loss, d_pred, ST);
}
int kp_bce_logits_fwd_bwd(const float* logits, int n, float label, float weight, float* loss, void* d_logits, void* stream) {
    KP_REQUIRE(n > 0, "%s: empty logits", __func__);
    KP_NONNULL(logits); KP_NONNULL(loss);
    return ew_bce_logits(logits, n, label, weight, loss, d_logits, ST);
}
int kp_adam_tf(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
               int t, float grad_scale, const float* lr_t_dev, void* stream) {
    KP_NONNEG(n);
    if (n == 0) return KP_OK;
    KP_REQUIRE(t >= 1, "%s: step t must be >= 1", __func__);
    KP_NONNULL(p); KP_NONNULL(g); KP_NONNULL(m); KP_NONNULL(v);
    return ew_adam_tf(p, g, m, v, n, lr, beta1, beta2, eps, t, grad_scale, lr_t_dev, ST);
}
int kp_pack_weights(const float* w, const kp_pack_desc* desc, const float* row_scale, void* dst, void* stream) {
    KP_NONNULL(w); KP_NONNULL(desc); KP_NONNULL(dst);
    return ew_pack_weights(w, desc, row_scale, dst, ST);
}
int kp_pack_job_blocks(const kp_pack_desc* desc) {
    if (desc == nullptr || desc->T < 1 || desc->T > KP_MAX_TAPS || desc->Kper <= 0 || desc->Ktot != desc->T * desc->Kper) return 0;
    return ew_pack_job_blocks(desc);
}
int kp_pack_weights_batch(const void* jobs_dev, int n_jobs, int total_blocks, void* stream) {
    KP_NONNULL(jobs_dev);
    return ew_pack_weights_batch(jobs_dev, n_jobs, total_blocks, ST);
}
int kp_conv1x1_f32(const void* x, const float* w, const float* bias, long long P, int Cin, int Cout, float* out, void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(x); KP_NONNULL(w); KP_NONNULL(out);
    return head1x1_launch(x, w, bias, P, Cin, Cout, out, ST);
}
int kp_channel_sum(const void* g, long long P, int C, float* out, void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(g); KP_NONNULL(out);
    return ew_channel_sum(g, P, C, out, 0, ST);
}
int kp_channel_sumsq(const void* g, long long P, int C, float* out, void* stream) {
    KP_NONNEG(P);
    if (P == 0) return KP_OK;
    KP_NONNULL(g); KP_NONNULL(out);
    return ew_channel_sum(g, P, C, out, 1, ST);
}

}  // extern "C"

extern "C" int kp_image_prep_unrolled(const float* x, int N, int H, int W, int KW, int pad_left, int Cpad, const float* a,
                                      const float* b, const int* perm, void* out, void* stream) {
    KP_REQUIRE(N >= 0 && H > 0 && W > 0, "%s: bad shape", __func__);
    if (N == 0) return KP_OK;
    KP_NONNULL(x); KP_NONNULL(a); KP_NONNULL(b); KP_NONNULL(perm); KP_NONNULL(out);
    for (int c = 0; c < 3; ++c) KP_REQUIRE(perm[c] >= 0 && perm[c] < 3, "%s: perm out of range", __func__);
    return ew_image_prep_unrolled(x, N, H, W, KW, pad_left, Cpad, a, b, perm, out, ST);
}
extern "C" int kp_image_prep_unrolled_bwd(const void* g, int N, int H, int W, int KW, int pad_left, int Cpad, const float* a,
                                          const int* perm, int accumulate, float* dx, void* stream) {
    KP_REQUIRE(N >= 0 && H > 0 && W > 0, "%s: bad shape", __func__);
    if (N == 0) return KP_OK;
    KP_NONNULL(g); KP_NONNULL(a); KP_NONNULL(perm); KP_NONNULL(dx);
    for (int c = 0; c < 3; ++c) KP_REQUIRE(perm[c] >= 0 && perm[c] < 3, "%s: perm out of range", __func__);
    return ew_image_prep_unrolled_bwd(g, N, H, W, KW, pad_left, Cpad, a, perm, accumulate, dx, ST);
}
