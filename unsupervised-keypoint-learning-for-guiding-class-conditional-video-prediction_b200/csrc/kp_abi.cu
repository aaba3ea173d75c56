// extern "C" surface declared in include/kp_b200.h: argument validation + dispatch only.
#include "../../include/kp_b200.h"
#include "kp_common.cuh"
#include "kp_internal.h"
#include <stdarg.h>
#include <stdio.h>
#include <atomic>

namespace kp {

static thread_local char g_err[512] = {0};
static std::atomic<unsigned long long> g_launches{0};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return KP_ERR_CUDA;
}

}  // namespace kp

using namespace kp;

#define KP_NONNULL(p) KP_REQUIRE((p) != nullptr, "%s: argument '%s' must not be NULL", __func__, #p)
#define KP_POS(v) KP_REQUIRE((v) > 0, "%s: argument '%s' must be positive (got %d)", __func__, #v, (int)(v))

extern "C" {

int kp_abi_version(void) { return 1; }

const char* kp_last_error(void) { return g_err; }

unsigned long long kp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int kp_softargmax_render_fwd(const float* logits, int B, int H, int W, int K, float* mu, float* prob_x, float* prob_y,
                             float* maps, int map_h, int map_w, float inv_std, void* stream) {
    KP_REQUIRE(B >= 0, "%s: negative batch %d", __func__, B);
    KP_POS(H); KP_POS(W); KP_POS(K);
    if (B == 0) return KP_OK;
    KP_NONNULL(logits); KP_NONNULL(mu);
    if (maps != nullptr) { KP_POS(map_h); KP_POS(map_w); }
    return k1_softargmax_render_fwd(logits, B, H, W, K, mu, prob_x, prob_y, maps, map_h, map_w, inv_std,
                                    static_cast<cudaStream_t>(stream));
}

int kp_softargmax_render_bwd(const float* d_maps, const float* d_mu_extra, const float* mu, const float* prob_x,
                             const float* prob_y, int B, int H, int W, int K, int map_h, int map_w, float inv_std,
                             float* d_logits, float* d_mu_scratch, void* stream) {
    KP_REQUIRE(B >= 0, "%s: negative batch %d", __func__, B);
    KP_POS(H); KP_POS(W); KP_POS(K);
    if (B == 0) return KP_OK;
    KP_REQUIRE(d_maps != nullptr || d_mu_extra != nullptr, "%s: need d_maps and/or d_mu_extra", __func__);
    KP_NONNULL(mu); KP_NONNULL(prob_x); KP_NONNULL(prob_y); KP_NONNULL(d_logits);
    if (d_maps != nullptr) { KP_POS(map_h); KP_POS(map_w); }
    return k1_softargmax_render_bwd(d_maps, d_mu_extra, mu, prob_x, prob_y, B, H, W, K, map_h, map_w, inv_std,
                                    d_logits, d_mu_scratch, static_cast<cudaStream_t>(stream));
}

int kp_render_fwd(const float* mu, int B, int K, int h, int w, float inv_std, float* maps, void* stream) {
    KP_REQUIRE(B >= 0, "%s: negative batch %d", __func__, B);
    KP_POS(K); KP_POS(h); KP_POS(w);
    if (B == 0) return KP_OK;
    KP_NONNULL(mu); KP_NONNULL(maps);
    return k1_render_fwd(mu, B, K, h, w, inv_std, maps, static_cast<cudaStream_t>(stream));
}

int kp_render_bwd(const float* d_maps, const float* d_mu_extra, const float* mu, int B, int K, int h, int w,
                  float inv_std, float* d_mu, void* stream) {
    KP_REQUIRE(B >= 0, "%s: negative batch %d", __func__, B);
    KP_POS(K); KP_POS(h); KP_POS(w);
    if (B == 0) return KP_OK;
    KP_NONNULL(d_maps); KP_NONNULL(mu); KP_NONNULL(d_mu);
    return k1_render_bwd(d_maps, d_mu_extra, mu, B, K, h, w, inv_std, d_mu, static_cast<cudaStream_t>(stream));
}

int kp_render_colorize_fwd(const float* mu, const float* colors, int B, int K, int h, int w, float inv_std, float* out,
                           void* stream) {
    KP_REQUIRE(B >= 0, "%s: negative batch %d", __func__, B);
    KP_POS(K); KP_POS(h); KP_POS(w);
    if (B == 0) return KP_OK;
    KP_NONNULL(mu); KP_NONNULL(colors); KP_NONNULL(out);
    return k1_render_colorize(mu, colors, B, K, h, w, inv_std, out, static_cast<cudaStream_t>(stream));
}

int kp_colorize_fwd(const float* maps, const float* colors, long long n_pixels, int K, float* out, void* stream) {
    KP_REQUIRE(n_pixels >= 0, "%s: negative pixel count", __func__);
    KP_POS(K);
    if (n_pixels == 0) return KP_OK;
    KP_NONNULL(maps); KP_NONNULL(colors); KP_NONNULL(out);
    return k1_colorize(maps, colors, n_pixels, K, out, static_cast<cudaStream_t>(stream));
}

int kp_tapconv_bf16(const kp_tapconv_desc* desc, const void* const* src, const void* wpacked, const float* bias,
                    void* out, float* stats_sum, float* stats_sq, void* stream) {
    KP_NONNULL(desc); KP_NONNULL(src); KP_NONNULL(wpacked); KP_NONNULL(out);
    return tapconv_launch(desc, src, wpacked, bias, out, stats_sum, stats_sq, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

extern "C" int kp_tapconv_wgrad_bf16(const kp_wgrad_desc* desc, const void* x, const void* dy, float* dw, void* stream) {
    KP_NONNULL(desc); KP_NONNULL(x); KP_NONNULL(dy); KP_NONNULL(dw);
    return kp::wgrad_launch(desc, x, dy, dw, static_cast<cudaStream_t>(stream));
}
