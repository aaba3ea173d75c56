// Input pipeline on the device: the per-frame Pillow chain of the reference's loaders
//   data/image_pair_dataloader.py:95-165   rotate -> resize -> crop -> flip -> apply_random_filter -> /255 -> *2-1
//   data/keypoint_dataloader.py:66-82      resize -> centre crop -> /255 -> *2-1, zero frames appended up to 663
//   utils/data.py:8-35                     apply_random_filter (6 ImageFilter kernels, 4 ImageEnhance blends)
// as ONE kernel over decoded uint8 frames resident in HBM.  Pillow 6.2.0 (the reference's pin) resamples rotate() and
// resize() with NEAREST, so the geometry of a frame is a composed gather: output (y, x) -> resized-frame coordinate
// (crop offset, flip: two 128-entry index tables built on the host exactly as Geometry.c ImagingScaleAffine builds them)
// -> rotated-frame coordinate -> source pixel through the 16.16 fixed-point inverse affine map of Geometry.c
// affine_fixed.  The 3x3 / 5x5 filters and the enhancement blends then run on the 128 x 128 tile in shared memory with
// Pillow's float32 operation order (explicit _rn intrinsics: no FMA contraction), and the result is written as fp32 in
// [-1, 1] with 16-byte stores.  Byte-exact against the oracle (oracle/pil_ops.py, pinned against the installed Pillow).
//
// HBM-bound byte work: 49 152 B gathered + 196 608 B written per frame.  A frame is split into four 32-row bands (one
// CTA each, 2-row halo) so that a training batch of 64 frames still fills the machine.
#include <math.h>
#include <string.h>

#include "augment_core.cuh"
#include "kp_common.cuh"
#include "kp_internal.h"

namespace kp {

namespace {

using namespace aug;
constexpr int THREADS = 256;

__constant__ FilterDef kFilters[6] = KP_AUG_FILTER_TABLE;

__global__ void __launch_bounds__(THREADS) augment_kernel(const uint8_t* __restrict__ src, const kp_frame_plan* __restrict__ plans,
                                                          float* __restrict__ out) {
    __shared__ __align__(16) uint8_t tile[ROWS * ROWB];      // gathered crop rows band-2 .. band+33
    __shared__ __align__(16) uint8_t res[BAND * ROWB];       // filtered band
    __shared__ __align__(16) kp_frame_plan plan;
    __shared__ float lut[256];
    __shared__ float kf[25];
    __shared__ unsigned int lsum;

    const int frame = blockIdx.y, band = blockIdx.x, tid = threadIdx.x;
    {   // the plan (568 B) once per CTA
        const int* g = reinterpret_cast<const int*>(plans + frame);
        int* s = reinterpret_cast<int*>(&plan);
        for (int i = tid; i < static_cast<int>(sizeof(kp_frame_plan) / 4); i += THREADS) s[i] = g[i];
    }
    lut[tid] = model_range(tid);
    if (tid == 0) lsum = 0;
    __syncthreads();
    const int fid = plan.filter_id;
    if (fid >= 0 && fid <= 6 && tid < 25) kf[tid] = filter_tap(kFilters, fid, tid);
    float4* out4 = reinterpret_cast<float4*>(out + (static_cast<long long>(frame) * S + band * BAND) * ROWB);
    if (plan.zero) {                                                   // keypoint_dataloader.py:77-80 zero frames -> -1.0
        const float m = lut[0];
        for (int i = tid; i < BAND * ROWB / 4; i += THREADS) out4[i] = make_float4(m, m, m, m);
        return;
    }
    phase_gather(src, plan, band, tile, tid, THREADS);
    if (fid == 9) {   // Contrast needs the mean luma of the WHOLE frame: every band recomputes it (one frame in ten)
        unsigned int part = phase_luma(src, plan, tid, THREADS);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
        if ((tid & 31) == 0) atomicAdd(&lsum, part);
    }
    __syncthreads();
    phase_filter(tile, plan, kf, lsum, band, res, tid, THREADS);
    __syncthreads();
    // store: 4 result bytes -> one float4
    const uchar4* res4 = reinterpret_cast<const uchar4*>(res);
    for (int i = tid; i < BAND * ROWB / 4; i += THREADS) {
        const uchar4 b = res4[i];
        out4[i] = make_float4(lut[b.x], lut[b.y], lut[b.z], lut[b.w]);
    }
}

// Geometry.c ImagingScaleAffine: xo = a0 / 2, then REPEATED additions of a0 = n_in / n_out; COORD() truncates.
void scale_table(int n_in, int n_out, int* tab) {
    const double a0 = static_cast<double>(n_in) / static_cast<double>(n_out);
    double xo = 0.0 + a0 * 0.5;
    for (int x = 0; x < n_out; ++x) {
        const int xin = xo < 0.0 ? -1 : static_cast<int>(xo);
        tab[x] = (xin >= 0 && xin < n_in) ? xin : -1;
        xo += a0;
    }
}

int fix16(double v) { return static_cast<int>(floor(v * 65536.0 + 0.5)); }     // Geometry.c FIX()

bool check_fixed(const double* a, int x, int y) {
    return fabs(x * a[0] + y * a[1] + a[2]) < 32768.0 && fabs(x * a[3] + y * a[4] + a[5]) < 32768.0;
}

// Python 3 round(): half to even (the default IEEE rounding mode of nearbyint)
int py_round(double v) { return static_cast<int>(nearbyint(v)); }

}  // namespace

int augment_launch(const uint8_t* src, const kp_frame_plan* plans, int n, float* out, cudaStream_t st) {
    augment_kernel<<<dim3(aug::S / aug::BAND, n), THREADS, 0, st>>>(src, plans, out);
    KP_LAUNCHED();
    return KP_OK;
}

}  // namespace kp

using namespace kp;

extern "C" {

int kp_augment_plan_host(kp_frame_plan* plan, long long src_offset, int src_w, int src_h, int resize_w, int resize_h,
                         double crop_left, double crop_top, int angle_deg, int flip, int filter_id, double factor) {
    KP_REQUIRE(plan != nullptr, "%s: argument 'plan' must not be NULL", __func__);
    KP_REQUIRE(src_offset >= 0 && src_w > 0 && src_h > 0 && resize_w > 0 && resize_h > 0, "%s: bad frame geometry", __func__);
    KP_REQUIRE(filter_id >= -1 && filter_id <= 9, "%s: filter_id %d outside -1..9", __func__, filter_id);
    if (src_w >= 32768 || src_h >= 32768 || resize_w >= 32768 || resize_h >= 32768) {
        set_error("%s: frames of 32768 pixels or more per side are not supported", __func__);
        return KP_ERR_UNSUPPORTED;
    }
    memset(plan, 0, sizeof(*plan));
    plan->src_offset = src_offset;
    plan->src_w = src_w;
    plan->src_h = src_h;
    plan->filter_id = filter_id;
    plan->factor = static_cast<float>(factor);                  // Image.blend passes a C float
    // Image.rotate(angle): expand 0, centre (w/2, h/2), resample NEAREST; angle % 360 == 0 returns a copy
    int ang = angle_deg % 360;
    if (ang < 0) ang += 360;
    if (ang != 0) {
        const double cx = src_w / 2.0, cy = src_h / 2.0, rad = -(ang * (M_PI / 180.0));
        auto round15 = [](double v) { return nearbyint(v * 1e15) / 1e15; };
        double m[6] = {round15(cos(rad)), round15(sin(rad)), 0.0, round15(-sin(rad)), round15(cos(rad)), 0.0};
        m[2] = m[0] * -cx + m[1] * -cy + m[2];
        m[5] = m[3] * -cx + m[4] * -cy + m[5];
        m[2] += cx;
        m[5] += cy;
        if (!(check_fixed(m, 0, 0) && check_fixed(m, src_w, src_h) && check_fixed(m, 0, src_h) && check_fixed(m, src_w, 0))) {
            set_error("%s: rotation outside the 16.16 fixed-point range", __func__);
            return KP_ERR_UNSUPPORTED;
        }
        plan->rotate = 1;
        plan->a[0] = fix16(m[0]);
        plan->a[1] = fix16(m[1]);
        plan->a[2] = fix16(m[2] + m[0] * 0.5 + m[1] * 0.5);
        plan->a[3] = fix16(m[3]);
        plan->a[4] = fix16(m[4]);
        plan->a[5] = fix16(m[5] + m[3] * 0.5 + m[4] * 0.5);
    }
    // Image.resize([resize_w, resize_h], NEAREST) then Image.crop((left, top, left + 128, top + 128)) [then FLIP_LEFT_RIGHT]
    int* xt = new int[resize_w];
    int* yt = new int[resize_h];
    scale_table(src_w, resize_w, xt);
    scale_table(src_h, resize_h, yt);
    const int x0 = py_round(crop_left), y0 = py_round(crop_top);
    for (int i = 0; i < KP_AUG_SIZE; ++i) {
        const int xo = flip ? KP_AUG_SIZE - 1 - i : i;
        const int xs = x0 + xo, ys = y0 + i;
        plan->xtab[i] = static_cast<short>((xs >= 0 && xs < resize_w) ? xt[xs] : -1);
        plan->ytab[i] = static_cast<short>((ys >= 0 && ys < resize_h) ? yt[ys] : -1);
    }
    delete[] xt;
    delete[] yt;
    return KP_OK;
}

int kp_augment_plan_zero_host(kp_frame_plan* plan) {
    KP_REQUIRE(plan != nullptr, "%s: argument 'plan' must not be NULL", __func__);
    memset(plan, 0, sizeof(*plan));
    plan->zero = 1;
    plan->filter_id = -1;
    return KP_OK;
}

int kp_augment_plan_batch_host(kp_frame_plan* plans, int n, const long long* src_offset, const int* src_w, const int* src_h,
                               const int* resize_w, const int* resize_h, const double* crop_left, const double* crop_top,
                               const int* angle_deg, const int* flip, const int* filter_id, const double* factor,
                               const int* zero) {
    KP_REQUIRE(n >= 0, "%s: argument 'n' must be non-negative", __func__);
    if (n == 0) return KP_OK;
    KP_REQUIRE(plans && src_offset && src_w && src_h && resize_w && resize_h && crop_left && crop_top && angle_deg && flip &&
                   filter_id && factor && zero,
               "%s: NULL argument", __func__);
    for (int i = 0; i < n; ++i) {
        const int rc = zero[i] ? kp_augment_plan_zero_host(plans + i)
                               : kp_augment_plan_host(plans + i, src_offset[i], src_w[i], src_h[i], resize_w[i], resize_h[i],
                                                      crop_left[i], crop_top[i], angle_deg[i], flip[i], filter_id[i], factor[i]);
        if (rc != KP_OK) return rc;
    }
    return KP_OK;
}

int kp_host_register(void* host_ptr, unsigned long long bytes) {
    KP_REQUIRE(host_ptr != nullptr && bytes > 0, "%s: empty range", __func__);
    const cudaError_t e = cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();      // a refused registration is an answer, not a fault: leave no sticky error behind
        set_error("%s: cudaHostRegister: %s", __func__, cudaGetErrorString(e));
        return KP_ERR_CUDA;
    }
    return KP_OK;
}

int kp_host_unregister(void* host_ptr) {
    KP_REQUIRE(host_ptr != nullptr, "%s: argument 'host_ptr' must not be NULL", __func__);
    const cudaError_t e = cudaHostUnregister(host_ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: cudaHostUnregister: %s", __func__, cudaGetErrorString(e));
        return KP_ERR_CUDA;
    }
    return KP_OK;
}

int kp_augment_frames(const unsigned char* src, const kp_frame_plan* plans, int n_frames, float* out, void* stream) {
    KP_REQUIRE(n_frames >= 0, "%s: argument 'n_frames' must be non-negative", __func__);
    if (n_frames == 0) return KP_OK;
    KP_REQUIRE(n_frames <= 65535, "%s: at most 65535 frames per call", __func__);
    KP_REQUIRE(plans != nullptr, "%s: argument 'plans' must not be NULL", __func__);
    KP_REQUIRE(out != nullptr, "%s: argument 'out' must not be NULL", __func__);
    KP_REQUIRE(src != nullptr, "%s: argument 'src' must not be NULL", __func__);
    return augment_launch(src, plans, n_frames, out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
