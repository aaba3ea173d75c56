// The per-band phases of the augmentation kernel (csrc/augment.cu) as host+device functions of (thread id, thread count):
// the kernel calls them between barriers with (threadIdx.x, blockDim.x); tests/host_emu/augment_host.cpp compiles the SAME
// code with g++ (-ffp-contract=off) and runs the phases serially, so that the arithmetic is checked against the oracle on a
// CPU-only box before it ever runs on a GPU.  Pillow's float32 operation order is spelled out with _rn intrinsics on the
// device (no FMA contraction).
#pragma once
#include <stdint.h>
#include "../../include/kp_b200.h"

#ifdef __CUDACC__
#define KP_HD __host__ __device__ __forceinline__
#else
#define KP_HD inline
#endif

namespace kp {
namespace aug {

constexpr int S = KP_AUG_SIZE;       // 128
constexpr int BAND = 32;             // output rows per CTA
constexpr int HALO = 2;              // 5x5 filters
constexpr int ROWS = BAND + 2 * HALO;
constexpr int ROWB = S * 3;          // bytes of a tile row

KP_HD float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
KP_HD float add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}

// ImageFilter built-ins in the order of utils/data.py:11-22 (r_id 0..5): size, scale, taps (row-major)
struct FilterDef {
    int size;
    int scale;
    signed char k[25];
};
#define KP_AUG_FILTER_TABLE                                                                                           \
    {                                                                                                                 \
        {3, 6, {0, -1, 0, -1, 10, -1, 0, -1, 0}},                                                /* DETAIL */            \
        {3, 2, {-1, -1, -1, -1, 10, -1, -1, -1, -1}},                                            /* EDGE_ENHANCE */      \
        {3, 13, {1, 1, 1, 1, 5, 1, 1, 1, 1}},                                                    /* SMOOTH */            \
        {5, 100, {1, 1, 1, 1, 1, 1, 5, 5, 5, 1, 1, 5, 44, 5, 1, 1, 5, 5, 5, 1, 1, 1, 1, 1, 1}}, /* SMOOTH_MORE */       \
        {3, 1, {-1, -1, -1, -1, 9, -1, -1, -1, -1}},                                             /* EDGE_ENHANCE_MORE */ \
        {5, 16, {1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 1}},   /* BLUR */              \
    }

// kernel taps divided by the scale in float32 (_imaging.c divides the float32 list in place); Sharpness (6) blends
// against SMOOTH (2)
KP_HD float filter_tap(const FilterDef* table, int fid, int i) {
    const FilterDef& f = table[fid == 6 ? 2 : fid];
    if (i >= f.size * f.size) return 0.0f;
#ifdef __CUDA_ARCH__
    return __fdiv_rn(static_cast<float>(f.k[i]), static_cast<float>(f.scale));
#else
    return static_cast<float>(f.k[i]) / static_cast<float>(f.scale);
#endif
}

// np.asarray(image) / 255.0 (float64) -> tf.float32 -> map_fn's * 2.0 - 1.0 in fp32
KP_HD float model_range(int v) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(__fmul_rn(__double2float_rn(__ddiv_rn(static_cast<double>(v), 255.0)), 2.0f), 1.0f);
#else
    const float x = static_cast<float>(static_cast<double>(v) / 255.0);
    const float y = x * 2.0f;
    return y - 1.0f;
#endif
}

KP_HD uint8_t clip8(float v) {      // Filter.c clip8 / Blend.c extrapolation clamp
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return static_cast<uint8_t>(static_cast<int>(v));
}

// source byte index of output (row, column) given the table entries (ys, xs); false = outside (Pillow leaves 0)
KP_HD bool source_index(const kp_frame_plan& p, int ys, int xs, long long* idx) {
    if ((ys | xs) < 0) return false;
    int xin = xs, yin = ys;
    if (p.rotate) {   // Geometry.c affine_fixed
        const long long xx = static_cast<long long>(p.a[2]) + static_cast<long long>(ys) * p.a[1] + static_cast<long long>(xs) * p.a[0];
        const long long yy = static_cast<long long>(p.a[5]) + static_cast<long long>(ys) * p.a[4] + static_cast<long long>(xs) * p.a[3];
        xin = static_cast<int>(xx >> 16);
        yin = static_cast<int>(yy >> 16);
        if (xin < 0 || xin >= p.src_w || yin < 0 || yin >= p.src_h) return false;
    }
    *idx = p.src_offset + (static_cast<long long>(yin) * p.src_w + xin) * 3;
    return true;
}

KP_HD int luma(const uint8_t* px) {                          // Convert.c rgb2l
    return (px[0] * 19595 + px[1] * 38470 + px[2] * 7471 + 0x8000) >> 16;
}

// Filter.c ImagingFilter3x3 / 5x5 at staged row r, column x, channel c: ss starts at offset + 0.5, kernel row j multiplies
// image row y + radius - j, the taps of a row are summed left to right, rows are added one at a time.
template <int SIZE>
KP_HD uint8_t filter_px(const uint8_t* tile, const float* kf, int r, int x, int c) {
    constexpr int R = SIZE / 2;
    float ss = 0.5f;
#pragma unroll
    for (int j = 0; j < SIZE; ++j) {
        const uint8_t* row = tile + (r + R - j) * ROWB + (x - R) * 3 + c;
        float acc = mul_rn(static_cast<float>(row[0]), kf[j * SIZE]);
#pragma unroll
        for (int i = 1; i < SIZE; ++i) acc = add_rn(acc, mul_rn(static_cast<float>(row[i * 3]), kf[j * SIZE + i]));
        ss = add_rn(ss, acc);
    }
    return clip8(ss);
}

KP_HD uint8_t blend_px(int deg, int img, float alpha) {     // Blend.c
    const float t = add_rn(static_cast<float>(deg), mul_rn(alpha, static_cast<float>(img - deg)));
    if (alpha >= 0.0f && alpha <= 1.0f) return static_cast<uint8_t>(static_cast<int>(t));
    return clip8(t);
}

// ---- phase 1: gather crop rows band*32-2 .. band*32+33 into the tile ------------------------------------------------------
// Latency-bound if done pixel by pixel: the indices of GATHER_UNROLL pixels are computed first, then all their loads are
// issued, then the tile is written.
constexpr int GATHER_UNROLL = 6;
KP_HD void phase_gather(const uint8_t* src, const kp_frame_plan& plan, int band, uint8_t* tile, int tid, int nthr) {
    const int row0 = band * BAND - HALO;
    for (int base = tid; base < ROWS * S; base += nthr * GATHER_UNROLL) {
        long long idx[GATHER_UNROLL];
        uint8_t v[GATHER_UNROLL][3];
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u) {
            const int i = base + u * nthr;
            const int r = i / S, x = i - r * S, y = row0 + r;
            idx[u] = -1;
            long long t;
            if (i < ROWS * S && y >= 0 && y < S && source_index(plan, plan.ytab[y], plan.xtab[x], &t)) idx[u] = t;
        }
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u) {
            v[u][0] = v[u][1] = v[u][2] = 0;
            if (idx[u] >= 0) v[u][0] = src[idx[u]], v[u][1] = src[idx[u] + 1], v[u][2] = src[idx[u] + 2];
        }
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u) {
            const int i = base + u * nthr;
            if (i < ROWS * S) {
                uint8_t* t = tile + i * 3;              // (i / S) * ROWB + (i % S) * 3
                t[0] = v[u][0], t[1] = v[u][1], t[2] = v[u][2];
            }
        }
    }
}

// ---- phase 1b (Contrast only): this thread's share of the luma sum of the WHOLE frame -----------------------------------
KP_HD unsigned int phase_luma(const uint8_t* src, const kp_frame_plan& plan, int tid, int nthr) {
    unsigned int part = 0;
    for (int i = tid; i < S * S; i += nthr) {
        const int y = i / S, x = i - y * S;
        uint8_t px[3] = {0, 0, 0};
        long long idx;
        if (source_index(plan, plan.ytab[y], plan.xtab[x], &idx)) px[0] = src[idx], px[1] = src[idx + 1], px[2] = src[idx + 2];
        part += luma(px);
    }
    return part;
}

// ---- phase 2: filter / enhance the band into res --------------------------------------------------------------------------
KP_HD void phase_filter(const uint8_t* tile, const kp_frame_plan& plan, const float* kf, unsigned int luma_sum, int band,
                        uint8_t* res, int tid, int nthr) {
    const int fid = plan.filter_id;
    const float alpha = plan.factor;
    // ImageEnhance.Contrast: int(ImageStat mean + 0.5); sum / 16384 is exact in double
    const int gray_mean = static_cast<int>((luma_sum + (S * S / 2)) / (S * S));
    for (int i = tid; i < BAND * S; i += nthr) {
        const int rr = i / S, x = i - rr * S, y = band * BAND + rr, r = rr + HALO;
        const uint8_t* px = tile + r * ROWB + x * 3;
        uint8_t o[3] = {px[0], px[1], px[2]};
        if (fid >= 0 && fid <= 6) {
            const bool five = (fid == 3 || fid == 5);
            const int rad = five ? 2 : 1;
            const bool inner = y >= rad && y < S - rad && x >= rad && x < S - rad;     // the frame is copied
            uint8_t f[3] = {px[0], px[1], px[2]};
            if (inner) {
#pragma unroll
                for (int c = 0; c < 3; ++c) f[c] = five ? filter_px<5>(tile, kf, r, x, c) : filter_px<3>(tile, kf, r, x, c);
            }
            if (fid == 6) {              // ImageEnhance.Sharpness: blend(SMOOTH(image), image, factor)
                if (alpha == 0.0f) {
                    o[0] = f[0], o[1] = f[1], o[2] = f[2];
                } else if (alpha != 1.0f) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) o[c] = blend_px(f[c], px[c], alpha);
                }
            } else {
                o[0] = f[0], o[1] = f[1], o[2] = f[2];
            }
        } else if (fid >= 7 && fid <= 9) {   // Brightness (black), Color (luma), Contrast (mean luma)
            const int deg = fid == 7 ? 0 : (fid == 8 ? luma(px) : gray_mean);
            if (alpha == 0.0f) {
                o[0] = o[1] = o[2] = static_cast<uint8_t>(deg);
            } else if (alpha != 1.0f) {
#pragma unroll
                for (int c = 0; c < 3; ++c) o[c] = blend_px(deg, px[c], alpha);
            }
        }
        uint8_t* d = res + rr * ROWB + x * 3;
        d[0] = o[0], d[1] = o[1], d[2] = o[2];
    }
}

}  // namespace aug
}  // namespace kp
