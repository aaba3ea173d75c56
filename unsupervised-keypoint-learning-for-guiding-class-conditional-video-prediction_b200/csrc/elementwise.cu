// Memory-bound companions of the convolution engine (sm_100a): batch-norm statistics/apply/backward,
// legacy-bilinear x2 upsampling (fused into the BN+ReLU pass), max-pool, image packing, mask compose,
// channel pack/unpack, losses and the TF-style Adam update.  Everything is bf16 NHWC with 16-byte
// (8-channel) vector accesses unless the reference boundary demands fp32.
//
// Reference ops replaced (all under /root/reference):
//   tf.contrib.layers.batch_norm + tf.nn.relu    models/networks/layers.py:13-14, networks/__init__.py:11-12...
//   tf.image.resize_images (legacy bilinear x2)  models/networks/__init__.py:63,98
//   tf.nn.max_pool 2x2 s2                        models/networks/vgg.py:45-46
//   im*mask + crude*(1-mask), clip               models/detector_translator_model.py:174, final_model.py:96-99
//   VGG preprocessing                            detector_translator_model.py:262-263, vgg.py:17-19
//   tf.concat for the joint embedding            detector_translator_model.py:170
//   L1 feature loss / BCE-with-logits            detector_translator_model.py:249-254,265-267,280-287
//   tf.train.AdamOptimizer                       detector_translator_model.py:198-202
#include "kp_common.cuh"
#include "kp_internal.h"
#include <math.h>

namespace kp {

struct bf8 {  // eight bf16 values = one 16-byte vector
    uint4 u;
};
__device__ __forceinline__ void bf8_unpack(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 bf8_pack(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
static inline int grid_for(long long work_items, int block, int max_blocks = 148 * 16) {
    long long b = (work_items + block - 1) / block;
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (int)b;
}

// =============================================================================================
// image packing: f32 [P,3] -> bf16 [P,16], out[c] = a[c]*x[perm[c]] + b[c] (c<3), 0 for c>=3
// =============================================================================================
struct ImgPrep {
    float a[3], b[3];
    int perm[3];
};
__global__ void image_prep_kernel(const float* __restrict__ x, long long P, ImgPrep q, __nv_bfloat16* __restrict__ out) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const float v0 = x[3 * p], v1 = x[3 * p + 1], v2 = x[3 * p + 2];
        const float in[3] = {v0, v1, v2};
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 3; ++c) f[c] = fmaf(q.a[c], in[q.perm[c]], q.b[c]);
        uint4* o = reinterpret_cast<uint4*>(out + 16 * p);
        o[0] = bf8_pack(f);
        o[1] = make_uint4(0, 0, 0, 0);
    }
}
// backward: g bf16 [P,16] -> dx f32 [P,3]: dx[perm[c]] (+)= a[c]*g[c]
__global__ void image_prep_bwd_kernel(const __nv_bfloat16* __restrict__ g, long long P, ImgPrep q, int accumulate,
                                      float* __restrict__ dx) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float f[8];
        bf8_unpack(*reinterpret_cast<const uint4*>(g + 16 * p), f);
        float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) o[q.perm[c]] = q.a[c] * f[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) dx[3 * p + c] = accumulate ? dx[3 * p + c] + o[c] : o[c];
    }
}

// =============================================================================================
// batch norm
// =============================================================================================
// finalize: stats of the PRE-bias accumulators -> scale/shift for y = conv+bias, saved mean/rstd, moving averages.
__global__ void bn_finalize_kernel(const float* __restrict__ ssum, const float* __restrict__ ssq, const float* __restrict__ bias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int C, float count,
                                   float eps, float decay, float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m0 = (double)ssum[c] / count;
    double var = (double)ssq[c] / count - m0 * m0;
    if (var < 0.0) var = 0.0;
    const float mean = (float)m0 + (bias ? bias[c] : 0.f);
    const float rstd = rsqrtf((float)var + eps);
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    if (save_mean) save_mean[c] = mean;
    if (save_rstd) save_rstd[c] = rstd;
    if (moving_mean) {
        const float unbiased = (float)(var * (count / fmax(count - 1.0, 1.0)));
        moving_mean[c] = moving_mean[c] * decay + mean * (1.f - decay);
        moving_var[c] = moving_var[c] * decay + unbiased * (1.f - decay);
    }
}

// y = relu(x*scale + shift) (scale == nullptr: identity), optionally followed by the legacy bilinear x2
// upsampling (out[2i] = in[i], out[2i+1] = (in[i] + in[min(i+1,n-1)])/2 on each axis).
// Per-channel scale/shift live in shared memory (two LDS.128 per 8-channel vector instead of 16 global loads).
// Plain variant: pure streaming, two independent vectors per loop trip.  UPSAMPLE variant: one thread per INPUT
// pixel-vector produces the 2x2 output block (4 normalised loads feed 4 stores, instead of up to 4 loads per store).
// With fin.ssum != nullptr the kernel also does the work of bn_finalize_kernel: every block derives scale/shift from the
// raw sums (C <= a few hundred values), block 0 publishes scale/shift/mean/rstd for the backward pass and updates the
// moving averages - one launch per batch-norm layer instead of two.
struct BnFin {
    const float *ssum, *ssq, *bias, *gamma, *beta;
    float count, eps, decay;
    float *moving_mean, *moving_var, *scale, *shift, *save_mean, *save_rstd;
};
template <bool UPSAMPLE>
__global__ void __launch_bounds__(256)
bn_act_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                    const BnFin fin, int relu, int N, int H, int W, int C, __nv_bfloat16* __restrict__ out) {
    extern __shared__ float sp[];   // [2][C]
    pdl_launch_dependents();
    pdl_wait();
    // Batch segments (gridDim.y of them, N images each) are normalised with their OWN statistics: the two calls of the
    // shared pose_encoder run as one launch on [image; future_image].  All per-segment vectors are [segments][C].
    const int seg = blockIdx.y;
    x += (long long)seg * N * H * W * C;
    out += (long long)seg * N * H * W * C * (UPSAMPLE ? 4 : 1);
    if (fin.ssum != nullptr) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double m0 = (double)fin.ssum[seg * C + c] / fin.count;
            double var = (double)fin.ssq[seg * C + c] / fin.count - m0 * m0;
            if (var < 0.0) var = 0.0;
            const float mean = (float)m0 + (fin.bias ? fin.bias[c] : 0.f);
            const float rstd = rsqrtf((float)var + fin.eps);
            const float sc = fin.gamma[c] * rstd;
            const float sh = fin.beta[c] - mean * sc;
            sp[c] = sc;
            sp[C + c] = sh;
            if (blockIdx.x == 0) {
                fin.scale[seg * C + c] = sc;
                fin.shift[seg * C + c] = sh;
                if (fin.save_mean) fin.save_mean[seg * C + c] = mean;
                if (fin.save_rstd) fin.save_rstd[seg * C + c] = rstd;
                if (fin.moving_mean && seg == 0) {
                    // one thread per channel applies the moving-average updates of ALL segments, in call order (the
                    // reference runs one update op per pose_encoder call)
                    float mm = fin.moving_mean[c], mv = fin.moving_var[c];
                    for (int g = 0; g < (int)gridDim.y; ++g) {
                        const double gm0 = (double)fin.ssum[g * C + c] / fin.count;
                        double gvar = (double)fin.ssq[g * C + c] / fin.count - gm0 * gm0;
                        if (gvar < 0.0) gvar = 0.0;
                        const float gmean = (float)gm0 + (fin.bias ? fin.bias[c] : 0.f);
                        const float unbiased = (float)(gvar * (fin.count / fmax(fin.count - 1.0, 1.0)));
                        mm = mm * fin.decay + gmean * (1.f - fin.decay);
                        mv = mv * fin.decay + unbiased * (1.f - fin.decay);
                    }
                    fin.moving_mean[c] = mm;
                    fin.moving_var[c] = mv;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            sp[i] = scale ? scale[seg * C + i] : 1.f;
            sp[C + i] = scale ? shift[seg * C + i] : 0.f;
        }
    }
    __syncthreads();
    const int CG = C >> 3;
    const long long total = (long long)N * H * W * CG;   // input vectors
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto norm = [&](const uint4& u, int cg, float* f) {
        bf8_unpack(u, f);
        const float4 s0 = *reinterpret_cast<const float4*>(sp + cg * 8), s1 = *reinterpret_cast<const float4*>(sp + cg * 8 + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(sp + C + cg * 8), h1 = *reinterpret_cast<const float4*>(sp + C + cg * 8 + 4);
        f[0] = fmaf(f[0], s0.x, h0.x); f[1] = fmaf(f[1], s0.y, h0.y); f[2] = fmaf(f[2], s0.z, h0.z); f[3] = fmaf(f[3], s0.w, h0.w);
        f[4] = fmaf(f[4], s1.x, h1.x); f[5] = fmaf(f[5], s1.y, h1.y); f[6] = fmaf(f[6], s1.z, h1.z); f[7] = fmaf(f[7], s1.w, h1.w);
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
        }
    };
    const uint4* x4 = reinterpret_cast<const uint4*>(x);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    if (!UPSAMPLE && (CG & (CG - 1)) == 0 && CG <= (int)blockDim.x) {
        // C/8 a power of two: the grid stride is a multiple of it, so a thread sees ONE channel group for its whole loop -
        // its 8 scale / shift pairs live in registers and nothing but the 16-byte vectors moves (the general path below
        // pays a 64-bit modulo and four shared-memory loads per vector; isolated it ran at 4.1-4.9 TB/s against 6.0 for a
        // plain copy of the same tensors).  Four independent vectors in flight per thread.
        const int cg = threadIdx.x & (CG - 1);
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = sp[cg * 8 + j];
            sh[j] = sp[C + cg * 8 + j];
        }
        const float lo = relu ? 0.f : -INFINITY;
        auto norm8 = [&](const uint4& u) {
            float f[8];
            bf8_unpack(u, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), lo);
            return bf8_pack(f);
        };
        long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; idx + 3 * stride < total; idx += 4 * stride) {
            const uint4 u0 = x4[idx], u1 = x4[idx + stride], u2 = x4[idx + 2 * stride], u3 = x4[idx + 3 * stride];
            o4[idx] = norm8(u0);
            o4[idx + stride] = norm8(u1);
            o4[idx + 2 * stride] = norm8(u2);
            o4[idx + 3 * stride] = norm8(u3);
        }
        for (; idx < total; idx += stride) o4[idx] = norm8(x4[idx]);
    } else if (!UPSAMPLE) {
        long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; idx + stride < total; idx += 2 * stride) {
            const uint4 u0 = x4[idx], u1 = x4[idx + stride];
            float f0[8], f1[8];
            norm(u0, (int)(idx % CG), f0);
            norm(u1, (int)((idx + stride) % CG), f1);
            o4[idx] = bf8_pack(f0);
            o4[idx + stride] = bf8_pack(f1);
        }
        if (idx < total) {
            float f0[8];
            norm(x4[idx], (int)(idx % CG), f0);
            o4[idx] = bf8_pack(f0);
        }
    } else {
        const int Wo = 2 * W;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
            const int cg = (int)(idx % CG);
            long long pix = idx / CG;
            const int w = (int)(pix % W);
            pix /= W;
            const int h = (int)(pix % H);
            const int n = (int)(pix / H);
            const int w1 = min(w + 1, W - 1), h1 = min(h + 1, H - 1);
            const long long rowb = ((long long)n * H + h) * W, rowc = ((long long)n * H + h1) * W;
            float a[8], b[8], c[8], d[8];
            norm(x4[(rowb + w) * CG + cg], cg, a);
            norm(x4[(rowb + w1) * CG + cg], cg, b);
            norm(x4[(rowc + w) * CG + cg], cg, c);
            norm(x4[(rowc + w1) * CG + cg], cg, d);
            float o01[8], o10[8], o11[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // TF computes the top/bottom row lerps along x first, then the lerp along y
                const float top = a[j] + (b[j] - a[j]) * 0.5f;
                const float bot = c[j] + (d[j] - c[j]) * 0.5f;
                o01[j] = top;
                o10[j] = a[j] + (c[j] - a[j]) * 0.5f;
                o11[j] = top + (bot - top) * 0.5f;
            }
            const long long ob = (((long long)n * 2 * H + 2 * h) * Wo + 2 * w) * CG + cg;
            o4[ob] = bf8_pack(a);
            o4[ob + CG] = bf8_pack(o01);
            o4[ob + (long long)Wo * CG] = bf8_pack(o10);
            o4[ob + (long long)Wo * CG + CG] = bf8_pack(o11);
        }
    }
}

// Gradient arriving at the (un-upsampled) activation a = relu(z): for UPSAMPLE the adjoint of the x2 resize is
// gathered from the 3x3 neighbourhood of dOut [N,2H,2W,C].
template <bool UPSAMPLE>
__device__ __forceinline__ void gather_dact(const __nv_bfloat16* __restrict__ dout, int n, int h, int w, int cg, int H, int W,
                                            int C, float* g) {
    if (!UPSAMPLE) {
        bf8_unpack(*reinterpret_cast<const uint4*>(dout + (((long long)n * H + h) * W + w) * C + cg * 8), g);
        return;
    }
    const int Ho = 2 * H, Wo = 2 * W;
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    // 1-D weights of out rows {2h-1, 2h, 2h+1} on in row h: 0.5 (if h>=1), 1, 0.5 (+0.5 more if h == H-1)
    float wy[3] = {h >= 1 ? 0.5f : 0.f, 1.f, h == H - 1 ? 1.f : 0.5f};
    float wx[3] = {w >= 1 ? 0.5f : 0.f, 1.f, w == W - 1 ? 1.f : 0.5f};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int oh = 2 * h - 1 + dy;
        if (oh < 0 || oh >= Ho || wy[dy] == 0.f) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int ow = 2 * w - 1 + dx;
            if (ow < 0 || ow >= Wo || wx[dx] == 0.f) continue;
            float t[8];
            bf8_unpack(*reinterpret_cast<const uint4*>(dout + (((long long)n * Ho + oh) * Wo + ow) * C + cg * 8), t);
            const float wgt = wy[dy] * wx[dx];
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = fmaf(wgt, t[j], g[j]);
        }
    }
}

// adjoint of the legacy bilinear x2 resize alone: dOut bf16 [N,2H,2W,C] -> dAct bf16 [N,H,W,C].  Used by the BN backward
// of the upsampling layers: gathering the 3x3 neighbourhood once (instead of inside BOTH backward passes) turned
// 150 + 69 us into ~55 us for the 128-channel 64x64 -> 128x128 layer.
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int N, int H, int W, int C, __nv_bfloat16* __restrict__ dact) {
    pdl_launch_dependents();
    pdl_wait();
    const int CG = C >> 3;
    const long long total = (long long)N * H * W * CG;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % CG);
        long long pix = idx / CG;
        const int w = (int)(pix % W);
        pix /= W;
        const int h = (int)(pix % H);
        const int n = (int)(pix / H);
        float g[8];
        gather_dact<true>(dout, n, h, w, cg, H, W, C, g);
        *reinterpret_cast<uint4*>(dact + idx * 8) = bf8_pack(g);
    }
}

// pass 1 of the BN+ReLU(+upsample) backward: dbeta[c] += sum g, dgamma[c] += sum g*xhat  (g = dAct * (z>0))
template <bool UPSAMPLE>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                         const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ rstd, int relu, int N, int H, int W, int C, float* __restrict__ dbeta,
                         float* __restrict__ dgamma) {
    pdl_launch_dependents();
    pdl_wait();
    {   // batch segment blockIdx.y (N images each): own statistics, own sums; all per-segment vectors are [segments][C]
        const int seg = blockIdx.y;
        dout += (long long)seg * N * H * W * C * (UPSAMPLE ? 4 : 1);
        x += (long long)seg * N * H * W * C;
        scale += seg * C; shift += seg * C; mean += seg * C; rstd += seg * C; dbeta += seg * C; dgamma += seg * C;
    }
    const int CG = C >> 3;                 // power of two, <= 256
    const int cg = threadIdx.x % CG;
    const int lanes = blockDim.x / CG;     // pixel lanes per block
    const int pl = threadIdx.x / CG;
    const long long P = (long long)N * H * W;
    float sc[8], sh[8], mu[8], rs[8], sb[8], sg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = scale[cg * 8 + j]; sh[j] = shift[cg * 8 + j];
        mu[j] = mean[cg * 8 + j]; rs[j] = rstd[cg * 8 + j];
        sb[j] = 0.f; sg[j] = 0.f;
    }
    auto accum = [&](const float* g, const float* xv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float z = fmaf(xv[j], sc[j], sh[j]);
            const float gj = (relu && z <= 0.f) ? 0.f : g[j];
            sb[j] += gj;
            sg[j] = fmaf(gj, (xv[j] - mu[j]) * rs[j], sg[j]);
        }
    };
    const long long pstride = (long long)gridDim.x * lanes;
    long long p = (long long)blockIdx.x * lanes + pl;
    if (!UPSAMPLE) {
        // four independent pixel-vectors per trip (8 x 16-byte loads in flight per thread)
        for (; p + 3 * pstride < P; p += 4 * pstride) {
            uint4 gq[4], xq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                gq[k] = *reinterpret_cast<const uint4*>(dout + (p + k * pstride) * C + cg * 8);
                xq[k] = *reinterpret_cast<const uint4*>(x + (p + k * pstride) * C + cg * 8);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float g[8], xv[8];
                bf8_unpack(gq[k], g); bf8_unpack(xq[k], xv); accum(g, xv);
            }
        }
        for (; p + pstride < P; p += 2 * pstride) {
            const uint4 g0 = *reinterpret_cast<const uint4*>(dout + p * C + cg * 8);
            const uint4 x0 = *reinterpret_cast<const uint4*>(x + p * C + cg * 8);
            const uint4 g1 = *reinterpret_cast<const uint4*>(dout + (p + pstride) * C + cg * 8);
            const uint4 x1 = *reinterpret_cast<const uint4*>(x + (p + pstride) * C + cg * 8);
            float g[8], xv[8];
            bf8_unpack(g0, g); bf8_unpack(x0, xv); accum(g, xv);
            bf8_unpack(g1, g); bf8_unpack(x1, xv); accum(g, xv);
        }
    }
    for (; p < P; p += pstride) {
        const int w = (int)(p % W);
        const int h = (int)((p / W) % H);
        const int n = (int)(p / ((long long)W * H));
        float g[8], xv[8];
        gather_dact<UPSAMPLE>(dout, n, h, w, cg, H, W, C, g);
        bf8_unpack(*reinterpret_cast<const uint4*>(x + p * C + cg * 8), xv);
        accum(g, xv);
    }
    // Reduce over the pixel lanes that share a channel group: shuffles inside the warp (lanes cg, cg+CG, ...), then
    // shared-memory atomics across the 8 warps, then ONE global atomic per channel per block.  (The first version let
    // the CG threads of pixel lane 0 walk all other lanes serially: 2 threads x 127 lanes for C = 16.)
    __shared__ float red[2][2048];
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i / C][i % C] = 0.f;
    __syncthreads();
    if (CG < 32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            for (int off = 16; off >= CG; off >>= 1) {
                sb[j] += __shfl_xor_sync(0xffffffffu, sb[j], off);
                sg[j] += __shfl_xor_sync(0xffffffffu, sg[j], off);
            }
        }
    }
    if ((threadIdx.x & 31) < CG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&red[0][cg * 8 + j], sb[j]);
            atomicAdd(&red[1][cg * 8 + j], sg[j]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(dbeta + c, red[0][c]);
        atomicAdd(dgamma + c, red[1][c]);
    }
}

// pass 2: dx = gamma*rstd*(g - dbeta/n - xhat*dgamma/n)   (bf16), gamma*rstd == scale
// With xhat = (x-mean)*rstd the formula is affine in (g, x): dx = sc*g + k1*x + k0, k1 = -sc*rstd*dgamma/n,
// k0 = -sc*dbeta/n - k1*mean.  The four per-channel coefficients (sc, sh for the mask, k1, k0) sit in shared memory.
template <bool UPSAMPLE>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                        const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                        int relu, int N, int H, int W, int C, __nv_bfloat16* __restrict__ dx,
                        float* __restrict__ gbeta_acc, float* __restrict__ ggamma_acc) {
    extern __shared__ float sp[];   // [4][C]: sc, sh, k1, k0
    pdl_launch_dependents();
    pdl_wait();
    {
        const int seg = blockIdx.y;
        dout += (long long)seg * N * H * W * C * (UPSAMPLE ? 4 : 1);
        x += (long long)seg * N * H * W * C;
        dx += (long long)seg * N * H * W * C;
        scale += seg * C; shift += seg * C; mean += seg * C; rstd += seg * C; dbeta += seg * C; dgamma += seg * C;
    }
    if (blockIdx.x == 0 && gbeta_acc != nullptr) {
        // fold this segment's dbeta / dgamma into the parameter-gradient buffers (the shared pose_encoder accumulates its two
        // calls; with several segments in one launch the blocks (0, seg) add concurrently, hence atomics)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            atomicAdd(gbeta_acc + c, dbeta[c]);
            atomicAdd(ggamma_acc + c, dgamma[c]);
        }
    }
    const float inv_n = 1.0f / (float)((long long)N * H * W);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = scale[c];
        const float k1 = -sc * rstd[c] * dgamma[c] * inv_n;
        sp[c] = sc;
        sp[C + c] = shift[c];
        sp[2 * C + c] = k1;
        sp[3 * C + c] = -sc * dbeta[c] * inv_n - k1 * mean[c];
    }
    __syncthreads();
    const int CG = C >> 3;
    const long long total = (long long)N * H * W * CG;
    if (!UPSAMPLE && (CG & (CG - 1)) == 0 && CG <= (int)blockDim.x) {
        // one channel group per thread for the whole loop (see bn_act_apply_kernel): the four coefficient vectors in
        // registers, two (dout, x) vector pairs in flight
        const int cg = threadIdx.x & (CG - 1);
        float csc[8], csh[8], ck1[8], ck0[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            csc[j] = sp[cg * 8 + j]; csh[j] = sp[C + cg * 8 + j]; ck1[j] = sp[2 * C + cg * 8 + j]; ck0[j] = sp[3 * C + cg * 8 + j];
        }
        const uint4* d4 = reinterpret_cast<const uint4*>(dout);
        const uint4* x4 = reinterpret_cast<const uint4*>(x);
        uint4* o4 = reinterpret_cast<uint4*>(dx);
        auto one = [&](const uint4& gu, const uint4& xu) {
            float g[8], xv[8], r[8];
            bf8_unpack(gu, g);
            bf8_unpack(xu, xv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float z = fmaf(xv[j], csc[j], csh[j]);
                const float gj = (relu && z <= 0.f) ? 0.f : g[j];
                r[j] = fmaf(csc[j], gj, fmaf(ck1[j], xv[j], ck0[j]));
            }
            return bf8_pack(r);
        };
        const long long stride = (long long)gridDim.x * blockDim.x;
        long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; idx + stride < total; idx += 2 * stride) {
            const uint4 g0 = d4[idx], x0 = x4[idx], g1 = d4[idx + stride], x1 = x4[idx + stride];
            o4[idx] = one(g0, x0);
            o4[idx + stride] = one(g1, x1);
        }
        if (idx < total) o4[idx] = one(d4[idx], x4[idx]);
        return;
    }
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % CG);
        float g[8], xv[8], r[8];
        if (UPSAMPLE) {
            const long long p = idx / CG;
            gather_dact<true>(dout, (int)(p / ((long long)W * H)), (int)((p / W) % H), (int)(p % W), cg, H, W, C, g);
        } else {
            bf8_unpack(reinterpret_cast<const uint4*>(dout)[idx], g);
        }
        bf8_unpack(reinterpret_cast<const uint4*>(x)[idx], xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cg * 8 + j;
            const float sc = sp[c];
            const float z = fmaf(xv[j], sc, sp[C + c]);
            const float gj = (relu && z <= 0.f) ? 0.f : g[j];
            r[j] = fmaf(sc, gj, fmaf(sp[2 * C + c], xv[j], sp[3 * C + c]));
        }
        reinterpret_cast<uint4*>(dx)[idx] = bf8_pack(r);
    }
}

// =============================================================================================
// activation-mask backward for bias+ReLU / bias+leaky layers (VGG, img_discr): g = dy * (y > 0 ? 1 : alpha)
// =============================================================================================
__global__ void act_mask_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y, float alpha,
                                    long long nvec, __nv_bfloat16* __restrict__ g) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        bf8_unpack(reinterpret_cast<const uint4*>(dy)[i], a);
        bf8_unpack(reinterpret_cast<const uint4*>(y)[i], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = b[j] > 0.f ? a[j] : alpha * a[j];
        reinterpret_cast<uint4*>(g)[i] = bf8_pack(a);
    }
}

// =============================================================================================
// max pool 2x2 stride 2 (even H, W)
// =============================================================================================
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                   __nv_bfloat16* __restrict__ out) {
    const int CG = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * Ho * Wo * CG;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % CG);
        long long pix = idx / CG;
        const int wo = (int)(pix % Wo);
        pix /= Wo;
        const int ho = (int)(pix % Ho);
        const int n = (int)(pix / Ho);
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float t[8];
                bf8_unpack(*reinterpret_cast<const uint4*>(x + (((long long)n * H + 2 * ho + dy) * W + 2 * wo + dx) * C + cg * 8), t);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], t[j]);
            }
        *reinterpret_cast<uint4*>(out + idx * 8) = bf8_pack(m);
    }
}
// dX[window] = dY routed to the first maximum of the window, times the ReLU mask of x (x > 0) when relu_mask.
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, int relu_mask,
                                   int N, int H, int W, int C, __nv_bfloat16* __restrict__ dx) {
    const int CG = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * Ho * Wo * CG;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % CG);
        long long pix = idx / CG;
        const int wo = (int)(pix % Wo);
        pix /= Wo;
        const int ho = (int)(pix % Ho);
        const int n = (int)(pix / Ho);
        float g[8], t[4][8];
        bf8_unpack(*reinterpret_cast<const uint4*>(dy + idx * 8), g);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            bf8_unpack(*reinterpret_cast<const uint4*>(x + (((long long)n * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1)) * C + cg * 8), t[k]);
        float o[4][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int best = 0;
            float bv = t[0][j];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (t[k][j] > bv) { bv = t[k][j]; best = k; }
            const float gj = ((relu_mask & 1) && bv <= 0.f) ? 0.f : g[j];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][j] = (k == best) ? gj : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint4* dst = reinterpret_cast<uint4*>(dx + (((long long)n * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1)) * C + cg * 8);
            if (relu_mask & 2) {  // accumulate into an existing gradient (a feature that also feeds the L1 loss)
                float old[8];
                bf8_unpack(*dst, old);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[k][j] += old[j];
            }
            *dst = bf8_pack(o[k]);
        }
    }
}

// =============================================================================================
// mask compose: heads f32 [P,4] = (crude rgb, mask in (0,1)), im f32 [P,3] -> final = im*m + crude*(1-m)
// =============================================================================================
__global__ void compose_fwd_kernel(const float* __restrict__ heads, const float* __restrict__ im, long long P, int clip,
                                   float* __restrict__ final_out, float* __restrict__ crude_out, float* __restrict__ mask_out) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const float4 hd = reinterpret_cast<const float4*>(heads)[p];
        const float m = hd.w;
        const float cr[3] = {hd.x, hd.y, hd.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float f = im[3 * p + c] * m + cr[c] * (1.f - m);
            float cc = cr[c];
            if (clip) {
                f = fminf(fmaxf(f, -1.f), 1.f);
                cc = fminf(fmaxf(cc, -1.f), 1.f);
            }
            final_out[3 * p + c] = f;
            if (crude_out) crude_out[3 * p + c] = cc;
        }
        if (mask_out) mask_out[p] = m;
    }
}
// d_final f32 [P,3] -> gradient w.r.t. the head PRE-activations, bf16 [P,16] (crude 3, mask logit 1, 12 zeros): 16 channels
// are one channel slot of the halo kernels (conv_halo2.cu data gradient, conv_wgrad2.cu); with 8 the heads' backward ran on
// the per-tap kernels at 0.8-1.1 TB/s
__global__ void compose_bwd_kernel(const float* __restrict__ d_final, const float* __restrict__ heads,
                                   const float* __restrict__ im, long long P, __nv_bfloat16* __restrict__ d_heads) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const float4 hd = reinterpret_cast<const float4*>(heads)[p];
        const float m = hd.w;
        const float cr[3] = {hd.x, hd.y, hd.z};
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float dm = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float g = d_final[3 * p + c];
            f[c] = g * (1.f - m);
            dm = fmaf(g, im[3 * p + c] - cr[c], dm);
        }
        f[3] = dm * m * (1.f - m);
        uint4* o = reinterpret_cast<uint4*>(d_heads + 16 * p);
        o[0] = bf8_pack(f);
        o[1] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// =============================================================================================
// channel pack / unpack (joint embedding): up to 3 sources (bf16 or f32) -> bf16 [P, Ctot] zero padded
// =============================================================================================
struct PackSrc {
    const void* ptr[3];
    int C[3];
    int is_f32[3];
    int n;
};
__global__ void pack_channels_kernel(PackSrc s, long long P, int Ctot, __nv_bfloat16* __restrict__ out) {
    const long long total = P * Ctot;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / Ctot;
        int c = (int)(idx - p * Ctot);
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i < s.n) {
                if (c >= 0 && c < s.C[i]) {
                    v = s.is_f32[i] ? reinterpret_cast<const float*>(s.ptr[i])[p * s.C[i] + c]
                                    : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s.ptr[i])[p * s.C[i] + c]);
                }
                c -= s.C[i];
            }
        }
        out[idx] = __float2bfloat16_rn(v);
    }
}
// 8-channel-group variants (every source a multiple of 8 channels, Ctot too): one 16-byte store per thread-item
__global__ void __launch_bounds__(256) pack_channels_vec_kernel(PackSrc s, long long P, int Ctot, __nv_bfloat16* __restrict__ out) {
    const int G = Ctot >> 3;
    const long long total = P * G;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / G;
        int c = (int)(idx - p * G) * 8;
        uint4 o = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i < s.n) {
                if (c >= 0 && c < s.C[i]) {
                    if (s.is_f32[i]) {
                        const float4* f = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(s.ptr[i]) + p * s.C[i] + c);
                        const float4 a = f[0], b = f[1];
                        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                        o = bf8_pack(v);
                    } else {
                        o = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(s.ptr[i]) + p * s.C[i] + c);
                    }
                }
                c -= s.C[i];
            }
        }
        *reinterpret_cast<uint4*>(out + idx * 8) = o;
    }
}
struct UnpackDst {
    void* ptr[3];
    int C[3];
    int is_f32[3];
    int n;
};
__global__ void unpack_channels_kernel(const __nv_bfloat16* __restrict__ g, long long P, int Ctot, UnpackDst d) {
    const long long total = P * Ctot;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / Ctot;
        int c = (int)(idx - p * Ctot);
        const __nv_bfloat16 v = g[idx];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i < d.n) {
                if (c >= 0 && c < d.C[i]) {
                    if (d.is_f32[i]) reinterpret_cast<float*>(d.ptr[i])[p * d.C[i] + c] = __bfloat162float(v);
                    else reinterpret_cast<__nv_bfloat16*>(d.ptr[i])[p * d.C[i] + c] = v;
                }
                c -= d.C[i];
            }
        }
    }
}

__global__ void __launch_bounds__(256) unpack_channels_vec_kernel(const __nv_bfloat16* __restrict__ g, long long P, int Ctot, UnpackDst d) {
    const int G = Ctot >> 3;
    const long long total = P * G;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / G;
        int c = (int)(idx - p * G) * 8;
        const uint4 v = *reinterpret_cast<const uint4*>(g + idx * 8);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (i < d.n) {
                if (c >= 0 && c < d.C[i]) {
                    if (d.is_f32[i]) {
                        float f[8];
                        bf8_unpack(v, f);
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(d.ptr[i]) + p * d.C[i] + c);
                        o[0] = make_float4(f[0], f[1], f[2], f[3]);
                        o[1] = make_float4(f[4], f[5], f[6], f[7]);
                    } else {
                        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.ptr[i]) + p * d.C[i] + c) = v;
                    }
                }
                c -= d.C[i];
            }
        }
    }
}

// =============================================================================================
// losses
// =============================================================================================
// L1 feature loss between the two halves of a bf16 feature tensor [2*half]: loss += weight * mean|gt - pred|,
// d_pred = weight/count * sign(pred - gt)   (bf16, may be nullptr)
__global__ void __launch_bounds__(256)
l1_pair_kernel(const __nv_bfloat16* __restrict__ feat_gt, const __nv_bfloat16* __restrict__ feat_pred, long long half_vec,
               float weight_over_count, float* __restrict__ loss, __nv_bfloat16* __restrict__ d_pred) {
    float acc = 0.f;
    const uint4* gt = reinterpret_cast<const uint4*>(feat_gt);
    const uint4* pr = reinterpret_cast<const uint4*>(feat_pred);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < half_vec; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8], g[8];
        bf8_unpack(gt[i], a);
        bf8_unpack(pr[i], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = b[j] - a[j];
            acc += fabsf(d);
            g[j] = d > 0.f ? weight_over_count : (d < 0.f ? -weight_over_count : 0.f);
        }
        if (d_pred) reinterpret_cast<uint4*>(d_pred)[i] = bf8_pack(g);
    }
    acc = warp_sum(acc);
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
        atomicAdd(loss, t * weight_over_count);
    }
}

// BCE with logits against a constant label z: loss += weight*mean(max(x,0) - x z + log1p(exp(-|x|)));
// d_logits bf16 [n,8] (channel 0 = weight/n * (sigmoid(x) - z), rest zero) for the D_logit conv backward.
__global__ void __launch_bounds__(256)
bce_logits_kernel(const float* __restrict__ x, int n, float z, float weight, float* __restrict__ loss,
                  __nv_bfloat16* __restrict__ d_logits) {
    float acc = 0.f;
    const float wn = weight / (float)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = x[i];
        acc += fmaxf(v, 0.f) - v * z + log1pf(expf(-fabsf(v)));
        if (d_logits) {
            float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            f[0] = wn * (1.f / (1.f + expf(-v)) - z);
            reinterpret_cast<uint4*>(d_logits)[i] = bf8_pack(f);
        }
    }
    acc = warp_sum(acc);
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
        atomicAdd(loss, t * wn);
    }
}

// =============================================================================================
// TF-style Adam over one flat f32 buffer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)
// =============================================================================================
__global__ void adam_tf_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                               float* __restrict__ v, long long n, float lr_t_host, const float* __restrict__ lr_t_dev,
                               float b1, float b2, float eps, float grad_scale) {
    // lr_t from device memory when given: lets a captured CUDA graph be replayed with a new step size
    const float lr_t = lr_t_dev != nullptr ? *lr_t_dev : lr_t_host;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
        m[i] = mi;
        v[i] = vi;
        p[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// per-channel sum over pixels of a bf16 [P,C] tensor (bias gradients of non-BN layers): out[c] += sum_p g[p,c]
// (SQ: sum of squares - the second batch-norm statistic of the stand-alone layers.batch_norm)
template <bool SQ>
__global__ void __launch_bounds__(256)
channel_sum_kernel(const __nv_bfloat16* __restrict__ g, long long P, int C, float* __restrict__ out) {
    const int CG = C >> 3;
    const int lanes = blockDim.x / CG;                       // threads beyond lanes*CG idle (C/8 need not divide 256)
    const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long p = (long long)blockIdx.x * lanes + pl; pl < lanes && p < P; p += (long long)gridDim.x * lanes) {
        float t[8];
        bf8_unpack(*reinterpret_cast<const uint4*>(g + p * C + cg * 8), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += SQ ? t[j] * t[j] : t[j];
    }
    // tree reduction (shuffles when the channel groups tile a warp, shared-memory atomics across warps, one global atomic
    // per channel per block) - the serial walk by pixel lane 0 cost 36 us per call for C = 8
    __shared__ float red[2048];
    for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const bool tiled = CG < 32 && (CG & (CG - 1)) == 0;
    if (tiled) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            for (int off = 16; off >= CG; off >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
    }
    if (pl < lanes && (!tiled || (threadIdx.x & 31) < CG)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&red[cg * 8 + j], s[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(out + c, red[c]);
}

// =============================================================================================
// launchers
// =============================================================================================
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int ew_image_prep(const float* x, long long P, const float* a, const float* b, const int* perm, void* out, cudaStream_t st) {
    ImgPrep q;
    for (int c = 0; c < 3; ++c) { q.a[c] = a[c]; q.b[c] = b[c]; q.perm[c] = perm[c]; }
    image_prep_kernel<<<grid_for(P, 256), 256, 0, st>>>(x, P, q, reinterpret_cast<__nv_bfloat16*>(out));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_image_prep_bwd(const void* g, long long P, const float* a, const int* perm, int accumulate, float* dx, cudaStream_t st) {
    ImgPrep q;
    for (int c = 0; c < 3; ++c) { q.a[c] = a[c]; q.b[c] = 0.f; q.perm[c] = perm[c]; }
    image_prep_bwd_kernel<<<grid_for(P, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), P, q, accumulate, dx);
    KP_LAUNCHED();
    return KP_OK;
}
int ew_bn_finalize(const float* ssum, const float* ssq, const float* bias, const float* gamma, const float* beta, int C,
                   double count, float eps, float decay, float* mm, float* mv, float* scale, float* shift, float* smean,
                   float* srstd, cudaStream_t st) {
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(ssum, ssq, bias, gamma, beta, C, (float)count, eps, decay, mm, mv,
                                                        scale, shift, smean, srstd);
    KP_LAUNCHED();
    return KP_OK;
}
// blocks of `threads` threads that fit the whole GPU at once for `kern` (one wave: the per-block prologue - statistics ->
// scale / shift in double precision - is then paid once per SM slot instead of once per 28 KB of data)
template <class K>
static int resident_blocks(K kern, int threads, size_t smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 4;
    return per_sm * 148;
}

static int bn_apply_launch(const void* x, const float* scale, const float* shift, const BnFin& fin, int relu, int upsample, int N,
                           int H, int W, int C, void* out, cudaStream_t st, int groups = 1) {
    KP_REQUIRE(C % 8 == 0, "bn_act_apply: C=%d must be a multiple of 8", C);
    KP_REQUIRE(C <= 4096, "bn_act_apply: C=%d too large", C);
    KP_REQUIRE(groups >= 1 && N % groups == 0, "bn_act_apply: %d images do not split into %d segments", N, groups);
    N /= groups;                                                 // images per segment; segments on grid.y
    const long long total = (long long)N * H * W * (C / 8);     // input vectors (one thread-item each)
    const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    const size_t smem = 2 * (size_t)C * sizeof(float);
    static int wave_up = 0, wave_plain = 0;     // (depends on the dynamic shared memory only weakly: C <= 512 floats x 2)
    if (wave_up == 0) wave_up = resident_blocks(bn_act_apply_kernel<true>, 256, 4096);
    if (wave_plain == 0) wave_plain = resident_blocks(bn_act_apply_kernel<false>, 256, 4096);
    if (upsample) KP_CUDA_CHECK(launch_pdl(bn_act_apply_kernel<true>, dim3(grid_for(total, 256, wave_up / groups), groups), dim3(256), smem, st, xi, scale, shift, fin, relu, N, H, W, C, o));
    else KP_CUDA_CHECK(launch_pdl(bn_act_apply_kernel<false>, dim3(grid_for((total + 3) / 4, 256, wave_plain / groups), groups), dim3(256), smem, st, xi, scale, shift, fin, relu, N, H, W, C, o));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_bn_act_apply(const void* x, const float* scale, const float* shift, int relu, int upsample, int N, int H, int W, int C,
                    void* out, cudaStream_t st) {
    BnFin fin;
    memset(&fin, 0, sizeof(fin));
    return bn_apply_launch(x, scale, shift, fin, relu, upsample, N, H, W, C, out, st);
}
int ew_bn_stats_apply(const float* ssum, const float* ssq, const float* bias, const float* gamma, const float* beta, double count,
                      float eps, float decay, float* mm, float* mv, float* scale, float* shift, float* smean, float* srstd,
                      const void* x, int relu, int upsample, int N, int H, int W, int C, void* out, int groups, cudaStream_t st) {
    BnFin fin = {ssum, ssq, bias, gamma, beta, (float)count, eps, decay, mm, mv, scale, shift, smean, srstd};
    return bn_apply_launch(x, nullptr, nullptr, fin, relu, upsample, N, H, W, C, out, st, groups);
}
int ew_bn_act_bwd(const void* dout, const void* x, const float* scale, const float* shift, const float* mean,
                  const float* rstd, int relu, int upsample, int N, int H, int W, int C, float* dbeta, float* dgamma,
                  void* dx, float* gbeta_acc, float* ggamma_acc, int prezeroed, int groups, cudaStream_t st) {
    KP_REQUIRE(C % 8 == 0 && pow2(C / 8) && C / 8 <= 256, "bn_act_bwd: C=%d must be 8 x a power of two <= 2048", C);
    KP_REQUIRE(groups >= 1 && N % groups == 0, "bn_act_bwd: %d images do not split into %d segments", N, groups);
    N /= groups;                                                 // images per segment; segments on grid.y
    const __nv_bfloat16* d = reinterpret_cast<const __nv_bfloat16*>(dout);
    const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
    if (!prezeroed) {
        KP_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, (size_t)groups * C * sizeof(float), st));
        KP_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, (size_t)groups * C * sizeof(float), st));
    }
    const long long P = (long long)N * H * W;
    const int lanes = 256 / (C / 8);
    // one wave of 256-thread blocks, 8 vectors in flight per thread; fewer blocks = fewer contended global atomics at the end
    static int wave_red = 0, wave_red_up = 0, wave_app = 0, wave_app_up = 0;
    if (wave_red == 0) {
        wave_red = resident_blocks(bn_act_bwd_reduce_kernel<false>, 256, 0);
        wave_red_up = resident_blocks(bn_act_bwd_reduce_kernel<true>, 256, 0);
        wave_app = resident_blocks(bn_act_bwd_apply_kernel<false>, 256, 8192);
        wave_app_up = resident_blocks(bn_act_bwd_apply_kernel<true>, 256, 8192);
    }
    const int rgrid = grid_for((P + 4 * lanes - 1) / (4 * lanes), 1, (upsample ? wave_red_up : wave_red) / groups);
    if (upsample)
        KP_CUDA_CHECK(launch_pdl(bn_act_bwd_reduce_kernel<true>, dim3(rgrid, groups), dim3(256), 0, st, d, xi, scale, shift, mean, rstd, relu, N, H, W, C, dbeta, dgamma));
    else
        KP_CUDA_CHECK(launch_pdl(bn_act_bwd_reduce_kernel<false>, dim3(rgrid, groups), dim3(256), 0, st, d, xi, scale, shift, mean, rstd, relu, N, H, W, C, dbeta, dgamma));
    KP_LAUNCHED();
    const long long total = P * (C / 8);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dx);
    if (upsample)
        KP_CUDA_CHECK(launch_pdl(bn_act_bwd_apply_kernel<true>, dim3(grid_for(total, 256, wave_app_up / groups), groups), dim3(256), 4 * (size_t)C * sizeof(float), st, d, xi, scale, shift, mean, rstd, dbeta, dgamma, relu, N, H, W, C, o, gbeta_acc, ggamma_acc));
    else
        KP_CUDA_CHECK(launch_pdl(bn_act_bwd_apply_kernel<false>, dim3(grid_for((total + 1) / 2, 256, wave_app / groups), groups), dim3(256), 4 * (size_t)C * sizeof(float), st, d, xi, scale, shift, mean, rstd, dbeta, dgamma, relu, N, H, W, C, o, gbeta_acc, ggamma_acc));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_upsample2x_bwd(const void* dout, int N, int H, int W, int C, void* dact, cudaStream_t st) {
    KP_REQUIRE(C % 8 == 0, "upsample2x_bwd: C=%d must be a multiple of 8", C);
    const long long total = (long long)N * H * W * (C / 8);
    upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dout), N, H, W, C,
                                                                  reinterpret_cast<__nv_bfloat16*>(dact));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_act_mask_bwd(const void* dy, const void* y, float alpha, long long n_elems, void* g, cudaStream_t st) {
    KP_REQUIRE(n_elems % 8 == 0, "act_mask_bwd: element count must be a multiple of 8");
    act_mask_bwd_kernel<<<grid_for(n_elems / 8, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy),
                                                                     reinterpret_cast<const __nv_bfloat16*>(y), alpha,
                                                                     n_elems / 8, reinterpret_cast<__nv_bfloat16*>(g));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_maxpool_fwd(const void* x, int N, int H, int W, int C, void* out, cudaStream_t st) {
    KP_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: needs C%%8==0 and even H,W");
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), N, H, W, C,
                                                             reinterpret_cast<__nv_bfloat16*>(out));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_maxpool_bwd(const void* dy, const void* x, int relu_mask, int N, int H, int W, int C, void* dx, cudaStream_t st) {
    KP_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: needs C%%8==0 and even H,W");
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy),
                                                             reinterpret_cast<const __nv_bfloat16*>(x), relu_mask, N, H, W, C,
                                                             reinterpret_cast<__nv_bfloat16*>(dx));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_compose_fwd(const float* heads, const float* im, long long P, int clip, float* final_out, float* crude_out,
                   float* mask_out, cudaStream_t st) {
    compose_fwd_kernel<<<grid_for(P, 256), 256, 0, st>>>(heads, im, P, clip, final_out, crude_out, mask_out);
    KP_LAUNCHED();
    return KP_OK;
}
int ew_compose_bwd(const float* d_final, const float* heads, const float* im, long long P, void* d_heads, cudaStream_t st) {
    compose_bwd_kernel<<<grid_for(P, 256), 256, 0, st>>>(d_final, heads, im, P, reinterpret_cast<__nv_bfloat16*>(d_heads));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_pack_channels(const void* const* src, const int* C, const int* is_f32, int n, long long P, int Ctot, void* out,
                     cudaStream_t st) {
    KP_REQUIRE(n >= 1 && n <= 3, "pack_channels: 1..3 sources");
    PackSrc s;
    s.n = n;
    int sum = 0;
    for (int i = 0; i < 3; ++i) {
        s.ptr[i] = i < n ? src[i] : nullptr;
        s.C[i] = i < n ? C[i] : 0;
        s.is_f32[i] = i < n ? is_f32[i] : 0;
        sum += s.C[i];
    }
    KP_REQUIRE(sum <= Ctot, "pack_channels: sources (%d channels) exceed Ctot=%d", sum, Ctot);
    bool vec = Ctot % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int i = 0; i < n; ++i) vec = vec && C[i] % 8 == 0 && (reinterpret_cast<uintptr_t>(src[i]) & 15) == 0;
    if (vec) pack_channels_vec_kernel<<<grid_for(P * (Ctot / 8), 256), 256, 0, st>>>(s, P, Ctot, reinterpret_cast<__nv_bfloat16*>(out));
    else pack_channels_kernel<<<grid_for(P * Ctot, 256), 256, 0, st>>>(s, P, Ctot, reinterpret_cast<__nv_bfloat16*>(out));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_unpack_channels(const void* g, long long P, int Ctot, void* const* dst, const int* C, const int* is_f32, int n,
                       cudaStream_t st) {
    KP_REQUIRE(n >= 1 && n <= 3, "unpack_channels: 1..3 destinations");
    UnpackDst d;
    d.n = n;
    for (int i = 0; i < 3; ++i) {
        d.ptr[i] = i < n ? dst[i] : nullptr;
        d.C[i] = i < n ? C[i] : 0;
        d.is_f32[i] = i < n ? is_f32[i] : 0;
    }
    bool vec = Ctot % 8 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    for (int i = 0; i < n; ++i) vec = vec && C[i] % 8 == 0 && (reinterpret_cast<uintptr_t>(dst[i]) & 15) == 0;
    if (vec) unpack_channels_vec_kernel<<<grid_for(P * (Ctot / 8), 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), P, Ctot, d);
    else unpack_channels_kernel<<<grid_for(P * Ctot, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), P, Ctot, d);
    KP_LAUNCHED();
    return KP_OK;
}
int ew_l1_pair(const void* feat_gt, const void* feat_pred, long long half_elems, float weight, float* loss, void* d_pred,
               cudaStream_t st) {
    KP_REQUIRE(half_elems % 8 == 0, "l1_pair: element count must be a multiple of 8");
    l1_pair_kernel<<<grid_for(half_elems / 8, 256, 148 * 8), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(feat_gt),
                                                                           reinterpret_cast<const __nv_bfloat16*>(feat_pred),
                                                                           half_elems / 8, weight / (float)half_elems, loss,
                                                                           reinterpret_cast<__nv_bfloat16*>(d_pred));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_bce_logits(const float* x, int n, float z, float weight, float* loss, void* d_logits, cudaStream_t st) {
    bce_logits_kernel<<<1, 256, 0, st>>>(x, n, z, weight, loss, reinterpret_cast<__nv_bfloat16*>(d_logits));
    KP_LAUNCHED();
    return KP_OK;
}
int ew_adam_tf(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, int t,
               float grad_scale, const float* lr_t_dev, cudaStream_t st) {
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t));
    adam_tf_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, g, m, v, n, (float)lr_t, lr_t_dev, b1, b2, eps, grad_scale);
    KP_LAUNCHED();
    return KP_OK;
}
int ew_channel_sum(const void* g, long long P, int C, float* out, int squares, cudaStream_t st) {
    KP_REQUIRE(C % 8 == 0 && C / 8 <= 256, "channel_sum: C=%d must be a multiple of 8, at most 2048", C);
    const int lanes = 256 / (C / 8);
    const int grid = grid_for((P + lanes - 1) / lanes, 1, 148 * 4);
    if (squares)
        channel_sum_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), P, C, out);
    else
        channel_sum_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), P, C, out);
    KP_LAUNCHED();
    return KP_OK;
}

}  // namespace kp

// =============================================================================================
// weight re-layout: fp32 HWIO kernel -> bf16 K-major GEMM operand (one launch per plan; pure data movement + cast)
//   fwd  : dst[co][t*Kper + kc]  = W[tap_flat[t]][ci(kc)][co] * row_scale[co]   (concat segments padded to CB)
//   dgrad: dst[ci-c0][t*Kper + co] = W[tap_flat[t]][ci][co]                      (co padded to CB)
// =============================================================================================
namespace kp {

__global__ void __launch_bounds__(256)
pack_weights_fwd_kernel(const float* __restrict__ w, kp_pack_desc d, const float* __restrict__ row_scale,
                        __nv_bfloat16* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int t = blockIdx.z;
    const int kc0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const float* wt = w + (long long)d.tap_flat[t] * d.cin * d.cout;
    // load: rows = kc (source channel), cols = co (contiguous in HWIO)
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int kc = kc0 + i, co = r0 + threadIdx.x;
        int ci = -1;
#pragma unroll
        for (int s = 0; s < 3; ++s)
            if (s < d.nseg && kc >= d.seg_kbase[s] && kc - d.seg_kbase[s] < d.seg_count[s]) ci = d.seg_start[s] + kc - d.seg_kbase[s];
        tile[i][threadIdx.x] = (ci >= 0 && ci < d.cin && co < d.cout) ? wt[(long long)ci * d.cout + co] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int co = r0 + i, kc = kc0 + threadIdx.x;
        if (co < d.rows_pad && kc < d.Kper) {
            float v = tile[threadIdx.x][i];
            if (row_scale != nullptr && co < d.cout) v *= row_scale[co];
            dst[(long long)co * d.Ktot + (long long)t * d.Kper + kc] = __float2bfloat16_rn(v);
        }
    }
}

__global__ void __launch_bounds__(256)
pack_weights_dgrad_kernel(const float* __restrict__ w, kp_pack_desc d, __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)d.rows_pad * d.Ktot;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / d.Ktot);
        const int k = (int)(idx - (long long)r * d.Ktot);
        const int t = k / d.Kper, kc = k - t * d.Kper;
        float v = 0.f;
        if (r < d.rows && d.c0 + r < d.cin && kc < d.cout) v = w[((long long)d.tap_flat[t] * d.cin + d.c0 + r) * d.cout + kc];
        dst[idx] = __float2bfloat16_rn(v);
    }
}

// ---- batched variant: one launch for a whole table of jobs (device memory) ----
// fwd mode: 64 (K) x 32 (rows) tiles through shared memory, 128-byte coalesced fp32 reads along cout and bf16x2 stores
// along K; dgrad mode: no transposition, 8 consecutive K elements (= cout) per thread as 2 x float4 -> one 16-byte store.
constexpr int PACK_DGRAD_PER_BLOCK = 256 * 8 * 4;

__device__ __forceinline__ void pack_fwd_tile(const float* __restrict__ w, const kp_pack_desc& d, __nv_bfloat16* __restrict__ dst,
                                              int bx, int by, int t, float (*tile)[33]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int kc0 = bx * 64, r0 = by * 32;
    const float* wt = w + (long long)d.tap_flat[t] * d.cin * d.cout;
    for (int i = ty; i < 64; i += 8) {
        const int kc = kc0 + i, co = r0 + tx;
        int ci = -1;
#pragma unroll
        for (int s = 0; s < 3; ++s)
            if (s < d.nseg && kc >= d.seg_kbase[s] && kc - d.seg_kbase[s] < d.seg_count[s]) ci = d.seg_start[s] + kc - d.seg_kbase[s];
        tile[i][tx] = (ci >= 0 && ci < d.cin && co < d.cout) ? wt[(long long)ci * d.cout + co] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int co = r0 + i, kc = kc0 + 2 * tx;             // Kper is a multiple of 16: kc and kc+1 are valid together
        if (co < d.rows_pad && kc < d.Kper) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(tile[2 * tx][i], tile[2 * tx + 1][i]);
            *reinterpret_cast<__nv_bfloat162*>(dst + (long long)co * d.Ktot + (long long)t * d.Kper + kc) = h;
        }
    }
}

__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const kp_pack_job* __restrict__ jobs, int n_jobs) {
    __shared__ float tile[64][33];
    __shared__ kp_pack_job job;
    __shared__ int job_idx;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_jobs - 1;                 // last job with block_begin <= blockIdx.x
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        job_idx = lo;
    }
    __syncthreads();
    {
        const int* src = reinterpret_cast<const int*>(jobs + job_idx);
        int* dstw = reinterpret_cast<int*>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(kp_pack_job) / sizeof(int)); i += blockDim.x) dstw[i] = src[i];
    }
    __syncthreads();
    const kp_pack_desc& d = job.d;
    const int b = (int)blockIdx.x - job.block_begin;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(job.dst);
    if (d.mode == 0) {
        const int gx = (d.Kper + 63) / 64, gy = (d.rows_pad + 31) / 32;
        const int t = b / (gx * gy), rem = b - t * gx * gy;
        pack_fwd_tile(job.w, d, dst, rem % gx, rem / gx, t, tile);
    } else {
        const long long total = (long long)d.rows_pad * d.Ktot;
        const long long base = (long long)b * PACK_DGRAD_PER_BLOCK;
        const bool vec = (d.cout & 7) == 0 && (reinterpret_cast<uintptr_t>(job.w) & 15) == 0;   // K groups of 8 never straddle cout
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long idx = base + ((long long)i * 256 + threadIdx.x) * 8;
            if (idx >= total) break;
            const int r = (int)(idx / d.Ktot);
            const int k = (int)(idx - (long long)r * d.Ktot);
            const int t = k / d.Kper, kc = k - t * d.Kper;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (r < d.rows && d.c0 + r < d.cin) {
                const float* src = job.w + ((long long)d.tap_flat[t] * d.cin + d.c0 + r) * d.cout + kc;
                if (vec) {
                    if (kc < d.cout) {
                        const float4 a = *reinterpret_cast<const float4*>(src), c = *reinterpret_cast<const float4*>(src + 4);
                        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (kc + j < d.cout) v[j] = src[j];
                }
            }
            *reinterpret_cast<uint4*>(dst + idx) = bf8_pack(v);
        }
    }
}

int ew_pack_job_blocks(const kp_pack_desc* d) {
    if (d->mode == 0) return ((d->Kper + 63) / 64) * ((d->rows_pad + 31) / 32) * d->T;
    const long long total = (long long)d->rows_pad * d->Ktot;
    return (int)((total + PACK_DGRAD_PER_BLOCK - 1) / PACK_DGRAD_PER_BLOCK);
}

int ew_pack_weights_batch(const void* jobs_dev, int n_jobs, int total_blocks, cudaStream_t st) {
    KP_REQUIRE(n_jobs > 0 && total_blocks > 0, "pack_weights_batch: empty table");
    static_assert(sizeof(kp_pack_job) % sizeof(int) == 0, "kp_pack_job must be int-copyable");
    pack_weights_batch_kernel<<<total_blocks, 256, 0, st>>>(reinterpret_cast<const kp_pack_job*>(jobs_dev), n_jobs);
    KP_LAUNCHED();
    return KP_OK;
}

int ew_pack_weights(const float* w, const kp_pack_desc* d, const float* row_scale, void* dst, cudaStream_t st) {
    KP_REQUIRE(d->T >= 1 && d->T <= KP_MAX_TAPS && d->Kper > 0 && d->Ktot == d->T * d->Kper, "pack_weights: bad descriptor");
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dst);
    if (d->mode == 0) {
        KP_REQUIRE(d->nseg >= 1 && d->nseg <= 3, "pack_weights: 1..3 segments");
        dim3 grid((unsigned)((d->Kper + 31) / 32), (unsigned)((d->rows_pad + 31) / 32), (unsigned)d->T);
        pack_weights_fwd_kernel<<<grid, dim3(32, 8), 0, st>>>(w, *d, row_scale, o);
    } else {
        const long long total = (long long)d->rows_pad * d->Ktot;
        pack_weights_dgrad_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, *d, o);
    }
    KP_LAUNCHED();
    return KP_OK;
}

}  // namespace kp

// =============================================================================================
// W-unrolled image packing for 3-channel first layers (encoder conv_1 7x7, VGG conv1_1 3x3):
//   out[n,h,w, kw*3 + c] = a[c]*x[n,h,w+kw-pl,perm[c]] + b[c]   (0 outside the image; channels KW*3.. zero)
// A KHxKW convolution over 3 channels becomes a KHx1 convolution over KW*3 channels of this tensor: KH taps of
// 32-64 bytes instead of KH*KW taps of 6 bytes, with the HWIO kernel [KH][KW][3][Cout] reinterpreted in place as
// [KH][1][KW*3][Cout].  The zero fill happens AFTER the affine map (TF pads the already-preprocessed image).
// =============================================================================================
namespace kp {

template <int CPAD>
__global__ void __launch_bounds__(256)
image_prep_unrolled_kernel(const float* __restrict__ x, int N, int H, int W, int KW, int pl, ImgPrep q,
                           __nv_bfloat16* __restrict__ out) {
    const long long P = (long long)N * H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(p % W);
        const long long row = p - w;     // first pixel of this image row
        float f[CPAD];
#pragma unroll
        for (int i = 0; i < CPAD; ++i) f[i] = 0.f;
#pragma unroll
        for (int kw = 0; kw < CPAD / 3; ++kw) {
            if (kw < KW) {
                const int ws = w + kw - pl;
                if (ws >= 0 && ws < W) {
                    const float* px = x + 3 * (row + ws);
                    const float in[3] = {px[0], px[1], px[2]};
#pragma unroll
                    for (int c = 0; c < 3; ++c) f[kw * 3 + c] = fmaf(q.a[c], in[q.perm[c]], q.b[c]);
                }
            }
        }
        uint4* o = reinterpret_cast<uint4*>(out + CPAD * p);
#pragma unroll
        for (int v = 0; v < CPAD / 8; ++v) o[v] = bf8_pack(f + 8 * v);
    }
}

// adjoint: g bf16 [N,H,W,CPAD] -> dx f32 [N,H,W,3]: dx[n,h,w',perm[c]] (+)= a[c] * sum_kw g[n,h,w'-kw+pl, kw*3+c]
template <int CPAD>
__global__ void __launch_bounds__(256)
image_prep_unrolled_bwd_kernel(const __nv_bfloat16* __restrict__ g, int N, int H, int W, int KW, int pl, ImgPrep q,
                               int accumulate, float* __restrict__ dx) {
    const long long P = (long long)N * H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(p % W);
        const long long row = p - w;
        float acc[3] = {0.f, 0.f, 0.f};
        for (int kw = 0; kw < KW; ++kw) {
            const int wd = w - kw + pl;
            if (wd < 0 || wd >= W) continue;
            const __nv_bfloat16* gp = g + CPAD * (row + wd) + kw * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] += __bfloat162float(gp[c]);
        }
        float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) o[q.perm[c]] = q.a[c] * acc[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) dx[3 * p + c] = accumulate ? dx[3 * p + c] + o[c] : o[c];
    }
}

int ew_image_prep_unrolled(const float* x, int N, int H, int W, int KW, int pl, int cpad, const float* a, const float* b,
                           const int* perm, void* out, cudaStream_t st) {
    KP_REQUIRE(KW >= 1 && KW * 3 <= cpad && (cpad == 16 || cpad == 32), "image_prep_unrolled: KW*3 must fit Cpad in {16,32}");
    ImgPrep q;
    for (int c = 0; c < 3; ++c) { q.a[c] = a[c]; q.b[c] = b[c]; q.perm[c] = perm[c]; }
    const long long P = (long long)N * H * W;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (cpad == 16) image_prep_unrolled_kernel<16><<<grid_for(P, 256), 256, 0, st>>>(x, N, H, W, KW, pl, q, o);
    else image_prep_unrolled_kernel<32><<<grid_for(P, 256), 256, 0, st>>>(x, N, H, W, KW, pl, q, o);
    KP_LAUNCHED();
    return KP_OK;
}

int ew_image_prep_unrolled_bwd(const void* g, int N, int H, int W, int KW, int pl, int cpad, const float* a, const int* perm,
                               int accumulate, float* dx, cudaStream_t st) {
    KP_REQUIRE(KW >= 1 && KW * 3 <= cpad && (cpad == 16 || cpad == 32), "image_prep_unrolled_bwd: KW*3 must fit Cpad in {16,32}");
    ImgPrep q;
    for (int c = 0; c < 3; ++c) { q.a[c] = a[c]; q.b[c] = 0.f; q.perm[c] = perm[c]; }
    const long long P = (long long)N * H * W;
    const __nv_bfloat16* gi = reinterpret_cast<const __nv_bfloat16*>(g);
    if (cpad == 16) image_prep_unrolled_bwd_kernel<16><<<grid_for(P, 256), 256, 0, st>>>(gi, N, H, W, KW, pl, q, accumulate, dx);
    else image_prep_unrolled_bwd_kernel<32><<<grid_for(P, 256), 256, 0, st>>>(gi, N, H, W, KW, pl, q, accumulate, dx);
    KP_LAUNCHED();
    return KP_OK;
}

}  // namespace kp
