// K1: fused marginal soft-argmax + Gaussian heat-map rendering (forward and backward) for sm_100a.
//
// Replaces the reference's get_coord x2 + stack + get_gaussian_maps chain
// (/root/reference/utils/model.py:49-70, call sites models/networks/__init__.py:68-71 and
//  models/detector_translator_model.py:168-169) with ONE pass over the fp32 logits.
//
// Data layout in HBM (all fp32, channels-last exactly like the reference's NHWC tensors):
//   logits [B,H,W,K]   mu [B,K,2] (x,y)   prob_x [B,W,K]   prob_y [B,H,K]   maps [B,hm,wm,K]
//
// Fast path (K % 8 == 0, W % 16 == 0): one CTA per frame, 2 CTAs per SM.
//   * a producer warp streams the frame row by row (W*K*4 bytes, 20 KB at penn.yaml shapes) into a
//     3-stage shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * K/8 consumer warps read each row once from shared memory (bank-conflict-free float4 mapping),
//     keep the column sums in registers, reduce the row sums with warp shuffles;
//   * 2K warp-level softmaxes + expectations give mu; separable exp tables give the maps, which are
//     written with coalesced float4 stores.
// Algorithmic HBM bytes per frame: H*W*K*4 read + hm*wm*K*4 + K*8 written (2 785 600 B at penn shapes).
#include "kp_common.cuh"
#include "kp_internal.h"
#include <math.h>

namespace kp {

static constexpr int K1_NSTAGE = 3;

__device__ __forceinline__ float lin_coord(int i, float step) { return fmaf((float)i, step, -1.0f); }
__host__ __device__ __forceinline__ float lin_step(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }

// =============================================================================================
// Forward, fast path
// =============================================================================================
template <int KG2, int NJ>
__global__ void __launch_bounds__((KG2 + 1) * 32, 2)
k1_fwd_fast(const float* __restrict__ logits, int H, float* __restrict__ mu, float* __restrict__ prob_x,
            float* __restrict__ prob_y, float* __restrict__ maps, int hm, int wm, float neg_s) {
    constexpr int K = 8 * KG2, W = 16 * NJ, KG = 2 * KG2, NCW = KG2, NCT = NCW * 32;
    constexpr uint32_t ROW_BYTES = W * K * 4;
    const int HP = H + 4, WP = W + 4;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    float* rT = ring + K1_NSTAGE * W * K;  // [K][H+4] row sums (later prob_y, later g_y table)
    float* qT = rT + K * HP;               // [K][W+4] col sums (later prob_x, later g_x table)
    float* mus = qT + K * WP;              // [2][K]: mu_x then mu_y
    uint64_t* full = reinterpret_cast<uint64_t*>(mus + 2 * K);
    uint64_t* empty = full + K1_NSTAGE;

    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src = logits + (size_t)b * H * W * K;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < K1_NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ---------------- TMA producer: one lane streams H rows through the ring ----------------
        if (lane == 0) {
            for (int h = 0; h < H; ++h) {
                const int s = h % K1_NSTAGE;
                const uint32_t it = h / K1_NSTAGE;
                if (h >= K1_NSTAGE) mbar_wait(&empty[s], (it - 1) & 1);
                mbar_arrive_expect_tx(&full[s], ROW_BYTES);
                bulk_g2s(ring + s * (W * K), src + (size_t)h * (W * K), ROW_BYTES, &full[s]);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    // lane -> (half, wl): the 8 lanes of a quarter-warp touch 8 distinct 16-byte bank groups.
    const int half = lane & 1, wl = lane >> 1, g = 2 * warp + half;
    float4 acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* ring4 = reinterpret_cast<const float4*>(ring);

    for (int h = 0; h < H; ++h) {
        const int s = h % K1_NSTAGE;
        const uint32_t it = h / K1_NSTAGE;
        mbar_wait(&full[s], it & 1);
        const float4* row = ring4 + s * (W * KG);
        float4 v[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) v[j] = row[(wl + 16 * j) * KG + g];
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[j] = f4_add(acc[j], v[j]);
        // pairwise row sum over this lane's NJ pixels
#pragma unroll
        for (int st = 1; st < NJ; st <<= 1) {
#pragma unroll
            for (int j = 0; j + st < NJ; j += 2 * st) v[j] = f4_add(v[j], v[j + st]);
        }
        float4 rs = v[0];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);  // all lanes' reads of this slot are done
#pragma unroll
        for (int m = 2; m < 32; m <<= 1) rs = f4_add(rs, f4_shfl_xor(rs, m));
        if (wl == 0) {
            rT[(4 * g + 0) * HP + h] = rs.x;
            rT[(4 * g + 1) * HP + h] = rs.y;
            rT[(4 * g + 2) * HP + h] = rs.z;
            rT[(4 * g + 3) * HP + h] = rs.w;
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int w = wl + 16 * j;
        qT[(4 * g + 0) * WP + w] = acc[j].x;
        qT[(4 * g + 1) * WP + w] = acc[j].y;
        qT[(4 * g + 2) * WP + w] = acc[j].z;
        qT[(4 * g + 3) * WP + w] = acc[j].w;
    }
    named_bar_sync(1, NCT);

    // ---------------- 2K softmax + expectation tasks, one warp each ----------------
    for (int t = warp; t < 2 * K; t += NCW) {
        const int axis = t / K, k = t - axis * K;          // axis 0 -> x (over W), 1 -> y (over H)
        const int N = axis ? H : W, NP = axis ? HP : WP;
        float* T = (axis ? rT : qT) + k * NP;
        const float inv = 1.0f / (float)(axis ? W : H);    // reduce_mean over the other axis
        const float step = lin_step(N);
        float e[8];
        float vmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = lane + 32 * i;
            e[i] = (n < N) ? T[n] * inv : -INFINITY;
            vmax = fmaxf(vmax, e[i]);
        }
        vmax = warp_max(vmax);
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = lane + 32 * i;
            e[i] = (n < N) ? expf(e[i] - vmax) : 0.f;
            se += e[i];
        }
        se = warp_sum(se);
        float sc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = lane + 32 * i;
            if (n < N) {
                const float p = e[i] / se;
                T[n] = p;
                sc = fmaf(p, lin_coord(n, step), sc);
            }
        }
        sc = warp_sum(sc);
        if (lane == 0) {
            mus[axis * K + k] = sc;
            mu[((size_t)b * K + k) * 2 + axis] = sc;
        }
    }
    named_bar_sync(1, NCT);

    const int tid = threadIdx.x;
    if (prob_x != nullptr) {
        float* px = prob_x + (size_t)b * W * K;
        for (int idx = tid; idx < W * K; idx += NCT) px[idx] = qT[(idx % K) * WP + idx / K];
    }
    if (prob_y != nullptr) {
        float* py = prob_y + (size_t)b * H * K;
        for (int idx = tid; idx < H * K; idx += NCT) py[idx] = rT[(idx % K) * HP + idx / K];
    }
    if (maps == nullptr) return;
    named_bar_sync(1, NCT);

    // ---------------- separable Gaussian tables (alias the marginal buffers) ----------------
    float* gy = rT;  // [hm][K]
    float* gx = qT;  // [wm][K]
    {
        const float sy = lin_step(hm), sx = lin_step(wm);
        for (int idx = tid; idx < hm * K; idx += NCT) {
            const int i = idx / K, k = idx - i * K;
            const float d = lin_coord(i, sy) - mus[K + k];
            gy[idx] = expf(neg_s * d * d);
        }
        for (int idx = tid; idx < wm * K; idx += NCT) {
            const int j = idx / K, k = idx - j * K;
            const float d = lin_coord(j, sx) - mus[k];
            gx[idx] = expf(neg_s * d * d);
        }
    }
    named_bar_sync(1, NCT);

    // ---------------- render: coalesced float4 stores, thread has a fixed k-group ----------------
    {
        const int tg = tid % KG, p0 = tid / KG;  // NCT == 16*KG
        const float4* gy4 = reinterpret_cast<const float4*>(gy);
        const float4* gx4 = reinterpret_cast<const float4*>(gx);
        float4* out4 = reinterpret_cast<float4*>(maps + (size_t)b * hm * wm * K);
        const int npix = hm * wm;
        for (int pix = p0; pix < npix; pix += 16) {
            const int i = pix / wm, j = pix - i * wm;
            const float4 a = gy4[i * KG + tg], c = gx4[j * KG + tg];
            out4[pix * KG + tg] = make_float4(a.x * c.x, a.y * c.y, a.z * c.z, a.w * c.w);
        }
    }
}

// =============================================================================================
// Forward, generic path (any H, W, K): correctness path for odd shapes. One CTA per frame.
// =============================================================================================
__global__ void __launch_bounds__(256)
k1_fwd_generic(const float* __restrict__ logits, int H, int W, int K, float* __restrict__ mu,
               float* __restrict__ prob_x, float* __restrict__ prob_y) {
    extern __shared__ float sm[];
    float* q = sm;          // [W][K]
    float* r = q + W * K;   // [H][K]
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float* src = logits + (size_t)b * H * W * K;
    for (int e = tid; e < W * K; e += NT) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += src[(size_t)h * W * K + e];
        q[e] = s / (float)H;
    }
    for (int e = tid; e < H * K; e += NT) {
        const int h = e / K, k = e - h * K;
        float s = 0.f;
        for (int w = 0; w < W; ++w) s += src[((size_t)h * W + w) * K + k];
        r[e] = s / (float)W;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    for (int t = warp; t < 2 * K; t += NW) {
        const int axis = t / K, k = t - axis * K;
        const int N = axis ? H : W;
        float* T = axis ? r : q;
        const float step = lin_step(N);
        float vmax = -INFINITY;
        for (int n = lane; n < N; n += 32) vmax = fmaxf(vmax, T[n * K + k]);
        vmax = warp_max(vmax);
        float se = 0.f;
        for (int n = lane; n < N; n += 32) {
            const float e = expf(T[n * K + k] - vmax);
            T[n * K + k] = e;
            se += e;
        }
        se = warp_sum(se);
        float sc = 0.f;
        for (int n = lane; n < N; n += 32) {
            const float p = T[n * K + k] / se;
            T[n * K + k] = p;
            sc = fmaf(p, lin_coord(n, step), sc);
        }
        sc = warp_sum(sc);
        if (lane == 0) mu[((size_t)b * K + k) * 2 + axis] = sc;
    }
    __syncthreads();
    if (prob_x != nullptr)
        for (int e = tid; e < W * K; e += NT) prob_x[(size_t)b * W * K + e] = q[e];
    if (prob_y != nullptr)
        for (int e = tid; e < H * K; e += NT) prob_y[(size_t)b * H * K + e] = r[e];
}

// =============================================================================================
// Standalone renderer: mu [B,K,2] -> maps [B,hm,wm,K]   (get_gaussian_maps, utils/model.py:49-60)
// =============================================================================================
__global__ void __launch_bounds__(256)
render_fwd_kernel(const float* __restrict__ mu, int K, int hm, int wm, float neg_s, float* __restrict__ maps) {
    extern __shared__ float sm[];
    float* gy = sm;            // [hm][K]
    float* gx = gy + hm * K;   // [wm][K]
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float* m = mu + (size_t)b * K * 2;
    const float sy = lin_step(hm), sx = lin_step(wm);
    for (int idx = tid; idx < hm * K; idx += NT) {
        const int i = idx / K, k = idx - i * K;
        const float d = lin_coord(i, sy) - m[2 * k + 1];
        gy[idx] = expf(neg_s * d * d);
    }
    for (int idx = tid; idx < wm * K; idx += NT) {
        const int j = idx / K, k = idx - j * K;
        const float d = lin_coord(j, sx) - m[2 * k];
        gx[idx] = expf(neg_s * d * d);
    }
    __syncthreads();
    float* out = maps + (size_t)b * hm * wm * K;
    if ((K & 3) == 0) {
        const int KG = K >> 2;
        const float4* gy4 = reinterpret_cast<const float4*>(gy);
        const float4* gx4 = reinterpret_cast<const float4*>(gx);
        float4* out4 = reinterpret_cast<float4*>(out);
        const int total = hm * wm * KG;
        for (int idx = tid; idx < total; idx += NT) {
            const int pix = idx / KG, tg = idx - pix * KG;
            const int i = pix / wm, j = pix - i * wm;
            const float4 a = gy4[i * KG + tg], c = gx4[j * KG + tg];
            out4[idx] = make_float4(a.x * c.x, a.y * c.y, a.z * c.z, a.w * c.w);
        }
    } else {
        const int total = hm * wm * K;
        for (int idx = tid; idx < total; idx += NT) {
            const int pix = idx / K, k = idx - pix * K;
            const int i = pix / wm, j = pix - i * wm;
            out[idx] = gy[i * K + k] * gx[j * K + k];
        }
    }
}

// d_maps [B,hm,wm,K], mu -> d_mu [B,K,2] (+= d_mu_extra if given). One CTA per frame, 256 threads.
__global__ void __launch_bounds__(256)
render_bwd_kernel(const float* __restrict__ d_maps, const float* __restrict__ d_mu_extra,
                  const float* __restrict__ mu, int K, int hm, int wm, float s, float* __restrict__ d_mu) {
    extern __shared__ float sm[];
    float* gy = sm;                 // [hm][K]
    float* gx = gy + hm * K;        // [wm][K]
    float* red = gx + wm * K;       // [2][K]
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float* m = mu + (size_t)b * K * 2;
    const float sy = lin_step(hm), sx = lin_step(wm);
    for (int idx = tid; idx < hm * K; idx += NT) {
        const int i = idx / K, k = idx - i * K;
        const float d = lin_coord(i, sy) - m[2 * k + 1];
        gy[idx] = expf(-s * d * d);
    }
    for (int idx = tid; idx < wm * K; idx += NT) {
        const int j = idx / K, k = idx - j * K;
        const float d = lin_coord(j, sx) - m[2 * k];
        gx[idx] = expf(-s * d * d);
    }
    for (int idx = tid; idx < 2 * K; idx += NT) red[idx] = 0.f;
    __syncthreads();
    // warp w handles keypoints k = w, w+NW, ...; lanes stride over pixels (gathered 4-byte loads:
    // this kernel is only the standalone get_gaussian_maps backward; the fused bwd below is the hot one).
    const int warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    const float* dm = d_maps + (size_t)b * hm * wm * K;
    for (int k = warp; k < K; k += NW) {
        const float mx = m[2 * k], my = m[2 * k + 1];
        float ax = 0.f, ay = 0.f;
        for (int pix = lane; pix < hm * wm; pix += 32) {
            const int i = pix / wm, j = pix - i * wm;
            const float t = dm[(size_t)pix * K + k] * gy[i * K + k] * gx[j * K + k];
            ax = fmaf(t, lin_coord(j, sx) - mx, ax);
            ay = fmaf(t, lin_coord(i, sy) - my, ay);
        }
        ax = warp_sum(ax);
        ay = warp_sum(ay);
        if (lane == 0) {
            float ox = 2.0f * s * ax, oy = 2.0f * s * ay;
            if (d_mu_extra != nullptr) {
                ox += d_mu_extra[((size_t)b * K + k) * 2 + 0];
                oy += d_mu_extra[((size_t)b * K + k) * 2 + 1];
            }
            d_mu[((size_t)b * K + k) * 2 + 0] = ox;
            d_mu[((size_t)b * K + k) * 2 + 1] = oy;
        }
    }
}

// =============================================================================================
// Backward, fast path: d_maps (+ d_mu_extra) -> d_logits [B,H,W,K] = dr[h,k]/W + dq[w,k]/H
// =============================================================================================
template <int KG2, int NJ>
__global__ void __launch_bounds__(KG2 * 32, 4)
k1_bwd_fast(const float* __restrict__ d_maps, const float* __restrict__ d_mu_extra, const float* __restrict__ mu,
            const float* __restrict__ prob_x, const float* __restrict__ prob_y, int H, int hm, int wm, float s,
            float* __restrict__ d_logits) {
    constexpr int K = 8 * KG2, W = 16 * NJ, KG = 2 * KG2, NT = KG2 * 32;  // NT == 16*KG
    extern __shared__ __align__(16) float smf[];
    float* gy = smf;                 // [hm][K]
    float* gx = gy + hm * K;         // [wm][K]
    float* red = gx + wm * K;        // [16][2][K] partial d_mu
    float* dmu = red + 16 * 2 * K;   // [2][K]
    float* mus = dmu + 2 * K;        // [2][K]
    float* dr = mus + 2 * K;         // [H][K]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int tg = tid % KG, p0 = tid / KG;

    for (int idx = tid; idx < 2 * K; idx += NT) {
        const int axis = idx / K, k = idx - axis * K;
        mus[idx] = mu[((size_t)b * K + k) * 2 + axis];
    }
    __syncthreads();
    float4 ax = make_float4(0.f, 0.f, 0.f, 0.f), ay = ax;
    if (d_maps != nullptr) {
        const float sy = lin_step(hm), sx = lin_step(wm);
        for (int idx = tid; idx < hm * K; idx += NT) {
            const int i = idx / K, k = idx - i * K;
            const float d = lin_coord(i, sy) - mus[K + k];
            gy[idx] = expf(-s * d * d);
        }
        for (int idx = tid; idx < wm * K; idx += NT) {
            const int j = idx / K, k = idx - j * K;
            const float d = lin_coord(j, sx) - mus[k];
            gx[idx] = expf(-s * d * d);
        }
        __syncthreads();
        const float4* gy4 = reinterpret_cast<const float4*>(gy);
        const float4* gx4 = reinterpret_cast<const float4*>(gx);
        const float4* dm4 = reinterpret_cast<const float4*>(d_maps + (size_t)b * hm * wm * K);
        const float4 mx = reinterpret_cast<const float4*>(mus)[tg];
        const float4 my = reinterpret_cast<const float4*>(mus + K)[tg];
        const int npix = hm * wm;
        for (int pix = p0; pix < npix; pix += 16) {
            const int i = pix / wm, j = pix - i * wm;
            const float4 a = gy4[i * KG + tg], c = gx4[j * KG + tg], d = dm4[pix * KG + tg];
            const float cx = lin_coord(j, sx), cy = lin_coord(i, sy);
            const float tx = d.x * a.x * c.x, ty = d.y * a.y * c.y, tz = d.z * a.z * c.z, tw = d.w * a.w * c.w;
            ax.x = fmaf(tx, cx - mx.x, ax.x); ay.x = fmaf(tx, cy - my.x, ay.x);
            ax.y = fmaf(ty, cx - mx.y, ax.y); ay.y = fmaf(ty, cy - my.y, ay.y);
            ax.z = fmaf(tz, cx - mx.z, ax.z); ay.z = fmaf(tz, cy - my.z, ay.z);
            ax.w = fmaf(tw, cx - mx.w, ax.w); ay.w = fmaf(tw, cy - my.w, ay.w);
        }
    }
    reinterpret_cast<float4*>(red + (p0 * 2 + 0) * K)[tg] = ax;
    reinterpret_cast<float4*>(red + (p0 * 2 + 1) * K)[tg] = ay;
    __syncthreads();
    for (int idx = tid; idx < 2 * K; idx += NT) {
        float sum = 0.f;
#pragma unroll
        for (int p = 0; p < 16; ++p) sum += red[p * 2 * K + idx];
        sum *= 2.0f * s;
        const int axis = idx / K, k = idx - axis * K;
        if (d_mu_extra != nullptr) sum += d_mu_extra[((size_t)b * K + k) * 2 + axis];
        dmu[idx] = sum;
    }
    __syncthreads();
    // dr[h][k] = p_y (c_h - mu_y) dmu_y / W ;  dq kept in registers per thread (fixed k-group, NJ pixels)
    {
        const float stepH = lin_step(H), invW = 1.0f / (float)W;
        const float* py = prob_y + (size_t)b * H * K;
        for (int idx = tid; idx < H * K; idx += NT) {
            const int h = idx / K, k = idx - h * K;
            dr[idx] = py[idx] * (lin_coord(h, stepH) - mus[K + k]) * dmu[K + k] * invW;
        }
    }
    float4 dq[NJ];
    {
        const float stepW = lin_step(W), invH = 1.0f / (float)H;
        const float4* px4 = reinterpret_cast<const float4*>(prob_x + (size_t)b * W * K);
        const float4 mx = reinterpret_cast<const float4*>(mus)[tg];
        const float4 dx = reinterpret_cast<const float4*>(dmu)[tg];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int w = p0 + 16 * j;
            const float4 p = px4[w * KG + tg];
            const float c = lin_coord(w, stepW);
            dq[j] = make_float4(p.x * (c - mx.x) * dx.x * invH, p.y * (c - mx.y) * dx.y * invH,
                                p.z * (c - mx.z) * dx.z * invH, p.w * (c - mx.w) * dx.w * invH);
        }
    }
    __syncthreads();
    const float4* dr4 = reinterpret_cast<const float4*>(dr);
    float4* out4 = reinterpret_cast<float4*>(d_logits + (size_t)b * H * W * K);
    for (int h = 0; h < H; ++h) {
        const float4 r = dr4[h * KG + tg];
        float4* orow = out4 + (size_t)h * W * KG;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int w = p0 + 16 * j;
            __stcs(&orow[w * KG + tg], f4_add(r, dq[j]));
        }
    }
}

// Backward, generic: d_mu [B,K,2] (already complete) -> d_logits. One CTA per frame.
__global__ void __launch_bounds__(256)
k1_bwd_generic(const float* __restrict__ d_mu, const float* __restrict__ mu, const float* __restrict__ prob_x,
               const float* __restrict__ prob_y, int H, int W, int K, float* __restrict__ d_logits) {
    extern __shared__ float sm[];
    float* dq = sm;          // [W][K]
    float* dr = dq + W * K;  // [H][K]
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float stepW = lin_step(W), stepH = lin_step(H);
    for (int e = tid; e < W * K; e += NT) {
        const int w = e / K, k = e - w * K;
        const size_t mk = ((size_t)b * K + k) * 2;
        dq[e] = prob_x[(size_t)b * W * K + e] * (lin_coord(w, stepW) - mu[mk]) * d_mu[mk] / (float)H;
    }
    for (int e = tid; e < H * K; e += NT) {
        const int h = e / K, k = e - h * K;
        const size_t mk = ((size_t)b * K + k) * 2 + 1;
        dr[e] = prob_y[(size_t)b * H * K + e] * (lin_coord(h, stepH) - mu[mk]) * d_mu[mk] / (float)W;
    }
    __syncthreads();
    float* out = d_logits + (size_t)b * H * W * K;
    const int WK = W * K;
    for (int h = 0; h < H; ++h)
        for (int e = tid; e < WK; e += NT) out[(size_t)h * WK + e] = dr[h * K + (e % K)] + dq[e];
}

// =============================================================================================
// Fused render + colourise (utils/model.py:42-46 applied to get_gaussian_maps output) — SURVEY §8(f).1
//   out[b,i,j,c] = max_k g_y[i,k] g_x[j,k] colour[k,c]
// =============================================================================================
__global__ void __launch_bounds__(256)
render_colorize_kernel(const float* __restrict__ mu, const float* __restrict__ colors, int K, int hm, int wm,
                       float neg_s, float* __restrict__ out) {
    extern __shared__ float sm[];
    float* gy = sm;            // [hm][K]
    float* gx = gy + hm * K;   // [wm][K]
    float* col = gx + wm * K;  // [K][3]
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float* m = mu + (size_t)b * K * 2;
    const float sy = lin_step(hm), sx = lin_step(wm);
    for (int idx = tid; idx < hm * K; idx += NT) {
        const int i = idx / K, k = idx - i * K;
        const float d = lin_coord(i, sy) - m[2 * k + 1];
        gy[idx] = expf(neg_s * d * d);
    }
    for (int idx = tid; idx < wm * K; idx += NT) {
        const int j = idx / K, k = idx - j * K;
        const float d = lin_coord(j, sx) - m[2 * k];
        gx[idx] = expf(neg_s * d * d);
    }
    for (int idx = tid; idx < 3 * K; idx += NT) col[idx] = colors[idx];
    __syncthreads();
    float* o = out + (size_t)b * hm * wm * 3;
    for (int pix = tid; pix < hm * wm; pix += NT) {
        const int i = pix / wm, j = pix - i * wm;
        float r0 = -INFINITY, r1 = -INFINITY, r2 = -INFINITY;
        for (int k = 0; k < K; ++k) {
            const float v = gy[i * K + k] * gx[j * K + k];
            r0 = fmaxf(r0, v * col[3 * k + 0]);
            r1 = fmaxf(r1, v * col[3 * k + 1]);
            r2 = fmaxf(r2, v * col[3 * k + 2]);
        }
        o[3 * pix + 0] = r0;
        o[3 * pix + 1] = r1;
        o[3 * pix + 2] = r2;
    }
}

// colorize_point_maps on materialised maps (utils/model.py:42-46): maps [P,K], colors [K,3] -> out [P,3]
__global__ void __launch_bounds__(256)
colorize_kernel(const float* __restrict__ maps, const float* __restrict__ colors, long long P, int K,
                float* __restrict__ out) {
    extern __shared__ float col[];
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) col[i] = colors[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < P; p += nwarps) {
        float r0 = -INFINITY, r1 = -INFINITY, r2 = -INFINITY;
        for (int k = lane; k < K; k += 32) {
            const float v = maps[p * K + k];
            r0 = fmaxf(r0, v * col[3 * k + 0]);
            r1 = fmaxf(r1, v * col[3 * k + 1]);
            r2 = fmaxf(r2, v * col[3 * k + 2]);
        }
        r0 = warp_max(r0); r1 = warp_max(r1); r2 = warp_max(r2);
        if (lane == 0) {
            out[3 * p + 0] = r0;
            out[3 * p + 1] = r1;
            out[3 * p + 2] = r2;
        }
    }
}

// =============================================================================================
// host launchers
// =============================================================================================
template <int KG2, int NJ>
static size_t k1_fwd_fast_smem(int H) {
    constexpr int K = 8 * KG2, W = 16 * NJ;
    return (size_t)(K1_NSTAGE * W * K + K * (H + 4) + K * (W + 4) + 2 * K) * 4 + 2 * K1_NSTAGE * 8;
}

template <int KG2, int NJ>
static int launch_k1_fwd_fast(const float* logits, int B, int H, float* mu, float* px, float* py, float* maps, int hm,
                              int wm, float inv_std, cudaStream_t st) {
    const size_t smem = k1_fwd_fast_smem<KG2, NJ>(H);
    static bool attr_done = false;  // per-instantiation; idempotent, so a benign race at worst
    if (!attr_done) {
        KP_CUDA_CHECK(cudaFuncSetAttribute(k1_fwd_fast<KG2, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        attr_done = true;
    }
    const float neg_s = -(float)((double)inv_std * (double)inv_std);
    k1_fwd_fast<KG2, NJ><<<B, (KG2 + 1) * 32, smem, st>>>(logits, H, mu, px, py, maps, hm, wm, neg_s);
    KP_LAUNCHED();
    return KP_OK;
}

static bool k1_fast_ok(int H, int W, int K, int hm, int wm, bool want_maps) {
    if (!(K == 40 && W == 128)) return false;            // instantiated shape (configs/penn.yaml: n_pts 40, 128x128)
    if (H < 1 || H > 256) return false;
    if (want_maps && (hm > H + 4 || wm > W + 4)) return false;
    return k1_fwd_fast_smem<5, 8>(H) <= 113 * 1024;
}

int k1_render_fwd(const float* mu, int B, int K, int hm, int wm, float inv_std, float* maps, cudaStream_t st) {
    if (B == 0) return KP_OK;
    const size_t smem = (size_t)(hm + wm) * K * 4;
    KP_REQUIRE(smem <= 200 * 1024, "kp_render_fwd: (h+w)*K too large for shared memory (%zu B)", smem);
    if (smem > 48 * 1024)
        KP_CUDA_CHECK(cudaFuncSetAttribute(render_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float neg_s = -(float)((double)inv_std * (double)inv_std);
    render_fwd_kernel<<<B, 256, smem, st>>>(mu, K, hm, wm, neg_s, maps);
    KP_LAUNCHED();
    return KP_OK;
}

int k1_render_bwd(const float* d_maps, const float* d_mu_extra, const float* mu, int B, int K, int hm, int wm,
                  float inv_std, float* d_mu, cudaStream_t st) {
    if (B == 0) return KP_OK;
    const size_t smem = (size_t)((hm + wm) * K + 2 * K) * 4;
    KP_REQUIRE(smem <= 200 * 1024, "kp_render_bwd: (h+w)*K too large for shared memory (%zu B)", smem);
    if (smem > 48 * 1024)
        KP_CUDA_CHECK(cudaFuncSetAttribute(render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float s = (float)((double)inv_std * (double)inv_std);
    render_bwd_kernel<<<B, 256, smem, st>>>(d_maps, d_mu_extra, mu, K, hm, wm, s, d_mu);
    KP_LAUNCHED();
    return KP_OK;
}

int k1_render_colorize(const float* mu, const float* colors, int B, int K, int hm, int wm, float inv_std, float* out,
                       cudaStream_t st) {
    if (B == 0) return KP_OK;
    const size_t smem = (size_t)((hm + wm) * K + 3 * K) * 4;
    KP_REQUIRE(smem <= 200 * 1024, "kp_render_colorize: (h+w)*K too large for shared memory (%zu B)", smem);
    if (smem > 48 * 1024)
        KP_CUDA_CHECK(
            cudaFuncSetAttribute(render_colorize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float neg_s = -(float)((double)inv_std * (double)inv_std);
    render_colorize_kernel<<<B, 256, smem, st>>>(mu, colors, K, hm, wm, neg_s, out);
    KP_LAUNCHED();
    return KP_OK;
}

int k1_colorize(const float* maps, const float* colors, long long P, int K, float* out, cudaStream_t st) {
    if (P == 0) return KP_OK;
    KP_REQUIRE(K <= 4096, "kp_colorize_fwd: K=%d too large", K);
    long long blocks = (P * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    colorize_kernel<<<(int)blocks, 256, 3 * K * sizeof(float), st>>>(maps, colors, P, K, out);
    KP_LAUNCHED();
    return KP_OK;
}

int k1_softargmax_render_fwd(const float* logits, int B, int H, int W, int K, float* mu, float* prob_x, float* prob_y,
                             float* maps, int hm, int wm, float inv_std, cudaStream_t st) {
    if (B == 0) return KP_OK;
    const bool want_maps = maps != nullptr;
    if (k1_fast_ok(H, W, K, hm, wm, want_maps) && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
        (!want_maps || (reinterpret_cast<uintptr_t>(maps) & 15) == 0)) {
        return launch_k1_fwd_fast<5, 8>(logits, B, H, mu, prob_x, prob_y, maps, hm, wm, inv_std, st);
    }
    const size_t smem = (size_t)(H + W) * K * 4;
    KP_REQUIRE(smem <= 200 * 1024, "kp_softargmax: (H+W)*K too large for shared memory (%zu B)", smem);
    if (smem > 48 * 1024)
        KP_CUDA_CHECK(cudaFuncSetAttribute(k1_fwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1_fwd_generic<<<B, 256, smem, st>>>(logits, H, W, K, mu, prob_x, prob_y);
    KP_LAUNCHED();
    if (want_maps) return k1_render_fwd(mu, B, K, hm, wm, inv_std, maps, st);
    return KP_OK;
}

int k1_softargmax_render_bwd(const float* d_maps, const float* d_mu_extra, const float* mu, const float* prob_x,
                             const float* prob_y, int B, int H, int W, int K, int hm, int wm, float inv_std,
                             float* d_logits, float* d_mu_scratch, cudaStream_t st) {
    if (B == 0) return KP_OK;
    const float s = (float)((double)inv_std * (double)inv_std);
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_logits) | reinterpret_cast<uintptr_t>(prob_x) |
                           reinterpret_cast<uintptr_t>(d_maps)) & 15) == 0;
    if (K == 40 && W == 128 && aligned) {
        const int hm_ = d_maps ? hm : 0, wm_ = d_maps ? wm : 0;
        const size_t smem = (size_t)((hm_ + wm_) * K + 16 * 2 * K + 4 * K + H * K) * 4;
        if (smem <= 56 * 1024) {
            static bool attr_done = false;
            if (!attr_done) {
                KP_CUDA_CHECK(
                    cudaFuncSetAttribute(k1_bwd_fast<5, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024));
                attr_done = true;
            }
            k1_bwd_fast<5, 8><<<B, 160, smem, st>>>(d_maps, d_mu_extra, mu, prob_x, prob_y, H, hm_, wm_, s, d_logits);
            KP_LAUNCHED();
            return KP_OK;
        }
    }
    // generic: d_mu first (needs scratch [B,K,2]), then the rank-structured scatter
    KP_REQUIRE(d_mu_scratch != nullptr, "kp_softargmax_render_bwd: generic path needs a d_mu scratch buffer [B,K,2]");
    if (d_maps != nullptr) {
        int rc = k1_render_bwd(d_maps, d_mu_extra, mu, B, K, hm, wm, inv_std, d_mu_scratch, st);
        if (rc != KP_OK) return rc;
    } else {
        KP_REQUIRE(d_mu_extra != nullptr, "kp_softargmax_render_bwd: neither d_maps nor d_mu given");
        KP_CUDA_CHECK(cudaMemcpyAsync(d_mu_scratch, d_mu_extra, (size_t)B * K * 2 * 4, cudaMemcpyDeviceToDevice, st));
    }
    const size_t smem = (size_t)(H + W) * K * 4;
    KP_REQUIRE(smem <= 200 * 1024, "kp_softargmax_render_bwd: (H+W)*K too large for shared memory (%zu B)", smem);
    if (smem > 48 * 1024)
        KP_CUDA_CHECK(cudaFuncSetAttribute(k1_bwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1_bwd_generic<<<B, 256, smem, st>>>(d_mu_scratch, mu, prob_x, prob_y, H, W, K, d_logits);
    KP_LAUNCHED();
    return KP_OK;
}

}  // namespace kp
