// Tap-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Serves every convolution of the stage-1 path (reference: layers.conv, models/networks/layers.py:4-10;
// Vgg19.conv_layer, models/networks/vgg.py:48-55), forward and data-gradient; see include/kp_b200.h.
//
// PERSISTENT kernel: grid = min(#tiles, SMs x CTAs/SM); every CTA walks output tiles (128 pixels x BN channels)
// tile = blockIdx.x, blockIdx.x + gridDim.x, ...  Three roles run concurrently and are decoupled by mbarriers:
//   warp 0     TMA producer  - per K step one 4-D box [TN][TH][TW][CB] of the (shifted) activation view
//                              (zero fill outside the image = TF zero padding) + one 2-D box [BN][CB] of
//                              the packed weights, both hardware-swizzled, into an S-stage smem ring that
//                              keeps rolling ACROSS tiles (the next tile's loads start while this tile computes);
//   warp 1     MMA issuer    - tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16, fp32 accumulators in
//                              TMEM, DOUBLE BUFFERED (2 x BN columns): tile i+1 accumulates while tile i drains;
//   warps 2-5  epilogue      - tcgen05.ld 32x32b (one output pixel per thread), optional per-channel
//                              sum / sum-of-squares (batch-norm statistics) by a warp transpose-reduce +
//                              atomics, bias + activation (+ accumulate), bf16/f32 NHWC stores.
// (ncu of the first, non-persistent version: tensor pipe 9-21 % active with DRAM/L2/L1 all far from saturated,
//  i.e. bound by the serial prologue -> main loop -> epilogue chain of each tile; profiles/README.md.)
#include "kp_tc.cuh"
#include "conv_epilogue.cuh"
#include "kp_internal.h"
#include <cudaTypedefs.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace kp {

struct alignas(64) TapConvKParams {
    CUtensorMap mapA[KP_MAX_MAPS];
    CUtensorMap mapB;
    int n_taps, n_src;
    int nblk[KP_MAX_MAPS];
    signed char dh[KP_MAX_TAPS], dw[KP_MAX_TAPS], mf[KP_MAX_TAPS];
    int TW, TH, TN, tiles_w, tiles_h;
    int Ho, Wo, N;
    int BN, n_tiles, total_tiles, tmem_cols, stages, total_iters;
    int ksplit, groups_per_split, bpt;      // split-K over CTAs (fp32 atomic epilogue), blocks per tap
    uint32_t a_bytes, b_bytes, stage_bytes;
    void* out;
    long long out_off, out_sw, out_sh, out_sn;
    int Cout, cout_pad, out_f32, act, accumulate;
    int sgroups, group_n;      // batch-norm statistics per batch segment of group_n images (tiles never straddle segments)
    float alpha, slope;
    const float* bias;
    float* ssum;
    float* ssq;
    unsigned long long* dbg;   // KP_TAPCONV_TRACE: per-tile timestamps of CTA 0 (debug only)
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KP_TTRACE(slot, tileidx)                                                                   \
    do {                                                                                          \
        if (p.dbg != nullptr && blockIdx.x == 0 && (tileidx) < 24) p.dbg[(tileidx) * 8 + (slot)] = gtime(); \
    } while (0)

__device__ __forceinline__ float apply_act(float x, int act, float alpha, bool last) {
    switch (act) {
        case KP_ACT_RELU: return fmaxf(x, 0.f);
        case KP_ACT_LEAKY: return x >= 0.f ? x : alpha * x;
        case KP_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
        case KP_ACT_SIGMOID_LAST: return last ? 1.f / (1.f + __expf(-x)) : x;
        default: return x;
    }
}

// EPI: 0 = general epilogue (epi_chunk), 1 = fast epilogue (conv_epilogue.cuh: branch-free bias + activation + bf16 stores),
// 2 = fast epilogue + batch-norm statistics.
template <int CB, int EPI>
__global__ void __launch_bounds__(192, 1) tapconv_kernel(const __grid_constant__ TapConvKParams p) {
    constexpr uint32_t ROW_BYTES = CB * 2;
    constexpr uint32_t SBO = 8 * ROW_BYTES;
    constexpr uint32_t LAYOUT = (CB == 64) ? 2u : (CB == 32) ? 4u : 6u;
    // One pipeline stage always carries 64 K-elements: G = 64/CB activation boxes (one per channel block, each with
    // the swizzle of its own row width) and ONE weight box [BN][64] with the 128-byte swizzle.  For the small-channel
    // layers (CB 16/32) this divides the barrier round trips of the single-thread TMA / MMA issuers by G.
    constexpr int G = 64 / CB;
    constexpr uint32_t A_BOX_BYTES = 128u * CB * 2u;

    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)p.stages * p.stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;   // [2] accumulator buffer ready for the epilogue
    uint64_t* tempty = tfull + 2;         // [2] accumulator buffer drained
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tempty + 2);
    // per-CTA copies for the epilogue: bias [Cout_pad] and the batch-norm partial sums [2][Cout_pad]
    float* s_bias = reinterpret_cast<float*>(tslot + 4);
    float* s_stat = s_bias + p.cout_pad;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);     // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 1) tmem_alloc(tslot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    pdl_wait();      // everything above overlaps the tail of the previous kernel; global memory is touched below only

    if (warp == 0) {
        // ------------------------------- TMA producer -------------------------------
        if (lane == 0) {
            for (int m = 0; m < KP_MAX_MAPS; ++m)
                if (p.nblk[m] > 0) tma_prefetch_desc(&p.mapA[m]);
            tma_prefetch_desc(&p.mapB);
            uint32_t git = 0;   // ring position (one per GROUP of G channel blocks), keeps counting across tiles
            int ltp = 0;
            for (int work = blockIdx.x; work < p.total_tiles * p.ksplit; work += gridDim.x, ++ltp) {
                const int tile = work / p.ksplit, ks = work - tile * p.ksplit;
                KP_TTRACE(0, ltp);
                const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
                const int w0 = (mt % p.tiles_w) * p.TW, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
                const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
                const int n_off = nt * p.BN;
                const int it_begin = ks * p.groups_per_split * G;
                const int it_end = min(p.total_iters, (ks + 1) * p.groups_per_split * G);
                // (tap, source, channel block) of the first K block of this split
                int t = it_begin / p.bpt, s = 0, cb = it_begin - t * p.bpt;
                while (cb >= p.nblk[p.mf[t] + s]) { cb -= p.nblk[p.mf[t] + s]; ++s; }
                for (int it0 = it_begin; it0 < it_end; it0 += G, ++git) {
                    const int nv = min(G, it_end - it0);
                    const uint32_t st = git % (uint32_t)S;
                    if (git >= (uint32_t)S) mbar_wait(&empty[st], ((git / (uint32_t)S) - 1) & 1);
                    uint8_t* a_dst = base + (size_t)st * p.stage_bytes;
                    mbar_arrive_expect_tx(&full[st], (uint32_t)nv * A_BOX_BYTES + p.b_bytes);
                    for (int g = 0; g < nv; ++g) {
                        const int m = p.mf[t] + s;
                        tma_load_4d(a_dst + (size_t)g * A_BOX_BYTES, &p.mapA[m], &full[st], cb * CB, w0 + p.dw[t], h0 + p.dh[t], n0);
                        if (++cb == p.nblk[m]) {
                            cb = 0;
                            if (++s == p.n_src) { s = 0; ++t; }
                        }
                    }
                    tma_load_2d(a_dst + p.a_bytes, &p.mapB, &full[st], it0 * CB, n_off);   // [BN][64] K-major, 128B swizzle
                }
                KP_TTRACE(1, ltp);
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer -------------------------------
        {
            const uint32_t leader = elect_one() ? 1u : 0u;
            const uint32_t idesc = umma_idesc_bf16(128, p.BN, 0, 0);
            const uint64_t a_hi = umma_smem_desc(0u, SBO, 16, LAYOUT);
            const uint64_t b_hi = umma_smem_desc(0u, 1024, 16, 2u);
            uint32_t git = 0;
            int lt = 0;
            for (int work = blockIdx.x; work < p.total_tiles * p.ksplit; work += gridDim.x, ++lt) {
                const int ks = work % p.ksplit;
                const int it_begin = ks * p.groups_per_split * G;
                const int it_end = min(p.total_iters, (ks + 1) * p.groups_per_split * G);
                const int acc = lt & 1;
                if (lt >= 2) mbar_wait(&tempty[acc], ((lt >> 1) - 1) & 1);   // epilogue drained this buffer
                tc_fence_after();
                if (lane == 0) KP_TTRACE(2, lt);
                const uint32_t d_tmem = tmem + (uint32_t)(acc * p.BN);
                uint32_t accum = 0u;
                for (int it0 = it_begin; it0 < it_end; it0 += G, ++git) {
                    const int nv = min(G, it_end - it0);
                    const uint32_t st = git % (uint32_t)S;
                    // A stage always carries 4 MMAs (64 K-elements).  Their descriptors are built BEFORE the wait - the
                    // single issuing warp runs on the uniform datapath at ~8 cycles per dependent instruction, and a
                    // dozen instructions between two UTCHMMA cost ~100 cycles per MMA (measured: T_mma ~ 100 + N/2
                    // cycles for N = 16..256) - so that the MMAs go out back to back once the data has landed.
                    const uint32_t a16 = (smem_base + st * p.stage_bytes) >> 4;
                    const uint32_t b16 = a16 + (p.a_bytes >> 4);
                    uint64_t da[4], db[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        constexpr int KPB = CB / 16;                 // MMAs per channel block
                        const int g = i / KPB, k = i % KPB;
                        // A: box g, rows of CB*2 bytes with the matching swizzle; B: 128-byte rows (64 K-elements),
                        // the 16-element slice of block g / step k sits (g*CB + k*16)*2 bytes into the row
                        da[i] = a_hi | (uint64_t)(a16 + (uint32_t)((g * (int)A_BOX_BYTES + k * 32) >> 4));
                        db[i] = b_hi | (uint64_t)(b16 + (uint32_t)(((g * CB + k * 16) * 2) >> 4));
                    }
                    const int n_mma = nv * (CB / 16);
                    mbar_wait(&full[st], (git / (uint32_t)S) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (i < n_mma) {
                            umma_bf16_if(leader, d_tmem, da[i], db[i], idesc, i == 0 ? accum : 1u);
                        }
                    }
                    accum = 1u;
                    umma_commit_if(leader, &empty[st]);
                }
                if (lane == 0) KP_TTRACE(3, lt);
                umma_commit_if(leader, &tfull[acc]);
            }
        }
    } else {
        // ------------------------------- epilogue -------------------------------
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int tw = row % p.TW, th = (row / p.TW) % p.TH, tn = row / (p.TW * p.TH);
        // Stage the bias in shared memory once (a global load per 16-column chunk was the top stall of the epilogue
        // on the short-K layers) and keep the batch-norm partial sums per CTA: they are flushed with ONE global atomic
        // per channel per CTA after the last tile instead of one per channel per warp per tile.
        const int et = threadIdx.x - 64;
        for (int i = et; i < p.cout_pad; i += 128) s_bias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
        if (p.ssum != nullptr)
            for (int i = et; i < 8 * p.sgroups * p.cout_pad; i += 128) s_stat[i] = 0.f;      // four warp-private copies of [G][2][cout_pad]
        named_bar_sync(1, 128);
        float* const s_stat_w0 = s_stat + (warp - 2) * p.sgroups * 2 * p.cout_pad;
        const int group_n = p.group_n, gstride = 2 * p.cout_pad;
        int lt = 0;
        if (EPI != 0) {
            // fast epilogue (ksplit == 1): launch constants in registers, branch-free chunks
            const EpiFast ef = {p.slope, p.cout_pad};
            __nv_bfloat16* const outp = reinterpret_cast<__nv_bfloat16*>(p.out) + p.out_off;
            const long long out_sn = p.out_sn, out_sh = p.out_sh, out_sw = p.out_sw;
            const int Ho = p.Ho, Wo = p.Wo, N = p.N, BN = p.BN, n_tiles = p.n_tiles, tiles_w = p.tiles_w, tiles_h = p.tiles_h;
            const int TW = p.TW, TH = p.TH, TN = p.TN;
            const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
            for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x, ++lt) {
                const int mt = work / n_tiles, nt = work - mt * n_tiles;
                const int w0 = (mt % tiles_w) * TW, h0 = ((mt / tiles_w) % tiles_h) * TH;
                const int n0 = (mt / (tiles_w * tiles_h)) * TN;
                const int n_off = nt * BN;
                const int uw = w0 + tw, uh = h0 + th, n = n0 + tn;
                const bool valid = (uw < Wo) && (uh < Ho) && (n < N);
                float* const s_stat_w = s_stat_w0 + (n0 / group_n) * gstride;     // the tile's statistics segment (warp-uniform)
                __nv_bfloat16* const o_p = outp + (long long)n * out_sn + (long long)uh * out_sh + (long long)uw * out_sw + n_off;
                const int acc = lt & 1;
                mbar_wait(&tfull[acc], (lt >> 1) & 1);
                tc_fence_after();
                const uint32_t t_row = t_lane + (uint32_t)(acc * BN);
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    float v[1][16];
                    tmem_ld16(t_row + (uint32_t)c0, v[0]);
                    if (c0 + 16 >= BN) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
                    if (EPI == 2) {
                        if (!valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[0][j] = 0.f;
                        }
                        epi_stats16<1>(v, lane, s_stat_w + n_off + c0, ef.cout_pad);
                    }
                    epi_chunk_fast(v[0], valid, o_p + c0, s_bias + n_off + c0, ef);
                }
            }
        } else
        for (int work = blockIdx.x; work < p.total_tiles * p.ksplit; work += gridDim.x, ++lt) {
            const int tile = work / p.ksplit, ks = work - tile * p.ksplit;
            const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
            const int w0 = (mt % p.tiles_w) * p.TW, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
            const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
            const int n_off = nt * p.BN;
            const int uw = w0 + tw, uh = h0 + th, n = n0 + tn;
            const bool valid = (uw < p.Wo) && (uh < p.Ho) && (n < p.N);
            float* const s_stat_w = s_stat_w0 + (n0 / group_n) * gstride;
            const long long pix = p.out_off + (long long)n * p.out_sn + (long long)uh * p.out_sh + (long long)uw * p.out_sw;
            const int acc = lt & 1;
            if (threadIdx.x == 64) KP_TTRACE(4, lt);
            mbar_wait(&tfull[acc], (lt >> 1) & 1);
            tc_fence_after();
            if (threadIdx.x == 64) KP_TTRACE(5, lt);
            const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                float v[16];
                __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-pixel predicated stores
                tmem_ld16(t_row + (uint32_t)c0, v);
                if (c0 == 0 && threadIdx.x == 64) KP_TTRACE(7, lt);
                if (c0 + 16 >= p.BN) {
                    // last TMEM read of this tile: hand the accumulator buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                }
                if (p.ssum != nullptr && !valid) {
                    // an output pixel outside the image still reads real input through its taps: keep it out of the
                    // batch-norm statistics (Wo / Ho not a multiple of the pixel tile)
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                epi_chunk(p, v, n_off + c0, valid, pix, lane, ks, s_bias, s_stat_w);
            }
            if (threadIdx.x == 64) KP_TTRACE(6, lt);
        }
        if (p.ssum != nullptr) {
            named_bar_sync(1, 128);
            const int cp = p.cout_pad, G = p.sgroups, ws = G * 2 * cp;      // ws: stride between the warps' copies
            for (int i = et; i < G * cp; i += 128) {
                const int g = i / cp, c = i - g * cp;
                const float* b = s_stat + g * 2 * cp + c;
                atomicAdd(p.ssum + i, b[0] + b[ws] + b[2 * ws] + b[3 * ws]);
                atomicAdd(p.ssq + i, b[cp] + b[ws + cp] + b[2 * ws + cp] + b[3 * ws + cp]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
            set_error("cuTensorMapEncodeTiled entry point unavailable (cuda error %d, query %d)", (int)e, (int)qres);
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static CUtensorMapSwizzle swizzle_for(int CB) {
    return CB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("KP_PDL");
        on = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    }
    return on == 1;
}

int device_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            sms <= 0)
            sms = 148;
    }
    return sms;
}

// Encode a strided [N][Hd][Wd][C] bf16 view as a 4-D tiled tensor map with box {CB, TW, TH, TN}.
int encode_view_map(CUtensorMap* out, const kp_tap_view& v, const void* src_base, int N, int CB, int TW, int TH, int TN,
                    const char* what) {
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) return KP_ERR_DRIVER;
    KP_REQUIRE(v.C > 0 && v.C % 8 == 0, "%s: channels %d must be a positive multiple of 8", what, v.C);
    KP_REQUIRE(v.sw % 8 == 0 && v.sh % 8 == 0 && v.sn % 8 == 0 && v.off % 8 == 0,
               "%s: strides/offset must be multiples of 8 elements (16 B)", what);
    KP_REQUIRE(src_base != nullptr, "%s: NULL source", what);
    const char* basep = reinterpret_cast<const char*>(src_base) + v.off * 2;
    KP_REQUIRE((reinterpret_cast<uintptr_t>(basep) & 15) == 0, "%s: base not 16-byte aligned", what);
    cuuint64_t gdim[4] = {(cuuint64_t)v.C, (cuuint64_t)v.Wd, (cuuint64_t)v.Hd, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
    cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(basep), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(CB), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed with %d (C=%d W=%d H=%d N=%d box %d,%d,%d,%d)", what, (int)r, v.C, v.Wd,
                  v.Hd, N, CB, TW, TH, TN);
        return KP_ERR_DRIVER;
    }
    return KP_OK;
}

// pick TW,TH,TN (product `pixels`, powers of two) minimising padded pixels
void choose_pixel_tile(int pixels, int Wo, int Ho, int N, int* TW, int* TH, int* TN) {
    long long best = -1;
    for (int tw = 1; tw <= pixels; tw <<= 1) {
        for (int th = 1; tw * th <= pixels; th <<= 1) {
            const int tn = pixels / (tw * th);
            const long long cover = (long long)((Wo + tw - 1) / tw) * tw * ((Ho + th - 1) / th) * th *
                                    ((N + tn - 1) / tn) * tn;
            if (best < 0 || cover < best || (cover == best && tw > *TW)) {
                best = cover;
                *TW = tw; *TH = th; *TN = tn;
            }
        }
    }
}

static int pow2_at_least(int v, int lo) {
    int r = lo;
    while (r < v) r <<= 1;
    return r;
}

int tapconv_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                   float* ssum, float* ssq, cudaStream_t st) {
    KP_REQUIRE(d->CB == 16 || d->CB == 32 || d->CB == 64, "kp_tapconv: CB must be 16, 32 or 64 (got %d)", d->CB);
    KP_REQUIRE(d->n_maps >= 1 && d->n_maps <= KP_MAX_MAPS, "kp_tapconv: n_maps %d out of range", d->n_maps);
    KP_REQUIRE(d->n_taps >= 1 && d->n_taps <= KP_MAX_TAPS, "kp_tapconv: n_taps %d out of range", d->n_taps);
    KP_REQUIRE(d->n_src >= 1 && d->n_src <= KP_MAX_MAPS, "kp_tapconv: n_src %d out of range", d->n_src);
    KP_REQUIRE(d->N > 0 && d->Ho > 0 && d->Wo > 0 && d->Cout > 0, "kp_tapconv: empty problem");
    KP_REQUIRE((ssum == nullptr) == (ssq == nullptr), "kp_tapconv: stats_sum and stats_sq go together");
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) return KP_ERR_DRIVER;
    // stride-1 multi-tap layers: halo-tile kernels (each input element is fetched once per tile instead of once per tap);
    // the TMA-staged one first, the cp.async gather kernel for what it does not take (8-channel slots)
    if (halo2_eligible(d, ssum)) return halo2_launch(d, src, wpacked, bias, out, ssum, ssq, st);
    const bool segmented = d->stat_groups > 1 && ssum != nullptr;      // only the two kernels below keep per-segment statistics
    if (!segmented && haloconv_eligible(d, ssum)) return haloconv_launch(d, src, wpacked, bias, out, ssum, ssq, st);
    // wide layers, experimental: CTA pairs (cta_group::2) halve the weight fill per SM
    if (!segmented && tapconv2_eligible(d)) return tapconv2_launch(d, src, wpacked, bias, out, ssum, ssq, st);

    TapConvKParams p;
    memset(&p, 0, sizeof(p));
    const int CB = d->CB;
    int TW = d->TW, TH = d->TH, TN = d->TN, BN = d->BN;
    // with per-segment statistics a tile must not straddle two segments: choose the image count per tile for one segment
    if (TW <= 0 || TH <= 0 || TN <= 0)
        choose_pixel_tile(128, d->Wo, d->Ho, (d->stat_groups > 1 && ssum != nullptr) ? d->N / d->stat_groups : d->N, &TW, &TH, &TN);
    KP_REQUIRE(TW * TH * TN == 128, "kp_tapconv: tile %dx%dx%d is not 128 pixels", TW, TH, TN);
    // Split-K needs an fp32 output that takes atomic adds: either a contiguous tensor this call may zero itself, or any
    // view the CALLER has zeroed (accumulate == 2: the four strided parity views of a stride-2 data gradient).
    const bool out_contiguous = d->out_off == 0 && d->out_sw == d->Cout && d->out_sh == (long long)d->Wo * d->Cout &&
                                d->out_sn == (long long)d->Ho * d->Wo * d->Cout;
    const bool splitk_ok = d->out_f32 && d->act == KP_ACT_NONE && ssum == nullptr &&
                           (d->accumulate == 2 || (!d->accumulate && out_contiguous));
    if (BN <= 0) {
        BN = d->Cout_pad <= 256 ? d->Cout_pad : (d->Cout_pad % 256 == 0 ? 256 : 128);
        // Under-filled launches (few pixel tiles: 16x16 layers, img_discr conv_3..5): halve the channel tile until the
        // tile count reaches ~the SM count; per-tile efficiency drops a little, idle SMs cost a lot.  When the K loop is long
        // and may be split over CTAs instead, the wide tile stays (every CTA then re-reads 4x less of the activations).
        const long long m_tiles = (long long)((d->Wo + TW - 1) / TW) * ((d->Ho + TH - 1) / TH) * ((d->N + TN - 1) / TN);
        const bool prefer_splitk = splitk_ok && d->Ktot >= 32 * 64 && m_tiles * (d->Cout_pad / BN) * 2 <= device_sm_count();
        while (!prefer_splitk && BN >= 128 && BN % 32 == 0 && m_tiles * (d->Cout_pad / BN) < (long long)(device_sm_count() * 3) / 4)
            BN /= 2;
        if (const char* e = getenv("KP_TAPCONV_BN_MAX")) {   // experiments: cap the channel tile
            const int cap = atoi(e);
            if (cap >= 16 && cap % 16 == 0 && BN > cap && d->Cout_pad % cap == 0) BN = cap;
        }
    }
    KP_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256 && d->Cout_pad % BN == 0,
               "kp_tapconv: BN=%d must be a multiple of 16 in [16,256] dividing Cout_pad=%d", BN, d->Cout_pad);
    KP_REQUIRE(d->Cout <= d->Cout_pad, "kp_tapconv: Cout > Cout_pad");

    // A maps
    int total_blocks_per_tap = -1;
    for (int m = 0; m < d->n_maps; ++m) {
        const kp_tap_view& v = d->map[m];
        KP_REQUIRE(v.src >= 0 && v.src < KP_MAX_MAPS, "kp_tapconv: map %d has no source", m);
        const int rc = encode_view_map(&p.mapA[m], v, src[v.src], d->N, CB, TW, TH, TN, "kp_tapconv A map");
        if (rc != KP_OK) return rc;
        p.nblk[m] = (v.C + CB - 1) / CB;
    }
    int total_iters = 0;
    for (int t = 0; t < d->n_taps; ++t) {
        KP_REQUIRE(d->map_first[t] >= 0 && d->map_first[t] + d->n_src <= d->n_maps, "kp_tapconv: tap %d maps out of range", t);
        int blocks = 0;
        for (int s = 0; s < d->n_src; ++s) blocks += p.nblk[d->map_first[t] + s];
        if (total_blocks_per_tap < 0) total_blocks_per_tap = blocks;
        KP_REQUIRE(blocks == total_blocks_per_tap, "kp_tapconv: taps must read the same number of channel blocks");
        total_iters += blocks;
        p.dh[t] = d->dh[t]; p.dw[t] = d->dw[t]; p.mf[t] = d->map_first[t];
    }
    KP_REQUIRE(d->Ktot == total_iters * CB, "kp_tapconv: Ktot=%d does not match taps x blocks x CB = %d", d->Ktot,
               total_iters * CB);
    {
        KP_REQUIRE((reinterpret_cast<uintptr_t>(wpacked) & 15) == 0, "kp_tapconv: packed weights not 16-byte aligned");
        cuuint64_t gdim[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Cout_pad};
        cuuint64_t gstr[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)BN};     // 64 K-elements (G channel blocks) per stage, 128-byte rows
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wpacked), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("kp_tapconv: cuTensorMapEncodeTiled(weights) failed with %d (Ktot=%d Cout_pad=%d)", (int)r, d->Ktot,
                      d->Cout_pad);
            return KP_ERR_DRIVER;
        }
    }
    p.n_taps = d->n_taps; p.n_src = d->n_src;
    p.TW = TW; p.TH = TH; p.TN = TN;
    p.tiles_w = (d->Wo + TW - 1) / TW;
    p.tiles_h = (d->Ho + TH - 1) / TH;
    const int tiles_n = (d->N + TN - 1) / TN;
    p.Ho = d->Ho; p.Wo = d->Wo; p.N = d->N;
    p.BN = BN;
    p.n_tiles = d->Cout_pad / BN;
    p.total_tiles = p.tiles_w * p.tiles_h * tiles_n * p.n_tiles;
    p.tmem_cols = pow2_at_least(2 * BN, 32);          // double-buffered accumulator
    p.total_iters = total_iters;
    p.bpt = total_blocks_per_tap;
    p.ksplit = 1;
    {
        // split-K for long-K launches with very few tiles (img_discr D_logit: 18 tiles x 288 K-groups): fp32 output,
        // no activation, contiguous output so that it can be zeroed here
        const int G = 64 / CB;
        const int n_groups = (total_iters + G - 1) / G;
        if (splitk_ok && n_groups >= 32 && p.total_tiles * 2 <= device_sm_count()) {
            int ks = device_sm_count() / p.total_tiles;
            if (ks > n_groups / 8) ks = n_groups / 8;
            if (ks > 1) {
                p.groups_per_split = (n_groups + ks - 1) / ks;
                p.ksplit = (n_groups + p.groups_per_split - 1) / p.groups_per_split;
                if (d->accumulate != 2)     // (accumulate == 2: the caller zeroed the whole output before its launches)
                    KP_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)d->N * d->Ho * d->Wo * d->Cout * sizeof(float), st));
            }
        }
        if (p.ksplit == 1) p.groups_per_split = n_groups;
    }
    {
        // G = 64/CB activation boxes of 128 x CB per stage - or fewer when the whole K loop is shorter than one stage
        // (1x1 head: one 4 KB box), which leaves room for more stages / CTAs
        const int G = 64 / CB;
        const int boxes = total_iters < G ? total_iters : G;
        p.a_bytes = ((uint32_t)boxes * 128u * (uint32_t)CB * 2u + 1023u) & ~1023u;
    }
    p.b_bytes = (uint32_t)BN * 64u * 2u;               // one weight box [BN][64]
    p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
    // CTAs per SM: two independent pipelines per SM hide the barrier round trips of the single-thread TMA / MMA
    // issuers; the 256-wide tiles need all 512 TMEM columns and most of the shared memory, so they run alone.
    // Short-K tiles (<= 8 K groups, narrow channel tile) are bound by the latency of the epilogue chain, not by the
    // tensor pipe: a third CTA per SM adds epilogue warps.
    const int SG = d->stat_groups > 1 ? d->stat_groups : 1;
    KP_REQUIRE(d->N % SG == 0 && (SG == 1 || (d->N / SG) % TN == 0),
               "kp_tapconv: %d images do not split into %d statistics segments of whole %d-image tiles", d->N, SG, TN);
    p.sgroups = SG; p.group_n = d->N / SG;
    // bias (+ four warp-private [G][2][Cout_pad] statistics)
    const uint32_t epi_bytes = (ssum != nullptr ? 1u + 8u * (uint32_t)SG : 1u) * (uint32_t)d->Cout_pad * sizeof(float);
    int ctas_per_sm = p.tmem_cols <= 256 ? 2 : 1;
    if (p.tmem_cols <= 128 && p.groups_per_split <= 8) ctas_per_sm = 3;
    if (p.tmem_cols <= 128 && p.groups_per_split <= 1) ctas_per_sm = 4;   // one K group per tile: pure epilogue/latency work
    if (const char* e = getenv("KP_TAPCONV_CTAS_PER_SM")) {
        const int want = atoi(e);
        ctas_per_sm = want >= 4 && p.tmem_cols <= 128 ? 4 : want >= 3 && p.tmem_cols <= 128 ? 3 : want >= 2 && p.tmem_cols <= 256 ? 2 : 1;
    }
    auto cta_budget = [](int c) { return (c == 4 ? 52u : c == 3 ? 70u : c == 2 ? 108u : 216u) * 1024u; };
    while (ctas_per_sm > 1 && cta_budget(ctas_per_sm) < epi_bytes + 2u * p.stage_bytes) --ctas_per_sm;   // wide Cout: the staged bias needs room
    KP_REQUIRE(cta_budget(ctas_per_sm) >= epi_bytes + 2u * p.stage_bytes, "kp_tapconv: tile does not fit shared memory (Cout_pad=%d)", d->Cout_pad);
    uint32_t budget = cta_budget(ctas_per_sm) - epi_bytes;
    if (const char* e = getenv("KP_TAPCONV_SMEM_KB")) budget = (uint32_t)atoi(e) * 1024u;
    int stages = (int)(budget / p.stage_bytes);
    if (stages < 2) stages = 2;
    if (stages > 8) stages = 8;
    p.stages = stages;
    p.out = out;
    p.out_off = d->out_off; p.out_sw = d->out_sw; p.out_sh = d->out_sh; p.out_sn = d->out_sn;
    p.Cout = d->Cout; p.cout_pad = d->Cout_pad; p.out_f32 = d->out_f32; p.act = d->act; p.alpha = d->alpha; p.accumulate = d->accumulate;
    p.bias = bias; p.ssum = ssum; p.ssq = ssq;

    const size_t smem = (size_t)stages * p.stage_bytes + (2 * stages + 4) * 8 + 16 + 1024 + epi_bytes;
    unsigned long long* trace = nullptr;
#ifdef KP_TRACE   // debug builds only (nvcc -DKP_TRACE): the shipped library never allocates device memory
    if (getenv("KP_TAPCONV_TRACE")) {
        cudaMalloc(&trace, 24 * 8 * sizeof(unsigned long long));
        cudaMemset(trace, 0, 24 * 8 * sizeof(unsigned long long));
    }
#endif
    p.dbg = trace;
    int grid = device_sm_count() * ctas_per_sm;
    if (grid > p.total_tiles * p.ksplit) grid = p.total_tiles * p.ksplit;
    const int epi = (p.ksplit == 1 && epi_fast_ok(d, out, 1)) ? (ssum != nullptr ? 2 : 1) : 0;
    p.slope = epi_fast_slope(d);
#define KP_LAUNCH_TAPCONV_E(CBV, E)                                                                                 \
    do {                                                                                                            \
        static bool attr_done = false;                                                                              \
        if (!attr_done) {                                                                                           \
            KP_CUDA_CHECK(cudaFuncSetAttribute(tapconv_kernel<CBV, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               227 * 1024));                                                        \
            attr_done = true;                                                                                       \
        }                                                                                                           \
        KP_CUDA_CHECK(launch_pdl(tapconv_kernel<CBV, E>, dim3(grid), dim3(192), smem, st, p));                      \
    } while (0)
#define KP_LAUNCH_TAPCONV(CBV)                                                                                      \
    do {                                                                                                            \
        if (epi == 2) KP_LAUNCH_TAPCONV_E(CBV, 2);                                                                  \
        else if (epi == 1) KP_LAUNCH_TAPCONV_E(CBV, 1);                                                             \
        else KP_LAUNCH_TAPCONV_E(CBV, 0);                                                                           \
    } while (0)
    if (CB == 64) KP_LAUNCH_TAPCONV(64);
    else if (CB == 32) KP_LAUNCH_TAPCONV(32);
    else KP_LAUNCH_TAPCONV(16);
#undef KP_LAUNCH_TAPCONV_E
#undef KP_LAUNCH_TAPCONV
    KP_LAUNCHED();
#ifdef KP_TRACE
    if (trace != nullptr) {
        // debug only: synchronise, print the per-tile timeline of CTA 0 (ns relative to the first event)
        unsigned long long h[24 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 24 * 8; ++i) if (h[i] != 0 && h[i] < t0) t0 = h[i];
        fprintf(stderr, "tapconv trace grid=%d ctas/sm=%d stages=%d BN=%d tiles=%d iters=%d: tile: prod_start prod_issued mma_start mma_issued epi_wait epi_go epi_done first_ld_done (ns)\n",
                grid, ctas_per_sm, stages, BN, p.total_tiles, total_iters);
        for (int t = 0; t < 24; ++t) {
            if (h[t * 8] == 0) break;
            fprintf(stderr, "  %2d:", t);
            for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", h[t * 8 + k] ? (long long)(h[t * 8 + k] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
        cudaFree(trace);
    }
#endif
    return KP_OK;
}

}  // namespace kp
