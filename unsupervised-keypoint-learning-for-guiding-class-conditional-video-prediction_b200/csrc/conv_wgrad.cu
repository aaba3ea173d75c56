// Weight gradient of the tap-GEMM convolutions on tcgen05 tensor cores (sm_100a).
//
//   dW[tap][ci][co] += sum_pixels X_tap[pixel][ci] * dY[pixel][co]
//
// M = ci (128 rows per CTA), N = co (BN <= 256), K = pixels.  Both operands arrive exactly as in the forward
// kernel - TMA 4-D boxes [KP pixels][CB channels], hardware swizzled - and are consumed as MN-major UMMA
// operands (channels contiguous), so no transposed copy of X or dY exists anywhere.
//
// TAP SHARING: one CTA accumulates T taps of the same pixel range at once (T accumulators of BN columns in TMEM):
// the dY tile is loaded once per pipeline stage and reused by the T shifted X tiles, which divides the dY traffic
// and the number of pipeline round trips per FLOP by T (the first version launched one CTA per tap and was bound
// by L2 re-reads of dY for the small-channel 128x128 layers and by its atomic epilogue everywhere).
// grid = (ci blocks x co blocks, tap groups, pixel splits); fp32 partial tiles are accumulated into the HWIO
// gradient with vectorised red.global.add.v4.f32.
#include "kp_tc.cuh"
#include "kp_internal.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace kp {

// pixels per pipeline stage: chosen so that every TMA box [KP pixels][CB channels] is 8 KB (KP = 64 / 128 / 256)
static inline int wg_kp(int CB) { return 4096 / CB; }

struct alignas(64) WgradKParams {
    CUtensorMap mapX[KP_MAX_MAPS];
    CUtensorMap mapDY;
    signed char dh[KP_MAX_TAPS], dw[KP_MAX_TAPS], mf[KP_MAX_TAPS];
    int tap_flat[KP_MAX_TAPS];
    int TW, TH, TN, tiles_w, tiles_h, total_tiles, tiles_per_split;
    int n_taps, T, R;                    // taps per CTA (group size); taps covered by ONE M=128 MMA (R > 1 when Cin <= 64)
    int Cin, Cout, BN, co_blocks, tmem_cols, stages;
    int n_a_max, n_b_max;
    uint32_t box_bytes, ybox_bytes, stage_bytes;
    float* dw_out;
    long long dw_off, dw_stap, dw_sci;
    unsigned long long* dbg;   // KP_TAPCONV_TRACE: per-stage timestamps of CTA (0,0,0) (debug only)
};

__device__ __forceinline__ unsigned long long wg_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KP_WTRACE(slot, idx)                                                                                              \
    do {                                                                                                                  \
        if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 24) p.dbg[(idx) * 8 + (slot)] = wg_gtime(); \
    } while (0)


__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// CB = channel block of the X operand (rows of dW), CBY = channel block of the dY operand (columns of dW): each follows
// its own channel count (64 if divisible by 64, else 32, else 16), so a 64->8 layer moves X in 128-byte rows even
// though dY only has 16-byte... 32-byte rows.  Pixels per stage follow X (its boxes are 8 KB).
template <int CB, int CBY>
__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradKParams p) {
    constexpr uint32_t ROW_BYTES = CB * 2;
    constexpr uint32_t SBO = 8 * ROW_BYTES;
    constexpr uint32_t LAYOUT = (CB == 64) ? 2u : (CB == 32) ? 4u : 6u;
    constexpr uint32_t KSTEP_BYTES = 16 * ROW_BYTES;  // 16 pixels per UMMA K-step
    constexpr uint32_t ROW_BYTES_Y = CBY * 2;
    constexpr uint32_t SBO_Y = 8 * ROW_BYTES_Y;
    constexpr uint32_t LAYOUT_Y = (CBY == 64) ? 2u : (CBY == 32) ? 4u : 6u;
    constexpr uint32_t KSTEP_BYTES_Y = 16 * ROW_BYTES_Y;
    constexpr int KP = 4096 / CB;                     // pixels per stage

    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)p.stages * p.stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const int ci0 = (blockIdx.x / p.co_blocks) * 128;
    const int co0 = (blockIdx.x % p.co_blocks) * p.BN;
    const int tap0 = blockIdx.y * p.T;
    const int nt = min(p.T, p.n_taps - tap0);
    const int t_begin = blockIdx.z * p.tiles_per_split;
    const int t_end = min(t_begin + p.tiles_per_split, p.total_tiles);
    const int n_iters = t_end - t_begin;  // host guarantees >= 1

    const int n_a = min(128 / CB, (p.Cin - ci0 + CB - 1) / CB);
    const int n_b = (min(p.BN, p.Cout - co0) + CBY - 1) / CBY;
    // stage layout: [dY: n_b_max boxes][X of tap 0: n_a_max boxes][X of tap 1] ...
    const uint32_t x_region = (uint32_t)p.n_a_max * p.box_bytes;
    const uint32_t dy_region = (uint32_t)p.n_b_max * p.ybox_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 1) tmem_alloc(tslot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    pdl_wait();

    if (warp == 0) {
        // TMA producer, the WHOLE warp: each lane issues the boxes b = lane, lane+32, ... of a stage.  One thread issuing
        // the 7-19 loads of a stage one after the other was the bottleneck of the weight gradient (timeline trace:
        // ~600 ns of issue per 56 KB stage = 30 B/clk/SM, well under the 50 B/clk/SM fill rate).
        if (lane == 0) {
            tma_prefetch_desc(&p.mapDY);
            for (int m = 0; m < KP_MAX_MAPS; ++m)
                if (m == 0 || p.mf[tap0] == m) tma_prefetch_desc(&p.mapX[m]);
        }
        const uint32_t tx = (uint32_t)(nt * n_a) * p.box_bytes + (uint32_t)n_b * p.ybox_bytes;
        const int n_boxes = n_b + nt * n_a;
        for (int it = 0; it < n_iters; ++it) {
            const int tile = t_begin + it;
            const int w0 = (tile % p.tiles_w) * p.TW;
            const int h0 = ((tile / p.tiles_w) % p.tiles_h) * p.TH;
            const int n0 = (tile / (p.tiles_w * p.tiles_h)) * p.TN;
            const int st = it % S;
            if (it >= S) mbar_wait(&empty[st], ((it / S) - 1) & 1);
            if (lane == 0) KP_WTRACE(0, it);
            uint8_t* dy_dst = base + (size_t)st * p.stage_bytes;
            if (lane == 0) mbar_arrive_expect_tx(&full[st], tx);
            __syncwarp();                                  // the expected byte count is registered before any copy can land
            for (int bx = lane; bx < n_boxes; bx += 32) {
                if (bx < n_b) {
                    tma_load_4d(dy_dst + (size_t)bx * p.ybox_bytes, &p.mapDY, &full[st], co0 + bx * CBY, w0, h0, n0);
                } else {
                    const int j = bx - n_b;
                    const int ti = j / n_a, c = j - ti * n_a;
                    const int t = tap0 + ti;
                    tma_load_4d(dy_dst + dy_region + (size_t)ti * x_region + (size_t)c * p.box_bytes, &p.mapX[p.mf[t]], &full[st],
                                ci0 + c * CB, w0 + p.dw[t], h0 + p.dh[t], n0);
                }
            }
            if (lane == 0) KP_WTRACE(1, it);
        }
    } else if (warp == 1) {
        {
            const uint32_t leader = elect_one() ? 1u : 0u;   // warp-uniform loop, one elected lane issues (see kp_tc.cuh)
            const uint32_t idesc = umma_idesc_bf16(128, p.BN, 1, 1);  // both operands MN-major
            for (int it = 0; it < n_iters; ++it) {
                const int st = it % S;
                mbar_wait(&full[st], (it / S) & 1);
                tc_fence_after();
                if (lane == 0) KP_WTRACE(2, it);
                const uint32_t dy_addr = smem_base + (uint32_t)st * p.stage_bytes;
                // One M=128 MMA reads 128/CB consecutive X boxes: with Cin <= 64 these are the boxes of R consecutive
                // taps (rows = (tap, channel)), so ONE accumulator serves R taps and the MMA count drops R-fold.
                for (int ti = 0; ti < nt; ti += p.R) {
                    const uint32_t x_addr = dy_addr + dy_region + (uint32_t)ti * x_region;
                    const uint32_t d_tmem = tmem + (uint32_t)((ti / p.R) * p.BN);
#pragma unroll
                    for (int kk = 0; kk < KP / 16; ++kk) {
                        const uint64_t da = umma_smem_desc(x_addr + kk * KSTEP_BYTES, SBO, p.box_bytes, LAYOUT);
                        const uint64_t db = umma_smem_desc(dy_addr + kk * KSTEP_BYTES_Y, SBO_Y, p.ybox_bytes, LAYOUT_Y);
                        umma_bf16_if(leader, d_tmem, da, db, idesc, (it | kk) != 0 ? 1u : 0u);
                    }
                }
                if (lane == 0) KP_WTRACE(3, it);
                umma_commit_if(leader, &empty[st]);
            }
            umma_commit_if(leader, tfull);
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int rows_per_tap = p.n_a_max * CB;          // accumulator rows that belong to one tap
        const int t_in_group = row / rows_per_tap;        // 0 unless R > 1
        const int ci = ci0 + row - t_in_group * rows_per_tap;
        mbar_wait(tfull, 0);
        tc_fence_after();
        for (int ti = 0; ti < nt; ti += p.R) {
            const int tap_i = ti + t_in_group;
            const bool valid = (t_in_group < p.R) && (tap_i < nt) && (ci < p.Cin);
            float* orow = p.dw_out + p.dw_off + (long long)p.tap_flat[tap0 + (valid ? tap_i : ti)] * p.dw_stap + (long long)ci * p.dw_sci + co0;
            const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((ti / p.R) * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(t_row + (uint32_t)c0, v);
                if (valid) {
                    const int nvalid = p.Cout - co0 - c0;
                    float* o = orow + c0;
                    if (nvalid >= 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) red_add_v4(o + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < nvalid) atomicAdd(o + j, v[j]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

int device_sm_count();

int wgrad_launch(const kp_wgrad_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    KP_REQUIRE(d->CB == 16 || d->CB == 32 || d->CB == 64, "kp_wgrad: CB must be 16, 32 or 64 (got %d)", d->CB);
    KP_REQUIRE(d->n_maps >= 1 && d->n_maps <= KP_MAX_MAPS, "kp_wgrad: n_maps %d out of range", d->n_maps);
    KP_REQUIRE(d->n_taps >= 1 && d->n_taps <= KP_MAX_TAPS, "kp_wgrad: n_taps %d out of range", d->n_taps);
    KP_REQUIRE(d->N > 0 && d->Ho > 0 && d->Wo > 0 && d->Cin > 0 && d->Cout > 0, "kp_wgrad: empty problem");
    // stride-1 KHxKW layers: halo-tile kernel (X travels once per pixel tile instead of once per tap)
    if (wgrad2_eligible(d)) return wgrad2_launch(d, x, dy, dw, st);
    WgradKParams p;
    memset(&p, 0, sizeof(p));
    // operand channel blocks follow the operands' own channel counts (desc->CB is only validated)
    auto blk = [](int c) { return c % 64 == 0 ? 64 : c % 32 == 0 ? 32 : 16; };
    const int CB = blk(d->Cin), CBY = blk(d->Cout);
    choose_pixel_tile(wg_kp(CB), d->Wo, d->Ho, d->N, &p.TW, &p.TH, &p.TN);
    for (int m = 0; m < d->n_maps; ++m) {
        const int rc = encode_view_map(&p.mapX[m], d->map[m], x, d->N, CB, p.TW, p.TH, p.TN, "kp_wgrad X map");
        if (rc != KP_OK) return rc;
    }
    {
        const int rc = encode_view_map(&p.mapDY, d->dy, dy, d->N, CBY, p.TW, p.TH, p.TN, "kp_wgrad dY map");
        if (rc != KP_OK) return rc;
    }
    for (int t = 0; t < d->n_taps; ++t) {
        KP_REQUIRE(d->map_first[t] >= 0 && d->map_first[t] < d->n_maps, "kp_wgrad: tap %d map out of range", t);
        p.dh[t] = d->dh[t]; p.dw[t] = d->dw[t]; p.mf[t] = d->map_first[t]; p.tap_flat[t] = d->tap_flat[t];
    }
    p.n_taps = d->n_taps;
    p.tiles_w = (d->Wo + p.TW - 1) / p.TW;
    p.tiles_h = (d->Ho + p.TH - 1) / p.TH;
    p.total_tiles = p.tiles_w * p.tiles_h * ((d->N + p.TN - 1) / p.TN);
    p.Cin = d->Cin; p.Cout = d->Cout;
    const int cout_pad = (d->Cout + 15) / 16 * 16;
    p.BN = cout_pad <= 256 ? cout_pad : 256;
    p.co_blocks = (cout_pad + p.BN - 1) / p.BN;
    const int ci_blocks = (d->Cin + 127) / 128;
    p.box_bytes = (uint32_t)wg_kp(CB) * CB * 2u;                          // 8 KB
    p.ybox_bytes = (uint32_t)wg_kp(CB) * CBY * 2u;                        // 2-32 KB (same pixels, dY's block)
    // X regions hold only the channel chunks that exist (the MMA still reads 128 rows = 128/CB chunks at LBO stride:
    // rows past the loaded chunks alias the next tap's region / the next stage / the slack below and only produce
    // accumulator rows that are never stored)
    p.n_a_max = ((d->Cin < 128 ? d->Cin : 128) + CB - 1) / CB;
    p.n_b_max = (p.BN + CBY - 1) / CBY;
    // taps per CTA: limited by TMEM (T accumulators of BN columns) and by shared memory (<= 11 boxes = 88 KB per stage)
    p.R = (128 / CB) / p.n_a_max;            // taps per M=128 MMA
    if (p.R < 1) p.R = 1;
    if (const char* e = getenv("KP_WGRAD_TAPS_PER_MMA")) { const int c = atoi(e); if (c >= 1 && c < p.R) p.R = c; }
    int t_max = (512 / p.BN) * p.R;
    const size_t slack = (size_t)(128 / CB - p.n_a_max) * p.box_bytes;
    const uint32_t dy_bytes = (uint32_t)p.n_b_max * p.ybox_bytes;
    uint32_t stage_cap = (uint32_t)((215u * 1024u - slack) / 2);               // two stages must fit ...
    if (stage_cap > 64u * 1024u) stage_cap = 64u * 1024u;   // ... but 3-4 stages hide the TMA latency much better
    const int by_smem = stage_cap > dy_bytes ? (int)((stage_cap - dy_bytes) / ((uint32_t)p.n_a_max * p.box_bytes)) : 0;
    if (t_max > by_smem) t_max = by_smem;
    if (t_max > d->n_taps) t_max = d->n_taps;
    if (const char* e = getenv("KP_WGRAD_TMAX")) { const int c = atoi(e); if (c >= 1 && c < t_max) t_max = c; }
    if (t_max < 1) t_max = 1;
    const int groups = (d->n_taps + t_max - 1) / t_max;
    p.T = (d->n_taps + groups - 1) / groups;
    int tm = 32;
    while (tm < ((p.T + p.R - 1) / p.R) * p.BN) tm <<= 1;
    p.tmem_cols = tm;
    p.stage_bytes = ((uint32_t)(p.T * p.n_a_max) * p.box_bytes + dy_bytes + 1023u) & ~1023u;
    int stages = (int)((215u * 1024u - slack) / p.stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    const int ctas_per_sm = ((size_t)stages * p.stage_bytes <= 100u * 1024u && tm <= 256) ? 2 : 1;
    int splits = d->splits;
    const int base_ctas = ci_blocks * p.co_blocks * groups;
    if (splits <= 0) {
        const int target = device_sm_count() * ctas_per_sm;
        splits = (target + base_ctas - 1) / base_ctas;
        const int max_by_work = (p.total_tiles + 7) / 8;   // at least ~8 pixel tiles per CTA (amortises the red epilogue)
        if (splits > max_by_work) splits = max_by_work;
    }
    if (splits < 1) splits = 1;
    if (splits > p.total_tiles) splits = p.total_tiles;
    p.tiles_per_split = (p.total_tiles + splits - 1) / splits;
    splits = (p.total_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    if (stages > p.tiles_per_split) stages = p.tiles_per_split < 1 ? 1 : p.tiles_per_split;
    p.stages = stages;
    p.dw_out = dw;
    p.dw_off = d->dw_off; p.dw_stap = d->dw_stap; p.dw_sci = d->dw_sci;

    const size_t smem = (size_t)stages * p.stage_bytes + (2 * stages + 1) * 8 + 16 + 1024 + slack;
    KP_REQUIRE(smem <= 227u * 1024u, "kp_wgrad: shared memory %zu exceeds the SM (internal tiling error)", smem);
    dim3 grid((unsigned)(ci_blocks * p.co_blocks), (unsigned)groups, (unsigned)splits);
    unsigned long long* trace = nullptr;
#ifdef KP_TRACE   // debug builds only (nvcc -DKP_TRACE): the shipped library never allocates device memory
    if (getenv("KP_TAPCONV_TRACE")) {
        cudaMalloc(&trace, 24 * 8 * sizeof(unsigned long long));
        cudaMemset(trace, 0, 24 * 8 * sizeof(unsigned long long));
    }
#endif
    p.dbg = trace;
#define KP_LAUNCH_WGRAD(CBX, CBYV)                                                                                  \
    do {                                                                                                            \
        static bool attr_done = false;                                                                              \
        if (!attr_done) {                                                                                           \
            KP_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel<CBX, CBYV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               227 * 1024));                                                        \
            attr_done = true;                                                                                       \
        }                                                                                                           \
        KP_CUDA_CHECK(launch_pdl(wgrad_kernel<CBX, CBYV>, grid, dim3(192), smem, st, p));                            \
    } while (0)
#define KP_LAUNCH_WGRAD_Y(CBX)                         \
    do {                                               \
        if (CBY == 64) KP_LAUNCH_WGRAD(CBX, 64);       \
        else if (CBY == 32) KP_LAUNCH_WGRAD(CBX, 32);  \
        else KP_LAUNCH_WGRAD(CBX, 16);                 \
    } while (0)
    if (CB == 64) KP_LAUNCH_WGRAD_Y(64);
    else if (CB == 32) KP_LAUNCH_WGRAD_Y(32);
    else KP_LAUNCH_WGRAD_Y(16);
#undef KP_LAUNCH_WGRAD_Y
#undef KP_LAUNCH_WGRAD
    KP_LAUNCHED();
#ifdef KP_TRACE
    if (trace != nullptr) {
        unsigned long long h[24 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 24 * 8; ++i) if (h[i] != 0 && h[i] < t0) t0 = h[i];
        fprintf(stderr, "wgrad trace grid=(%u,%u,%u) CB=%d CBY=%d T=%d R=%d BN=%d stages=%d stage_bytes=%u tiles/split=%d: stage: prod_go prod_issued "
                "mma_data mma_issued (ns)\n", grid.x, grid.y, grid.z, CB, CBY, p.T, p.R, p.BN, stages, p.stage_bytes, p.tiles_per_split);
        for (int t = 0; t < 24; ++t) {
            if (h[t * 8] == 0) break;
            fprintf(stderr, "  %2d:", t);
            for (int k = 0; k < 4; ++k) fprintf(stderr, " %7lld", h[t * 8 + k] ? (long long)(h[t * 8 + k] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
        cudaFree(trace);
    }
#endif
    return KP_OK;
}

}  // namespace kp
