// Halo-tile weight gradient on tcgen05 tensor cores (sm_100a): the stride-1 KHxKW layers.
//
//   dW[tap][ci][co] += sum_pixels X[pixel + tap][ci] * dY[pixel][co]          (backward of layers.conv w.r.t. its kernel,
//                                                                               reference models/networks/layers.py:4-10)
//
// The first weight-gradient kernel (conv_wgrad.cu) loads one shifted X box per tap and pixel tile: for a 3x3 layer every
// activation travels L2 -> shared memory 9 times (measured 1.4 GB moved for 201 MB of operands on 128->64 @128x128) and the
// kernel runs at 300-450 TFLOP/s on the wide layers, 100-200 on the narrow ones.  Here the X operand of ALL taps comes from
// ONE halo box per pixel tile, exactly as in the forward halo kernel (conv_halo2.cu):
//   * pixel tile = 8 x 16 pixels of one image; X halo box [16+KH-1][8+KW-1] pixels x nch channels by one TMA tile-mode
//     load per 16/32/64-channel slot (hardware swizzle of the row width), dY box [16][8] pixels x Cout;
//   * GEMM per tap: M = ci (rows of dW), N = co, K = pixels; both operands are MN-major (channels contiguous), K runs over
//     pixel rows.  One K step = 16 pixels = two 8-pixel tile rows: the descriptor's stride between 8-row groups (SBO) is
//     the halo pitch, the tap (dh,dw) and the K step are start-address offsets (the tensor core applies the swizzle to the
//     absolute shared-memory address, see conv_halo2.cu / scripts/micro/sw128_shift.cu);
//   * M = 128 rows per MMA: with Cin >= 128 these are two 64-channel slots of one tap (LBO = slot stride); with Cin <= 64
//     the 128/nch channel blocks of one MMA are CONSECUTIVE TAPS of one kernel row - block b starts one pixel (= one
//     shared-memory row) after block b-1, so LBO = row bytes and a 3-wide kernel row of a 32-channel layer is ONE MMA;
//   * every tap group keeps its accumulators in TMEM over the CTA's whole pixel range; the epilogue runs once and adds the
//     fp32 tile into the HWIO gradient with red.global.add.v4.f32.
// grid = (ci blocks x co blocks, MMA groups of taps, pixel splits); warp 0 = TMA, warp 1 = MMA issue, warps 2-5 = epilogue.
#include "kp_tc.cuh"
#include "kp_internal.h"
#include <cudaTypedefs.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace kp {

constexpr int WG2_MAX_MG = 24;      // MMA groups (accumulators) per launch

struct alignas(64) Wgrad2KParams {
    CUtensorMap mapX, mapDY;
    int n_mg;                                   // MMA groups of the layer: one accumulator each
    unsigned short mg_off[WG2_MAX_MG];          // pixel offset (dh*pitch + dw) of the group's first tap inside the halo
    unsigned char mg_tap0[WG2_MAX_MG], mg_ntaps[WG2_MAX_MG];   // first tap (row-major index) and taps covered (blocks used)
    int tap_flat[KP_MAX_TAPS];
    int T;                                      // MMA groups per CTA
    int nch, n_a, CBY, n_b;                     // X slot width / slots per ci block; dY block width / blocks per co block
    int pitch, R, dh_min, dw_min;
    int tiles_w, tiles_h, total_tiles, tiles_per_split;
    int Cin, Cout, BN, co_blocks, tmem_cols, stages;
    uint32_t a_slot_bytes, ybox_bytes, stage_bytes, lbo_a;
    float* dw_out;
    long long dw_off, dw_stap, dw_sci;
};

__device__ __forceinline__ void wg2_red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(192, 1) wgrad2_kernel(const __grid_constant__ Wgrad2KParams p) {
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
    const int mg0 = blockIdx.y * p.T;
    const int nmg = min(p.T, p.n_mg - mg0);
    const int t_begin = blockIdx.z * p.tiles_per_split;
    const int t_end = min(t_begin + p.tiles_per_split, p.total_tiles);
    const int n_iters = t_end - t_begin;  // host guarantees >= 1
    const int tiles_per_image = p.tiles_w * p.tiles_h;
    // stage layout: [X slots: n_a boxes of a_slot_bytes][dY: n_b boxes of ybox_bytes]
    const uint32_t x_region = (uint32_t)p.n_a * p.a_slot_bytes;

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
        // ------------------------------- TMA producer -------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&p.mapX);
            tma_prefetch_desc(&p.mapDY);
            const int na = min(p.n_a, (p.Cin - ci0 + p.nch - 1) / p.nch);
            const int nb = min(p.n_b, (p.Cout - co0 + p.CBY - 1) / p.CBY);
            const uint32_t tx = (uint32_t)na * (uint32_t)(p.R * p.pitch * p.nch * 2) + (uint32_t)nb * p.ybox_bytes;
            for (int it = 0; it < n_iters; ++it) {
                const int tile = t_begin + it;
                const int n = tile / tiles_per_image, r = tile - n * tiles_per_image;
                const int tr = r / p.tiles_w;
                const int h0 = tr * 16, w0 = (r - tr * p.tiles_w) * 8;
                const int st = it % S;
                if (it >= S) mbar_wait(&empty[st], ((it / S) - 1) & 1);
                uint8_t* dst = base + (size_t)st * p.stage_bytes;
                mbar_arrive_expect_tx(&full[st], tx);
                for (int a = 0; a < na; ++a)
                    tma_load_4d(dst + (size_t)a * p.a_slot_bytes, &p.mapX, &full[st], ci0 + a * p.nch, w0 + p.dw_min, h0 + p.dh_min, n);
                for (int b = 0; b < nb; ++b)
                    tma_load_4d(dst + x_region + (size_t)b * p.ybox_bytes, &p.mapDY, &full[st], co0 + b * p.CBY, w0, h0, n);
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer (warp-uniform, one elected lane) -------------------------------
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t idesc = umma_idesc_bf16(128, p.BN, 1, 1);          // both operands MN-major
        const uint32_t nch = (uint32_t)p.nch, rb = nch * 2u, rby = (uint32_t)p.CBY * 2u;
        const uint32_t lay_a = nch == 64u ? 2u : nch == 32u ? 4u : 6u;
        const uint32_t lay_b = p.CBY == 64 ? 2u : p.CBY == 32 ? 4u : 6u;
        // 8-pixel groups of the X halo are `pitch` rows apart; dY rows are dense
        const uint32_t a_hi = (uint32_t)(umma_smem_desc(0u, (uint32_t)p.pitch * rb, p.lbo_a, lay_a) >> 32);
        const uint32_t a_lbo = ((p.lbo_a >> 4) & 0x3FFFu) << 16;
        const uint32_t b_hi = (uint32_t)(umma_smem_desc(0u, 8u * rby, p.ybox_bytes, lay_b) >> 32);
        const uint32_t b_lbo = ((p.ybox_bytes >> 4) & 0x3FFFu) << 16;
        const uint32_t a_kstep16 = (2u * (uint32_t)p.pitch * rb) >> 4;      // one K step = two tile rows of the halo
        const uint32_t b_kstep16 = (16u * rby) >> 4;
        for (int it = 0; it < n_iters; ++it) {
            const int st = it % S;
            mbar_wait(&full[st], (it / S) & 1);
            tc_fence_after();
            const uint32_t x16 = (smem_base + (uint32_t)st * p.stage_bytes) >> 4;
            const uint32_t y16 = x16 + (x_region >> 4);
#pragma unroll 1
            for (int g = 0; g < nmg; ++g) {
                const uint32_t a0 = (x16 + (((uint32_t)p.mg_off[mg0 + g] * rb) >> 4)) | a_lbo;
                const uint32_t d_tmem = tmem + (uint32_t)(g * p.BN);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_bf16_if_split(leader, d_tmem, a0 + kk * a_kstep16, a_hi, (y16 + kk * b_kstep16) | b_lbo, b_hi, idesc,
                                       (it | kk) != 0 ? 1u : 0u);
            }
            umma_commit_if(leader, &empty[st]);
        }
        umma_commit_if(leader, tfull);
    } else {
        // ------------------------------- epilogue: once, after the last pixel tile -------------------------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int rows_per_tap = p.n_a * p.nch;           // accumulator rows of one tap (128 when Cin >= 128)
        const int blk = row / rows_per_tap;               // tap inside the MMA group
        const int ci = ci0 + row - blk * rows_per_tap;
        mbar_wait(tfull, 0);
        tc_fence_after();
        for (int g = 0; g < nmg; ++g) {
            const bool valid = blk < (int)p.mg_ntaps[mg0 + g] && ci < p.Cin;
            const int tap = (int)p.mg_tap0[mg0 + g] + (valid ? blk : 0);
            float* orow = p.dw_out + p.dw_off + (long long)p.tap_flat[tap] * p.dw_stap + (long long)ci * p.dw_sci + co0;
            const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(t_row + (uint32_t)c0, v);
                if (valid) {
                    const int nvalid = p.Cout - co0 - c0;
                    float* o = orow + c0;
                    if (nvalid >= 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) wg2_red_add_v4(o + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
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

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMapSwizzle wg2_swizzle(int nch) {
    return nch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : nch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// stride-1 layers with a dense KHxKW tap grid in row-major order, channel counts the slots can carry
bool wgrad2_eligible(const kp_wgrad_desc* d) {
    if (const char* e = getenv("KP_WGRAD_HALO"))
        if (atoi(e) == 0) return false;
    if (d->n_maps != 1 || d->n_taps < 2 || d->Ho < 16 || d->Wo < 8) return false;
    if (d->Cin % 16 != 0 || d->Cout % 16 != 0) return false;
    if (d->Cin > 64 && d->Cin % 64 != 0) return false;
    if (d->Cin < 64 && d->Cin != 16 && d->Cin != 32) return false;
    if (d->Cout > 64 && d->Cout % 64 != 0) return false;
    if (d->Cout < 64 && d->Cout != 16 && d->Cout != 32) return false;
    const kp_tap_view& v = d->map[0];
    if (v.C != d->Cin && v.C < d->Cin) return false;
    if (v.sw % 8 || v.sh % 8 || v.sn % 8 || v.off % 8) return false;
    if (d->dy.sw % 8 || d->dy.sh % 8 || d->dy.sn % 8 || d->dy.off % 8) return false;
    int dhmin = 127, dhmax = -128, dwmin = 127, dwmax = -128;
    for (int t = 0; t < d->n_taps; ++t) {
        if (d->map_first[t] != 0) return false;
        dhmin = d->dh[t] < dhmin ? d->dh[t] : dhmin; dhmax = d->dh[t] > dhmax ? d->dh[t] : dhmax;
        dwmin = d->dw[t] < dwmin ? d->dw[t] : dwmin; dwmax = d->dw[t] > dwmax ? d->dw[t] : dwmax;
    }
    const int KH = dhmax - dhmin + 1, KW = dwmax - dwmin + 1;
    if (KH > 9 || KW > 9 || KH * KW != d->n_taps) return false;
    for (int t = 0; t < d->n_taps; ++t)
        if (d->dh[t] != dhmin + t / KW || d->dw[t] != dwmin + t % KW) return false;
    if (const char* e = getenv("KP_WGRAD_HALO"))
        if (atoi(e) >= 2) return true;                       // tests: every eligible shape
    // Measured at batch 32 / 64 (scripts/wgrad_bench.py): 1.2-3.2x faster than the per-tap kernel wherever a launch has a
    // few hundred pixel tiles; KHx1 kernels of narrow layers (one tap per MMA, three quarters of its rows wasted) and the
    // 16x16 layers (128 tiles at batch 64) stay with the per-tap kernel.
    if (KW == 1 && d->Cin < 128) return false;
    const long long tiles = (long long)d->N * ((d->Wo + 7) / 8) * ((d->Ho + 15) / 16);
    return tiles >= 256;
}

int wgrad2_launch(const kp_wgrad_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    Wgrad2KParams p;
    memset(&p, 0, sizeof(p));
    void* fn = nullptr;
    {
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        KP_REQUIRE(e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr,
                   "kp_wgrad(halo): cuTensorMapEncodeTiled entry point unavailable");
    }
    EncodeTiledFn4 encode = reinterpret_cast<EncodeTiledFn4>(fn);
    int dhmin = 127, dhmax = -128, dwmin = 127, dwmax = -128;
    for (int t = 0; t < d->n_taps; ++t) {
        dhmin = d->dh[t] < dhmin ? d->dh[t] : dhmin; dhmax = d->dh[t] > dhmax ? d->dh[t] : dhmax;
        dwmin = d->dw[t] < dwmin ? d->dw[t] : dwmin; dwmax = d->dw[t] > dwmax ? d->dw[t] : dwmax;
        p.tap_flat[t] = d->tap_flat[t];
    }
    const int KH = dhmax - dhmin + 1, KW = dwmax - dwmin + 1;
    p.dh_min = dhmin; p.dw_min = dwmin;
    p.R = 16 + KH - 1;
    p.pitch = 8 + KW - 1;
    // X: slots of nch channels; a ci block of 128 rows is two 64-channel slots of one tap, or 128/nch consecutive taps
    p.nch = d->Cin >= 64 ? 64 : d->Cin;
    p.n_a = d->Cin >= 128 ? 2 : 1;
    const int ci_blocks = (d->Cin + 127) / 128;
    const int taps_per_mma = p.n_a == 2 ? 1 : 128 / p.nch;
    p.a_slot_bytes = ((uint32_t)(p.R * p.pitch * p.nch * 2) + 1023u) & ~1023u;
    p.lbo_a = p.n_a == 2 ? p.a_slot_bytes : (uint32_t)p.nch * 2u;
    // MMA groups: per kernel row, runs of taps_per_mma consecutive dw taps
    int n_mg = 0;
    for (int kh = 0; kh < KH; ++kh)
        for (int kw0 = 0; kw0 < KW; kw0 += taps_per_mma) {
            KP_REQUIRE(n_mg < WG2_MAX_MG, "kp_wgrad(halo): too many MMA groups");
            p.mg_off[n_mg] = (unsigned short)(kh * p.pitch + kw0);
            p.mg_tap0[n_mg] = (unsigned char)(kh * KW + kw0);
            p.mg_ntaps[n_mg] = (unsigned char)((KW - kw0) < taps_per_mma ? (KW - kw0) : taps_per_mma);
            ++n_mg;
        }
    p.n_mg = n_mg;
    // dY: blocks of CBY channels, BN columns per CTA
    p.CBY = d->Cout >= 64 ? 64 : d->Cout;
    const int cout_pad = (d->Cout + 15) / 16 * 16;
    p.BN = cout_pad <= 128 ? cout_pad : 128;
    p.co_blocks = (cout_pad + p.BN - 1) / p.BN;
    p.n_b = (p.BN + p.CBY - 1) / p.CBY;
    p.ybox_bytes = 128u * (uint32_t)p.CBY * 2u;
    p.Cin = d->Cin; p.Cout = d->Cout;
    // accumulators per CTA: TMEM holds 512 columns
    p.tiles_w = (d->Wo + 7) / 8;
    p.tiles_h = (d->Ho + 15) / 16;
    p.total_tiles = d->N * p.tiles_w * p.tiles_h;
    int T = 512 / p.BN;
    if (T > n_mg) T = n_mg;
    // few (ci, co) blocks: spread the tap groups over more CTAs until the machine is full (each group re-reads the
    // operands, but from L2) - a launch of 64 CTAs with all taps each loses to one of 192 CTAs with a third of the taps
    const int ctas_wanted = device_sm_count();
    while (T > 1 && (long long)ci_blocks * p.co_blocks * ((n_mg + T - 1) / T) * ((p.total_tiles + 3) / 4) < ctas_wanted) --T;
    int groups = (n_mg + T - 1) / T;
    T = (n_mg + groups - 1) / groups;
    p.T = T;
    int tm = 32;
    while (tm < T * p.BN) tm <<= 1;
    p.tmem_cols = tm;
    p.stage_bytes = ((uint32_t)p.n_a * p.a_slot_bytes + (uint32_t)p.n_b * p.ybox_bytes + 1023u) & ~1023u;
    // The last channel blocks of an M=128 MMA may lie beyond the slots that exist (Cin = 64: block 1 = next pixel row, fine;
    // Cin < 128 in a 2-slot layout cannot happen).  With n_a == 1 the MMA reads up to 128/nch - 1 pixel rows past the tap:
    // always inside the halo box for the taps that count; for the junk blocks of the last group it may run up to 7 rows past
    // the slot into the dY region / next stage - finite data, rows never stored.
    int stages = (int)((200u * 1024u) / p.stage_bytes);
    if (stages > 4) stages = 4;
    KP_REQUIRE(stages >= 2, "kp_wgrad(halo): stage of %u bytes does not fit twice", p.stage_bytes);
    {
        const kp_tap_view& v = d->map[0];
        const char* basep = reinterpret_cast<const char*>(x) + v.off * 2;
        KP_REQUIRE((reinterpret_cast<uintptr_t>(basep) & 15) == 0, "kp_wgrad(halo): X not 16-byte aligned");
        cuuint64_t gdim[4] = {(cuuint64_t)v.C, (cuuint64_t)v.Wd, (cuuint64_t)v.Hd, (cuuint64_t)d->N};
        cuuint64_t gstr[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.nch, (cuuint32_t)p.pitch, (cuuint32_t)p.R, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&p.mapX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(basep), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, wg2_swizzle(p.nch), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("kp_wgrad(halo): cuTensorMapEncodeTiled(X) failed with %d", (int)r);
            return KP_ERR_DRIVER;
        }
    }
    {
        const kp_tap_view& v = d->dy;
        const char* basep = reinterpret_cast<const char*>(dy) + v.off * 2;
        KP_REQUIRE((reinterpret_cast<uintptr_t>(basep) & 15) == 0, "kp_wgrad(halo): dY not 16-byte aligned");
        cuuint64_t gdim[4] = {(cuuint64_t)v.C, (cuuint64_t)v.Wd, (cuuint64_t)v.Hd, (cuuint64_t)d->N};
        cuuint64_t gstr[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.CBY, 8u, 16u, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&p.mapDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(basep), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, wg2_swizzle(p.CBY), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("kp_wgrad(halo): cuTensorMapEncodeTiled(dY) failed with %d", (int)r);
            return KP_ERR_DRIVER;
        }
    }
    int splits = d->splits;
    const int base_ctas = ci_blocks * p.co_blocks * groups;
    // two CTAs per SM when a CTA needs less than half of the SM (narrow layers: their MMAs are issue-bound, two issuing
    // warps per SM help)
    const bool two = (size_t)stages * p.stage_bytes <= 100u * 1024u && tm <= 256;
    if (splits <= 0) {
        splits = (device_sm_count() * (two ? 2 : 1) + base_ctas - 1) / base_ctas;
        const int max_by_work = (p.total_tiles + 3) / 4;   // at least ~4 pixel tiles per CTA (amortises the red epilogue)
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

    const size_t smem = (size_t)stages * p.stage_bytes + (2 * stages + 1) * 8 + 16 + 1024 + 2048;
    KP_REQUIRE(smem <= 227u * 1024u, "kp_wgrad(halo): shared memory %zu exceeds the SM (internal tiling error)", smem);
    dim3 grid((unsigned)(ci_blocks * p.co_blocks), (unsigned)groups, (unsigned)splits);
    static bool attr_done = false;
    if (!attr_done) {
        KP_CUDA_CHECK(cudaFuncSetAttribute(wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    KP_CUDA_CHECK(launch_pdl(wgrad2_kernel, grid, dim3(192), smem, st, p));
    KP_LAUNCHED();
    return KP_OK;
}

}  // namespace kp
