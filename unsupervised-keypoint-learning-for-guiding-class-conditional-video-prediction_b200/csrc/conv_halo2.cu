// TMA-staged halo-tile implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): the stride-1 KHxKW layers, forward
// and data gradient (reference: layers.conv, models/networks/layers.py:4-10; Vgg19.conv_layer, vgg.py:48-55).
//
// Why a second halo kernel: the TMA-tap kernel (conv_tc.cu) moves every input element L2 -> shared memory once PER TAP
// (9x for a 3x3 layer) and sits on the L2 -> SM fill rate (~45 B/clk/SM, profiles/README.md third pass); the first halo
// kernel (conv_halo.cu) removed the re-fill but gathers its un-swizzled plane layout with 16-byte cp.async at only
// ~16 B/clk/SM.  Here the input halo of a tile travels ONCE, by ONE TMA tile-mode box per (source, 16/32/64-channel
// slot) with the hardware swizzle of its row width, and the taps are start-address offsets of the shared-memory matrix
// descriptor:
//   * a CTA owns an 8 x (16*halves) pixel tile of one image; the halo box is [R = 16*halves + KH-1][pitch = 8 + KW-1]
//     pixels x nch channels, i.e. R*pitch rows of RB = nch*2 bytes (32/64/128) written with SWIZZLE_32B/64B/128B;
//   * the A operand of tap (dh,dw) is the K-major swizzled matrix whose 16 8-row groups start at pixel
//     (dh + g)*pitch + dw: descriptor start address = slot + (dh*pitch + dw)*RB, SBO = pitch*RB, base-offset field 0.
//     The tensor core applies the swizzle XOR to the ABSOLUTE shared-memory address bits - exactly what the TMA unit
//     did when writing - so a start address that is not aligned to the 8-row swizzle atom, and 8-row groups whose phase
//     differs from group to group (pitch = 10 rows), read back the right 16-byte chunks.  Verified bit-exactly for
//     all three swizzle widths and pitches 10/12/16 by scripts/micro/sw128_shift.cu (the base-offset field must stay 0;
//     setting it to the atom phase gives wrong results).
//   * halves = 2: two M=128 accumulators (tile rows 0-15 / 16-31) share every weight box;
//   * weights: [BN][64] boxes of the packed matrix (128-byte swizzle) - all of them resident in shared memory for the
//     life of the CTA when they fit, otherwise streamed through a ring, one tap group per stage;
//   * warp roles: 0 = activation TMA, 1 = weight TMA, 2/3 = MMA issuers, one per accumulator half (warp-uniform issue,
//     elected lane), 4-7 = epilogue
//     (conv_epilogue.cuh, TMEM accumulators double buffered whenever they fit); 256 threads, up to two CTAs per SM so
//     that the small-channel layers have two independent issue/epilogue pipelines per SM.
// L2 -> SM bytes per output pixel of a 3x3 layer: (1.33*Cin + Ktot*BN/256)*2 instead of (9*Cin + Ktot*BN/128)*2.
#include "kp_tc.cuh"
#include "conv_epilogue.cuh"
#include "kp_internal.h"
#include <cudaTypedefs.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace kp {

constexpr int H2_MAX_SLOTS = 24;
constexpr int H2_THREADS = 256;

struct alignas(64) Halo2KParams {
    CUtensorMap mapA[KP_MAX_MAPS];
    CUtensorMap mapB;
    CUtensorMap mapO;                                // fast epilogue: output store boxes {CW channels, 8, 4, 1}
    int st_nbuf;                                     // staging sets per epilogue warp (1 or 2)
    int n_slots;
    int sl_kofs[H2_MAX_SLOTS];
    short sl_c0[H2_MAX_SLOTS], sl_nch[H2_MAX_SLOTS];
    unsigned char sl_map[H2_MAX_SLOTS];
    int n_taps;
    unsigned short tapoff[KP_MAX_TAPS];              // dh*pitch + dw of each tap (pixels inside the halo)
    int dh_min, dw_min, Kper;
    int R, pitch;
    uint32_t a_slot_bytes, b_bytes, b_stage_bytes;
    int NSA, NSB, halves, acc_bufs;
    uint32_t rcp_ntiles, rcp_tpi, rcp_tw;            // ceil(2^32/d): exact x/d for x*d < 2^32
    int TB, resident;
    int t9;                                          // dense 3x3 taps at pitch 10: 1 forward order, 2 flipped (data gradient), 0 other
    int tiles_w, tiles_h, n_tiles, total_tiles;
    int Ho, Wo, N, BN, tmem_cols;
    int sgroups, group_n;                            // batch-norm statistics per batch segment of group_n images
    void* out;
    long long out_off, out_sw, out_sh, out_sn;
    int Cout, cout_pad, out_f32, act, accumulate, ksplit;
    float alpha, slope;
    const float* bias;
    float* ssum;
    float* ssq;
    unsigned long long* dbg;   // KP_TRACE builds: per-tile %globaltimer stamps of CTA 0
};

#ifdef KP_TRACE
__device__ __forceinline__ unsigned long long h2_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KP_H2TRACE(slot, idx)                                                                          \
    do {                                                                                               \
        if (p.dbg != nullptr && blockIdx.x == 0 && (idx) < 24) p.dbg[(idx) * 8 + (slot)] = h2_gtime(); \
    } while (0)
#else
#define KP_H2TRACE(slot, idx) do { } while (0)
#endif

__device__ __forceinline__ int h2_fdiv(int x, uint32_t rcp) { return rcp == 0u ? x : (int)__umulhi((uint32_t)x, rcp); }

// MMA issue loop of ONE accumulator half (HF), run by a whole warp with one elected lane issuing.  A separate function
// template per half keeps every operand provably warp-uniform for the compiler (anything derived from threadIdx would
// be routed through vector registers and R2UR before each UTCHMMA).
template <int HALVES, int KS, int HF, int T9>
__device__ __forceinline__ void h2_issue(const Halo2KParams& p, const uint32_t tmem, const uint32_t smem_base, const uint32_t a_base,
                                         uint64_t* fullA, uint64_t* emptyA, uint64_t* fullB, uint64_t* emptyB, uint64_t* tfull,
                                         uint64_t* tempty) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc_bf16(128, p.BN, 0, 0);
    const uint32_t b_hi = (uint32_t)(umma_smem_desc(0u, 1024u, 16u, 2u) >> 32);
    const int n_slots = p.n_slots, n_taps = p.n_taps, TB = p.TB, NSA = p.NSA, NSB = p.NSB, acc_bufs = p.acc_bufs;
    const uint32_t BN = (uint32_t)p.BN, b_box16 = p.b_bytes >> 4, pitch = (uint32_t)p.pitch;
    const bool resident = p.resident != 0;
    uint32_t ga = 0, gb = 0;
    bool b_ready = false;
    int lt = 0;
    for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x, ++lt) {
        const int acc = acc_bufs == 2 ? (lt & 1) : 0;
        if (lt >= acc_bufs) mbar_wait(&tempty[acc], ((lt / acc_bufs) - 1) & 1);
        tc_fence_after();
        if (HF == 0 && leader) KP_H2TRACE(2, lt);
        const uint32_t d0 = tmem + (uint32_t)(acc * HALVES + HF) * BN;
        uint32_t accum = 0;
        for (int s = 0; s < n_slots; ++s, ++ga) {
            const uint32_t stA = ga % (uint32_t)NSA;
            // slot geometry: rows of RB = nch*2 bytes with the swizzle of that width; 8-pixel groups pitch rows apart
            const uint32_t nch = KS > 0 ? 16u * (uint32_t)KS : (uint32_t)p.sl_nch[s];
            const uint32_t rb16 = nch >> 3;                                   // row bytes / 16
            const uint32_t lay = nch == 64u ? 2u : nch == 32u ? 4u : 6u;
            const uint32_t a_hi = (uint32_t)(umma_smem_desc(0u, pitch * nch * 2u, 16u, lay) >> 32);
            const int ksteps = KS > 0 ? KS : (int)(nch >> 4);
            mbar_wait(&fullA[stA], (ga / (uint32_t)NSA) & 1);
            tc_fence_after();
            if (HF == 0 && leader && s == 0) KP_H2TRACE(3, lt);
            // low descriptor word: start address (16-byte units) | LBO field (unused for swizzled K-major: 1)
            const uint32_t a_lo0 = (((a_base + stA * p.a_slot_bytes) >> 4) + (uint32_t)HF * 16u * pitch * rb16) | (1u << 16);
            if (T9) {
                // resident weights + the dense 3x3 tap grid at pitch 10 (every 3x3 layer, forward or flipped): fully unrolled,
                // tap offsets are immediates, one uniform multiply-add per tap and one add per operand per MMA
                if (!b_ready) {
                    mbar_wait(&fullB[0], 0);
                    tc_fence_after();
                    b_ready = true;
                }
                const uint32_t b_lo0 = ((smem_base >> 4) + (uint32_t)(s * 9) * b_box16) | (1u << 16);
                const bool rev = p.t9 == 2;
                const uint32_t a_org = rev ? a_lo0 + 22u * rb16 : a_lo0;
                const int astep = rev ? -(int)rb16 : (int)rb16;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const uint32_t a_lo = a_org + (uint32_t)(astep * ((t / 3) * 10 + (t % 3)));
                    const uint32_t b_lo = b_lo0 + (uint32_t)t * b_box16;
#pragma unroll
                    for (int kk = 0; kk < (KS > 0 ? KS : 1); ++kk)
                        umma_bf16_if_split(leader, d0, a_lo + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc, (t | kk) == 0 ? accum : 1u);
                }
                accum = 1u;
            } else
            for (int t0 = 0; t0 < n_taps; t0 += TB) {
                uint32_t stB = 0, b_lo;
                if (resident) {
                    if (!b_ready) {
                        mbar_wait(&fullB[0], 0);
                        tc_fence_after();
                        b_ready = true;
                    }
                    b_lo = ((smem_base >> 4) + (uint32_t)(s * n_taps + t0) * b_box16) | (1u << 16);
                } else {
                    stB = gb % (uint32_t)NSB;
                    mbar_wait(&fullB[stB], (gb / (uint32_t)NSB) & 1);
                    tc_fence_after();
                    b_lo = ((smem_base + stB * p.b_stage_bytes) >> 4) | (1u << 16);
                }
                const int t1 = min(t0 + TB, n_taps);
#pragma unroll 1      // keep the loop on the uniform datapath (unrolled, it runs out of uniform registers -> R2UR per MMA)
                for (int t = t0; t < t1; ++t, b_lo += b_box16) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)p.tapoff[t] * rb16;
                    if (KS > 0) {
#pragma unroll
                        for (int kk = 0; kk < (KS > 0 ? KS : 1); ++kk)
                            umma_bf16_if_split(leader, d0, a_lo + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc, kk == 0 ? accum : 1u);
                    } else {
                        for (int kk = 0; kk < ksteps; ++kk)
                            umma_bf16_if_split(leader, d0, a_lo + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc, kk == 0 ? accum : 1u);
                    }
                    accum = 1u;
                }
                if (!resident) {
                    umma_commit_if(leader, &emptyB[stB]);
                    ++gb;
                }
            }
            umma_commit_if(leader, &emptyA[stA]);
        }
        umma_commit_if(leader, &tfull[acc]);
        if (HF == 0 && leader) KP_H2TRACE(4, lt);
    }
}

// KS: K steps (16 channels each) per slot when all slots are alike (1, 2, 4), 0 = per slot.
// EPI: 0 = general epilogue (epi_chunk); fast epilogues (TMA store): 1 = bias + activation, 2 = bias + batch-norm statistics,
// 3 / 4 = the same with the statistics accumulated in registers (BN = 16 / 32).
// T9: 1 = the unrolled 3x3 issue loop (resident weights, see h2_issue).
template <int HALVES, int KS, int EPI, int T9>
__global__ void __launch_bounds__(H2_THREADS, 2) halo2_kernel(const __grid_constant__ Halo2KParams p) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    // [B region (1024-aligned)] [A slots (1024-aligned each)] [store staging (fast epilogue)] [barriers] [tmem slot] [bias] [stats]
    const uint32_t b_region = p.resident ? (((uint32_t)(p.n_slots * p.n_taps) * p.b_bytes + 1023u) & ~1023u)
                                         : (uint32_t)p.NSB * p.b_stage_bytes;
    const uint32_t a_base = smem_base + b_region;
    const uint32_t stage_bytes = EPI != 0 ? 4u * (uint32_t)p.st_nbuf * HALVES * 32u * (uint32_t)(p.BN < 64 ? p.BN : 64) * 2u : 0u;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(base + (size_t)b_region + (size_t)p.NSA * p.a_slot_bytes + stage_bytes);
    uint64_t* emptyA = fullA + p.NSA;
    uint64_t* fullB = emptyA + p.NSA;
    uint64_t* emptyB = fullB + p.NSB;
    uint64_t* tfull = emptyB + p.NSB;
    uint64_t* tempty = tfull + 2;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>(tslot + 4);
    float* s_stat = s_bias + p.cout_pad;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TILE_ROWS = 16 * HALVES;
    const int tiles_per_image = p.tiles_w * p.tiles_h;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.NSA; ++s) {
            mbar_init(&fullA[s], 1);
            mbar_init(&emptyA[s], HALVES);      // one commit per MMA-issuing warp
        }
        for (int s = 0; s < p.NSB; ++s) {
            mbar_init(&fullB[s], 1);
            mbar_init(&emptyB[s], HALVES);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], HALVES);
            mbar_init(&tempty[a], 4);
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 2) tmem_alloc(tslot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    pdl_wait();

    if (warp == 0) {
        // ------------------------------- activation halo boxes (TMA) -------------------------------
        if (lane == 0) {
            for (int m = 0; m < KP_MAX_MAPS; ++m) tma_prefetch_desc(&p.mapA[m]);
            uint32_t ga = 0;
            int ltp = 0;
            for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x, ++ltp) {
                const int mt = h2_fdiv(work, p.rcp_ntiles);
                const int n = h2_fdiv(mt, p.rcp_tpi), r = mt - n * tiles_per_image;
                const int tr = h2_fdiv(r, p.rcp_tw);
                const int y0 = tr * TILE_ROWS + p.dh_min, x0 = (r - tr * p.tiles_w) * 8 + p.dw_min;
                for (int s = 0; s < p.n_slots; ++s, ++ga) {
                    const uint32_t st = ga % (uint32_t)p.NSA;
                    if (s == 0) KP_H2TRACE(0, ltp);
                    if (ga >= (uint32_t)p.NSA) mbar_wait(&emptyA[st], ((ga / (uint32_t)p.NSA) - 1) & 1);
                    if (s == 0) KP_H2TRACE(1, ltp);
                    mbar_arrive_expect_tx(&fullA[st], (uint32_t)(p.R * p.pitch) * (uint32_t)p.sl_nch[s] * 2u);
                    tma_load_4d(base + (size_t)b_region + (size_t)st * p.a_slot_bytes, &p.mapA[p.sl_map[s]], &fullA[st],
                                (int)p.sl_c0[s], x0, y0, n);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------- weight boxes (TMA) -------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&p.mapB);
            if (p.resident) {
                mbar_arrive_expect_tx(&fullB[0], (uint32_t)(p.n_slots * p.n_taps) * p.b_bytes);
                for (int s = 0; s < p.n_slots; ++s)
                    for (int t = 0; t < p.n_taps; ++t)
                        tma_load_2d(base + (size_t)(s * p.n_taps + t) * p.b_bytes, &p.mapB, &fullB[0], t * p.Kper + p.sl_kofs[s], 0);
            } else {
                uint32_t gb = 0;
                for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x) {
                    const int n_off = (work - h2_fdiv(work, p.rcp_ntiles) * p.n_tiles) * p.BN;
                    for (int s = 0; s < p.n_slots; ++s) {
                        for (int t0 = 0; t0 < p.n_taps; t0 += p.TB, ++gb) {
                            const int nt = min(p.TB, p.n_taps - t0);
                            const uint32_t st = gb % (uint32_t)p.NSB;
                            if (gb >= (uint32_t)p.NSB) mbar_wait(&emptyB[st], ((gb / (uint32_t)p.NSB) - 1) & 1);
                            mbar_arrive_expect_tx(&fullB[st], (uint32_t)nt * p.b_bytes);
                            for (int i = 0; i < nt; ++i)
                                tma_load_2d(base + (size_t)st * p.b_stage_bytes + (size_t)i * p.b_bytes, &p.mapB, &fullB[st],
                                            (t0 + i) * p.Kper + p.sl_kofs[s], n_off);
                        }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------- MMA issuers -------------------------------
        // One warp per accumulator half (the halves are independent M=128 tiles over the same operands).  The issue loop
        // is what bounds the short-K layers (an N <= 64 MMA occupies the tensor pipe for 8-32 cycles, one thread issues one
        // per ~40 cycles at best), hence: two issuing warps, K steps per slot as a template parameter (no branches between
        // MMAs), descriptors kept as (lo, hi) halves so that a step is one uniform add per operand.
        h2_issue<HALVES, KS, 0, T9>(p, tmem, smem_base, a_base, fullA, emptyA, fullB, emptyB, tfull, tempty);
    } else if (warp == 3) {
        if (HALVES == 2) h2_issue<HALVES, KS, 1, T9>(p, tmem, smem_base, a_base, fullA, emptyA, fullB, emptyB, tfull, tempty);
    } else if (warp >= 4) {
        // ------------------------------- epilogue -------------------------------
        const int q = warp & 3;                   // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const int th = row >> 3, tw = row & 7;
        const int et = threadIdx.x - 128;
        for (int i = et; i < p.cout_pad; i += 128) s_bias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
        if (p.ssum != nullptr)
            for (int i = et; i < 8 * p.sgroups * p.cout_pad; i += 128) s_stat[i] = 0.f;     // four warp-private copies of [G][2][cout_pad]
        named_bar_sync(1, 128);
        float* const s_stat_w0 = s_stat + (warp - 4) * p.sgroups * 2 * p.cout_pad;
        const int group_n = p.group_n, gstride = 2 * p.cout_pad;
        if (EPI != 0) {
            // fast epilogue: everything that is constant for the launch sits in registers, the chunk is branch-free.
            // The bf16 tile leaves through shared memory: each warp stages its 4 x 8 pixels of both halves, CW = min(BN, 64)
            // channels at a time, in the swizzled layout of a TMA box {CW, 8, 4, 1} (conflict-free 16-byte writes) and one
            // lane stores the boxes.  A direct store would touch 32 different 128-byte lines per warp instruction (a pixel
            // per thread) and the epilogue, not the tensor pipe, bounded every layer with <= 64 output channels.
            const EpiFast ef = {p.slope, p.cout_pad};
            const int Ho = p.Ho, Wo = p.Wo, BN = p.BN, n_tiles = p.n_tiles, tiles_w = p.tiles_w, acc_bufs = p.acc_bufs;
            const uint32_t rcp_ntiles = p.rcp_ntiles, rcp_tpi = p.rcp_tpi, rcp_tw = p.rcp_tw;
            const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
            const int CW = BN < 64 ? BN : 64;
            const uint32_t rb = (uint32_t)CW * 2u, unit = 32u * rb, set_bytes = HALVES * unit;
            const uint32_t nbuf = (uint32_t)p.st_nbuf;
            const uint32_t stage_w = a_base + (uint32_t)p.NSA * p.a_slot_bytes + (uint32_t)(warp - 4) * nbuf * set_bytes;
            // 16-byte piece j of staging row `lane` sits at piece j ^ swz (SWIZZLE_128B / 64B / 32B of the row width)
            const uint32_t swz = rb == 128u ? (uint32_t)(lane & 7) : rb == 64u ? (uint32_t)((lane >> 1) & 3) : (uint32_t)((lane >> 2) & 1);
            const uint32_t row_addr = (uint32_t)lane * rb;
            uint32_t sbuf = 0;
            // Batch-norm statistics.  EPI 2: one transpose-reduce per 16-column chunk (epi_stats16).  EPI 3/4 (BN = 16/32, one
            // channel tile): every thread keeps running sums of its pixels' columns in registers across ALL its tiles and the
            // transpose-reduce happens once per CTA (and when the tiles move to the next statistics segment): 4 packed
            // operations per column pair and tile instead of ~170 instructions per chunk.
            constexpr int SRN = EPI == 3 ? 1 : EPI == 4 ? 2 : 0;
            constexpr bool STATS = EPI >= 2;
            float2 as2[SRN > 0 ? SRN : 1][8], aq2[SRN > 0 ? SRN : 1][8];
#pragma unroll
            for (int ci = 0; ci < (SRN > 0 ? SRN : 1); ++ci)
#pragma unroll
                for (int j = 0; j < 8; ++j) as2[ci][j] = aq2[ci][j] = make_float2(0.f, 0.f);
            int cur_seg = -1;
            int lt = 0;
            for (int work = blockIdx.x;; work += gridDim.x, ++lt) {
                const bool more = work < p.total_tiles;
                const int mt = h2_fdiv(work, rcp_ntiles), nt = work - mt * n_tiles;
                const int n = h2_fdiv(mt, rcp_tpi), r = mt - n * tiles_per_image;
                if (SRN > 0) {
                    const int seg = more ? n / group_n : -2;
                    if (seg != cur_seg) {
                        if (cur_seg >= 0) {
                            float* const sw = s_stat_w0 + cur_seg * gstride;
#pragma unroll
                            for (int ci = 0; ci < SRN; ++ci) {
                                float sv[16], qv[16];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    sv[2 * j] = as2[ci][j].x; sv[2 * j + 1] = as2[ci][j].y;
                                    qv[2 * j] = aq2[ci][j].x; qv[2 * j + 1] = aq2[ci][j].y;
                                    as2[ci][j] = aq2[ci][j] = make_float2(0.f, 0.f);
                                }
                                const float s1 = warp_colsum16(sv, lane), s2 = warp_colsum16(qv, lane);
                                if ((lane & 1) == 0) {
                                    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                                    sw[ci * 16 + col] += s1;
                                    sw[ef.cout_pad + ci * 16 + col] += s2;
                                }
                            }
                            __syncwarp();
                        }
                        cur_seg = seg;
                    }
                }
                if (!more) break;
                const int tr = h2_fdiv(r, rcp_tw);
                const int h0 = tr * TILE_ROWS, w0 = (r - tr * tiles_w) * 8;
                const int n_off = nt * BN;
                const int acc = acc_bufs == 2 ? (lt & 1) : 0;
                float* const s_stat_w = s_stat_w0 + (n / group_n) * gstride;       // this image's statistics segment
                if (et == 0) KP_H2TRACE(5, lt);
                mbar_wait(&tfull[acc], (lt / acc_bufs) & 1);
                tc_fence_after();
                if (et == 0) KP_H2TRACE(6, lt);
                // chunk-major: the same 16 columns of both accumulator halves are in registers together
                const bool edge = (h0 + TILE_ROWS > Ho) || (w0 + 8 > Wo);      // warp-uniform: the tile overhangs the image
                bool valid[HALVES];
#pragma unroll
                for (int hf = 0; hf < HALVES; ++hf) valid[hf] = (h0 + hf * 16 + th < Ho) && (w0 + tw < Wo);
                const uint32_t t_row = t_lane + (uint32_t)(acc * HALVES * BN);
                auto chunk = [&](const int c0, const int ci) {
                    float v[HALVES][16];
                    if (HALVES == 2) {
                        // both halves' loads in flight before the first use
                        uint32_t r0[16], r1[16];
                        tmem_ld16_issue(t_row + (uint32_t)c0, r0);
                        tmem_ld16_issue(t_row + (uint32_t)(BN + c0), r1);
                        tmem_ld_wait2(r0, r1);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            v[0][j] = __uint_as_float(r0[j]);
                            v[HALVES - 1][j] = __uint_as_float(r1[j]);
                        }
                    } else {
                        tmem_ld16(t_row + (uint32_t)c0, v[0]);
                    }
                    if (c0 + 16 >= BN) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
                    const int cl = c0 & (CW - 1);                  // channel inside the current store box
                    if (cl == 0) {
                        // the staging set about to be rewritten: its previous store must have read it
                        if (lane == 0) {
                            if (nbuf == 2u) bulk_wait_group_read<1>();
                            else bulk_wait_group_read<0>();
                        }
                        __syncwarp();
                    }
                    const uint32_t set = stage_w + sbuf * set_bytes + row_addr;
                    const uint32_t pc = (uint32_t)(cl >> 3);
#pragma unroll
                    for (int hf = 0; hf < HALVES; ++hf)
                        epi_chunk_fast_smem<!STATS>(v[hf], set + hf * unit + ((pc ^ swz) << 4), set + hf * unit + (((pc + 1u) ^ swz) << 4),
                                                    s_bias + n_off + c0, ef);
                    if (cl + 16 == CW) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
#pragma unroll
                            for (int hf = 0; hf < HALVES; ++hf)
                                tma_store_4d(&p.mapO, stage_w + sbuf * set_bytes + hf * unit, n_off + c0 - cl, w0, h0 + hf * 16 + q * 4, n);
                            bulk_commit_group();
                        }
                        sbuf ^= nbuf - 1u;
                    }
                    if (STATS) {
                        // an output pixel outside the image still saw real halo pixels: keep it out of the statistics
                        if (edge) {
#pragma unroll
                            for (int hf = 0; hf < HALVES; ++hf)
                                if (!valid[hf]) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[hf][j] = 0.f;
                                }
                        }
                        if (SRN > 0) {
#pragma unroll
                            for (int hf = 0; hf < HALVES; ++hf)
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float2 x = make_float2(v[hf][2 * j], v[hf][2 * j + 1]);
                                    as2[ci][j] = __fadd2_rn(as2[ci][j], x);
                                    aq2[ci][j] = __ffma2_rn(x, x, aq2[ci][j]);
                                }
                        } else {
                            epi_stats16<HALVES>(v, lane, s_stat_w + n_off + c0, ef.cout_pad);
                        }
                    }
                };
                if (SRN > 0) {
#pragma unroll
                    for (int ci = 0; ci < SRN; ++ci) chunk(ci * 16, ci);
                } else {
                    for (int c0 = 0; c0 < BN; c0 += 16) chunk(c0, 0);
                }
                if (et == 0) KP_H2TRACE(7, lt);
            }
            if (lane == 0) bulk_wait_group<0>();       // the stores read shared memory: keep the CTA alive until they are done
        } else {
            int lt = 0;
            for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x, ++lt) {
                const int mt = h2_fdiv(work, p.rcp_ntiles), nt = work - mt * p.n_tiles;
                const int n = h2_fdiv(mt, p.rcp_tpi), r = mt - n * tiles_per_image;
                const int tr = h2_fdiv(r, p.rcp_tw);
                const int h0 = tr * TILE_ROWS, w0 = (r - tr * p.tiles_w) * 8;
                const int n_off = nt * p.BN;
                float* const s_stat_w = s_stat_w0 + (n / group_n) * gstride;
                const int acc = p.acc_bufs == 2 ? (lt & 1) : 0;
                if (et == 0) KP_H2TRACE(5, lt);
                mbar_wait(&tfull[acc], (lt / p.acc_bufs) & 1);
                tc_fence_after();
                if (et == 0) KP_H2TRACE(6, lt);
    #pragma unroll
                for (int hf = 0; hf < HALVES; ++hf) {
                    const int uh = h0 + hf * 16 + th, uw = w0 + tw;
                    const bool valid = (uh < p.Ho) && (uw < p.Wo);
                    const long long pix = p.out_off + (long long)n * p.out_sn + (long long)uh * p.out_sh + (long long)uw * p.out_sw;
                    const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * HALVES + hf) * p.BN);
                    for (int c0 = 0; c0 < p.BN; c0 += 16) {
                        float v[16];
                        __syncwarp();
                        tmem_ld16(t_row + (uint32_t)c0, v);
                        if (hf == HALVES - 1 && c0 + 16 >= p.BN) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty[acc]);
                        }
                        if (p.ssum != nullptr && !valid) {
                            // an output pixel outside the image still sees real halo pixels: keep it out of the statistics
    #pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0.f;
                        }
                        epi_chunk(p, v, n_off + c0, valid, pix, lane, 0, s_bias, s_stat_w);
                    }
                }
                if (et == 0) KP_H2TRACE(7, lt);
            }
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
    if (warp == 2) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int device_sm_count();

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int h2_mode() {
    // KP_TAPCONV_HALO2: 0 = off, 1 = heuristics (default), 2 = every eligible layer (tests force every branch)
    const char* e = getenv("KP_TAPCONV_HALO2");      // read per launch: the tests switch kernels inside one process
    return e != nullptr ? atoi(e) : 1;
}

// Stride-1 multi-tap layers whose taps all read the same un-strided views, sources in slots of 16/32/64 channels.
bool halo2_eligible(const kp_tapconv_desc* d, const float* ssum) {
    (void)ssum;
    const int mode = h2_mode();
    if (mode == 0) return false;
    if (d->TW > 0 || d->TH > 0 || d->TN > 0 || d->BN > 0) return false;          // explicit tile request: TMA-tap kernel
    if (d->n_maps != d->n_src || d->n_taps < 2 || d->Ho < 16 || d->Wo < 8) return false;
    if (d->Cout_pad > 256 && d->Cout_pad % 128 != 0) return false;
    int dhmin = 127, dhmax = -128, dwmin = 127, dwmax = -128;
    for (int t = 0; t < d->n_taps; ++t) {
        if (d->map_first[t] != 0) return false;
        dhmin = d->dh[t] < dhmin ? d->dh[t] : dhmin; dhmax = d->dh[t] > dhmax ? d->dh[t] : dhmax;
        dwmin = d->dw[t] < dwmin ? d->dw[t] : dwmin; dwmax = d->dw[t] > dwmax ? d->dw[t] : dwmax;
    }
    if (dhmax - dhmin > 8 || dwmax - dwmin > 8) return false;
    int slots = 0, kinds = 0, cin = 0;
    for (int m = 0; m < d->n_maps; ++m) {
        const kp_tap_view& v = d->map[m];
        if (v.C < 16 || v.C % 16 != 0 || v.sw % 8 != 0 || v.sh % 8 != 0 || v.sn % 8 != 0 || v.off % 8 != 0) return false;
        // one tensor map per (source, slot width): a source is cut into 64-channel slots plus at most one narrower rest
        const int rest = v.C % 64;
        if (rest != 0 && rest != 16 && rest != 32) return false;
        slots += v.C / 64 + (rest ? 1 : 0);
        kinds += (v.C >= 64 ? 1 : 0) + (rest ? 1 : 0);
        cin += v.C;
    }
    if (slots > H2_MAX_SLOTS || kinds > KP_MAX_MAPS) return false;
    // Measured on B200 at batch 32 (scripts/halo2_bench.py, profiles/README.md round 2): with the fast epilogue this kernel
    // is 1.04x (512->512 @16x16) to 2.3x (16->32 @128x128) faster than the TMA-tap kernel on every stride-1 layer of the
    // stage-1 graph, so it takes everything eligible; KP_HALO2_MIN_TILES / KP_HALO2_MAX_COUT restrict it for experiments.
    if (mode >= 2) return true;
    const long long tiles = (long long)d->N * ((d->Wo + 7) / 8) * ((d->Ho + 31) / 32);
    if (const char* e = getenv("KP_HALO2_MIN_TILES")) if (tiles < atoi(e)) return false;
    if (const char* e = getenv("KP_HALO2_MAX_COUT")) if (d->Cout_pad > atoi(e)) return false;
    return true;
}

int halo2_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                 float* ssum, float* ssq, cudaStream_t st) {
    Halo2KParams p;
    memset(&p, 0, sizeof(p));
    KP_REQUIRE(d->Ktot % d->n_taps == 0, "kp_tapconv(halo2): Ktot %d is not a multiple of the tap count %d", d->Ktot, d->n_taps);
    p.Kper = d->Ktot / d->n_taps;
    p.halves = d->Ho >= 32 ? 2 : 1;
    int BN = d->Cout_pad;
    const int bn_cap = p.halves == 2 ? 128 : 256;
    if (BN > bn_cap) {
        BN = bn_cap;
        while (BN > 16 && d->Cout_pad % BN != 0) BN >>= 1;
    }
    KP_REQUIRE(BN % 16 == 0 && d->Cout_pad % BN == 0, "kp_tapconv(halo2): Cout_pad=%d does not tile by %d", d->Cout_pad, BN);
    p.BN = BN;
    p.n_tiles = d->Cout_pad / BN;
    int tm = 32;
    p.acc_bufs = 2 * p.halves * BN <= 512 ? 2 : 1;
    while (tm < p.acc_bufs * p.halves * BN) tm <<= 1;
    p.tmem_cols = tm;

    int dhmin = 127, dhmax = -128, dwmin = 127, dwmax = -128;
    for (int t = 0; t < d->n_taps; ++t) {
        dhmin = d->dh[t] < dhmin ? d->dh[t] : dhmin; dhmax = d->dh[t] > dhmax ? d->dh[t] : dhmax;
        dwmin = d->dw[t] < dwmin ? d->dw[t] : dwmin; dwmax = d->dw[t] > dwmax ? d->dw[t] : dwmax;
    }
    p.n_taps = d->n_taps;
    p.dh_min = dhmin; p.dw_min = dwmin;
    p.R = 16 * p.halves + (dhmax - dhmin);
    p.pitch = 8 + (dwmax - dwmin);
    for (int t = 0; t < d->n_taps; ++t) p.tapoff[t] = (unsigned short)((d->dh[t] - dhmin) * p.pitch + (d->dw[t] - dwmin));

    void* fn = nullptr;
    {
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        KP_REQUIRE(e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr,
                   "kp_tapconv(halo2): cuTensorMapEncodeTiled entry point unavailable");
    }
    EncodeTiledFn3 encode = reinterpret_cast<EncodeTiledFn3>(fn);

    // sources -> channel slots (K offsets follow the packed-weight layout: per tap, per source segment padded to CB) and
    // one 4-D tensor map {C, W, H, N} with box {nch, pitch, R, 1} per (source, slot width)
    int kbase = 0, ns = 0, nm = 0, max_nch = 16;
    for (int m = 0; m < d->n_maps; ++m) {
        const kp_tap_view& v = d->map[m];
        KP_REQUIRE(v.src >= 0 && v.src < KP_MAX_MAPS && src[v.src] != nullptr, "kp_tapconv(halo2): map %d has no source", m);
        const char* basep = reinterpret_cast<const char*>(src[v.src]) + v.off * 2;
        KP_REQUIRE((reinterpret_cast<uintptr_t>(basep) & 15) == 0, "kp_tapconv(halo2): source %d not 16-byte aligned", m);
        int c0 = 0, last_nch = 0, last_map = -1;
        while (c0 < v.C) {
            int nch = 64;
            while (nch > v.C - c0) nch >>= 1;
            KP_REQUIRE(nch >= 16, "kp_tapconv(halo2): channel slot narrower than 16 (C=%d)", v.C);
            KP_REQUIRE(ns < H2_MAX_SLOTS, "kp_tapconv(halo2): too many channel slots");
            if (nch != last_nch) {
                KP_REQUIRE(nm < KP_MAX_MAPS, "kp_tapconv(halo2): too many tensor maps");
                cuuint64_t gdim[4] = {(cuuint64_t)v.C, (cuuint64_t)v.Wd, (cuuint64_t)v.Hd, (cuuint64_t)d->N};
                cuuint64_t gstr[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
                cuuint32_t box[4] = {(cuuint32_t)nch, (cuuint32_t)p.pitch, (cuuint32_t)p.R, 1u};
                cuuint32_t estr[4] = {1, 1, 1, 1};
                const CUtensorMapSwizzle swz = nch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : nch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                                  : CU_TENSOR_MAP_SWIZZLE_32B;
                CUresult r = encode(&p.mapA[nm], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(basep), gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    set_error("kp_tapconv(halo2): cuTensorMapEncodeTiled(activations) failed with %d (C=%d W=%d H=%d N=%d box %d,%d,%d)",
                              (int)r, v.C, v.Wd, v.Hd, d->N, nch, p.pitch, p.R);
                    return KP_ERR_DRIVER;
                }
                last_map = nm++;
                last_nch = nch;
            }
            p.sl_map[ns] = (unsigned char)last_map; p.sl_c0[ns] = (short)c0; p.sl_nch[ns] = (short)nch; p.sl_kofs[ns] = kbase + c0;
            max_nch = nch > max_nch ? nch : max_nch;
            ++ns;
            c0 += nch;
        }
        kbase += (v.C + d->CB - 1) / d->CB * d->CB;
    }
    for (int m = nm; m < KP_MAX_MAPS; ++m) p.mapA[m] = p.mapA[0];      // prefetch-safe fillers
    KP_REQUIRE(kbase == p.Kper, "kp_tapconv(halo2): channel segments (%d) do not add up to Kper=%d", kbase, p.Kper);
    p.n_slots = ns;

    p.a_slot_bytes = ((uint32_t)(p.R * p.pitch) * (uint32_t)max_nch * 2u + 1023u) & ~1023u;
    p.b_bytes = (uint32_t)BN * 128u;
    auto rcp32 = [](int d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + (unsigned long long)d - 1ull) / (unsigned long long)d); };
    const int G = d->stat_groups > 1 ? d->stat_groups : 1;
    KP_REQUIRE(d->N % G == 0, "kp_tapconv(halo2): %d images do not split into %d statistics segments", d->N, G);
    p.sgroups = G; p.group_n = d->N / G;
    // bias (+ four warp-private [G][2][Cout_pad] statistics)
    const uint32_t epi_bytes = (ssum != nullptr ? 1u + 8u * (uint32_t)G : 1u) * (uint32_t)d->Cout_pad * sizeof(float);
    const uint32_t all_b = (uint32_t)(ns * d->n_taps) * p.b_bytes;
    const uint32_t fixed = epi_bytes + 1024u + 512u;
    // statistics together with an activation (no layer of the stage-1 graph) take the general epilogue
    const bool fast = epi_fast_ok(d, out, 1) && !(ssum != nullptr && d->act != KP_ACT_NONE);
    // fast epilogue: store staging, per epilogue warp st_nbuf sets of [halves][32 pixels][CW channels] bf16.  Two sets
    // (a store in flight while the next box is written) when the CTA owns the SM and the boxes are small.
    const uint32_t CW = BN < 64 ? (uint32_t)BN : 64u;
    auto nbuf_of = [&](bool two_) {
        if (const char* e = getenv("KP_HALO2_STAGE_NBUF")) return atoi(e) >= 2 ? 2u : 1u;
        return (two_ || BN > 64) ? 1u : 2u;
    };
    bool one_set = false;                              // two sets did not fit: fall back to one
    auto budget_of = [&](bool two_) {
        const uint32_t stage = fast ? 4u * (one_set ? 1u : nbuf_of(two_)) * (uint32_t)p.halves * 32u * CW * 2u : 0u;
        return (two_ ? 111u : 222u) * 1024u - fixed - stage;
    };
    // Two CTAs per SM whenever one CTA fits half of the SM's shared memory and TMEM: two independent TMA / issue /
    // epilogue pipelines per SM for the short-K layers.
    bool two = p.tmem_cols <= 256;
    uint32_t budget = budget_of(two);
    p.resident = (p.n_tiles == 1 && all_b <= (two ? 56u : 100u) * 1024u) ? 1 : 0;
    if (two && !p.resident && p.n_tiles == 1 && all_b <= 100u * 1024u) {      // resident weights beat a second CTA
        two = false;
        budget = budget_of(false);
        p.resident = 1;
    }
    if (const char* e = getenv("KP_HALO_RESIDENT")) if (atoi(e) == 0) p.resident = 0;
    uint32_t b_region;
    for (int attempt = 0;; ++attempt) {
        if (p.resident) {
            p.TB = d->n_taps; p.b_stage_bytes = (uint32_t)d->n_taps * p.b_bytes; p.NSB = 1;
            b_region = (all_b + 1023u) & ~1023u;
            p.NSA = b_region < budget ? (int)((budget - b_region) / p.a_slot_bytes) : 0;
            if (p.NSA > 4) p.NSA = 4;
        } else {
            // weight stages of ~32 KB (TB taps under one barrier), the rest goes to activation slots (2-3)
            p.TB = (int)(32u * 1024u / p.b_bytes);
            if (p.TB < 1) p.TB = 1;
            if (p.TB > d->n_taps) p.TB = d->n_taps;
            const int groups = (d->n_taps + p.TB - 1) / p.TB;
            p.TB = (d->n_taps + groups - 1) / groups;
            p.b_stage_bytes = (uint32_t)p.TB * p.b_bytes;
            p.NSA = 2;
            p.NSB = budget > 2u * p.a_slot_bytes ? (int)((budget - 2u * p.a_slot_bytes) / p.b_stage_bytes) : 0;
            if (p.NSB > 6) p.NSB = 6;
            b_region = (uint32_t)p.NSB * p.b_stage_bytes;
            if (p.NSB >= 4 && budget - b_region >= 3u * p.a_slot_bytes) p.NSA = 3;
        }
        const bool fits = p.NSA >= 2 && p.NSB >= 1 && (p.resident || p.NSB >= 2);
        if (fits) break;
        if (fast && !one_set && nbuf_of(two) == 2u) {
            one_set = true;
            budget = budget_of(two);
            --attempt;
            continue;
        }
        KP_REQUIRE(two && attempt == 0, "kp_tapconv(halo2): tile does not fit shared memory");
        two = false;                                   // retry with the whole SM
        budget = budget_of(false);
    }
    {
        KP_REQUIRE((reinterpret_cast<uintptr_t>(wpacked) & 15) == 0, "kp_tapconv(halo2): packed weights not 16-byte aligned");
        cuuint64_t gdim[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Cout_pad};
        cuuint64_t gstr[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wpacked), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("kp_tapconv(halo2): cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
            return KP_ERR_DRIVER;
        }
    }
    p.tiles_w = (d->Wo + 7) / 8;
    p.tiles_h = (d->Ho + 16 * p.halves - 1) / (16 * p.halves);
    p.total_tiles = d->N * p.tiles_w * p.tiles_h * p.n_tiles;
    p.rcp_ntiles = rcp32(p.n_tiles); p.rcp_tpi = rcp32(p.tiles_w * p.tiles_h); p.rcp_tw = rcp32(p.tiles_w);
    KP_REQUIRE((long long)p.total_tiles * (p.tiles_w * p.tiles_h > p.n_tiles ? p.tiles_w * p.tiles_h : p.n_tiles) < (1ll << 32),
               "kp_tapconv(halo2): too many tiles for the 32-bit work decode");
    p.Ho = d->Ho; p.Wo = d->Wo; p.N = d->N;
    p.out = out;
    p.out_off = d->out_off; p.out_sw = d->out_sw; p.out_sh = d->out_sh; p.out_sn = d->out_sn;
    p.Cout = d->Cout; p.cout_pad = d->Cout_pad; p.out_f32 = d->out_f32; p.act = d->act; p.alpha = d->alpha;
    p.accumulate = d->accumulate; p.ksplit = 1;
    p.bias = bias; p.ssum = ssum; p.ssq = ssq;
    p.slope = epi_fast_slope(d);
    p.st_nbuf = one_set ? 1 : (int)nbuf_of(two);
    const uint32_t stage_bytes = fast ? 4u * (uint32_t)p.st_nbuf * (uint32_t)p.halves * 32u * CW * 2u : 0u;
    if (fast) {
        // epi_fast_ok: bf16 output, Cout == Cout_pad (multiple of 16), view offset and strides multiples of 8 elements
        cuuint64_t gdim[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->N};
        cuuint64_t gstr[3] = {(cuuint64_t)d->out_sw * 2, (cuuint64_t)d->out_sh * 2, (cuuint64_t)d->out_sn * 2};
        cuuint32_t box[4] = {CW, 8u, 4u, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUtensorMapSwizzle swz = CW == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CW == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        CUresult r = encode(&p.mapO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, reinterpret_cast<char*>(out) + d->out_off * 2, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("kp_tapconv(halo2): cuTensorMapEncodeTiled(output) failed with %d (C=%d W=%d H=%d N=%d strides %lld,%lld,%lld)", (int)r,
                      d->Cout, d->Wo, d->Ho, d->N, (long long)d->out_sw, (long long)d->out_sh, (long long)d->out_sn);
            return KP_ERR_DRIVER;
        }
    } else {
        p.mapO = p.mapB;
    }

    size_t smem = (size_t)b_region + (size_t)p.NSA * p.a_slot_bytes + stage_bytes + (size_t)(2 * p.NSA + 2 * p.NSB + 4) * 8 + 16 + epi_bytes +
                  1024;
    // a CTA that needs more than half of the TMEM columns must not share its SM (the second allocation would block)
    if (!two && smem < 116u * 1024u) smem = 116u * 1024u;
    KP_REQUIRE(smem <= 227u * 1024u, "kp_tapconv(halo2): shared memory %zu exceeds the SM", smem);
    int grid = device_sm_count() * (two ? 2 : 1);
    if (grid > p.total_tiles) grid = p.total_tiles;
    unsigned long long* trace = nullptr;
#ifdef KP_TRACE   // debug builds only (nvcc -DKP_TRACE): the shipped library never allocates device memory
    if (getenv("KP_TAPCONV_TRACE")) {
        cudaMalloc(&trace, 24 * 8 * sizeof(unsigned long long));
        cudaMemset(trace, 0, 24 * 8 * sizeof(unsigned long long));
    }
#endif
    p.dbg = trace;
    int ks = p.sl_nch[0] / 16;                       // all slots alike -> K steps per slot known at compile time
    for (int s = 1; s < ns; ++s)
        if (p.sl_nch[s] != p.sl_nch[0]) ks = 0;
    if (ks != 1 && ks != 2 && ks != 4) ks = 0;
    // fast epilogues: 1 = bias + activation, 2..4 = bias + batch-norm statistics (no activation; 3/4 = register-accumulated
    // for BN 16/32 with both halves)
    int epi = !fast ? 0 : ssum == nullptr ? 1 : 2;
    if (epi == 2 && p.halves == 2 && p.n_tiles == 1 && (BN == 16 || BN == 32) && !getenv("KP_HALO2_NO_STATREG")) epi = BN == 16 ? 3 : 4;
    if (epi == 0) ks = 0;                            // the general epilogue is only instantiated with the general issue loop
    // dense 3x3 tap grid at pitch 10 in forward or flipped order + resident weights: unrolled issue loop
    p.t9 = 0;
    if (d->n_taps == 9 && p.pitch == 10 && p.resident && ks > 0 && p.halves == 2 && !getenv("KP_HALO2_NO_T9")) {
        bool fwd = true, rev = true;
        for (int t = 0; t < 9; ++t) {
            const int f = (t / 3) * 10 + t % 3;
            fwd = fwd && p.tapoff[t] == f;
            rev = rev && p.tapoff[t] == 22 - f;
        }
        p.t9 = fwd ? 1 : rev ? 2 : 0;
    }
    const int t9 = p.t9 != 0 ? 1 : 0;
    typedef void (*KernelFn)(const Halo2KParams);
    KernelFn fnk = nullptr;
#define KP_H2_PICK(H, K, E, T) if (p.halves == H && ks == K && epi == E && t9 == T) fnk = halo2_kernel<H, K, E, T>
    KP_H2_PICK(1, 0, 0, 0); KP_H2_PICK(2, 0, 0, 0);
    KP_H2_PICK(1, 0, 1, 0); KP_H2_PICK(1, 1, 1, 0); KP_H2_PICK(1, 2, 1, 0); KP_H2_PICK(1, 4, 1, 0);
    KP_H2_PICK(1, 0, 2, 0); KP_H2_PICK(1, 1, 2, 0); KP_H2_PICK(1, 2, 2, 0); KP_H2_PICK(1, 4, 2, 0);
    KP_H2_PICK(2, 0, 1, 0); KP_H2_PICK(2, 1, 1, 0); KP_H2_PICK(2, 2, 1, 0); KP_H2_PICK(2, 4, 1, 0);
    KP_H2_PICK(2, 0, 2, 0); KP_H2_PICK(2, 1, 2, 0); KP_H2_PICK(2, 2, 2, 0); KP_H2_PICK(2, 4, 2, 0);
    KP_H2_PICK(2, 1, 1, 1); KP_H2_PICK(2, 2, 1, 1); KP_H2_PICK(2, 4, 1, 1);
    KP_H2_PICK(2, 1, 2, 1); KP_H2_PICK(2, 2, 2, 1); KP_H2_PICK(2, 4, 2, 1);
    KP_H2_PICK(2, 0, 3, 0); KP_H2_PICK(2, 1, 3, 0); KP_H2_PICK(2, 2, 3, 0); KP_H2_PICK(2, 4, 3, 0);
    KP_H2_PICK(2, 0, 4, 0); KP_H2_PICK(2, 1, 4, 0); KP_H2_PICK(2, 2, 4, 0); KP_H2_PICK(2, 4, 4, 0);
    KP_H2_PICK(2, 1, 3, 1); KP_H2_PICK(2, 2, 3, 1); KP_H2_PICK(2, 4, 3, 1);
    KP_H2_PICK(2, 1, 4, 1); KP_H2_PICK(2, 2, 4, 1); KP_H2_PICK(2, 4, 4, 1);
#undef KP_H2_PICK
    KP_REQUIRE(fnk != nullptr, "kp_tapconv(halo2): no kernel variant for halves=%d ks=%d epi=%d t9=%d", p.halves, ks, epi, t9);
    {
        // once per variant: allow the full shared memory
        static KernelFn done[64];
        static int n_done = 0;
        bool seen = false;
        for (int i = 0; i < n_done; ++i) seen = seen || done[i] == fnk;
        if (!seen) {
            KP_CUDA_CHECK(cudaFuncSetAttribute(reinterpret_cast<const void*>(fnk), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            if (n_done < 64) done[n_done++] = fnk;
        }
    }
    KP_CUDA_CHECK(launch_pdl(fnk, dim3(grid), dim3(H2_THREADS), smem, st, p));
    KP_LAUNCHED();
#ifdef KP_TRACE
    if (trace != nullptr) {
        unsigned long long h[24 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 24 * 8; ++i) if (h[i] != 0 && h[i] < t0) t0 = h[i];
        fprintf(stderr, "halo2 trace grid=%d two=%d halves=%d ks=%d slots=%d taps=%d BN=%d NSA=%d NSB=%d resident=%d tiles=%d: tile: prod_wait prod_go "
                "mma_start mma_data mma_issued epi_wait epi_go epi_done (ns)\n", grid, (int)two, p.halves, ks, p.n_slots, p.n_taps, p.BN, p.NSA, p.NSB,
                p.resident, p.total_tiles);
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
