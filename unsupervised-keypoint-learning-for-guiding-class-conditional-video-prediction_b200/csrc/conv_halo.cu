// Halo-tile implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): the stride-1 KHxKW layers.
//
// Why: the TMA-tap kernel (conv_tc.cu) fetches one shifted activation box per tap, i.e. every input element travels
// L2 -> shared memory KH*KW times (and the weights once per 128-pixel tile).  Measured on B200, every 3x3 layer whose
// FLOPs do not hide it runs at the L2 -> SM rate of ~8 TB/s on 9x the input.  Here a CTA owns a (16*halves) x 8 pixel
// tile of ONE image and loads the input halo ((16*halves+KH-1) x (8+KW-1) pixels) ONCE per 64-channel slot; all taps
// read it in place:
//   * shared-memory layout of a slot: one PLANE per 8 channels, [halo pixel][16 bytes].  For a K-major operand WITHOUT
//     swizzle a UMMA core matrix is 8 rows x 16 bytes stored contiguously, so 8 consecutive pixels of a plane ARE a
//     core matrix; the 8-pixel groups of the 16 tile rows are `pitch*16` bytes apart (SBO), the two planes of a
//     16-channel K step are `plane_bytes` apart (LBO).  A tap (dh,dw) is nothing but a start-address offset of
//     (dh*pitch + dw)*16 bytes - no re-load, no swizzle phase to respect.
//   * halves = 2: two M=128 accumulators (rows 0-15 and 16-31 of the tile) share every weight box, which halves the
//     weight traffic per output pixel as well.
//   * the gather that produces the plane layout is done by 4 producer warps with 16-byte cp.async (zero-fill form for
//     the TF SAME padding and for image borders); weights arrive by TMA exactly as in conv_tc.cu ([BN][64] boxes of
//     the same packed matrix, 128-byte swizzle); one thread issues the MMAs; 4 warps run the shared epilogue
//     (conv_epilogue.cuh) on TMEM accumulators that are double buffered whenever 2*halves*BN <= 512 columns.
// L2 -> SM bytes per output pixel drop from (9*Cin + Ktot*BN/128)*2 to (1.33*Cin + Ktot*BN/256)*2 for a 3x3 layer.
#include "kp_tc.cuh"
#include "conv_epilogue.cuh"
#include "kp_internal.h"
#include <cudaTypedefs.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace kp {

constexpr int HALO_MAX_SLOTS = 24;
constexpr int HALO_THREADS = 320;        // 4 gather warps, 1 weight-TMA warp, 1 MMA warp, 4 epilogue warps

struct alignas(64) HaloKParams {
    CUtensorMap mapB;
    const __nv_bfloat16* src[KP_MAX_MAPS];
    long long s_sn[KP_MAX_MAPS], s_sh[KP_MAX_MAPS], s_sw[KP_MAX_MAPS];
    int s_H[KP_MAX_MAPS], s_W[KP_MAX_MAPS];
    int n_slots;
    int sl_kofs[HALO_MAX_SLOTS];
    short sl_c0[HALO_MAX_SLOTS], sl_nch[HALO_MAX_SLOTS];
    unsigned char sl_src[HALO_MAX_SLOTS];
    int n_taps;
    unsigned char dh[KP_MAX_TAPS], dw[KP_MAX_TAPS];   // offsets inside the halo (>= 0)
    int dh_min, dw_min, Kper;
    int R, pitch;
    uint32_t plane_bytes, a_slot_bytes, b_bytes, b_stage_bytes, pitch_rcp;
    int NSA, NSB, halves, acc_bufs;
    int gather_depth;
    uint32_t rcp_ntiles, rcp_tpi, rcp_tw;            // ceil(2^32/d): exact x/d for x*d < 2^32 (work decode without divisions)
    unsigned short tapoff[KP_MAX_TAPS];              // dh*pitch + dw of each tap (16-byte units inside the halo)
    int TB, tap_groups, resident;      // taps per weight stage, stages per slot; resident: all weights loaded once
    int tiles_w, tiles_h, n_tiles, total_tiles;
    int Ho, Wo, N, BN, tmem_cols;
    void* out;
    long long out_off, out_sw, out_sh, out_sn;
    int Cout, cout_pad, out_f32, act, accumulate, ksplit;
    float alpha;
    const float* bias;
    float* ssum;
    float* ssq;
    unsigned long long* dbg;   // KP_TAPCONV_TRACE: per-item timestamps of CTA 0 (debug only)
};

__device__ __forceinline__ unsigned long long halo_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KP_HTRACE(slot, idx)                                                                            \
    do {                                                                                                \
        if (p.dbg != nullptr && blockIdx.x == 0 && (idx) < 24) p.dbg[(idx) * 8 + (slot)] = halo_gtime(); \
    } while (0)

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// wait until at most `n` (0..3) of this thread's most recent cp.async groups are still in flight
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    if (n <= 0) cp_async_wait<0>();
    else if (n == 1) cp_async_wait<1>();
    else if (n == 2) cp_async_wait<2>();
    else cp_async_wait<3>();
}

__device__ __forceinline__ int fdiv(int x, uint32_t rcp) { return rcp == 0u ? x : (int)__umulhi((uint32_t)x, rcp); }   // rcp 0 = divide by 1

template <int HALVES>
__global__ void __launch_bounds__(HALO_THREADS, 1) haloconv_kernel(const __grid_constant__ HaloKParams p) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    // [B slots (1024-aligned)] [A slots] [barriers] [tmem slot] [bias] [stats]
    const uint32_t b_base = smem_base;
    const uint32_t b_region = p.resident ? (uint32_t)(p.n_slots * p.n_taps) * p.b_bytes : (uint32_t)p.NSB * p.b_stage_bytes;
    const uint32_t a_base = b_base + b_region;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(base + (size_t)b_region + (size_t)p.NSA * p.a_slot_bytes);
    uint64_t* emptyA = fullA + p.NSA;
    uint64_t* fullB = emptyA + p.NSA;
    uint64_t* emptyB = fullB + p.NSB;
    uint64_t* tfull = emptyB + p.NSB;
    uint64_t* tempty = tfull + 2;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>(tslot + 4);
    float* s_stat = s_bias + p.cout_pad;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_rows = 16 * HALVES;
    const int tiles_per_image = p.tiles_w * p.tiles_h;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.NSA; ++s) {
            mbar_init(&fullA[s], 128);    // every gather thread arrives after its copies have landed
            mbar_init(&emptyA[s], 1);
        }
        for (int s = 0; s < p.NSB; ++s) {
            mbar_init(&fullB[s], 1);
            mbar_init(&emptyB[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 5) tmem_alloc(tslot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    pdl_wait();

    if (warp < 4) {
        // ------------------------------- activation gather (cp.async) -------------------------------
        const int tid = threadIdx.x;
        const int halo_px = p.R * p.pitch;
        uint32_t ga = 0;
        // `depth` slots are in flight per thread, signalled in issue order, oldest first.  Measured: depth 2-3 does not
        // help the small-channel layers (the gather is not their bottleneck) and costs 64-channel slots 15-40 %
        // (more cp.async in flight than the LSU queue absorbs), so the default is 1.
        const int depth = min(p.gather_depth, p.NSA - 1);
        int pending = 0;
        int ltg = 0;
        for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x, ++ltg) {
            const int mt = fdiv(work, p.rcp_ntiles);
            const int n = fdiv(mt, p.rcp_tpi), r = mt - n * tiles_per_image;
            const int tr = fdiv(r, p.rcp_tw);
            const int y0 = tr * tile_rows + p.dh_min, x0 = (r - tr * p.tiles_w) * 8 + p.dw_min;
            for (int s = 0; s < p.n_slots; ++s, ++ga) {
                const int st = (int)(ga % (uint32_t)p.NSA);
                if (ga >= (uint32_t)p.NSA) {
                    const uint32_t par = ((ga / (uint32_t)p.NSA) - 1) & 1;
                    if (!mbar_try_wait(&emptyA[st], par)) {
                        // ring full: the MMA warp is the slow side.  Hand over everything still in flight before
                        // blocking, otherwise the consumer would wait for slots that have long landed.
                        while (pending > 0) {
                            cp_async_wait_dyn(pending - 1);
                            fence_proxy_async_smem();
                            mbar_arrive(&fullA[(int)((ga - (uint32_t)pending) % (uint32_t)p.NSA)]);
                            --pending;
                        }
                        mbar_wait(&emptyA[st], par);
                    }
                }
                if (tid == 0 && s == 0) KP_HTRACE(0, ltg);
                // (with one slot per item - every small-channel layer - s is always 0 and these loads are loop invariant)
                const int sq = p.n_slots == 1 ? 0 : s;
                const int m = p.sl_src[sq], nch = p.sl_nch[sq];
                const int nj = nch >> 3;                       // 16-byte chunks per pixel: 1, 2, 4 or 8
                const int j = tid & (nj - 1);
                const __nv_bfloat16* sbase = p.src[m] + (long long)n * p.s_sn[m] + p.sl_c0[sq] + 8 * j;
                const uint32_t dst0 = a_base + (uint32_t)st * p.a_slot_bytes + (uint32_t)j * p.plane_bytes;
                const int H = p.s_H[m], W = p.s_W[m];
                const int sh = (int)p.s_sh[m], sw = (int)p.s_sw[m];    // per-image offsets fit 32 bits (host checks)
                const int pstep = 128 / nj;
                for (int px = tid / nj; px < halo_px; px += pstep) {
                    const int py = (int)(((uint32_t)px * p.pitch_rcp) >> 16), pxx = px - py * p.pitch;
                    const int y = y0 + py, x = x0 + pxx;
                    const bool inb = ((unsigned)y < (unsigned)H) && ((unsigned)x < (unsigned)W);
                    const __nv_bfloat16* g = sbase + (inb ? y * sh + x * sw : 0);
                    cp_async_16_zfill(dst0 + (uint32_t)px * 16u, g, inb ? 16u : 0u);
                }
                if (nj == 1) {
                    // 8-channel slot: the 16-wide K step also reads plane 1 - keep it zero (the packed weights of the
                    // padded channels are zero, but 0 x stale NaN would poison the accumulator)
                    const uint32_t z0 = a_base + (uint32_t)st * p.a_slot_bytes + p.plane_bytes;
                    for (int px = tid; px < halo_px; px += 128)
                        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(z0 + (uint32_t)px * 16u), "r"(0u) : "memory");
                }
                cp_async_commit();
                if (tid == 0 && s == p.n_slots - 1) KP_HTRACE(1, ltg);
                ++pending;
                if (pending > depth) {
                    cp_async_wait_dyn(depth);     // the oldest slot in flight has landed
                    fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core's async proxy
                    mbar_arrive(&fullA[(int)((ga + 1u - (uint32_t)pending) % (uint32_t)p.NSA)]);
                    --pending;
                }
            }
        }
        while (pending > 0) {
            cp_async_wait_dyn(pending - 1);
            fence_proxy_async_smem();
            mbar_arrive(&fullA[(int)((ga - (uint32_t)pending) % (uint32_t)p.NSA)]);
            --pending;
        }
    } else if (warp == 4) {
        // ------------------------------- weight TMA -------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&p.mapB);
            if (p.resident) {
                // all (slot, tap) weight boxes fit: load them once for the whole life of the CTA
                mbar_arrive_expect_tx(&fullB[0], (uint32_t)(p.n_slots * p.n_taps) * p.b_bytes);
                for (int s = 0; s < p.n_slots; ++s)
                    for (int t = 0; t < p.n_taps; ++t)
                        tma_load_2d(base + (size_t)(s * p.n_taps + t) * p.b_bytes, &p.mapB, &fullB[0], t * p.Kper + p.sl_kofs[s], 0);
            } else {
                uint32_t gb = 0;
                for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x) {
                    const int n_off = (work - fdiv(work, p.rcp_ntiles) * p.n_tiles) * p.BN;
                    for (int s = 0; s < p.n_slots; ++s) {
                        for (int t0 = 0; t0 < p.n_taps; t0 += p.TB, ++gb) {
                            const int nt = min(p.TB, p.n_taps - t0);
                            const int st = (int)(gb % (uint32_t)p.NSB);
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
    } else if (warp == 5) {
        // ------------------------------- MMA issuer -------------------------------
        // One thread issues every MMA.  With N <= 128 an MMA lasts 8..64 cycles, so the scalar code between two issues
        // must be a handful of instructions: everything that does not change per MMA is folded into `a_hi` / `b_hi` (the
        // descriptors minus the 14-bit start address) and into 16-byte-unit offsets kept in registers, tap offsets come
        // from a small shared-memory table, the K-step loop is unrolled.  (First version: ~70 instructions per MMA pair,
        // ncu: tensor pipe 15 % active with both the gather and the epilogue waiting on this warp.)
        {
            const uint32_t leader = elect_one() ? 1u : 0u;
            const uint32_t idesc = umma_idesc_bf16(128, p.BN, 0, 0);
            const uint32_t sbo = (uint32_t)p.pitch * 16u;
            const uint64_t a_hi = umma_smem_desc(0u, sbo, p.plane_bytes, 0u);
            const uint64_t b_hi = umma_smem_desc(0u, 1024u, 16u, 2u);
            const uint32_t plane2 = (2u * p.plane_bytes) >> 4, half16 = (16u * sbo) >> 4;
            const int n_slots = p.n_slots, n_taps = p.n_taps, TB = p.TB, NSA = p.NSA, NSB = p.NSB, acc_bufs = p.acc_bufs;
            const uint32_t BN = (uint32_t)p.BN, b_box16 = p.b_bytes >> 4;
            const bool resident = p.resident != 0;
            uint32_t ga = 0, gb = 0;
            bool b_ready = false;
            int lt = 0;
            for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x, ++lt) {
                const int acc = acc_bufs == 2 ? (lt & 1) : 0;
                if (lt >= acc_bufs) mbar_wait(&tempty[acc], ((lt / acc_bufs) - 1) & 1);
                tc_fence_after();
                if (lane == 0) KP_HTRACE(2, lt);
                const uint32_t d0 = tmem + (uint32_t)acc * (uint32_t)HALVES * BN;
                uint32_t accum = 0;
                for (int s = 0; s < n_slots; ++s, ++ga) {
                    const int stA = (int)(ga % (uint32_t)NSA);
                    mbar_wait(&fullA[stA], (ga / (uint32_t)NSA) & 1);
                    tc_fence_after();
                    if (lane == 0 && s == n_slots - 1) KP_HTRACE(3, lt);
                    const uint32_t a16 = (a_base + (uint32_t)stA * p.a_slot_bytes) >> 4;
                    const int ksteps = (p.sl_nch[s] + 15) >> 4;
                    for (int t0 = 0; t0 < n_taps; t0 += TB) {
                        int stB = 0;
                        uint32_t b16;
                        if (resident) {
                            if (!b_ready) {
                                mbar_wait(&fullB[0], 0);
                                tc_fence_after();
                                b_ready = true;
                            }
                            b16 = (b_base >> 4) + (uint32_t)(s * n_taps + t0) * b_box16;
                        } else {
                            stB = (int)(gb % (uint32_t)NSB);
                            mbar_wait(&fullB[stB], (gb / (uint32_t)NSB) & 1);
                            tc_fence_after();
                            b16 = (b_base + (uint32_t)stB * p.b_stage_bytes) >> 4;
                        }
                        const int t1 = min(t0 + TB, n_taps);
                        for (int t = t0; t < t1; ++t, b16 += b_box16) {
                            const uint32_t at16 = a16 + (uint32_t)p.tapoff[t];
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                if (kk < ksteps) {
                                    const uint64_t db = b_hi | (uint64_t)(b16 + 2u * kk);
                                    const uint64_t da = a_hi | (uint64_t)(at16 + (uint32_t)kk * plane2);
                                    umma_bf16_if(leader, d0, da, db, idesc, accum);
                                    if (HALVES == 2) umma_bf16_if(leader, d0 + BN, da + half16, db, idesc, accum);
                                    accum = 1u;
                                }
                            }
                        }
                        if (!resident) {
                            umma_commit_if(leader, &emptyB[stB]);
                            ++gb;
                        }
                    }
                    umma_commit_if(leader, &emptyA[stA]);
                }
                if (lane == 0) KP_HTRACE(4, lt);
                umma_commit_if(leader, &tfull[acc]);
            }
        }
    } else {
        // ------------------------------- epilogue -------------------------------
        const int q = warp & 3;                   // TMEM lane quarter (warps 6..9 -> 2,3,0,1)
        const int row = q * 32 + lane;
        const int th = row >> 3, tw = row & 7;
        const int et = threadIdx.x - 192;
        if (p.bias != nullptr)
            for (int i = et; i < p.cout_pad; i += 128) s_bias[i] = __ldg(p.bias + i);
        if (p.ssum != nullptr)
            for (int i = et; i < 2 * p.cout_pad; i += 128) s_stat[i] = 0.f;
        named_bar_sync(1, 128);
        int lt = 0;
        for (int work = blockIdx.x; work < p.total_tiles; work += gridDim.x, ++lt) {
            const int mt = fdiv(work, p.rcp_ntiles), nt = work - mt * p.n_tiles;
            const int n = fdiv(mt, p.rcp_tpi), r = mt - n * tiles_per_image;
            const int tr = fdiv(r, p.rcp_tw);
            const int h0 = tr * tile_rows, w0 = (r - tr * p.tiles_w) * 8;
            const int n_off = nt * p.BN;
            const int acc = p.acc_bufs == 2 ? (lt & 1) : 0;
            mbar_wait(&tfull[acc], (lt / p.acc_bufs) & 1);
            tc_fence_after();
            if (et == 0) KP_HTRACE(5, lt);
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
                    epi_chunk(p, v, n_off + c0, valid, pix, lane, 0, s_bias, s_stat);
                }
            }
            if (et == 0) KP_HTRACE(6, lt);
        }
        if (p.ssum != nullptr) {
            named_bar_sync(1, 128);
            for (int i = et; i < p.cout_pad; i += 128) {
                atomicAdd(p.ssum + i, s_stat[i]);
                atomicAdd(p.ssq + i, s_stat[p.cout_pad + i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int device_sm_count();

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Stride-1 layers whose taps all read the same (un-strided) views; at least two taps (a 1x1 convolution re-reads
// nothing) and at least 16 output rows (the tile is 16 or 32 rows tall).
bool haloconv_eligible(const kp_tapconv_desc* d, const float* ssum) {
    (void)ssum;
    if (const char* e = getenv("KP_TAPCONV_HALO"))
        if (atoi(e) == 0) return false;
    if (d->TW > 0 || d->TH > 0 || d->TN > 0) return false;          // explicit tile request: TMA-tap kernel
    if (d->n_maps != d->n_src || d->n_taps < 2 || d->Ho < 16 || d->Wo < 8) return false;
    if (d->Cout_pad > 256 && d->Cout_pad % 128 != 0) return false;
    int dhmin = 127, dhmax = -128, dwmin = 127, dwmax = -128;
    for (int t = 0; t < d->n_taps; ++t) {
        if (d->map_first[t] != 0) return false;
        dhmin = d->dh[t] < dhmin ? d->dh[t] : dhmin; dhmax = d->dh[t] > dhmax ? d->dh[t] : dhmax;
        dwmin = d->dw[t] < dwmin ? d->dw[t] : dwmin; dwmax = d->dw[t] > dwmax ? d->dw[t] : dwmax;
    }
    if (dhmax - dhmin > 8 || dwmax - dwmin > 8) return false;
    int slots = 0, cin = 0;
    for (int m = 0; m < d->n_maps; ++m) {
        const kp_tap_view& v = d->map[m];
        if (v.C <= 0 || v.C % 8 != 0 || v.sw % 8 != 0 || v.sh % 8 != 0 || v.sn % 8 != 0 || v.off % 8 != 0) return false;
        slots += v.C / 64;
        cin += v.C;
        for (int rem = v.C % 64, b = 32; b >= 8; b >>= 1)
            if (rem & b) ++slots;
    }
    // Measured (B200, scripts/grad_probe.py perf, after both kernels got the warp-uniform MMA issue): the halo kernel wins
    // where the output tile is narrow (Cout_pad <= 16: 16->16 27 vs 45 us, 48->16 42 vs 71 us, 64->8 41 vs 57 us at
    // 32 x 128 x 128) - there the TMA-tap kernel moves 9 activation boxes for a few tiny MMAs; from 32 output channels
    // on the TMA-tap kernel (2-3 CTAs per SM) is as fast or faster, so it keeps those layers.
    int max_cout = 16, max_cin = 1 << 20;
    if (const char* e = getenv("KP_HALO_MAX_COUT")) max_cout = atoi(e);
    if (const char* e = getenv("KP_HALO_MAX_CIN")) max_cin = atoi(e);
    if (d->Cout_pad > max_cout || cin > max_cin) return false;
    return slots <= HALO_MAX_SLOTS;
}

int haloconv_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                    float* ssum, float* ssq, cudaStream_t st) {
    HaloKParams p;
    memset(&p, 0, sizeof(p));
    KP_REQUIRE(d->Ktot % d->n_taps == 0, "kp_tapconv(halo): Ktot %d is not a multiple of the tap count %d", d->Ktot, d->n_taps);
    p.Kper = d->Ktot / d->n_taps;
    p.halves = d->Ho >= 32 ? 2 : 1;
    int BN = d->Cout_pad;
    const int bn_cap = p.halves == 2 ? 128 : 256;
    if (BN > bn_cap) {
        BN = bn_cap;
        while (BN > 16 && d->Cout_pad % BN != 0) BN >>= 1;
    }
    KP_REQUIRE(BN % 16 == 0 && d->Cout_pad % BN == 0, "kp_tapconv(halo): Cout_pad=%d does not tile by %d", d->Cout_pad, BN);
    p.BN = BN;
    p.n_tiles = d->Cout_pad / BN;
    int tm = 32;
    p.acc_bufs = 2 * p.halves * BN <= 512 ? 2 : 1;
    while (tm < p.acc_bufs * p.halves * BN) tm <<= 1;
    p.tmem_cols = tm;

    // sources and channel slots (K offsets follow the packed-weight layout: per tap, per source segment padded to CB)
    int kbase = 0, ns = 0;
    for (int m = 0; m < d->n_maps; ++m) {
        const kp_tap_view& v = d->map[m];
        KP_REQUIRE(v.src >= 0 && v.src < KP_MAX_MAPS && src[v.src] != nullptr, "kp_tapconv(halo): map %d has no source", m);
        p.src[m] = reinterpret_cast<const __nv_bfloat16*>(src[v.src]) + v.off;
        KP_REQUIRE((reinterpret_cast<uintptr_t>(p.src[m]) & 15) == 0, "kp_tapconv(halo): source %d not 16-byte aligned", m);
        p.s_sn[m] = v.sn; p.s_sh[m] = v.sh; p.s_sw[m] = v.sw; p.s_H[m] = v.Hd; p.s_W[m] = v.Wd;
        int c0 = 0;
        while (c0 < v.C) {
            int nch = 64;
            while (nch > v.C - c0) nch >>= 1;
            KP_REQUIRE(ns < HALO_MAX_SLOTS, "kp_tapconv(halo): too many channel slots");
            p.sl_src[ns] = (unsigned char)m; p.sl_c0[ns] = (short)c0; p.sl_nch[ns] = (short)nch; p.sl_kofs[ns] = kbase + c0;
            ++ns;
            c0 += nch;
        }
        kbase += (v.C + d->CB - 1) / d->CB * d->CB;
    }
    KP_REQUIRE(kbase == p.Kper, "kp_tapconv(halo): channel segments (%d) do not add up to Kper=%d", kbase, p.Kper);
    p.n_slots = ns;

    int dhmin = 127, dhmax = -128, dwmin = 127, dwmax = -128;
    for (int t = 0; t < d->n_taps; ++t) {
        dhmin = d->dh[t] < dhmin ? d->dh[t] : dhmin; dhmax = d->dh[t] > dhmax ? d->dh[t] : dhmax;
        dwmin = d->dw[t] < dwmin ? d->dw[t] : dwmin; dwmax = d->dw[t] > dwmax ? d->dw[t] : dwmax;
    }
    p.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps; ++t) {
        p.dh[t] = (unsigned char)(d->dh[t] - dhmin);
        p.dw[t] = (unsigned char)(d->dw[t] - dwmin);
    }
    p.dh_min = dhmin; p.dw_min = dwmin;
    p.R = 16 * p.halves + (dhmax - dhmin);
    p.pitch = 8 + (dwmax - dwmin);
    p.plane_bytes = (uint32_t)((p.R * p.pitch) | 1) * 16u;       // odd number of 16-byte units: conflict-free gather stores
    int max_planes = 2;
    for (int s = 0; s < ns; ++s) max_planes = p.sl_nch[s] / 8 > max_planes ? p.sl_nch[s] / 8 : max_planes;
    p.a_slot_bytes = ((uint32_t)max_planes * p.plane_bytes + 127u) & ~127u;
    p.b_bytes = (uint32_t)BN * 128u;
    p.gather_depth = 1;
    if (const char* e = getenv("KP_HALO_GATHER_DEPTH")) { const int c = atoi(e); if (c >= 1 && c <= 3) p.gather_depth = c; }
    auto rcp32 = [](int d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + (unsigned long long)d - 1ull) / (unsigned long long)d); };
    p.pitch_rcp = (65536u + (uint32_t)p.pitch - 1u) / (uint32_t)p.pitch;
    for (int m = 0; m < d->n_maps; ++m)
        KP_REQUIRE((long long)p.s_H[m] * p.s_sh[m] + (long long)p.s_W[m] * p.s_sw[m] < (1ll << 31),
                   "kp_tapconv(halo): image of source %d too large for 32-bit offsets", m);
    const uint32_t epi_bytes = 3u * (uint32_t)d->Cout_pad * sizeof(float);
    const uint32_t budget = 222u * 1024u - epi_bytes - 1024u - 1024u;
    const uint32_t all_b = (uint32_t)(ns * d->n_taps) * p.b_bytes;
    p.resident = (p.n_tiles == 1 && all_b <= 80u * 1024u) ? 1 : 0;
    if (const char* e = getenv("KP_HALO_RESIDENT")) if (atoi(e) == 0) p.resident = 0;
    uint32_t b_region;
    if (p.resident) {
        p.TB = d->n_taps; p.tap_groups = 1; p.b_stage_bytes = (uint32_t)d->n_taps * p.b_bytes; p.NSB = 1;
        b_region = all_b;
        p.NSA = (int)((budget - b_region) / p.a_slot_bytes);
        if (p.NSA > 4) p.NSA = 4;
    } else {
        // weight stages of ~32 KB (TB taps under one barrier), activations double buffered, the rest goes to weight stages
        p.TB = (int)(32u * 1024u / p.b_bytes);
        if (p.TB < 1) p.TB = 1;
        if (p.TB > d->n_taps) p.TB = d->n_taps;
        p.tap_groups = (d->n_taps + p.TB - 1) / p.TB;
        p.TB = (d->n_taps + p.tap_groups - 1) / p.tap_groups;
        p.b_stage_bytes = (uint32_t)p.TB * p.b_bytes;
        p.NSA = 2;
        p.NSB = (int)((budget - 2u * p.a_slot_bytes) / p.b_stage_bytes);
        if (p.NSB > 6) p.NSB = 6;
        b_region = (uint32_t)p.NSB * p.b_stage_bytes;
        if (p.NSB >= 4 && budget - b_region >= 3u * p.a_slot_bytes) p.NSA = 3;
    }
    KP_REQUIRE(p.NSA >= 2 && p.NSB >= 1 && (p.resident || p.NSB >= 2), "kp_tapconv(halo): tile does not fit shared memory");
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        KP_REQUIRE(e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr,
                   "kp_tapconv(halo): cuTensorMapEncodeTiled entry point unavailable");
        KP_REQUIRE((reinterpret_cast<uintptr_t>(wpacked) & 15) == 0, "kp_tapconv(halo): packed weights not 16-byte aligned");
        cuuint64_t gdim[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Cout_pad};
        cuuint64_t gstr[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = reinterpret_cast<EncodeTiledFn2>(fn)(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wpacked),
                                                          gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("kp_tapconv(halo): cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
            return KP_ERR_DRIVER;
        }
    }
    p.tiles_w = (d->Wo + 7) / 8;
    p.tiles_h = (d->Ho + 16 * p.halves - 1) / (16 * p.halves);
    p.total_tiles = d->N * p.tiles_w * p.tiles_h * p.n_tiles;
    p.rcp_ntiles = rcp32(p.n_tiles); p.rcp_tpi = rcp32(p.tiles_w * p.tiles_h); p.rcp_tw = rcp32(p.tiles_w);
    KP_REQUIRE((long long)p.total_tiles * (p.tiles_w * p.tiles_h > p.n_tiles ? p.tiles_w * p.tiles_h : p.n_tiles) < (1ll << 32),
               "kp_tapconv(halo): too many tiles for the 32-bit work decode");
    for (int t = 0; t < d->n_taps; ++t) p.tapoff[t] = (unsigned short)(p.dh[t] * p.pitch + p.dw[t]);
    p.Ho = d->Ho; p.Wo = d->Wo; p.N = d->N;
    p.out = out;
    p.out_off = d->out_off; p.out_sw = d->out_sw; p.out_sh = d->out_sh; p.out_sn = d->out_sn;
    p.Cout = d->Cout; p.cout_pad = d->Cout_pad; p.out_f32 = d->out_f32; p.act = d->act; p.alpha = d->alpha;
    p.accumulate = d->accumulate; p.ksplit = 1;
    p.bias = bias; p.ssum = ssum; p.ssq = ssq;

    const size_t smem = (size_t)b_region + (size_t)p.NSA * p.a_slot_bytes + (size_t)(2 * p.NSA + 2 * p.NSB + 4) * 8 + 16 +
                        epi_bytes + KP_MAX_TAPS * 4 + 1024;
    KP_REQUIRE(smem <= 227u * 1024u, "kp_tapconv(halo): shared memory %zu exceeds the SM", smem);
    int grid = device_sm_count();
    if (grid > p.total_tiles) grid = p.total_tiles;
    unsigned long long* trace = nullptr;
#ifdef KP_TRACE   // debug builds only (nvcc -DKP_TRACE): the shipped library never allocates device memory
    if (getenv("KP_TAPCONV_TRACE")) {
        cudaMalloc(&trace, 24 * 8 * sizeof(unsigned long long));
        cudaMemset(trace, 0, 24 * 8 * sizeof(unsigned long long));
    }
#endif
    p.dbg = trace;
    static bool attr_done = false;
    if (!attr_done) {
        KP_CUDA_CHECK(cudaFuncSetAttribute(haloconv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KP_CUDA_CHECK(cudaFuncSetAttribute(haloconv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    if (p.halves == 2) KP_CUDA_CHECK(launch_pdl(haloconv_kernel<2>, dim3(grid), dim3(HALO_THREADS), smem, st, p));
    else KP_CUDA_CHECK(launch_pdl(haloconv_kernel<1>, dim3(grid), dim3(HALO_THREADS), smem, st, p));
    KP_LAUNCHED();
#ifdef KP_TRACE
    if (trace != nullptr) {
        unsigned long long h[24 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 24 * 8; ++i) if (h[i] != 0 && h[i] < t0) t0 = h[i];
        fprintf(stderr, "haloconv trace grid=%d halves=%d slots=%d taps=%d BN=%d NSA=%d NSB=%d resident=%d items=%d: item: gather_start gather_issued "
                "mma_start mma_data mma_issued epi_go epi_done (ns)\n", grid, p.halves, p.n_slots, p.n_taps, p.BN, p.NSA, p.NSB, p.resident, p.total_tiles);
        for (int t = 0; t < 24; ++t) {
            if (h[t * 8] == 0) break;
            fprintf(stderr, "  %2d:", t);
            for (int k = 0; k < 7; ++k) fprintf(stderr, " %7lld", h[t * 8 + k] ? (long long)(h[t * 8 + k] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
        cudaFree(trace);
    }
#endif
    return KP_OK;
}

}  // namespace kp
