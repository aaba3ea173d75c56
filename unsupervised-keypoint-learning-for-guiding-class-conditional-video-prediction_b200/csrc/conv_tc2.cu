// Tap-GEMM convolution with CTA PAIRS (tcgen05 cta_group::2), sm_100a - the wide layers (Cout_pad >= 128, 64-channel blocks).
//
// Why: the single-CTA kernel (conv_tc.cu) is bound by the L2 -> shared-memory fill rate (~50 B/clk/SM, profiles/README.md):
// per 64 K-elements it fills 16 KB of activations + BN x 128 B of weights per SM.  Here two CTAs of a cluster (the two SMs
// of a TPC) work on ONE 256-pixel x BN tile: each CTA fills its own 128-pixel activation box and only HALF of the weight
// box (BN/2 rows); one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), the tensor cores of both SMs
// read both halves.  Weight fill per SM halves: 85 -> 128 FLOP per filled byte at BN = 256, 64 -> 85 at BN = 128.
//
// Synchronisation (all mbarriers live at the same shared-memory offset in both CTAs):
//   full[s]    leader's copy only, 1 arrival (the leader's arrive + expect_tx of BOTH CTAs' bytes);
//              both CTAs' TMA loads complete_tx on the LEADER's barrier (address via mapa)
//   empty[s]   own copy, released for both CTAs by tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11)
//   tfull[a]   own copy, same multicast commit;  tempty[a]  leader's copy, 8 arrivals (4 epilogue warps x 2 CTAs)
// Selected by KP_TAPCONV_2CTA=1 (experimental); same C entry point kp_tapconv_bf16, same epilogue (conv_epilogue.cuh).
#include "kp_tc.cuh"
#include "conv_epilogue.cuh"
#include "kp_internal.h"
#include <cudaTypedefs.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace kp {

struct alignas(64) Tap2KParams {
    CUtensorMap mapA[KP_MAX_MAPS];
    CUtensorMap mapB;                 // box {64, BN/2}
    int n_taps, n_src;
    int nblk[KP_MAX_MAPS];
    signed char dh[KP_MAX_TAPS], dw[KP_MAX_TAPS], mf[KP_MAX_TAPS];
    int TW, TH, TN, tiles_w, tiles_h;
    int Ho, Wo, N;
    int BN, n_tiles, m_tiles, total_work, tmem_cols, stages, total_iters, bpt;
    uint32_t a_bytes, b_bytes, stage_bytes;      // per CTA: one activation box, HALF a weight box
    void* out;
    long long out_off, out_sw, out_sh, out_sn;
    int Cout, cout_pad, out_f32, act, accumulate, ksplit;
    float alpha;
    const float* bias;
    float* ssum;
    float* ssq;
    unsigned long long* dbg;   // KP_TAPCONV_TRACE: per-stage timestamps of cluster 0 (debug only)
};

__device__ __forceinline__ unsigned long long t2_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KP_T2TRACE(slot, idx)                                                                              \
    do {                                                                                                   \
        if (p.dbg != nullptr && (blockIdx.x >> 1) == 0 && (idx) < 24) p.dbg[(idx) * 8 + (slot)] = t2_gtime(); \
    } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's shared memory, the bytes are counted on `bar_cluster`
// (a shared::cluster address - the leader's barrier)
__device__ __forceinline__ void tma2_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2,
                                             int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_bf16_if(uint32_t leader, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_if(uint32_t leader, uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "mov.b16 m, 3;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(leader)
        : "memory");
}

__global__ void __launch_bounds__(192, 1) tapconv2_kernel(const __grid_constant__ Tap2KParams p) {
    constexpr int CB = 64;
    constexpr uint32_t SBO = 8 * CB * 2;
    constexpr uint32_t A_BOX_BYTES = 128u * CB * 2u;

    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)p.stages * p.stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>(tslot + 4);
    float* s_stat = s_bias + p.cout_pad;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader = rank == 0;
    const int n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);       // the leader arrives (and expects BOTH CTAs' bytes); the peer only sends bytes
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 8);     // 4 epilogue warps of each CTA
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tslot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                   // barriers of both CTAs initialised, TMEM of both allocated
    tc_fence_after();
    const uint32_t tmem = *tslot;

    if (warp == 0) {
        // ------------------------------- TMA producer (both CTAs) -------------------------------
        if (lane == 0) {
            for (int m = 0; m < KP_MAX_MAPS; ++m)
                if (p.nblk[m] > 0) tma_prefetch_desc(&p.mapA[m]);
            tma_prefetch_desc(&p.mapB);
            uint32_t git = 0;
            for (int work = cid; work < p.total_work; work += n_clusters) {
                const int pair = work / p.n_tiles, nt = work - pair * p.n_tiles;
                const int mt = 2 * pair + (int)rank;          // may be one past the last tile: TMA zero-fills, nothing is stored
                const int w0 = (mt % p.tiles_w) * p.TW, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
                const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
                const int n_off = nt * p.BN + (int)rank * (p.BN >> 1);
                int t = 0, s = 0, cb = 0;
                for (int it = 0; it < p.total_iters; ++it, ++git) {
                    const uint32_t st = git % (uint32_t)S;
                    if (git >= (uint32_t)S) mbar_wait(&empty[st], ((git / (uint32_t)S) - 1) & 1);
                    KP_T2TRACE(is_leader ? 0 : 4, git);
                    const uint32_t a_dst = smem_base + st * p.stage_bytes;
                    const uint32_t full_leader = mapa_rank(smem_u32(&full[st]), 0u);
                    // The peer does not arrive: its bytes may land before the leader has armed the phase (the count goes
                    // negative for a moment), the phase still cannot complete before the leader's arrive + expect_tx.
                    // (A remote mbarrier.arrive.release.cluster per stage cost the peer ~600 ns and set the pace.)
                    if (is_leader) mbar_arrive_expect_tx(&full[st], 2u * (A_BOX_BYTES + p.b_bytes));
                    const int m = p.mf[t] + s;
                    tma2_load_4d(a_dst, &p.mapA[m], full_leader, cb * CB, w0 + p.dw[t], h0 + p.dh[t], n0);
                    tma2_load_2d(a_dst + p.a_bytes, &p.mapB, full_leader, it * CB, n_off);
                    KP_T2TRACE(is_leader ? 1 : 5, git);
                    if (++cb == p.nblk[m]) {
                        cb = 0;
                        if (++s == p.n_src) { s = 0; ++t; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer (leader CTA only) -------------------------------
        if (is_leader) {
            const uint32_t leader = elect_one() ? 1u : 0u;
            const uint32_t idesc = umma_idesc_bf16(256, p.BN, 0, 0);
            const uint64_t a_hi = umma_smem_desc(0u, SBO, 16, 2u);
            const uint64_t b_hi = umma_smem_desc(0u, 1024, 16, 2u);
            uint32_t git = 0;
            int lt = 0;
            for (int work = cid; work < p.total_work; work += n_clusters, ++lt) {
                const int acc = lt & 1;
                if (lt >= 2) mbar_wait(&tempty[acc], ((lt >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t)(acc * p.BN);
                uint32_t accum = 0u;
                for (int it = 0; it < p.total_iters; ++it, ++git) {
                    const uint32_t st = git % (uint32_t)S;
                    const uint32_t a16 = (smem_base + st * p.stage_bytes) >> 4;
                    const uint32_t b16 = a16 + (p.a_bytes >> 4);
                    mbar_wait(&full[st], (git / (uint32_t)S) & 1);
                    tc_fence_after();
                    if (lane == 0) KP_T2TRACE(2, git);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma2_bf16_if(leader, d_tmem, a_hi | (uint64_t)(a16 + 2u * k), b_hi | (uint64_t)(b16 + 2u * k), idesc,
                                      k == 0 ? accum : 1u);
                    }
                    accum = 1u;
                    if (lane == 0) KP_T2TRACE(3, git);
                    umma2_commit_if(leader, &empty[st]);
                }
                umma2_commit_if(leader, &tfull[acc]);
            }
        }
    } else {
        // ------------------------------- epilogue (both CTAs, own 128 pixels) -------------------------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int tw = row % p.TW, th = (row / p.TW) % p.TH, tn = row / (p.TW * p.TH);
        const int et = threadIdx.x - 64;
        if (p.bias != nullptr)
            for (int i = et; i < p.cout_pad; i += 128) s_bias[i] = __ldg(p.bias + i);
        if (p.ssum != nullptr)
            for (int i = et; i < 2 * p.cout_pad; i += 128) s_stat[i] = 0.f;
        named_bar_sync(1, 128);
        int lt = 0;
        for (int work = cid; work < p.total_work; work += n_clusters, ++lt) {
            const int pair = work / p.n_tiles, nt = work - pair * p.n_tiles;
            const int mt = 2 * pair + (int)rank;
            const int w0 = (mt % p.tiles_w) * p.TW, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
            const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
            const int n_off = nt * p.BN;
            const int uw = w0 + tw, uh = h0 + th, n = n0 + tn;
            const bool valid = (mt < p.m_tiles) && (uw < p.Wo) && (uh < p.Ho) && (n < p.N);
            const long long pix = p.out_off + (long long)n * p.out_sn + (long long)uh * p.out_sh + (long long)uw * p.out_sw;
            const int acc = lt & 1;
            mbar_wait(&tfull[acc], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(t_row + (uint32_t)c0, v);
                if (c0 + 16 >= p.BN) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (is_leader) mbar_arrive(&tempty[acc]);
                        else mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0u));
                    }
                }
                if (p.ssum != nullptr && !valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                epi_chunk(p, v, n_off + c0, valid, pix, lane, 0, s_bias, s_stat);
            }
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
    cluster_sync_all();                   // nobody frees TMEM / leaves while the partner may still signal or read
    if (warp == 1) tmem_dealloc2(tmem, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int device_sm_count();

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool tapconv2_eligible(const kp_tapconv_desc* d) {
    const char* e = getenv("KP_TAPCONV_2CTA");
    if (e == nullptr || atoi(e) == 0) return false;
    if (d->CB != 64 || d->TW > 0 || d->TH > 0 || d->TN > 0 || d->BN > 0) return false;
    if (d->Cout_pad < 128 || (d->Cout_pad > 256 && d->Cout_pad % 256 != 0) || d->Cout_pad % 32 != 0) return false;
    if (d->Ktot / 64 < 8) return false;                               // short K loops are epilogue bound anyway
    int TW, TH, TN;
    choose_pixel_tile(128, d->Wo, d->Ho, d->N, &TW, &TH, &TN);
    const long long m_tiles = (long long)((d->Wo + TW - 1) / TW) * ((d->Ho + TH - 1) / TH) * ((d->N + TN - 1) / TN);
    const int BN = d->Cout_pad <= 256 ? d->Cout_pad : 256;
    return ((m_tiles + 1) / 2) * (d->Cout_pad / BN) >= device_sm_count() / 4;   // enough pair tiles to be worth it
}

int tapconv2_launch(const kp_tapconv_desc* d, const void* const* src, const void* wpacked, const float* bias, void* out,
                    float* ssum, float* ssq, cudaStream_t st) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t ee = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    KP_REQUIRE(ee == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr,
               "kp_tapconv(2cta): cuTensorMapEncodeTiled entry point unavailable");
    Tap2KParams p;
    memset(&p, 0, sizeof(p));
    const int CB = 64;
    int TW, TH, TN;
    choose_pixel_tile(128, d->Wo, d->Ho, d->N, &TW, &TH, &TN);
    const int BN = d->Cout_pad <= 256 ? d->Cout_pad : 256;
    int total_blocks_per_tap = -1;
    for (int m = 0; m < d->n_maps; ++m) {
        const kp_tap_view& v = d->map[m];
        KP_REQUIRE(v.src >= 0 && v.src < KP_MAX_MAPS, "kp_tapconv(2cta): map %d has no source", m);
        const int rc = encode_view_map(&p.mapA[m], v, src[v.src], d->N, CB, TW, TH, TN, "kp_tapconv(2cta) A map");
        if (rc != KP_OK) return rc;
        p.nblk[m] = (v.C + CB - 1) / CB;
    }
    int total_iters = 0;
    for (int t = 0; t < d->n_taps; ++t) {
        int blocks = 0;
        for (int s = 0; s < d->n_src; ++s) blocks += p.nblk[d->map_first[t] + s];
        if (total_blocks_per_tap < 0) total_blocks_per_tap = blocks;
        KP_REQUIRE(blocks == total_blocks_per_tap, "kp_tapconv(2cta): taps must read the same number of channel blocks");
        total_iters += blocks;
        p.dh[t] = d->dh[t]; p.dw[t] = d->dw[t]; p.mf[t] = d->map_first[t];
    }
    KP_REQUIRE(d->Ktot == total_iters * CB, "kp_tapconv(2cta): Ktot=%d does not match taps x blocks x 64 = %d", d->Ktot,
               total_iters * CB);
    {
        cuuint64_t gdim[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Cout_pad};
        cuuint64_t gstr[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)(BN / 2)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = reinterpret_cast<EncodeTiledFn3>(fn)(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wpacked), gdim,
                                                          gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("kp_tapconv(2cta): cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
            return KP_ERR_DRIVER;
        }
    }
    p.n_taps = d->n_taps; p.n_src = d->n_src;
    p.TW = TW; p.TH = TH; p.TN = TN;
    p.tiles_w = (d->Wo + TW - 1) / TW;
    p.tiles_h = (d->Ho + TH - 1) / TH;
    p.m_tiles = p.tiles_w * p.tiles_h * ((d->N + TN - 1) / TN);
    p.Ho = d->Ho; p.Wo = d->Wo; p.N = d->N;
    p.BN = BN;
    p.n_tiles = d->Cout_pad / BN;
    p.total_work = ((p.m_tiles + 1) / 2) * p.n_tiles;
    int tm = 32;
    while (tm < 2 * BN) tm <<= 1;
    p.tmem_cols = tm;
    p.total_iters = total_iters;
    p.bpt = total_blocks_per_tap;
    p.a_bytes = 128u * 64u * 2u;
    p.b_bytes = (uint32_t)(BN / 2) * 128u;
    p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
    const uint32_t epi_bytes = 3u * (uint32_t)d->Cout_pad * sizeof(float);
    // two CTA pairs per TPC when the accumulators (2 x BN columns) and a 4-5 stage ring fit twice: hides the per-stage
    // latency chain of the narrower tiles
    int pairs_per_tpc = (tm <= 256) ? 2 : 1;
    if (const char* e = getenv("KP_TAPCONV_2CTA_PAIRS")) pairs_per_tpc = atoi(e) >= 2 && tm <= 256 ? 2 : 1;
    const uint32_t budget = (pairs_per_tpc == 2 ? 106u : 212u) * 1024u - epi_bytes;
    int stages = (int)(budget / p.stage_bytes);
    if (stages > 8) stages = 8;
    KP_REQUIRE(stages >= 2, "kp_tapconv(2cta): stage does not fit");
    p.stages = stages;
    p.out = out;
    p.out_off = d->out_off; p.out_sw = d->out_sw; p.out_sh = d->out_sh; p.out_sn = d->out_sn;
    p.Cout = d->Cout; p.cout_pad = d->Cout_pad; p.out_f32 = d->out_f32; p.act = d->act; p.alpha = d->alpha;
    p.accumulate = d->accumulate; p.ksplit = 1;
    p.bias = bias; p.ssum = ssum; p.ssq = ssq;

    const size_t smem = (size_t)stages * p.stage_bytes + (2 * stages + 4) * 8 + 16 + 1024 + epi_bytes;
    int clusters = device_sm_count() / 2 * pairs_per_tpc;
    if (clusters > p.total_work) clusters = p.total_work;
    static bool attr_done = false;
    if (!attr_done) {
        KP_CUDA_CHECK(cudaFuncSetAttribute(tapconv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    unsigned long long* trace = nullptr;
#ifdef KP_TRACE   // debug builds only (nvcc -DKP_TRACE): the shipped library never allocates device memory
    if (getenv("KP_TAPCONV_TRACE")) {
        cudaMalloc(&trace, 24 * 8 * sizeof(unsigned long long));
        cudaMemset(trace, 0, 24 * 8 * sizeof(unsigned long long));
    }
#endif
    p.dbg = trace;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    KP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tapconv2_kernel, p));
    KP_LAUNCHED();
#ifdef KP_TRACE
    if (trace != nullptr) {
        unsigned long long h[24 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 24 * 8; ++i) if (h[i] != 0 && h[i] < t0) t0 = h[i];
        fprintf(stderr, "tapconv2 trace clusters=%d stages=%d BN=%d work=%d iters=%d: stage: L_prod_go L_prod_issued mma_data mma_issued P_prod_go "
                "P_prod_issued (ns)\n", clusters, stages, BN, p.total_work, total_iters);
        for (int t = 0; t < 24; ++t) {
            if (h[t * 8] == 0) break;
            fprintf(stderr, "  %2d:", t);
            for (int k = 0; k < 6; ++k) fprintf(stderr, " %7lld", h[t * 8 + k] ? (long long)(h[t * 8 + k] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
        cudaFree(trace);
    }
#endif
    return KP_OK;
}

}  // namespace kp
