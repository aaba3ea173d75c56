// Epilogue of the tap-GEMM convolution kernels, shared by the TMA-tap kernel (conv_tc.cu) and the halo-tile kernel
// (conv_halo.cu): one 16-column chunk of one accumulator row (= one output pixel) held in registers.
#pragma once
#include "kp_tc.cuh"

namespace kp {

// 32 lanes x 16 columns -> per-column totals: lane l ends with the total of column
// (bit4*8 + bit3*4 + bit2*2 + bit1) of l; 16 shuffles instead of 80.
__device__ __forceinline__ float warp_colsum16(const float* v, int lane) {
    float a[8];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float keep = up ? v[j + 8] : v[j], send = up ? v[j] : v[j + 8];
            a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float b[4];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float keep = up ? a[j + 4] : a[j], send = up ? a[j] : a[j + 4];
            b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float c[2];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float keep = up ? b[j + 2] : b[j], send = up ? b[j] : b[j + 2];
            c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float d;
    {
        const bool up = lane & 2;
        const float keep = up ? c[1] : c[0], send = up ? c[0] : c[1];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// P: kernel parameter struct with the fields out, out_off (unused here), Cout, cout_pad, out_f32, act, accumulate, alpha,
// bias, ssum, ssq, ksplit.  v: the 16 accumulator columns ch0..ch0+15 of this thread's pixel (all 32 lanes of the warp
// must call: the statistics use warp shuffles); pix: element offset of the pixel in the output view; valid: the pixel
// exists; ks: split-K index; s_bias / s_stat: the per-CTA shared-memory copies (bias [cout_pad], sums [2][cout_pad]).
template <class P>
__device__ __forceinline__ void epi_chunk(const P& p, float (&v)[16], const int ch0, const bool valid, const long long pix,
                                          const int lane, const int ks, const float* s_bias, float* s_stat) {
    if (p.ssum != nullptr) {
        float sq[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
        const float s1 = warp_colsum16(v, lane);
        const float s2 = warp_colsum16(sq, lane);
        if ((lane & 1) == 0) {
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            atomicAdd(s_stat + ch0 + col, s1);
            atomicAdd(s_stat + p.cout_pad + ch0 + col, s2);
        }
    }
    // Everything below is fully unrolled with compile-time indices so v[] stays in registers, and the
    // activation is selected by ONE warp-uniform branch per chunk (the first version indexed v[]
    // dynamically -> local memory, and evaluated the activation switch per element: ~8000 instructions
    // per warp per tile, which made the whole kernel epilogue-bound).
    if (p.ksplit > 1) {
        // split-K partial tile: fp32 atomic accumulation into the zeroed output (host guarantees f32 output,
        // no activation); the bias is contributed once, by split 0
        if (valid && ch0 < p.Cout) {
            float* o = reinterpret_cast<float*>(p.out) + pix + ch0;
            const int nvalid = p.Cout - ch0;
            if (ks == 0 && p.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += s_bias[ch0 + j];
            }
            if (nvalid >= 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                // vector reductions: 4 instead of 16 atomic operations per chunk
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                                 "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < nvalid) atomicAdd(o + j, v[j]);
            }
        }
    } else if (valid && ch0 < p.Cout) {
        if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + ch0);   // 64-byte aligned chunk, broadcast
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 b = b4[j];
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
        }
        if (p.act == KP_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == KP_ACT_LEAKY) {
            const float al = p.alpha;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] >= 0.f ? v[j] : al * v[j];
        } else if (p.act == KP_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        } else if (p.act == KP_ACT_SIGMOID_LAST) {
            const int last = p.Cout - 1 - ch0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j == last) v[j] = 1.f / (1.f + __expf(-v[j]));
        }
        const int nvalid = p.Cout - ch0;   // >= 1; the chunk is complete when >= 16
        // Vector stores per group of 4 floats / 8 bf16 wherever the group is complete (Cout = 40 leaves a
        // half chunk: scalar 2-byte stores at a 80-byte lane stride throttled the LSU), scalars for the rest.
        if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + pix + ch0;
            const bool al = (reinterpret_cast<uintptr_t>(o) & 15) == 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (al && 4 * j + 4 <= nvalid) {
                    float4* o4 = reinterpret_cast<float4*>(o) + j;
                    if (p.accumulate) {
                        const float4 b = *o4;
                        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                    }
                    *o4 = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (4 * j + e < nvalid) o[4 * j + e] = p.accumulate ? o[4 * j + e] + v[4 * j + e] : v[4 * j + e];
                }
            }
        } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix + ch0;
            const bool al = (reinterpret_cast<uintptr_t>(o) & 15) == 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (al && 8 * h + 8 <= nvalid) {
                    uint4* o4 = reinterpret_cast<uint4*>(o) + h;
                    if (p.accumulate) {
                        const uint4 u = *o4;
                        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 a = __bfloat1622float2(hp[j]);
                            v[8 * h + 2 * j] += a.x; v[8 * h + 2 * j + 1] += a.y;
                        }
                    }
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * h + 2 * j], v[8 * h + 2 * j + 1]);
                        w[j] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    *o4 = make_uint4(w[0], w[1], w[2], w[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (8 * h + e < nvalid)
                            o[8 * h + e] = __float2bfloat16_rn(p.accumulate ? __bfloat162float(o[8 * h + e]) + v[8 * h + e]
                                                                            : v[8 * h + e]);
                }
            }
        }
    }

}

// ---------------------------------------------------------------------------------------------------------------
// Fast path: bf16 output, complete 16-channel chunks (Cout % 16 == 0), 16-byte aligned pixels, no accumulate, no
// split-K, activation in {none, relu, leaky}.  The general epi_chunk above decides ~20 run-time flags per chunk; each is
// a constant-bank load + uniform compare + branch on the slow uniform datapath, measured ~800 cycles per chunk on B200
// (ncu source page of halo2_kernel, round 2: the epilogue, not the tensor pipe, set the pace of every layer with K <=
// 576).  Here everything is decided once per launch on the host (epi_fast_ok) and the chunk is branch-free:
//   stats (template) -> + bias (zeros staged when absent) -> max(v, slope*v) (slope 1 / 0 / alpha) -> two 16-byte stores.
// ---------------------------------------------------------------------------------------------------------------

// Per-channel sum / sum of squares of one or two 16-column chunks (the same columns of the two accumulator halves) into the
// warp-PRIVATE shared-memory accumulators s_stat_w[0..] / s_stat_w[cout_pad..]: one transpose-reduce per statistic for both
// halves together, then a plain read-modify-write by the 16 owning lanes (no atomics: nobody else touches this warp's copy;
// the four copies are summed once at the end of the kernel).
template <int NV>
__device__ __forceinline__ void epi_stats16(const float (&v)[NV][16], const int lane, float* __restrict__ s_stat_w, const int cout_pad) {
    float s[16], q[16];
    // packed fp32 pairs (add/mul/fma.f32x2 of sm_100): same rounding as the scalar operations, half the instructions
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 x = make_float2(v[0][2 * j], v[0][2 * j + 1]);
        float2 sj = x, qj = __fmul2_rn(x, x);
#pragma unroll
        for (int h = 1; h < NV; ++h) {
            const float2 y = make_float2(v[h][2 * j], v[h][2 * j + 1]);
            sj = __fadd2_rn(sj, y);
            qj = __ffma2_rn(y, y, qj);
        }
        s[2 * j] = sj.x; s[2 * j + 1] = sj.y;
        q[2 * j] = qj.x; q[2 * j + 1] = qj.y;
    }
    const float s1 = warp_colsum16(s, lane);
    const float s2 = warp_colsum16(q, lane);
    if ((lane & 1) == 0) {
        const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        s_stat_w[col] += s1;
        s_stat_w[cout_pad + col] += s2;
    }
    __syncwarp();
}

struct EpiFast {
    float slope;            // activation as max(v, slope*v): none 1, relu 0, leaky alpha
    int cout_pad;
};

// bias + activation + bf16 store of one 16-column chunk of one pixel (statistics are taken before, see epi_stats16)
__device__ __forceinline__ void epi_chunk_fast(float (&v)[16], const bool valid, __nv_bfloat16* __restrict__ o,
                                               const float* __restrict__ s_bias_c, const EpiFast f) {
    const float4* b4 = reinterpret_cast<const float4*>(s_bias_c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 b = b4[j];
        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], f.slope * v[j]);
    if (valid) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        uint4* o4 = reinterpret_cast<uint4*>(o);
        o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o4[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// the same chunk into a swizzled shared-memory staging row (two 16-byte pieces at a0 / a1) for a TMA store; pixels outside
// the image are clipped by the store itself.  ACT = false: bias only (the batch-norm layers).  Packed fp32 pairs.
template <bool ACT>
__device__ __forceinline__ void epi_chunk_fast_smem(const float (&v)[16], const uint32_t a0, const uint32_t a1,
                                                    const float* __restrict__ s_bias_c, const EpiFast f) {
    const float4* b4 = reinterpret_cast<const float4*>(s_bias_c);
    const float2 sl = make_float2(f.slope, f.slope);
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 b = b4[j];
        float2 x0 = __fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(b.x, b.y));
        float2 x1 = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(b.z, b.w));
        if (ACT) {
            const float2 t0 = __fmul2_rn(x0, sl), t1 = __fmul2_rn(x1, sl);
            x0 = make_float2(fmaxf(x0.x, t0.x), fmaxf(x0.y, t0.y));
            x1 = make_float2(fmaxf(x1.x, t1.x), fmaxf(x1.y, t1.y));
        }
        __nv_bfloat162 h0 = __floats2bfloat162_rn(x0.x, x0.y), h1 = __floats2bfloat162_rn(x1.x, x1.y);
        w[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
        w[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
    }
    st_shared_v4(a0, w[0], w[1], w[2], w[3]);
    st_shared_v4(a1, w[4], w[5], w[6], w[7]);
}

// host: does this launch qualify for the fast epilogue?
inline bool epi_fast_ok(const kp_tapconv_desc* d, const void* out, int ksplit) {
    if (d->out_f32 || d->accumulate || ksplit != 1) return false;
    if (d->Cout != d->Cout_pad || d->Cout_pad % 16 != 0) return false;
    if (d->act != KP_ACT_NONE && d->act != KP_ACT_RELU && d->act != KP_ACT_LEAKY) return false;
    if (d->out_off % 8 || d->out_sw % 8 || d->out_sh % 8 || d->out_sn % 8) return false;
    return (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}
inline float epi_fast_slope(const kp_tapconv_desc* d) {
    return d->act == KP_ACT_RELU ? 0.f : d->act == KP_ACT_LEAKY ? d->alpha : 1.f;
}

}  // namespace kp
