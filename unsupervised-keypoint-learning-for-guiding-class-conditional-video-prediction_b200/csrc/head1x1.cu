// 1x1 convolution with a narrow bf16 input and an fp32 output: the detector head (reference pose_encoder 'conv_0',
// models/networks/__init__.py:57-59: 16 -> n_pts logits at 128x128, no batch norm / activation).
//
// 7 FLOP per byte: purely HBM-bound (33 MB in, 168 MB out for 64 frames).  On the tensor-core tap kernel its fp32 epilogue
// wrote each pixel's 160-byte row from one thread (16-byte stores at a 160-byte lane stride) and ran at 1.3 TB/s.  Here a warp
// owns 32 consecutive pixels: each lane computes the Cout outputs of its pixel on the CUDA cores (weights broadcast from shared
// memory), the warp transposes through a padded shared-memory tile and writes the 32 x Cout block - contiguous in NHWC - with
// fully coalesced 16-byte stores.
#include "kp_common.cuh"
#include "kp_internal.h"

namespace kp {

template <int CIN, int COUT4>      // COUT4 = Cout / 4
__global__ void __launch_bounds__(256) head1x1_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, long long P, float* __restrict__ out) {
    constexpr int COUT = COUT4 * 4;
    constexpr int PITCH = COUT + 4;                     // floats per pixel row of the staging tile (16-byte aligned, bank-shifted)
    __shared__ float4 s_w[CIN * COUT4];                 // [ci][co/4]
    __shared__ float4 s_b[COUT4];
    __shared__ __align__(16) float s_tile[8][32 * PITCH];
    pdl_launch_dependents();
    pdl_wait();
    // the fp32 master weights are rounded to bf16 here, like the packed copies the tensor-core kernels read
    for (int i = threadIdx.x; i < CIN * COUT4; i += blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(w)[i];
        v.x = __bfloat162float(__float2bfloat16_rn(v.x)); v.y = __bfloat162float(__float2bfloat16_rn(v.y));
        v.z = __bfloat162float(__float2bfloat16_rn(v.z)); v.w = __bfloat162float(__float2bfloat16_rn(v.w));
        s_w[i] = v;
    }
    for (int i = threadIdx.x; i < COUT4; i += blockDim.x)
        s_b[i] = bias != nullptr ? reinterpret_cast<const float4*>(bias)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = s_tile[warp];
    const long long n_groups = (P + 31) / 32;
    for (long long g = (long long)blockIdx.x * 8 + warp; g < n_groups; g += (long long)gridDim.x * 8) {
        const long long p = g * 32 + lane;
        float xv[CIN];
        if (p < P) {
#pragma unroll
            for (int c8 = 0; c8 < CIN / 8; ++c8) {
                const uint4 u = *reinterpret_cast<const uint4*>(x + p * CIN + c8 * 8);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h[j]);
                    xv[c8 * 8 + 2 * j] = f.x;
                    xv[c8 * 8 + 2 * j + 1] = f.y;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CIN; ++c) xv[c] = 0.f;
        }
#pragma unroll
        for (int o = 0; o < COUT4; ++o) {
            float4 acc = s_b[o];
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float4 wv = s_w[c * COUT4 + o];
                acc.x = fmaf(xv[c], wv.x, acc.x); acc.y = fmaf(xv[c], wv.y, acc.y);
                acc.z = fmaf(xv[c], wv.z, acc.z); acc.w = fmaf(xv[c], wv.w, acc.w);
            }
            *reinterpret_cast<float4*>(tile + lane * PITCH + o * 4) = acc;
        }
        __syncwarp();
        // the 32 x COUT block is contiguous in the output: lane-interleaved 16-byte stores
        float* ob = out + g * 32 * COUT;
        const long long valid4 = (min(P - g * 32, 32LL) * COUT) / 4;
#pragma unroll
        for (int i = 0; i < COUT4; ++i) {
            const int q = i * 32 + lane;                 // float4 index inside the block
            const int px = q / COUT4, o = q - px * COUT4;
            if (q < valid4) reinterpret_cast<float4*>(ob)[q] = *reinterpret_cast<const float4*>(tile + px * PITCH + o * 4);
        }
        __syncwarp();
    }
}

int head1x1_launch(const void* x, const float* w, const float* bias, long long P, int Cin, int Cout, float* out, cudaStream_t st) {
    KP_REQUIRE(Cin == 16 && Cout % 4 == 0 && Cout >= 4 && Cout <= 40, "kp_conv1x1_f32: Cin=%d Cout=%d unsupported (Cin 16, Cout 4..40 step 4)",
               Cin, Cout);
    KP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
               "kp_conv1x1_f32: pointers must be 16-byte aligned");
    const long long groups = (P + 31) / 32;
    long long blocks = (groups + 7) / 8;
    const long long cap = 148LL * 8;
    if (blocks > cap) blocks = cap;
    const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
#define KP_HEAD(C4) case C4: KP_CUDA_CHECK(launch_pdl(head1x1_kernel<16, C4>, dim3((unsigned)blocks), dim3(256), 0, st, xi, w, bias, P, out)); break
    switch (Cout / 4) {
        KP_HEAD(1); KP_HEAD(2); KP_HEAD(3); KP_HEAD(4); KP_HEAD(5); KP_HEAD(6); KP_HEAD(7); KP_HEAD(8); KP_HEAD(9); KP_HEAD(10);
        default: KP_REQUIRE(false, "kp_conv1x1_f32: Cout=%d", Cout);
    }
#undef KP_HEAD
    KP_LAUNCHED();
    return KP_OK;
}

}  // namespace kp
