// Micro-experiment: can a tap shift of a halo tile be expressed as a START-ADDRESS offset of a 128-byte-swizzled K-major
// shared-memory descriptor?  A [HH][P][64] bf16 tile (one 64-channel block with halo, pixel pitch P) is loaded by ONE TMA
// tile-mode box with SWIZZLE_128B; the MMA then reads M = 128 rows = 16 tile rows x 8 pixels starting at pixel (dh, dw):
//   start = base + (dh*P + dw)*128 B,  SBO = P*128 B (distance between 8-pixel groups),
// with the descriptor's base-offset field either 0 or (start >> 7) & 7.  Compared against a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<pkg>/csrc -Iinclude -o sw128_shift scripts/micro/sw128_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "kp_b200.h"
#include "kp_tc.cuh"

using namespace kp;

namespace kp {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -1; }
void note_launch() {}
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

struct Params {
    CUtensorMap mapA, mapB;
    int P, HH, dh, dw, bo_mode, N, CB;
    uint32_t a_bytes, b_bytes;
    float* out;   // [128][N]
};

__global__ void __launch_bounds__(128, 1) shift_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tslot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const uint32_t a_smem = smem_base, b_smem = smem_base + ((p.a_bytes + 1023u) & ~1023u);
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar[0], p.a_bytes + p.b_bytes);
        tma_load_3d(base, &p.mapA, &bar[0], 0, -1, -1);      // halo box starts at pixel (-1,-1): OOB rows are zero-filled
        tma_load_2d(base + (b_smem - smem_base), &p.mapB, &bar[0], 0, 0);
    }
    if (warp == 1) {
        const uint32_t leader = elect_one() ? 1u : 0u;
        mbar_wait(&bar[0], 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, p.N, 0, 0);
        const uint32_t RB = (uint32_t)p.CB * 2u;
        const uint32_t lay = p.CB == 64 ? 2u : p.CB == 32 ? 4u : 6u;
        const uint32_t start = a_smem + (uint32_t)(p.dh * p.P + p.dw) * RB;
        uint64_t a_hi = umma_smem_desc(0u, (uint32_t)p.P * RB, 16, lay);
        if (p.bo_mode == 1) a_hi |= (uint64_t)((start >> 7) & 7u) << 49;
        const uint64_t b_hi = umma_smem_desc(0u, 8u * RB, 16, lay);
        for (int k = 0; k < p.CB / 16; ++k)
            umma_bf16_if(leader, tmem, a_hi | (uint64_t)(((start + 32u * k) >> 4) & 0x3FFF), b_hi | (uint64_t)(((b_smem + 32u * k) >> 4) & 0x3FFF),
                         idesc, k == 0 ? 0u : 1u);
        umma_commit_if(leader, &bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tc_fence_after();
    {
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int j = 0; j < 16; ++j) p.out[row * p.N + c0 + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || ptr == nullptr) {
        printf("no encode entry point\n");
        return 1;
    }
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(ptr);
  for (int C : {64, 32, 16}) {
    const int IH = 16, IW = 8, N = 64;   // the image IS one 8x16 output tile; halo pixels outside are zero padding
    std::vector<__nv_bfloat16> hx(IH * IW * C), hw(N * C);
    std::vector<float> fx(IH * IW * C), fw(N * C);
    srand(1);
    for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 7 - 3); hx[i] = __float2bfloat16(fx[i]); }
    for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)(rand() % 5 - 2); hw[i] = __float2bfloat16(fw[i]); }
    __nv_bfloat16 *dx, *dwt;
    float* dout;
    cudaMalloc(&dx, hx.size() * 2);
    cudaMalloc(&dwt, hw.size() * 2);
    cudaMalloc(&dout, 128 * N * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dwt, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int bad_total = 0;
    for (int P : {10, 16, 12}) {
        const int HH = IH + 2;
        Params p;
        memset(&p, 0, sizeof(p));
        {
            cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)IW, (cuuint64_t)IH};
            cuuint64_t gstr[2] = {(cuuint64_t)C * 2, (cuuint64_t)IW * C * 2};
            cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)P, (cuuint32_t)HH};
            const CUtensorMapSwizzle swz = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = encode(&p.mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode A failed %d (P=%d)\n", (int)r, P); continue; }
            cuuint64_t gd2[2] = {(cuuint64_t)C, (cuuint64_t)N};
            cuuint64_t gs2[1] = {(cuuint64_t)C * 2};
            cuuint32_t bx2[2] = {(cuuint32_t)C, (cuuint32_t)N};
            cuuint32_t es2[2] = {1, 1};
            r = encode(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dwt, gd2, gs2, bx2, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); continue; }
        }
        p.P = P; p.HH = HH; p.N = N; p.CB = C;
        p.a_bytes = (uint32_t)(P * HH * C * 2);
        p.b_bytes = (uint32_t)(N * C * 2);
        p.out = dout;
        for (int bo = 0; bo < 2; ++bo) {
            for (int dh = 0; dh < 3; ++dh) {
                for (int dw = 0; dw < 3; ++dw) {
                    p.dh = dh; p.dw = dw; p.bo_mode = bo;
                    cudaMemset(dout, 0xff, 128 * N * 4);
                    shift_kernel<<<1, 128, 90 * 1024>>>(p);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("P=%d bo=%d dh=%d dw=%d: CUDA error %s\n", P, bo, dh, dw, cudaGetErrorString(e)); return 1; }
                    std::vector<float> ho(128 * N);
                    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
                    // reference: output pixel (h, w) of the 16x8 tile reads input pixel (h + dh - 1, w + dw - 1), zero outside
                    double maxd = 0;
                    int bad = 0;
                    for (int h = 0; h < 16; ++h)
                        for (int w = 0; w < 8; ++w)
                            for (int n = 0; n < N; ++n) {
                                const int ih = h + dh - 1, iw = w + dw - 1;
                                float ref = 0.f;
                                if (ih >= 0 && ih < IH && iw >= 0 && iw < IW)
                                    for (int c = 0; c < C; ++c) ref += fx[(ih * IW + iw) * C + c] * fw[n * C + c];
                                const double d = fabs((double)ref - (double)ho[(h * 8 + w) * N + n]);
                                if (d > maxd) maxd = d;
                                if (d > 1e-3) ++bad;
                            }
                    printf("CB=%d P=%2d base_offset_mode=%d tap(dh=%d,dw=%d): max|diff| = %g, mismatches %d/%d %s\n", C, P, bo, dh, dw, maxd, bad,
                           128 * N, bad ? "WRONG" : "ok");
                    bad_total += bad ? 1 : 0;
                }
            }
        }
    }
    cudaFree(dx); cudaFree(dwt); cudaFree(dout);
    printf("CB=%d wrong variants: %d\n", C, bad_total);
  }
    return 0;
}
