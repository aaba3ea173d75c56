// Micro-benchmark: how many cycles does one tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) cost when a single
// warp issues a long run of them on operands that are already resident in shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<pkg>/csrc -I include -o mma_rate scripts/micro/mma_rate.cu
// Variants: N in {16..256}; one or two accumulators (alternating); a tcgen05.commit every `commit_every` MMAs.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "kp_b200.h"
#include "kp_tc.cuh"

using namespace kp;

__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int N, int n_mma, int n_acc, int commit_every, long long* out_cycles) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (smem_base - smem_u32(smem_dyn));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // operands: zeros are fine (timing only): A 16 KB, B up to 32 KB
    for (int i = threadIdx.x; i < (16 + 32) * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(&tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 1) {
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint64_t a_hi = umma_smem_desc(0u, 1024, 16, 2u), b_hi = umma_smem_desc(0u, 1024, 16, 2u);
        const uint32_t a16 = smem_base >> 4, b16 = (smem_base + 16 * 1024) >> 4;
        __syncwarp();
        const long long t0 = clock64();
        uint32_t phase = 0;
        for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_if(leader, tmem + (uint32_t)(((i / 4) % n_acc) * N), a_hi | (uint64_t)(a16 + 2u * k), b_hi | (uint64_t)(b16 + 2u * k),
                             idesc, 1u);
            if (commit_every > 0 && ((i / 4) % commit_every) == commit_every - 1) {
                umma_commit_if(leader, &bar[0]);
                if (commit_every >= 1000) {}   // (never waits inside the loop)
            }
        }
        umma_commit_if(leader, &bar[1]);
        const long long t_issue = clock64();
        mbar_wait(&bar[1], 0);
        const long long t1 = clock64();
        (void)phase;
        if (lane == 0) {
            out_cycles[0] = t_issue - t0;
            out_cycles[1] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int n_mma = 4096;
    printf("N  n_acc commit_every  issue_cyc/mma  total_cyc/mma  floor(N/2)\n");
    for (int N : {16, 32, 64, 128, 256}) {
        for (int n_acc : {1, 2}) {
            if (n_acc * N > 512) continue;
            for (int ce : {0, 1, 4}) {
                long long h[2];
                for (int rep = 0; rep < 2; ++rep) {
                    mma_rate_kernel<<<1, 64, 50 * 1024>>>(N, n_mma, n_acc, ce, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("%3d  %d  %d   %7.1f  %7.1f   %d\n", N, n_acc, ce, (double)h[0] / n_mma, (double)h[1] / n_mma, N / 2);
            }
        }
    }
    return 0;
}
