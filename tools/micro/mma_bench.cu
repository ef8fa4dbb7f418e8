// Microbenchmark: cycles per tcgen05.mma.kind::tf32 (M=128, K=8) as a function of N, accumulator
// rotation and operand reuse.  One CTA per SM, one issuing thread; operands are whatever is in smem.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

// mode bit0: rotate accumulators over `nacc`; bit1: vary the A start address per MMA (8 different rows offsets)
__global__ void __launch_bounds__(128, 1) bench(int N, int nmma, int nacc, int vary, int nw, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i & 255);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t dh = ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = smem_u32(smem + 32 * 1024) >> 4;
    long long t0 = 0, t1 = 0;
    const int w = threadIdx.x >> 5;
    if (w < nw) {
        t0 = clock64();
        if (elect_one()) {
            for (int m = 0; m < nmma; m++) {
                const uint32_t d = tbase + (uint32_t)((w * nacc + (m % nacc)) * N);
                const uint32_t ao = vary ? (uint32_t)((m & 7) * 8 + (m & 3) * 2) : 0u;
                mma(d, dh | (uint64_t)(a0 + ao), dh | (uint64_t)(b0 + (m & 3) * 2), idesc, m >= nacc ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[w])), "r"(0u) : "memory");
        t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u));
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int nmma = 4096;
    printf("%6s %5s %5s %10s %12s\n", "N", "nacc", "vary", "cyc/MMA", "MAC/clk/SM");
    for (int grid : {1, 148})
        for (int N : {32, 64, 128})
            for (int nacc : {1})
                for (int vary : {0})
                  for (int nw : {1, 2, 4}) {
                    if (nw * nacc * N > 512) continue;
                    bench<<<grid, 128, 50 * 1024>>>(N, nmma, nacc, vary, nw, d);
                    long long c = 0;
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("g%-4d N=%4d nacc=%d issuers=%d  cyc per MMA (per issuer) %8.1f   aggregate MAC/clk/SM %8.0f\n", grid, N, nacc, nw, (double)c / nmma, nw * 128.0 * N * 8 * nmma / (double)c);
                }
    return 0;
}
