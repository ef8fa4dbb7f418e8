// Micro-experiment for round 2 (NOT part of the library, not yet run on hardware): a 3xTF32 GEMM on the legacy
// tensor path (mma.sync.m16n8k8) whose accumulator is flushed into an fp32 round-to-nearest running sum every 8 K-steps
// (64 channels).  tools/emu_tc_accuracy.py predicts that this chunking brings a K = 2304 contraction to ~2.5e-6 max error
// (below an fp32 FMA chain) where the plain tensor-core accumulation gives ~8e-5 -- the accuracy the text encoder needs
// (duration cliff).  This program measures error against fp64 on the host and the throughput.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma3_gemm tools/micro/mma3_gemm.cu
//   /tmp/mma3_gemm [M=16384] [N=192] [K=2304] [flush=8]      (flush=0: never flush = plain tensor-core accumulation)
//
// C[M][N] = A[M][K] * B[K][N], all row-major fp32.  CTA tile 128 x 64, 8 warps as 4 (M) x 2 (N), warp tile 32 x 32 =
// 2 x 4 m16n8k8 fragments; K in blocks of 32 through a cp.async double buffer.  Fragment layout (g = lane >> 2,
// t = lane & 3):  A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B: b0 (k=t, n=g) b1 (k=t+4, n=g);
// C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int BM = 128, BN = 64, BK = 32, AS = 36, BS = 72;     // smem strides (floats): conflict-free fragment reads

__device__ __forceinline__ unsigned tf32(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float* c, const unsigned* a, const unsigned* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp16(void* s, const void* g) {
    unsigned a = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g));
}

__global__ void __launch_bounds__(256) mma3_gemm(const float* __restrict__ A, const float* __restrict__ B,
                                                 float* __restrict__ C, int M, int N, int K, int flush) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;                        // [2][BM][AS]
    float* Bs = sm + 2 * BM * AS;          // [2][BK][BS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    auto load = [&](int kb, int buf) {
        for (int i = tid; i < BM * (BK / 4); i += 256) {                    // A tile: 128 rows x 8 float4
            const int r = i >> 3, c4 = i & 7;
            cp16(As + (buf * BM + r) * AS + c4 * 4, A + (size_t)(m0 + r) * K + kb * BK + c4 * 4);
        }
        for (int i = tid; i < BK * (BN / 4); i += 256) {                    // B tile: 32 rows x 16 float4
            const int r = i >> 4, c4 = i & 15;
            cp16(Bs + (buf * BK + r) * BS + c4 * 4, B + (size_t)(kb * BK + r) * N + n0 + c4 * 4);
        }
        asm volatile("cp.async.commit_group;");
    };

    float run[2][4][4], acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) { run[i][j][e] = 0.f; acc[i][j][e] = 0.f; }

    const int nkb = K / BK;
    int ksteps = 0;
    load(0, 0);
    for (int kb = 0; kb < nkb; kb++) {
        const int buf = kb & 1;
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (kb + 1 < nkb) load(kb + 1, buf ^ 1);
        const float* Ab = As + (buf * BM + wm) * AS;
        const float* Bb = Bs + (buf * BK) * BS + wn;
#pragma unroll
        for (int k8 = 0; k8 < BK / 8; k8++) {
            unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const float* p = Ab + (i * 16 + g) * AS + k8 * 8 + t;
                const float v[4] = {p[0], p[8 * AS], p[4], p[8 * AS + 4]};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    ah[i][e] = tf32(v[e]);
                    al[i][e] = tf32(v[e] - __uint_as_float(ah[i][e]));
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float* p = Bb + (k8 * 8 + t) * BS + j * 8 + g;
                const float v[2] = {p[0], p[4 * BS]};
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    bh[j][e] = tf32(v[e]);
                    bl[j][e] = tf32(v[e] - __uint_as_float(bh[j][e]));
                }
            }
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    mma_tf32(acc[i][j], ah[i], bh[j]);
                    mma_tf32(acc[i][j], al[i], bh[j]);
                    mma_tf32(acc[i][j], ah[i], bl[j]);
                }
            ksteps++;
            if (flush && ksteps % flush == 0) {                              // fp32 round-to-nearest running sum
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int e = 0; e < 4; e++) { run[i][j][e] += acc[i][j][e]; acc[i][j][e] = 0.f; }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = m0 + wm + i * 16 + g, c = n0 + wn + j * 8 + 2 * t;
            const float o0 = run[i][j][0] + acc[i][j][0], o1 = run[i][j][1] + acc[i][j][1];
            const float o2 = run[i][j][2] + acc[i][j][2], o3 = run[i][j][3] + acc[i][j][3];
            *reinterpret_cast<float2*>(C + (size_t)r * N + c) = make_float2(o0, o1);
            *reinterpret_cast<float2*>(C + (size_t)(r + 8) * N + c) = make_float2(o2, o3);
        }
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 16384, N = argc > 2 ? atoi(argv[2]) : 192, K = argc > 3 ? atoi(argv[3]) : 2304;
    const int flush = argc > 4 ? atoi(argv[4]) : 8;
    if (M % BM || N % BN || K % BK) { fprintf(stderr, "M %% 128, N %% 64, K %% 32 must be 0\n"); return 2; }
    std::vector<float> hA((size_t)M * K), hB((size_t)K * N), hC((size_t)M * N);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };   // U(-1, 1)
    for (auto& v : hA) v = rnd() * 1.7320508f;                                 // unit variance
    const float wsc = 1.7320508f / sqrtf((float)K);
    for (auto& v : hB) v = rnd() * wsc;
    float *dA, *dB, *dC;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dC, hC.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * BM * AS + 2 * BK * BS) * 4;
    cudaFuncSetAttribute(mma3_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(M / BM, N / BN);
    mma3_gemm<<<grid, 256, smem>>>(dA, dB, dC, M, N, K, flush);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; i++) mma3_gemm<<<grid, 256, smem>>>(dA, dB, dC, M, N, K, flush);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double emax = 0, fmax_ = 0;                                                // fp64 reference and an fp32 FMA chain, 256 rows
    for (int r = 0; r < 256 && r < M; r++)
        for (int c = 0; c < N; c++) {
            double ref = 0; float f = 0.f;
            for (int k = 0; k < K; k++) {
                ref += (double)hA[(size_t)r * K + k] * (double)hB[(size_t)k * N + c];
                f = fmaf(hA[(size_t)r * K + k], hB[(size_t)k * N + c], f);
            }
            emax = fmax(emax, fabs((double)hC[(size_t)r * N + c] - ref));
            fmax_ = fmax(fmax_, fabs((double)f - ref));
        }
    printf("M %d N %d K %d flush %d : %.3f ms  %.1f TFLOP/s (useful)  max|err| %.3e  (fp32 FMA chain %.3e)\n", M, N, K, flush, ms,
           2.0 * M * N * K / (ms * 1e-3) / 1e12, emax, fmax_);
    return 0;
}
