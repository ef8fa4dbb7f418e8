// tcgen05 implicit-GEMM Conv1d with error-compensated 3xTF32 (sm_100a).
//
// Same contract as conv_simt.cu (ConvArgs): out[q][n] = epi(bias[n] + sum_t sum_c f(x[q+off_t][c]) w[t][c][n]).
// GEMM view per CTA: D[128 rows x NT cols] (fp32, in TMEM) += A_t[128 x 32] . W_t[32 x NT] over
// (32-channel K-block, tap).  A plain TF32 MMA misses the 1e-3 waveform tolerance (measured 1.9e-3 /
// 3.6e-3 on the medium / high voice), so every operand is split v = hi + lo with hi = v & 0xffffe000
// and three MMAs accumulate hi*hi + lo*hi + hi*lo  (error ~2e-6 end to end, oracle-level).
//
// Data path (no tensor maps needed):
//   * activations: producer warps read the (128 + span)-row WINDOW of the K-block from HBM with
//     coalesced 128-bit loads, apply the leaky-ReLU prologue, split hi/lo and store both images in the
//     canonical K-major SWIZZLE_128B layout (row r at r*128 B, 16-B chunk c at (c ^ (r & 7))).  A tap is
//     just a descriptor whose start address is shifted by off_t rows, so a k-tap conv stages its input
//     ONCE and issues k x 12 MMAs on it;
//   * weights: pre-split, pre-swizzled tile images written at voice-load time; one cp.async.bulk
//     (UBLKCP) per (K-block, tap) stage into an mbarrier-tracked ring;
//   * MMA: one elected thread issues tcgen05.mma.kind::tf32 (UTCxMMA), tcgen05.commit frees ring slots;
//   * epilogue: tcgen05.ld 32x32b.x32 (LDTM) -> bias / gate / residual / scale / accumulate -> HBM.
// Warp roles: w0 = MMA issuer + TMEM alloc, w1 = weight producer, w2..7 = activation producers,
// all 8 warps run the epilogue.  Several CTAs are resident per SM so one tile's epilogue overlaps
// another's loads and MMAs.  Every mbarrier wait carries a watchdog that traps instead of hanging.
#include "common.cuh"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace sb200 {

namespace {

constexpr int TC_THREADS = 256;
constexpr int TC_PRODUCERS = 192;     // warps 2..7
constexpr int TC_MAX_WSTAGES = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try(bar, parity)) {
        if ((++spins & 1023u) != 0) continue;
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > 2000000000ull) {   // 2 s: a pipeline bug must fail loudly, never hang the GPU
            printf("conv_tc: mbarrier watchdog (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, bar, parity);
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// start>>4 | LBO(1)<<16 | SBO(1024>>4)<<32 | version 1<<46 | base_offset<<49 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int mode = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7) << 49;   // 'matrix base offset' variant (experiment)
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcLaunch {
    int nt;          // columns per CTA (multiple of 32, <= 256)
    int win;         // window rows (multiple of 8)
    int na;          // activation buffers (1 or 2)
    int ws;          // weight ring stages
    int tmem_cols;   // power of two >= 32
    uint32_t idesc;
    int desc_mode;
};

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvArgs a, const TcLaunch L) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A hi/lo x na][W ring x ws][barriers]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_img = (uint32_t)L.win * 128u;              // bytes of one hi (or lo) window image
    const uint32_t a_buf = ((2u * a_img + 1023u) / 1024u) * 1024u;
    const uint32_t w_img = (uint32_t)L.nt * 128u;
    const uint32_t w_stage = 2u * w_img;
    uint8_t* A0 = smem;
    uint8_t* W0 = A0 + (size_t)L.na * a_buf;
    uint64_t* bars = reinterpret_cast<uint64_t*>(W0 + (size_t)L.ws * w_stage);
    // barrier indices
    uint64_t* w_full = bars;                          // [ws]
    uint64_t* w_empty = bars + TC_MAX_WSTAGES;        // [ws]
    uint64_t* a_full = bars + 2 * TC_MAX_WSTAGES;     // [2]
    uint64_t* a_empty = a_full + 2;                   // [2]
    uint64_t* acc_full = a_empty + 2;                 // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 128;
    const int n0 = blockIdx.y * L.nt;
    const int nkb = a.cin / 32;

    if (tid == 0) {
        for (int s = 0; s < L.ws; s++) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(smem_u32(&a_full[s]), TC_PRODUCERS); mbar_init(smem_u32(&a_empty[s]), 1); }
        mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)L.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int ws = 0; uint32_t wpar = 0;
            for (int kb = 0; kb < nkb; kb++) {
                const int buf = kb % L.na;
                mbar_wait(smem_u32(&a_full[buf]), (uint32_t)((kb / L.na) & 1));
                tc_fence_after();
                const uint32_t ahi = smem_u32(A0 + (size_t)buf * a_buf);
                const uint32_t alo = ahi + a_img;
                for (int t = 0; t < a.ntaps; t++) {
                    mbar_wait(smem_u32(&w_full[ws]), wpar);
                    tc_fence_after();
                    const uint32_t whi = smem_u32(W0 + (size_t)ws * w_stage);
                    const uint32_t wlo = whi + w_img;
                    const uint32_t rowoff = (uint32_t)(a.tap_off[t] - a.min_off) * 128u;
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) {
                        const uint64_t dah = make_desc(ahi + rowoff + k4 * 32, L.desc_mode);
                        const uint64_t dal = make_desc(alo + rowoff + k4 * 32, L.desc_mode);
                        const uint64_t dwh = make_desc(whi + k4 * 32);
                        const uint64_t dwl = make_desc(wlo + k4 * 32);
                        const uint32_t first = (kb | t | k4) ? 1u : 0u;
                        tc_mma_tf32(tmem_base, dah, dwh, L.idesc, first);
                        tc_mma_tf32(tmem_base, dal, dwh, L.idesc, 1u);
                        tc_mma_tf32(tmem_base, dah, dwl, L.idesc, 1u);
                    }
                    tc_commit(smem_u32(&w_empty[ws]));     // ring slot reusable once these MMAs retire
                    if (++ws == L.ws) { ws = 0; wpar ^= 1; }
                }
                tc_commit(smem_u32(&a_empty[buf]));
            }
            tc_commit(smem_u32(acc_full));
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== weight producer =====================
        if (lane == 0) {
            int ws = 0; uint32_t wpar = 0;
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wtc) +
                                  (size_t)blockIdx.y * nkb * a.ntaps * w_stage;
            for (int it = 0; it < nkb * a.ntaps; it++) {
                if (it >= L.ws) mbar_wait(smem_u32(&w_empty[ws]), wpar ^ 1);
                mbar_expect_tx(smem_u32(&w_full[ws]), w_stage);
                bulk_g2s(smem_u32(W0 + (size_t)ws * w_stage), wsrc + (size_t)it * w_stage, w_stage, smem_u32(&w_full[ws]));
                if (++ws == L.ws) { ws = 0; wpar ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== activation producers (192 threads) =====================
        const int pt = tid - 64;
        const int nchunk = L.win * 8;
        const int rbase = q0 + a.min_off;
        const float slope = a.in_slope;
        for (int kb = 0; kb < nkb; kb++) {
            const int buf = kb % L.na;
            if (kb >= L.na) mbar_wait(smem_u32(&a_empty[buf]), (uint32_t)(((kb / L.na) - 1) & 1));
            uint8_t* hi = A0 + (size_t)buf * a_buf;
            uint8_t* lo = hi + a_img;
            const float* xk = a.x + kb * 32;
            for (int base = pt; base < nchunk; base += TC_PRODUCERS * 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int idx = base + u * TC_PRODUCERS;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < nchunk) {
                        const int gr = rbase + (idx >> 3);
                        if (gr >= 0 && gr < a.rows_in)
                            v[u] = *reinterpret_cast<const float4*>(xk + (size_t)gr * a.ldx + (idx & 7) * 4);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int idx = base + u * TC_PRODUCERS;
                    if (idx >= nchunk) continue;
                    const int r = idx >> 3, c = idx & 7;
                    float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                    float h[4], l[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float f = e[i] > 0.f ? e[i] : e[i] * slope;
                        h[i] = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
                        l[i] = f - h[i];
                    }
                    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
                    *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
            fence_async_smem();                       // generic-proxy stores -> visible to the tensor core
            mbar_arrive(smem_u32(&a_full[buf]));
        }
    }

    // ===================== epilogue (all 8 warps) =====================
    mbar_wait(smem_u32(acc_full), 0);
    tc_fence_after();
    {
        const int quad = warp & 3, halfsel = warp >> 2;
        const int row = quad * 32 + lane;
        const int q = q0 + row;
        const bool inrange = q < a.rows_q;
        const bool valid = inrange && row_valid(a.map, q);
        const size_t orow = (size_t)q * a.orow_mul + a.orow_add;
        for (int ch = halfsel; ch < L.nt / 32; ch += 2) {
            float o[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ch * 32), o);
            const int n = n0 + ch * 32;
            if (!inrange || n >= a.cout) continue;
            if (a.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(a.bias + n + j);
                    o[j] += b.x; o[j + 1] += b.y; o[j + 2] += b.z; o[j + 3] += b.w;
                }
            }
            if (a.act == ACT_GATE) {
                float* dst = a.y0 + orow * a.ldy0 + (n >> 1);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float g[4];
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        g[e] = valid ? tanhf(o[2 * (j + e)]) * (1.f / (1.f + expf(-o[2 * (j + e) + 1]))) * a.scale : 0.f;
                    *reinterpret_cast<float4*>(dst + j) = make_float4(g[0], g[1], g[2], g[3]);
                }
                continue;
            }
            if (a.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; j++) o[j] = fmaxf(o[j], 0.f);
            }
            if (a.res && valid) {
                const float* rp = a.res + orow * a.ldres + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 r = *reinterpret_cast<const float4*>(rp + j);
                    o[j] += r.x; o[j + 1] += r.y; o[j + 2] += r.z; o[j + 3] += r.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 32; j++) o[j] *= a.scale;
            float* dst; int accum;
            if (n < a.split) { dst = a.y0 + orow * a.ldy0 + n; accum = a.acc0; }
            else { dst = a.y1 + orow * a.ldy1 + (n - a.split); accum = a.acc1; }
            if (accum) {
                if (!valid) continue;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 p = *reinterpret_cast<const float4*>(dst + j);
                    o[j] += p.x; o[j + 1] += p.y; o[j + 2] += p.z; o[j + 3] += p.w;
                }
            } else if (!valid) {
#pragma unroll
                for (int j = 0; j < 32; j++) o[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L.tmem_cols)
                     : "memory");
    }
}

bool plan(const ConvArgs& a, TcLaunch& L, size_t& smem) {
    if (!a.wtc || a.tc_nt <= 0) return false;
    L.nt = a.tc_nt;
    { const char* e = getenv("SB200_TC_DESC_MODE"); L.desc_mode = e ? atoi(e) : 0; }
    L.win = (128 + a.span + 7) & ~7;
    L.tmem_cols = 32;
    while (L.tmem_cols < L.nt) L.tmem_cols <<= 1;
    L.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(L.nt >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const size_t a_buf = ((size_t)2 * L.win * 128 + 1023) / 1024 * 1024;
    const size_t w_stage = (size_t)2 * L.nt * 128;
    const int nkb = a.cin / 32;
    const int iters = nkb * a.ntaps;
    // prefer small footprints (2-3 CTAs per SM) for short K loops, deeper buffering for long ones
    L.na = nkb > 1 ? 2 : 1;
    L.ws = iters < TC_MAX_WSTAGES ? iters : TC_MAX_WSTAGES;
    auto total = [&]() { return (size_t)L.na * a_buf + (size_t)L.ws * w_stage + 256 + 1024; };
    while (total() > 200 * 1024 && L.ws > 2) L.ws--;
    if (total() > 200 * 1024 && L.na > 1) L.na = 1;
    if (total() > 220 * 1024) return false;
    smem = total();
    return true;
}

}  // namespace

bool conv_tc_supported(const ConvArgs& a) {
    TcLaunch L; size_t smem;
    if (a.cin % 32 || a.cout % 32 || a.ntaps > SB_MAX_TAPS) return false;
    return plan(a, L, smem);
}

void launch_conv_tc(const ConvArgs& a, cudaStream_t st) {
    TcLaunch L; size_t smem;
    if (!plan(a, L, smem)) { launch_conv_simt(a, st); return; }
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_done = true;
    }
    dim3 grid((a.rows_q + 127) / 128, (a.cout + L.nt - 1) / L.nt);
    conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a, L);
    g_launch_count++;
}

// Host-side weight image builder: [n-tile][K-block][tap]{hi image, lo image}, each image nt rows x 128 B
// in the K-major SWIZZLE_128B layout the tensor core reads (row n at n*128, chunk c at (c ^ (n & 7))).
size_t conv_tc_weight_floats(int cin, int cout, int ntaps, int nt) {
    const int ntiles = (cout + nt - 1) / nt;
    return (size_t)ntiles * (cin / 32) * ntaps * 2 * nt * 32;
}

void conv_tc_build_weights(const float* wt /*[ntaps][cin][ldw]*/, int ldw, int cin, int cout, int ntaps, int nt,
                           float* out) {
    const int ntiles = (cout + nt - 1) / nt;
    const int nkb = cin / 32;
    size_t o = 0;
    for (int j = 0; j < ntiles; j++)
        for (int kb = 0; kb < nkb; kb++)
            for (int t = 0; t < ntaps; t++) {
                float* hi = out + o;
                float* lo = hi + (size_t)nt * 32;
                for (int n = 0; n < nt; n++)
                    for (int c = 0; c < 32; c++) {
                        const int col = j * nt + n;
                        const float v = col < cout ? wt[((size_t)t * cin + kb * 32 + c) * ldw + col] : 0.f;
                        uint32_t bits; memcpy(&bits, &v, 4);
                        bits &= 0xffffe000u;
                        float h; memcpy(&h, &bits, 4);
                        const size_t idx = (size_t)n * 32 + (size_t)(((c >> 2) ^ (n & 7)) << 2) + (c & 3);
                        hi[idx] = h;
                        lo[idx] = v - h;
                    }
                o += (size_t)2 * nt * 32;
            }
}

}  // namespace sb200
