// tcgen05 "tap-stacked, transposed" Conv1d for the 32-channel vocoder layers (C_in = C_out = 32), sm_100a.
//
// Why a second formulation.  conv_tc.cu maps one (tap, K-step, split product) to one M=128 x N=32 MMA: a k=7
// layer needs 42 MMAs per 128 output rows, and an M=128 MMA costs the tensor pipe the same ~64-83 cycles
// whether N is 32 or 128 (tools/micro/mma_bench.cu, profiles/notes_r01.md §5) -- the 32-channel layers ran
// at a quarter of the pipe's width and were bound by MMA count, not by HBM.  Here the roles are swapped:
//
//   A (M = 128) = the WEIGHTS of four taps stacked along M: row m = s*32 + co holds w[tap 4g+s][ci = 0..31][co]
//   B (N = 128) = 128 rows of the activation WINDOW starting at window row 4*g*dil, K = ci
//   D[m][n]     = sum_g sum_ci w[4g+s][ci][co] * x[window row n + 4*g*dil][ci]        (TMEM: lane m, column n)
//
// Taps are uniformly spaced (off_t = min_off + t*dil), so the second tap group is the SAME product shifted by
// 4*dil window rows: it is a B descriptor whose start address is 4*dil rows further and it accumulates into the
// same TMEM columns.  What is left of the tap sum -- the four slots s of D -- is a TMEM COLUMN offset when the
// epilogue reads the accumulator back:
//
//   out[q0 + j][co] = bias[co] + sum_{s<4} D[s*32 + co][j + s*dil]
//
// A k-tap layer issues ceil(k/4) * 6 full-width MMAs (N = 128) per tile of TQ = 128 - 3*dil output rows
// instead of k * 6 quarter-width ones per 128 rows: 3.5x fewer for k = 7.  Precision is the same bf16x2 split
// as conv_tc.cu (hi*hi + lo*hi + hi*lo, fp32 accumulate).
//
// Epilogue: warp `quad` owns TMEM lanes quad*32..+31, i.e. slot s = quad, lane = co.  It reads 32 columns at
// the shifted offset (tcgen05.ld 32x32b.x32), the four warps of the group exchange their partials through
// shared memory (128-bit accesses, XOR-swizzled 16-B chunks: conflict-free both ways), then each warp
// finishes 8 of the 32 rows: lane = co, so the read-modify-write loads and the stores are full 128-byte rows.
//
// Warps: w0/w1 MMA issuers (even / odd tiles), w2 weight + raw-window loader (one TMA bulk copy per
// tile, running ahead through a ring of raw fp32 windows), w3-6 converters (raw -> [hi|lo] bf16 image ring),
// w7-10 / w11-14 two epilogue groups (even / odd tiles) over four TMEM accumulator stages.  The raw window
// stays in shared memory until its tile's epilogue is done: in a ResBlock the residual IS the conv input, so
// the epilogue takes it from there instead of reading it from HBM a second time.  Persistent, one CTA per SM.
// With ~227 KB of the SM's 256 KB given to shared memory almost nothing is left for L1: register spills go to
// L2 (measured: 31 % L1 miss on spill loads, the epilogue then ran 5x slower), so the epilogue is written to
// stay inside the register budget (15 warps = 480 threads -> 128 registers per thread, no spills).
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace sb200 {

namespace {

using namespace tcx;

struct TsLaunch {
    int tq;          // output rows per tile (128 - 3*dil, or less for k < 4)
    int ng;          // tap groups (1 or 2)
    int dil;         // tap spacing in rows
    int win;         // window rows (multiple of 8, >= tq + span and >= 128 + 4*dil*(ng-1))
    int nr;          // raw fp32 window ring stages
    int ni;          // converted [hi|lo] image ring stages
    int ntiles;
    uint32_t idesc;
    int res_raw;     // the residual IS the conv input (ResBlock): read it from the raw window in smem
};

constexpr int TS_THREADS = 480;              // 15 warps: up to 128 registers per thread, no spills
constexpr int TS_PROD0 = 3, TS_NPROD = 128, TS_EPI0 = 7;
constexpr int TS_MAXNR = 8, TS_MAXNI = 4, TS_MAXPIECE = 6;    // win <= 192 rows = 768 pieces / 128 threads
constexpr uint32_t TS_WIMG = 128u * 128u;     // one stacked weight image
constexpr int TS_RED_FLOATS = 4 * 32 * 32;    // one exchange buffer: [slot][co][32 rows]

__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void stg32_if(const float* p, float v, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(TS_THREADS, 1) conv_ts_kernel(const ConvArgs a, const TsLaunch L) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // Everything below is addressed through 32-bit shared-window offsets from ONE base register: 64-bit generic
    // pointers kept live across the role branches were what pushed the kernel into (L2-latency) spills.
    const uint32_t sb = smem_u32(smem);
    const uint32_t a_buf = (uint32_t)L.win * 128u;              // one window: raw fp32 or [hi|lo] image
    const uint32_t I0 = sb;                                     // [ni] converted window images (1024-B aligned:
                                                                //  a_buf is a multiple of 1024)
    const uint32_t W0 = I0 + (uint32_t)L.ni * a_buf;            // [ng] stacked weight images
    const uint32_t R0 = W0 + (uint32_t)L.ng * TS_WIMG;          // [nr] raw fp32 windows
    const uint32_t RED = R0 + (uint32_t)L.nr * a_buf;           // [2 groups][2 buffers][4][32][32] floats
    const uint32_t bars = RED + 4u * TS_RED_FLOATS * 4u;
    const uint32_t raw_full = bars;                             // [TS_MAXNR]  loader -> converters
    const uint32_t raw_empty = raw_full + TS_MAXNR * 8;         // [TS_MAXNR]  epilogue group -> loader
    const uint32_t a_full = raw_empty + TS_MAXNR * 8;           // [TS_MAXNI]  converters -> MMA
    const uint32_t a_empty = a_full + TS_MAXNI * 8;             // [TS_MAXNI]  MMA -> converters
    const uint32_t acc_full = a_empty + TS_MAXNI * 8;           // [4]
    const uint32_t acc_empty = acc_full + 4 * 8;                // [4]
    const uint32_t w_full = acc_empty + 4 * 8;                  // [1]
    const uint32_t tmem_slot = w_full + 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (L.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < TS_MAXNR; s++) { mbar_init((raw_full + (uint32_t)(s) * 8u), 1); mbar_init((raw_empty + (uint32_t)(s) * 8u), 128); }
        for (int s = 0; s < TS_MAXNI; s++) { mbar_init((a_full + (uint32_t)(s) * 8u), 1); mbar_init((a_empty + (uint32_t)(s) * 8u), 1); }
        for (int s = 0; s < 4; s++) { mbar_init((acc_full + (uint32_t)(s) * 8u), 1); mbar_init((acc_empty + (uint32_t)(s) * 8u), 128); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < 2) {
        // ===================== MMA issuer of tiles lt = warp, warp + 2, ... =====================
        const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint32_t wimg = W0 >> 4;
        mbar_wait(w_full, 0);
        tc_fence_after();
        for (int lt = warp; lt < my_tiles; lt += 2) {
            const int s = lt & 3;
            mbar_wait((acc_empty + (uint32_t)(s) * 8u), (uint32_t)(((lt >> 2) & 1) ^ 1));
            const int as = lt % L.ni;
            mbar_wait((a_full + (uint32_t)(as) * 8u), (uint32_t)((lt / L.ni) & 1));
            tc_fence_after();
            if (warp == 0 && lane == 0) TC_TRACE(a, lt, 3);
            const uint32_t ximg = (I0 + (uint32_t)as * a_buf) >> 4;
            const uint32_t dcol = tmem_base + (uint32_t)s * 128u;
            if (elect_one()) {
                for (int g = 0; g < L.ng; g++) {
                    const uint32_t wg = wimg + (uint32_t)g * (TS_WIMG >> 4);
                    const uint32_t xg = ximg + (uint32_t)(g * 4 * L.dil) * 8u;      // rows * 128 B >> 4
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) {
                        const uint64_t dwh = desc_hi | (uint64_t)(wg + ks * 2);
                        const uint64_t dwl = desc_hi | (uint64_t)(wg + 4 + ks * 2);
                        const uint64_t dxh = desc_hi | (uint64_t)(xg + ks * 2);
                        const uint64_t dxl = desc_hi | (uint64_t)(xg + 4 + ks * 2);
                        tc_mma_bf16(dcol, dwh, dxh, L.idesc, (g | ks) ? 1u : 0u);
                        tc_mma_bf16(dcol, dwl, dxh, L.idesc, 1u);
                        tc_mma_bf16(dcol, dwh, dxl, L.idesc, 1u);
                    }
                }
                tc_commit((a_empty + (uint32_t)(as) * 8u));
                tc_commit((acc_full + (uint32_t)(s) * 8u));
            }
            __syncwarp();
            if (warp == 0 && lane == 0) TC_TRACE(a, lt, 4);
        }
    } else if (warp == 2) {
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)L.ng * TS_WIMG);
            bulk_g2s(W0, a.wts, (uint32_t)L.ng * TS_WIMG, w_full);
        }
        __syncwarp();
        // ===================== raw-window loader: runs ahead as far as the raw ring allows =====================
        // Rows are contiguous (ldx == 32), so a window is ONE block: one TMA bulk copy per tile.  The first /
        // last window of the array sticks out of it: those are filled by the 32 lanes with zero-filling
        // cp.async and published by hand.
        for (int j = 0; j < my_tiles; j++) {
            const int rs = j % L.nr;
            const int rbase = ((int)blockIdx.x + j * (int)gridDim.x) * L.tq + a.min_off;
            mbar_wait((raw_empty + (uint32_t)(rs) * 8u), (uint32_t)(((j / L.nr) & 1) ^ 1));
            if (lane == 0) TC_TRACE(a, j, 0);
            const uint32_t dst = (R0 + (uint32_t)rs * a_buf);
            if (rbase >= 0 && rbase + L.win <= a.rows_in) {
                if (lane == 0) {
                    mbar_expect_tx((raw_full + (uint32_t)(rs) * 8u), a_buf);
                    bulk_g2s(dst, a.x + (size_t)rbase * 32, a_buf, (raw_full + (uint32_t)(rs) * 8u));
                }
            } else {
                for (int i = lane; i < L.win * 8; i += 32) {     // 16-B pieces
                    const int gr = rbase + (i >> 3);
                    const bool ok = gr >= 0 && gr < a.rows_in;
                    const float* src = ok ? a.x + (size_t)gr * 32 + (i & 7) * 4 : a.x;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)i * 16u), "l"(src),
                                 "r"(ok ? 16u : 0u) : "memory");
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive((raw_full + (uint32_t)(rs) * 8u));
            }
            __syncwarp();
        }
    } else if (warp < TS_EPI0) {
        // ===================== converters (128 threads): raw fp32 -> [hi|lo] bf16 image =====================
        // One warp polls the mbarriers, the others park at a hardware barrier (polling warps cost issue slots and
        // shared-memory transactions the epilogue needs); one thread publishes the image after a second barrier.
        const int gt = tid - TS_PROD0 * 32;
        const float slope = a.in_slope;
        const int npiece = L.win * 4;
        for (int j = 0; j < my_tiles; j++) {
            const int rs = j % L.nr, as = j % L.ni;
            if (warp == TS_PROD0) {
                mbar_wait((raw_full + (uint32_t)(rs) * 8u), (uint32_t)((j / L.nr) & 1));
                if (gt == 0) TC_TRACE(a, j, 1);
                mbar_wait((a_empty + (uint32_t)(as) * 8u), (uint32_t)(((j / L.ni) & 1) ^ 1));
            }
            bar_sync_named(3, TS_NPROD);
            const uint32_t raw = (R0 + (uint32_t)rs * a_buf);
            const uint32_t img = (I0 + (uint32_t)as * a_buf);
#pragma unroll
            for (int u = 0; u < TS_MAXPIECE; u++) {
                const int idx = gt + u * TS_NPROD;
                if (idx < npiece) {
                    const float4 v0 = lds128(raw + (uint32_t)idx * 32u);
                    const float4 v1 = lds128(raw + (uint32_t)idx * 32u + 16u);
                    const int r = idx >> 2, c = idx & 3;
                    float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    if (slope != 1.f) {
#pragma unroll
                        for (int i = 0; i < 8; i++) e[i] = fmaxf(e[i], e[i] * slope);
                    }
                    uint4 hi, lo;
                    hi.x = split2(e[0], e[1], lo.x);
                    hi.y = split2(e[2], e[3], lo.y);
                    hi.z = split2(e[4], e[5], lo.z);
                    hi.w = split2(e[6], e[7], lo.w);
                    const uint32_t rowb = (uint32_t)r * 128u;
                    const uint32_t sw = (uint32_t)(r & 7);
                    sts128u(img + rowb + (((uint32_t)c ^ sw) << 4), hi);
                    sts128u(img + rowb + (((uint32_t)(c + 4) ^ sw) << 4), lo);
                }
            }
            fence_async_smem();                        // generic-proxy stores -> visible to the tensor core
            bar_sync_named(3, TS_NPROD);
            if (gt == 0) { mbar_arrive((a_full + (uint32_t)(as) * 8u)); TC_TRACE(a, j, 2); }
        }
    } else {
        // ===================== epilogue group e: tiles lt = e, e + 2, ... =====================
        const int e = (warp - TS_EPI0) >> 2;
        const int quad = warp & 3;                     // hardware rule: a warp reaches TMEM lanes 32*(warp%4)..+31;
                                                       // each group of four consecutive warps covers all four
        const uint32_t red = RED + (uint32_t)e * (2u * TS_RED_FLOATS * 4u);
        const int nchunks = (L.tq + 31) >> 5;
        const float bias = a.bias ? a.bias[lane] : 0.f;
        const float scale = a.scale;
        // slots beyond the tap count hold zero weights (their D rows are exactly 0): read them unshifted so that
        // the read stays inside the accumulator stage
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(quad < a.ntaps ? quad * L.dil : 0);
        const bool res_ldg = a.res != nullptr && !L.res_raw;
        const bool need_ldg = res_ldg || a.acc0;
        const uint32_t swz = (uint32_t)(lane & 7);
        // exchange-buffer addresses: writer = this warp's slot row [quad][lane]; reader = slot rows [s][lane]
        const uint32_t wr_base = red + (uint32_t)((quad * 32 + lane) * 128);
        const uint32_t rd_base = red + (uint32_t)(lane * 128);
        const uint32_t rd_c0 = ((uint32_t)(quad * 2) ^ swz) << 4, rd_c1 = ((uint32_t)(quad * 2 + 1) ^ swz) << 4;
        // Addends that come from HBM (the read-modify-write operand; a residual that is not the conv input) are
        // fetched one chunk ahead with branch-free, independent loads (a dependent validity-check -> load chain
        // per row costs a full memory latency per row).  Everything in the chunk loop is 32-bit index math on
        // uniform 64-bit bases: a single warp per scheduler runs this, so instruction count IS its latency.
        float pf[8];
        uint32_t nbuf = 0;
        const int last_row = a.rows_q - 1;
        const int ldy = a.ldy0, ldr = a.ldres;
        const float* __restrict__ resp = a.res;
        float* __restrict__ yp = a.y0;
        const int yoff = a.orow_add * ldy + lane, roff = a.orow_add * ldr + lane;
        auto prefetch = [&](int qrow /* first of the 8 rows */) {
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const int q = min(qrow + jj, last_row);
                float v = 0.f;
                if (res_ldg) v = __ldg(resp + (q * ldr + roff)) * scale;
                if (a.acc0) v += yp[q * ldy + yoff];
                pf[jj] = v;
            }
        };
        // Row validity of a whole tile as four 32-bit masks (bit = row), one segment-table load per lane and
        // mask, issued a tile ahead: a per-chunk validity load would stall every chunk for an L2 round trip.
        int se[4];
        uint32_t vm0, vm1, vm2, vm3;
        auto seg_loads = [&](int q0) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int qc = min(q0 + 32 * i + lane, last_row);
                se[i] = a.map.seg_end[qc / a.map.gran] * a.map.seg_mul;
            }
        };
        auto seg_masks = [&](int q0) {
            const int q = q0 + lane;
            vm0 = __ballot_sync(0xffffffffu, q < se[0]);
            vm1 = __ballot_sync(0xffffffffu, q + 32 < se[1]);
            vm2 = __ballot_sync(0xffffffffu, q + 64 < se[2]);
            vm3 = __ballot_sync(0xffffffffu, q + 96 < se[3]);
        };
        auto tile_q0 = [&](int lt) { return ((int)blockIdx.x + lt * (int)gridDim.x) * L.tq; };
        auto chunk_start = [&](int c) { const int c0 = c * 32; return c0 + 32 <= L.tq ? c0 : L.tq - 32; };
        if (e < my_tiles) {
            seg_loads(tile_q0(e)); seg_masks(tile_q0(e));
            if (need_ldg) prefetch(tile_q0(e) + quad * 8);
        }
        for (int lt = e; lt < my_tiles; lt += 2) {
            const int q0 = tile_q0(lt);
            const int rs = lt % L.nr;
            const int s = lt & 3;
            const uint32_t raw = (R0 + (uint32_t)rs * a_buf) + (uint32_t)((-a.min_off) * 128 + lane * 4);
            const uint32_t t0 = tlane + (uint32_t)s * 128u;
            const uint32_t m0 = vm0, m1 = vm1, m2 = vm2, m3 = vm3;
            const int q0n = lt + 2 < my_tiles ? tile_q0(lt + 2) : q0;
            if (lt + 2 < my_tiles) seg_loads(q0n);
            if (quad == 0) mbar_wait((acc_full + (uint32_t)(s) * 8u), (uint32_t)((lt >> 2) & 1));
            bar_sync_named(1 + e, 128);
            tc_fence_after();
            if (quad == 0 && lane == 0) TC_TRACE(a, lt, 5);
            // Software pipeline over the 32-row chunks: the TMEM load of chunk c+1 is in flight while chunk c is
            // exchanged and finished; the exchange buffer is double-buffered, so one barrier per chunk suffices
            // (a buffer is rewritten two chunks later, after every warp has passed the barrier in between).
            float acc[32];
            tmem_ld32_issue(t0, acc);
            for (int c = 0; c < nchunks; c++) {
                const int c0 = c * 32;
                const int cs = chunk_start(c);         // the last chunk is re-based so that every TMEM read stays
                                                       // inside the 128-column stage
                const int r0 = cs + quad * 8;          // first of the 8 rows this warp finishes
                const int w = r0 >> 5;
                const uint32_t wa = w == 0 ? m0 : w == 1 ? m1 : w == 2 ? m2 : m3;
                const uint32_t wb = w == 0 ? m1 : w == 1 ? m2 : w == 2 ? m3 : 0u;
                const uint32_t valid = __funnelshift_r(wa, wb, r0 & 31);
                // rows to store: inside the array, and not below c0 (re-based chunk)
                const int nlive = min(max(a.rows_q - (q0 + r0), 0), 8), nskip = min(max(c0 - r0, 0), 8);
                const uint32_t lmask = ((1u << nlive) - 1u) & ~((1u << nskip) - 1u);
                const uint32_t okmask = valid & lmask;
                const uint32_t stmask = a.acc0 ? okmask : lmask;     // accumulated buffers keep their gap zeros
                float rr[8];                                          // residual rows from the raw window
                if (L.res_raw) {
                    const uint32_t rrow = raw + (uint32_t)r0 * 128u;
#pragma unroll
                    for (int jj = 0; jj < 8; jj++) rr[jj] = lds32(rrow + (uint32_t)jj * 128u);
                }
                const uint32_t xb = (nbuf & 1u) * (uint32_t)(TS_RED_FLOATS * 4);
                nbuf++;
                tmem_ld32_wait(acc);
                if (c == nchunks - 1) {                // accumulator stage fully read: hand it back
                    tc_fence_before();
                    mbar_arrive((acc_empty + (uint32_t)(s) * 8u));
                    if (quad == 0 && lane == 0) TC_TRACE(a, lt, 6);
                }
#pragma unroll
                for (int i = 0; i < 8; i++)
                    sts128f(wr_base + xb + (((uint32_t)i ^ swz) << 4), acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                if (c + 1 < nchunks) tmem_ld32_issue(t0 + (uint32_t)chunk_start(c + 1), acc);
                bar_sync_named(1 + e, 128);
                float o[8];
                {
                    const float4 x0 = lds128(rd_base + xb + rd_c0), x1 = lds128(rd_base + xb + rd_c1);
                    o[0] = x0.x; o[1] = x0.y; o[2] = x0.z; o[3] = x0.w; o[4] = x1.x; o[5] = x1.y; o[6] = x1.z; o[7] = x1.w;
                }
#pragma unroll
                for (int sl = 1; sl < 4; sl++) {
                    const float4 x0 = lds128(rd_base + xb + (uint32_t)(sl * 4096) + rd_c0), x1 = lds128(rd_base + xb + (uint32_t)(sl * 4096) + rd_c1);
                    o[0] += x0.x; o[1] += x0.y; o[2] += x0.z; o[3] += x0.w; o[4] += x1.x; o[5] += x1.y; o[6] += x1.z; o[7] += x1.w;
                }
                const int yrow = (q0 + r0) * ldy + yoff;
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    float v = o[jj] + bias;
                    if (a.act == ACT_RELU) v = fmaxf(v, 0.f);
                    if (L.res_raw) v += rr[jj];
                    v = need_ldg ? fmaf(v, scale, pf[jj]) : v * scale;
                    stg32_if(yp + (yrow + jj * ldy), ((okmask >> jj) & 1u) ? v : 0.f, (stmask >> jj) & 1u);
                }
                // next chunk's addends (possibly of this group's next tile) go in flight now
                if (need_ldg) prefetch(c + 1 < nchunks ? q0 + chunk_start(c + 1) + quad * 8 : q0n + quad * 8);
            }
            mbar_arrive((raw_empty + (uint32_t)(rs) * 8u));     // raw window (residual source) no longer needed
            if (lt + 2 < my_tiles) seg_masks(tile_q0(lt + 2));
            if (quad == 0 && lane == 0) TC_TRACE(a, lt, 7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

bool plan_ts(const ConvArgs& a, TsLaunch& L, size_t& smem) {
    if (!a.wts || a.cin != 32 || a.cout != 32 || a.ldx != 32 || a.ntaps > 8 || a.ntaps < 1) return false;
    if (a.orow_mul != 1 || a.phase_cols || a.act == ACT_GATE || a.split < a.cout) return false;
    if (a.min_off > 0 || a.min_off + a.span < 0) return false;      // output rows must lie inside the window
    if ((long long)(a.rows_q + a.orow_add + 1) * (a.ldy0 > a.ldres ? a.ldy0 : a.ldres) >= (1ll << 31)) return false;   // 32-bit index math
    if (!getenv("SB200_TS")) return false;       // experimental: opt-in (measured slower than conv_tc)
    // uniformly spaced taps: off_t = min_off + t * dil
    L.dil = a.ntaps > 1 ? a.tap_off[1] - a.tap_off[0] : 1;
    if (L.dil < 1) return false;
    for (int t = 0; t < a.ntaps; t++) if (a.tap_off[t] != a.min_off + t * L.dil) return false;
    L.ng = (a.ntaps + 3) / 4;
    const int slots = a.ntaps < 4 ? a.ntaps : 4;
    L.tq = 128 - (slots - 1) * L.dil;
    if (L.tq < 32) return false;
    { const char* e = getenv("SB200_TS_TQ"); if (e && atoi(e) >= 32 && atoi(e) <= L.tq) L.tq = atoi(e); }
    int need = L.tq + a.span;
    if (128 + 4 * L.dil * (L.ng - 1) > need) need = 128 + 4 * L.dil * (L.ng - 1);   // rows the MMAs touch
    L.win = (need + 7) & ~7;
    if (L.win * 4 > TS_MAXPIECE * TS_NPROD) return false;
    L.ntiles = (a.rows_q + L.tq - 1) / L.tq;
    L.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    L.res_raw = (a.res != nullptr && a.res == a.x && a.ldres == a.ldx && a.orow_add == 0 && !getenv("SB200_TS_NORESRAW")) ? 1 : 0;
    L.ni = 2;
    { const char* e = getenv("SB200_TS_NI"); if (e && atoi(e) >= 2 && atoi(e) <= TS_MAXNI) L.ni = atoi(e); }
    const size_t a_buf = (size_t)L.win * 128;
    const size_t fixed = (size_t)L.ni * a_buf + (size_t)L.ng * TS_WIMG + (size_t)4 * TS_RED_FLOATS * 4 +
                         (2 * TS_MAXNR + 2 * TS_MAXNI + 9) * 8 + 16 + 2048;
    L.nr = TS_MAXNR;
    { const char* e = getenv("SB200_TS_NR"); if (e && atoi(e) >= 2 && atoi(e) <= TS_MAXNR) L.nr = atoi(e); }
    while (L.nr > 2 && fixed + (size_t)L.nr * a_buf > 227 * 1024) L.nr--;
    smem = fixed + (size_t)L.nr * a_buf;
    return smem <= 227 * 1024;
}

uint16_t bf16_rn_h(float f) {
    uint32_t b; memcpy(&b, &f, 4);
    b += 0x7fffu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
}
float bf16_to_f(uint16_t h) {
    uint32_t b = (uint32_t)h << 16; float f; memcpy(&f, &b, 4); return f;
}

}  // namespace

bool conv_ts_supported(const ConvArgs& a) {
    TsLaunch L; size_t smem;
    return plan_ts(a, L, smem);
}

void launch_conv_ts(const ConvArgs& a, cudaStream_t st) {
    TsLaunch L; size_t smem;
    if (!plan_ts(a, L, smem)) { launch_conv_simt(a, st); return; }
    static PerDeviceOnce once;
    once.run([] { cudaFuncSetAttribute(conv_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int grid = L.ntiles < sms ? L.ntiles : sms;
    conv_ts_kernel<<<grid, TS_THREADS, smem, st>>>(a, L);
    g_launch_count++;
    check_launch("conv_ts");
}

// Stacked weight images: [group g][128 rows x 128 B], row m = s*32 + co = [hi: 32 ci bf16 | lo: 32 ci bf16] of
// w[tap 4g+s][ci][co] (zero rows for taps >= ntaps), K-major SWIZZLE_128B.  Size in floats.
size_t conv_ts_weight_floats(int ntaps) { return (size_t)((ntaps + 3) / 4) * 128 * 32; }

void conv_ts_build_weights(const float* wt /*[ntaps][32][ldw]*/, int ldw, int ntaps, float* out) {
    const int ng = (ntaps + 3) / 4;
    uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
    memset(o16, 0, (size_t)ng * 128 * 64 * 2);
    for (int g = 0; g < ng; g++)
        for (int s = 0; s < 4; s++) {
            const int t = 4 * g + s;
            if (t >= ntaps) continue;
            for (int co = 0; co < 32; co++) {
                const int m = s * 32 + co;
                uint16_t* row = o16 + ((size_t)g * 128 + m) * 64;
                for (int ci = 0; ci < 32; ci++) {
                    const float v = wt[((size_t)t * 32 + ci) * ldw + co];
                    const uint16_t h = bf16_rn_h(v);
                    const uint16_t l = bf16_rn_h(v - bf16_to_f(h));
                    const int ch = ci >> 3, el = ci & 7;
                    row[(size_t)((ch ^ (m & 7)) << 3) + el] = h;
                    row[(size_t)(((ch + 4) ^ (m & 7)) << 3) + el] = l;
                }
            }
        }
}

}  // namespace sb200
