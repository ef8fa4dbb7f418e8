// Non-GEMM stages of the Piper/VITS path as sm_100a kernels: embedding, channel LayerNorm
// (warp-shuffle reductions), DDSConv depthwise stage, relative-position attention, the
// duration-predictor spline flow, the duration ceil/scan, the monotonic-alignment expansion
// (generate_path restated as a gather), conv_post+tanh and the Philox noise source.
// The reference executes all of these inside onnxruntime (piper/src/lib.rs:362-379); the
// arithmetic restated here follows oracle/vits_oracle.py function by function.
#include "common.cuh"
#include <math.h>

namespace sb200 {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// ------------------------------------------------------------------ embedding
// oracle: text_encoder()  x = emb[ids] * sqrt(H)
__global__ void embed_kernel(const int* __restrict__ ids, const float* __restrict__ emb, float scale,
                             float* __restrict__ x, int rows, int H) {
    pdl_trigger(); pdl_wait();
    const int h4 = H / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * h4) return;
    const int r = (int)(i / h4), c = (int)(i % h4);
    const int id = ids[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (id >= 0) {
        v = reinterpret_cast<const float4*>(emb + (size_t)id * H)[c];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    reinterpret_cast<float4*>(x + (size_t)r * H)[c] = v;
}

// ------------------------------------------------------------------ channel LayerNorm (one warp per row)
// oracle: _layer_norm()  (eps = 1e-5, biased variance, two-pass)
template <int NV>
__global__ void __launch_bounds__(256) ln_kernel(const float* __restrict__ x, const float* __restrict__ res1,
                                                 const float* __restrict__ res2, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float* __restrict__ out, int act,
                                                 RowMap map) {
    pdl_trigger(); pdl_wait();
    constexpr int C = NV * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= map.rows) return;
    float* o = out + (size_t)r * C;
    if (!row_valid(map, r)) {
#pragma unroll
        for (int j = 0; j < NV; j++) o[lane + 32 * j] = 0.f;
        return;
    }
    float v[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        float t = x[(size_t)r * C + lane + 32 * j];
        if (res1) t += res1[(size_t)r * C + lane + 32 * j];
        v[j] = t;
        s += t;
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) { const float d = v[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-5f);
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = lane + 32 * j;
        float y = (v[j] - mean) * rstd * gamma[c] + beta[c];
        if (act == 1) y = gelu_exact(y);
        if (res2) y += res2[(size_t)r * C + c];
        o[c] = y;
    }
}

// ------------------------------------------------------------------ DDSConv: depthwise conv + LN + GELU
// oracle: _dds()  y = gelu(LN(conv_sep(x)))
template <int NV>
__global__ void __launch_bounds__(256) dw_ln_gelu_kernel(const float* __restrict__ x, const float* __restrict__ wdw,
                                                         const float* __restrict__ bdw, int k, int dil,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ out,
                                                         RowMap map) {
    pdl_trigger(); pdl_wait();
    constexpr int C = NV * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= map.rows) return;
    float* o = out + (size_t)r * C;
    if (!row_valid(map, r)) {
#pragma unroll
        for (int j = 0; j < NV; j++) o[lane + 32 * j] = 0.f;
        return;
    }
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) v[j] = bdw[lane + 32 * j];
    const int half = (k - 1) / 2;
    for (int t = 0; t < k; t++) {
        const int rr = r + (t - half) * dil;
        if (rr < 0 || rr >= map.rows) continue;   // gap rows hold zeros, so no per-segment test needed
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = lane + 32 * j;
            v[j] = fmaf(wdw[t * C + c], x[(size_t)rr * C + c], v[j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) s += v[j];
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) { const float d = v[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-5f);
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = lane + 32 * j;
        o[c] = gelu_exact((v[j] - mean) * rstd * gamma[c] + beta[c]);
    }
}

// ------------------------------------------------------------------ relative-position attention
// oracle: _mha()   one CTA = (16 query rows, head, segment); 2*D threads.
//   scores[i][j] = (q_i/sqrt(D)).k_j + [|j-i|<=w] (q_i/sqrt(D)).E_k[j-i+w]
//   out_i = softmax(scores_i) . V + sum_{|d|<=w} p[i][i+d] E_v[d+w]
constexpr int ATT_QT = 16;

template <int D>
__global__ void __launch_bounds__(2 * D) attention_kernel(const float* __restrict__ qkv, int ldq,
                                                          const float* __restrict__ relk,
                                                          const float* __restrict__ relv, int window,
                                                          float* __restrict__ out, int ldo, int H,
                                                          const SegInfo* __restrict__ segs, int tpad) {
    pdl_trigger(); pdl_wait();
    const SegInfo sg = segs[blockIdx.z];
    const int T = sg.len;
    const int i0 = blockIdx.x * ATT_QT;
    if (i0 >= T) return;
    const int head = blockIdx.y;
    const int tid = threadIdx.x;
    constexpr int NT = 2 * D;
    const int nrel = 2 * window + 1;

    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                               // [QT][D]
    float* S = Qs + ATT_QT * D;                   // [QT][tpad]
    float* qE = S + ATT_QT * tpad;                // [QT][nrel]
    float* inv = qE + ATT_QT * nrel;              // [QT]
    float* Osum = inv + ATT_QT;                   // [4][QT][D] partial outputs of the four key-range quarters

    const float* qbase = qkv + (size_t)sg.off * ldq + head * D;
    const float* kbase = qbase + H;
    const float* vbase = qbase + 2 * H;
    const float qscale = rsqrtf((float)D);

    for (int idx = tid; idx < ATT_QT * D; idx += NT) {
        const int i = idx / D, c = idx % D;
        Qs[idx] = (i0 + i < T) ? qbase[(size_t)(i0 + i) * ldq + c] * qscale : 0.f;
    }
    __syncthreads();
    // relative-key logits
    for (int idx = tid; idx < ATT_QT * nrel; idx += NT) {
        const int i = idx / nrel, d = idx % nrel;
        float s = 0.f;
        for (int c = 0; c < D; c++) s = fmaf(Qs[i * D + c], relk[d * D + c], s);
        qE[idx] = s;
    }
    // scores: KPT keys per thread, all ATT_QT queries in registers.  Every Q value is a broadcast LDS that feeds one
    // FMA per key held by the thread: with one key per thread the phase issued one LDS.128 per four FMAs and was
    // bound by the load/store unit, not by the FP32 pipe.
    constexpr int KPT = 3;
    for (int j0 = tid * KPT; j0 < T; j0 += NT * KPT) {
        float acc[KPT][ATT_QT];
#pragma unroll
        for (int k = 0; k < KPT; k++)
#pragma unroll
            for (int i = 0; i < ATT_QT; i++) acc[k][i] = 0.f;
        const float4* kr[KPT];
#pragma unroll
        for (int k = 0; k < KPT; k++)
            kr[k] = reinterpret_cast<const float4*>(kbase + (size_t)min(j0 + k, T - 1) * ldq);
#pragma unroll 2
        for (int c4 = 0; c4 < D / 4; c4++) {
            float4 kv[KPT];
#pragma unroll
            for (int k = 0; k < KPT; k++) kv[k] = kr[k][c4];
#pragma unroll
            for (int i = 0; i < ATT_QT; i++) {
                const float4 qv = *reinterpret_cast<const float4*>(Qs + i * D + c4 * 4);
#pragma unroll
                for (int k = 0; k < KPT; k++) {
                    acc[k][i] = fmaf(qv.x, kv[k].x, acc[k][i]);
                    acc[k][i] = fmaf(qv.y, kv[k].y, acc[k][i]);
                    acc[k][i] = fmaf(qv.z, kv[k].z, acc[k][i]);
                    acc[k][i] = fmaf(qv.w, kv[k].w, acc[k][i]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KPT; k++)
            if (j0 + k < T) {
#pragma unroll
                for (int i = 0; i < ATT_QT; i++) S[i * tpad + j0 + k] = acc[k][i];
            }
    }
    __syncthreads();
    // add relative logits, softmax (one warp per query row, round-robin)
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < ATT_QT; i += NT / 32) {
        const int ia = i0 + i;
        if (ia >= T) { if (lane == 0) inv[i] = 0.f; continue; }
        float* Sr = S + i * tpad;
        if (lane < nrel) {
            const int j = ia + lane - window;
            if (j >= 0 && j < T) Sr[j] += qE[i * nrel + lane];
        }
        __syncwarp();
        float m = -INFINITY;
        for (int j = lane; j < T; j += 32) m = fmaxf(m, Sr[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < T; j += 32) { const float e = expf(Sr[j] - m); Sr[j] = e; s += e; }
        s = warp_sum(s);
        if (lane == 0) inv[i] = 1.f / s;
    }
    // zero the tail so the float4 loop below may over-read up to tpad
    for (int idx = tid; idx < ATT_QT * (tpad - T); idx += NT) {
        const int i = idx / (tpad - T), j = T + idx % (tpad - T);
        S[i * tpad + j] = 0.f;
    }
    __syncthreads();
    // P.V : thread = (quarter of the key range, channel PAIR): two channels per thread halve the broadcast LDS of the
    // probabilities per FMA (same load/store-unit argument as in the score phase)
    {
        const int part = tid / (D / 2), cp = tid % (D / 2);
        float acc[ATT_QT][2];
#pragma unroll
        for (int i = 0; i < ATT_QT; i++) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
        const int nj4 = tpad / 4;
        for (int j4 = part; j4 < nj4; j4 += 4) {
            float2 vv[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int j = j4 * 4 + e;
                vv[e] = j < T ? *reinterpret_cast<const float2*>(vbase + (size_t)j * ldq + cp * 2) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < ATT_QT; i++) {
                const float4 p = *reinterpret_cast<const float4*>(S + i * tpad + j4 * 4);
                acc[i][0] = fmaf(p.x, vv[0].x, acc[i][0]); acc[i][1] = fmaf(p.x, vv[0].y, acc[i][1]);
                acc[i][0] = fmaf(p.y, vv[1].x, acc[i][0]); acc[i][1] = fmaf(p.y, vv[1].y, acc[i][1]);
                acc[i][0] = fmaf(p.z, vv[2].x, acc[i][0]); acc[i][1] = fmaf(p.z, vv[2].y, acc[i][1]);
                acc[i][0] = fmaf(p.w, vv[3].x, acc[i][0]); acc[i][1] = fmaf(p.w, vv[3].y, acc[i][1]);
            }
        }
#pragma unroll
        for (int i = 0; i < ATT_QT; i++)
            *reinterpret_cast<float2*>(Osum + (part * ATT_QT + i) * D + cp * 2) = make_float2(acc[i][0], acc[i][1]);
    }
    __syncthreads();
    for (int idx = tid; idx < ATT_QT * D; idx += NT) {
        const int i = idx / D, c = idx % D;
        const int ia = i0 + i;
        if (ia >= T) continue;
        float o = (Osum[i * D + c] + Osum[(ATT_QT + i) * D + c]) + (Osum[(2 * ATT_QT + i) * D + c] + Osum[(3 * ATT_QT + i) * D + c]);
        for (int d = 0; d < nrel; d++) {
            const int j = ia + d - window;
            if (j >= 0 && j < T) o = fmaf(S[i * tpad + j], relv[d * D + c], o);
        }
        out[(size_t)(sg.off + ia) * ldo + head * D + c] = o * inv[i];
    }
}

// ------------------------------------------------------------------ relative-position softmax (tensor-core attention)
// Sits between the two grouped GEMMs of conv_tf.cu.  S[head][row][key] = (q.k)/sqrt(D) arrives from the first GEMM;
// one warp per (row, head) adds the relative-key logits on the |j - i| <= window band, takes the softmax over the
// utterance's keys IN PLACE, zero-fills the row up to the next multiple of 32 keys (the K extent of the second GEMM) and
// writes the relative-value term  orel[row][head*D + c] = sum_d p[i][i+d] E_v[d+w][c],  which the second GEMM adds as
// its residual.  The whole row lives in registers (NREG x 32 keys).  oracle: _mha()
template <int D, int NREG>
__global__ void __launch_bounds__(256) attn_softmax_kernel(float* __restrict__ S, int Tp, const float* __restrict__ qkv,
                                                           int ldq, const float* __restrict__ relk,
                                                           const float* __restrict__ relv, int window,
                                                           float* __restrict__ orel, int ldo, int RX,
                                                           const SegInfo* __restrict__ segs,
                                                           const int* __restrict__ seg_of_gran, int gran) {
    pdl_trigger(); pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + warp;
    const int head = blockIdx.y;
    if (q >= RX) return;
    const SegInfo sg = segs[seg_of_gran[q / gran]];
    const int i = q - sg.off, T = sg.len;
    if (i < 0 || i >= T) return;                       // gap row
    float* Sr = S + ((size_t)head * RX + q) * Tp;
    const int nrel = 2 * window + 1;
    constexpr int NV = D / 32;
    const float qs = rsqrtf((float)D);
    float qv[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) qv[k] = qkv[(size_t)q * ldq + head * D + lane + 32 * k] * qs;
    float mine = 0.f;                                  // lane d keeps the logit of relative offset d - window
    for (int d = 0; d < nrel; d++) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; k++) s = fmaf(qv[k], relk[d * D + lane + 32 * k], s);
        s = warp_sum(s);
        if (lane == d) mine = s;
    }
    if (lane < nrel) {
        const int j = i + lane - window;
        if (j >= 0 && j < T) Sr[j] += mine;
    }
    __syncwarp();
    float s[NREG];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < NREG; k++) {
        const int j = lane + 32 * k;
        s[k] = j < T ? Sr[j] : -INFINITY;
        m = fmaxf(m, s[k]);
    }
    m = warp_max(m);
    float l = 0.f;
#pragma unroll
    for (int k = 0; k < NREG; k++) {
        const float e = (lane + 32 * k) < T ? expf(s[k] - m) : 0.f;
        s[k] = e;
        l += e;
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
    const int Tz = (T + 31) & ~31;
#pragma unroll
    for (int k = 0; k < NREG; k++) {
        const int j = lane + 32 * k;
        if (j < Tz) Sr[j] = s[k] * inv;
    }
    __syncwarp();
    float pb = 0.f;
    if (lane < nrel) {
        const int j = i + lane - window;
        if (j >= 0 && j < T) pb = Sr[j];
    }
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.f;
    for (int d = 0; d < nrel; d++) {
        const float p = __shfl_sync(0xffffffffu, pb, d);
#pragma unroll
        for (int k = 0; k < NV; k++) acc[k] = fmaf(p, relv[d * D + lane + 32 * k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < NV; k++) orel[(size_t)q * ldo + head * D + lane + 32 * k] = acc[k];
}

// ------------------------------------------------------------------ duration-predictor flow pieces
// oracle: _conv_flow_reverse()  h = pre(z0) + g
__global__ void flow_pre_kernel(const float* __restrict__ z, int zcol, const float* __restrict__ w,
                                const float* __restrict__ b, const float* __restrict__ g, float* __restrict__ h,
                                int C, RowMap map) {
    pdl_trigger(); pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)map.rows * C) return;
    const int r = (int)(i / C), c = (int)(i % C);
    h[i] = row_valid(map, r) ? fmaf(w[c], z[2 * r + zcol], b[c]) + g[i] : 0.f;
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// oracle: _rqs_inverse()  (tails = linear, bound 5, min bin w/h 1e-3, min derivative 1e-3)
template <int NB>
__global__ void spline_kernel(const float* __restrict__ h29, int ldh, float* __restrict__ z, int tcol,
                              float inv_sqrt_filter, RowMap map) {
    pdl_trigger(); pdl_wait();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= map.rows) return;
    if (!row_valid(map, r)) { z[2 * r + tcol] = 0.f; return; }
    const float x = z[2 * r + tcol];
    const float B = 5.0f;
    if (!(x >= -B && x <= B)) return;   // identity outside the tails
    const float* h = h29 + (size_t)r * ldh;
    float cw[NB + 1], ch[NB + 1], dv[NB + 1];
    // widths
    {
        float u[NB], m = -INFINITY, s = 0.f;
#pragma unroll
        for (int k = 0; k < NB; k++) { u[k] = h[k] * inv_sqrt_filter; m = fmaxf(m, u[k]); }
#pragma unroll
        for (int k = 0; k < NB; k++) { u[k] = expf(u[k] - m); s += u[k]; }
        float c = 0.f;
        cw[0] = -B;
#pragma unroll
        for (int k = 0; k < NB; k++) {
            const float wk = 1e-3f + (1.f - 1e-3f * NB) * (u[k] / s);
            c += wk;
            cw[k + 1] = 2.f * B * c + (-B);
        }
        cw[NB] = B;
    }
    {
        float u[NB], m = -INFINITY, s = 0.f;
#pragma unroll
        for (int k = 0; k < NB; k++) { u[k] = h[NB + k] * inv_sqrt_filter; m = fmaxf(m, u[k]); }
#pragma unroll
        for (int k = 0; k < NB; k++) { u[k] = expf(u[k] - m); s += u[k]; }
        float c = 0.f;
        ch[0] = -B;
#pragma unroll
        for (int k = 0; k < NB; k++) {
            const float hk = 1e-3f + (1.f - 1e-3f * NB) * (u[k] / s);
            c += hk;
            ch[k + 1] = 2.f * B * c + (-B);
        }
        ch[NB] = B;
    }
    {
        const float cst = logf(expf(1.f - 1e-3f) - 1.f);
        dv[0] = 1e-3f + softplus_f(cst);
        dv[NB] = dv[0];
#pragma unroll
        for (int k = 1; k < NB; k++) dv[k] = 1e-3f + softplus_f(h[2 * NB + k - 1]);
    }
    int bin = -1;
#pragma unroll
    for (int k = 0; k <= NB; k++) {
        const float loc = (k == NB) ? ch[k] + 1e-6f : ch[k];
        bin += (x >= loc) ? 1 : 0;
    }
    bin = min(max(bin, 0), NB - 1);
    float in_cw = 0, in_w = 0, in_ch = 0, in_h = 0, in_d = 0, in_d1 = 0;
#pragma unroll
    for (int k = 0; k < NB; k++)
        if (k == bin) {
            in_cw = cw[k]; in_w = cw[k + 1] - cw[k];
            in_ch = ch[k]; in_h = ch[k + 1] - ch[k];
            in_d = dv[k]; in_d1 = dv[k + 1];
        }
    const float delta = in_h / in_w;
    const float t = x - in_ch;
    const float e = in_d + in_d1 - 2.f * delta;
    const float a = t * e + in_h * (delta - in_d);
    const float b = in_h * in_d - t * e;
    const float c = -delta * t;
    const float disc = b * b - 4.f * a * c;
    const float root = (2.f * c) / (-b - sqrtf(disc));
    z[2 * r + tcol] = root * in_w + in_cw;
}

// z[r][0..1] = eps[r][0..1] * s   (oracle: sdp_reverse  z = eps_w * noise_w)
__global__ void scale_copy2_kernel(const float* __restrict__ eps, float s, float* __restrict__ z, RowMap map) {
    pdl_trigger(); pdl_wait();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= map.rows) return;
    const bool v = row_valid(map, r) && eps != nullptr;
    z[2 * r] = v ? eps[2 * r] * s : 0.f;
    z[2 * r + 1] = v ? eps[2 * r + 1] * s : 0.f;
}

// ------------------------------------------------------------------ durations: ceil + inclusive scan
// oracle: sdp_reverse() tail (ElementwiseAffine^-1) + durations()
__global__ void __launch_bounds__(256) durations_kernel(const float* __restrict__ z, float m0, float logs0,
                                                        float length_scale, const SegInfo* __restrict__ segs,
                                                        float* __restrict__ logw, int* __restrict__ cum,
                                                        int* __restrict__ y_len) {
    pdl_trigger(); pdl_wait();
    const SegInfo sg = segs[blockIdx.x];
    __shared__ int warp_tot[8];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    const float einv = expf(-logs0);
    for (int base = 0; base < sg.len; base += 256) {
        const int i = base + tid;
        int w = 0;
        if (i < sg.len) {
            const float lw = (z[2 * (sg.off + i)] - m0) * einv;
            logw[sg.off + i] = lw;
            w = (int)ceilf(expf(lw) * length_scale);
        }
        int s = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        int pre = carry_s;
        for (int k = 0; k < warp; k++) pre += warp_tot[k];
        if (i < sg.len) cum[sg.off + i] = pre + s;
        __syncthreads();
        if (tid == 255) carry_s = pre + s;
        __syncthreads();
    }
    if (tid == 0) y_len[blockIdx.x] = max(carry_s, 1);
}

// ------------------------------------------------------------------ alignment expansion (one warp per frame)
// oracle: expand()   frame j takes token i iff cum[i-1] <= j < cum[i]
__global__ void __launch_bounds__(256) expand_kernel(const float* __restrict__ stats, int ldst, int I,
                                                     const int* __restrict__ cum, const float* __restrict__ eps,
                                                     float noise_scale, float* __restrict__ zp,
                                                     const FrameSeg* __restrict__ fsegs,
                                                     const int* __restrict__ ftile_seg, RowMap ymap) {
    pdl_trigger(); pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= ymap.rows) return;
    float* o = zp + (size_t)r * I;
    if (!row_valid(ymap, r)) {
        for (int c = lane; c < I; c += 32) o[c] = 0.f;
        return;
    }
    const FrameSeg fs = fsegs[ftile_seg[r / ymap.gran]];
    const int j = r - fs.off;
    const int* cm = cum + fs.xoff;
    int lo = 0, hi = fs.xlen;          // first i with cm[i] > j
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cm[mid] > j) hi = mid; else lo = mid + 1;
    }
    if (lo >= fs.xlen) {               // only when sum(w_ceil) == 0 and y_len was clamped to 1
        for (int c = lane; c < I; c += 32) o[c] = (eps && noise_scale != 0.f) ? eps[(size_t)r * I + c] * noise_scale : 0.f;
        return;
    }
    const float* st = stats + (size_t)(fs.xoff + lo) * ldst;
    for (int c = lane; c < I; c += 32) {
        float v = st[c];
        if (eps && noise_scale != 0.f) v += eps[(size_t)r * I + c] * expf(st[I + c]) * noise_scale;
        o[c] = v;
    }
}

// ------------------------------------------------------------------ conv_post + tanh (C -> 1, k = 7)
// oracle: decoder() tail.  HBM-bound in principle (read 4*C bytes, write 4 bytes per sample); the first version (one
// sample per thread, 7*C scalar shared-memory loads each) was bound by shared-memory wavefronts at 0.70 ms on C2 against
// a 0.3 ms HBM floor.  Here a CTA stages 512 rows CHANNEL-MAJOR (xs[c][row], leaky-relu applied) and every thread
// produces FOUR consecutive samples: per channel three 128-bit loads bring the 12 staged rows its 4 x 7 taps touch, so a
// staged value is read ~once instead of seven times.  Layout: 16-byte unit u (4 rows) of channel c lives at unit
// (u ^ ((c >> 2) & 7)) -- with that XOR both the transposing stores of the load phase (8 channel groups x 4 rows per
// warp) and the unit-strided 128-bit loads of the compute phase are bank-conflict free.  256 threads: all of them load
// (four independent 128-bit loads in flight each), pairs of lanes split the channels of one 4-sample group.
template <int C>
__global__ void __launch_bounds__(256) conv_post_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ wav, const FrameSeg* __restrict__ fsegs,
                                                        const int* __restrict__ ftile_seg, int U, RowMap map, int vec_ok) {
    pdl_trigger(); pdl_wait();
    constexpr int NT = 256, ROWS = 512, LEAD = 4, SR = ROWS + 8;   // staged rows r0-4 .. r0+515
    constexpr int S = 544;                               // floats per channel row: 136 units, a multiple of 8 units >= SR/4 + 7
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                       // [C][S]
    float* ws = xs + C * S;               // [C][8]: taps 0..6 of channel c, then 0
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * ROWS;
    for (int i = tid; i < C * 8; i += NT) {
        const int c = i >> 3, t = i & 7;
        ws[i] = t < 7 ? w[t * C + c] : 0.f;
    }
    constexpr int NF4 = SR * (C / 4);
    for (int base = tid; base < NF4; base += NT * 4) {    // four independent 128-bit loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = base + k * NT;
            const int rr = i / (C / 4), c4 = i % (C / 4);
            const int gr = r0 - LEAD + rr;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < NF4 && gr >= 0 && gr < map.rows) v[k] = reinterpret_cast<const float4*>(x + (size_t)gr * C)[c4];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = base + k * NT;
            if (i >= NF4) continue;
            const int rr = i / (C / 4), c4 = i % (C / 4);
            const int pos = (((rr >> 2) ^ (c4 & 7)) << 2) | (rr & 3);
            float* d = xs + (size_t)(c4 * 4) * S + pos;
            d[0] = v[k].x > 0.f ? v[k].x : 0.01f * v[k].x;
            d[S] = v[k].y > 0.f ? v[k].y : 0.01f * v[k].y;
            d[2 * S] = v[k].z > 0.f ? v[k].z : 0.01f * v[k].z;
            d[3 * S] = v[k].w > 0.f ? v[k].w : 0.01f * v[k].w;
        }
    }
    __syncthreads();
    // thread pair (2g, 2g+1) owns the four rows r0 + 4g ..; each lane of the pair sums half of the channels
    const int g = tid >> 1, half = tid & 1;
    const int r = r0 + 4 * g;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int u0 = g + 1;                                     // staged unit of the group's own four rows
#pragma unroll 4
    for (int k = 0; k < C / 2; k++) {
        const int c = half * (C / 2) + k;
        const int sw = (c >> 2) & 7;
        const float* xc = xs + (size_t)c * S;
        const float4 a = *reinterpret_cast<const float4*>(xc + (((u0 - 1) ^ sw) << 2));
        const float4 b = *reinterpret_cast<const float4*>(xc + ((u0 ^ sw) << 2));
        const float4 d = *reinterpret_cast<const float4*>(xc + (((u0 + 1) ^ sw) << 2));
        const float4 w0 = *reinterpret_cast<const float4*>(ws + c * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + c * 8 + 4);
        const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
        const float wt[7] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z};
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int t = 0; t < 7; t++) acc[j] = fmaf(v[j + 1 + t], wt[t], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
    if (half != 0 || r >= map.rows || !row_valid(map, r)) return;     // a group of four rows never straddles a segment end
    const FrameSeg fs = fsegs[ftile_seg[r / map.gran]];
    float* dst = wav + fs.out_off + (long long)(r - (long long)fs.off * U);
    const float o0 = tanhf(acc[0]), o1 = tanhf(acc[1]), o2 = tanhf(acc[2]), o3 = tanhf(acc[3]);
    if (vec_ok) *reinterpret_cast<float4*>(dst) = make_float4(o0, o1, o2, o3);
    else { dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3; }
}

// ------------------------------------------------------------------ f32 -> i16 with per-utterance peak normalisation
// crates/audio/ops/src/samples.rs:51-75 (`to_i16_vec`): scale = 32767 / max(|x|_max, f32::EPSILON);
// y = trunc(clamp(x * scale, -32768, 32767)).  Bit-exact with the host version (same fp32 operations in the same
// order); halves the device->host bytes of a synthesis result.  `PcmPost` folds in what the reference does to a
// chunk before that conversion: trimming the overlap frames of a streamed chunk (piper/src/lib.rs:811-826),
// crossfade(42) (samples.rs:144-157; the sine table is computed by the host so both sides use the same floats) and
// the linear volume gain of AudioOutputConfig (synth/src/lib.rs:84-86).
__device__ __forceinline__ float pcm_value(const float* __restrict__ x, long long i, long long m, const PcmPost& p) {
    float v = x[i];
    if (p.fade_n > 0) {
        if (i < p.fade_n) v = __fmul_rn(v, p.tab[i]);
        else if (i >= m - p.fade_n) v = __fmul_rn(v, p.tab[m - 1 - i]);
    }
    return p.gain == 1.f ? v : __fmul_rn(v, p.gain);
}

__global__ void i16_absmax_kernel(const float* __restrict__ wav, const FrameSeg* __restrict__ fsegs, int hop,
                                  unsigned* __restrict__ maxbits, const PcmPost post) {
    pdl_trigger(); pdl_wait();
    const FrameSeg fs = fsegs[blockIdx.y];
    const long long n = (long long)fs.len * hop - post.trim_lo - post.trim_hi;
    const float* x = wav + fs.out_off + post.trim_lo;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(pcm_value(x, i, n, post)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(maxbits + blockIdx.y, __float_as_uint(m));   // non-negative floats
                                                                                                  // order like their bits
}

__global__ void i16_convert_kernel(const float* __restrict__ wav, const FrameSeg* __restrict__ fsegs, int hop,
                                   const unsigned* __restrict__ maxbits, short* __restrict__ out, const PcmPost post) {
    pdl_trigger(); pdl_wait();
    const FrameSeg fs = fsegs[blockIdx.y];
    const long long n = (long long)fs.len * hop - post.trim_lo - post.trim_hi;
    const float* x = wav + fs.out_off + post.trim_lo;
    short* y = out + fs.out_off;
    const float amax = fmaxf(__uint_as_float(maxbits[blockIdx.y]), 1.1920928955078125e-07f);
    const float scale = __fdiv_rn(32767.0f, amax);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = fminf(fmaxf(__fmul_rn(pcm_value(x, i, n, post), scale), -32768.0f), 32767.0f);
        y[i] = (short)(int)v;                      // truncating cast
    }
}

// ------------------------------------------------------------------ Philox4x32-10 -> N(0,1)
__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0,
                                             unsigned k1) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
    const unsigned n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    const unsigned n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__global__ void randn_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                             unsigned long long stream_id) {
    pdl_trigger(); pdl_wait();
    const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 * 4 >= n) return;
    unsigned c0 = (unsigned)i4, c1 = (unsigned)(i4 >> 32), c2 = (unsigned)stream_id, c3 = (unsigned)(stream_id >> 32);
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const float u0 = ((float)c0 + 0.5f) * 2.3283064365386963e-10f;
    const float u1 = ((float)c1 + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)c2 + 0.5f) * 2.3283064365386963e-10f;
    const float u3 = ((float)c3 + 0.5f) * 2.3283064365386963e-10f;
    const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
    float s0, cs0, s1, cs1;
    sincosf(6.283185307179586f * u1, &s0, &cs0);
    sincosf(6.283185307179586f * u3, &s1, &cs1);
    const float v[4] = {r0 * cs0, r0 * s0, r1 * cs1, r1 * s1};
    for (int e = 0; e < 4; e++)
        if (i4 * 4 + e < n) out[i4 * 4 + e] = v[e];
}

// speaker conditioning: one warp per row of the stacked 1x1 conditioning convs
__global__ void __launch_bounds__(256) cond_bias_kernel(const float* __restrict__ w, const float* __restrict__ base,
                                                        const float* __restrict__ g, int rows, int gin,
                                                        float* __restrict__ out) {
    pdl_trigger(); pdl_wait();
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int k = lane; k < gin; k += 32) s = fmaf(w[(size_t)r * gin + k], g[k], s);
    s = warp_sum(s);
    if (lane == 0) out[r] = base[r] + s;
}

__global__ void fill_zero_kernel(float4* p, long long n4) {
    pdl_trigger(); pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <typename K>
void set_smem(K kern, size_t bytes) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

// ====================================================================== launchers
void launch_embed(const int* ids_rows, const float* emb, float scale, float* x, int rows, int H, cudaStream_t st) {
    const long long n = (long long)rows * (H / 4);
    launch_pdl(embed_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, ids_rows, emb, scale, x, rows, H);
    g_launch_count++;
}

void launch_ln(const float* x, const float* res1, const float* res2, const float* gamma, const float* beta,
               float* out, int C, int act, RowMap map, cudaStream_t st) {
    const unsigned grid = (map.rows + 7) / 8;
    switch (C) {
        case 96: launch_pdl(ln_kernel<3>, dim3(grid), dim3(256), 0, st, x, res1, res2, gamma, beta, out, act, map); break;
        case 192: launch_pdl(ln_kernel<6>, dim3(grid), dim3(256), 0, st, x, res1, res2, gamma, beta, out, act, map); break;
        case 256: launch_pdl(ln_kernel<8>, dim3(grid), dim3(256), 0, st, x, res1, res2, gamma, beta, out, act, map); break;
        default: throw_launch_error("LayerNorm: unsupported channel count (96 / 192 / 256)");
    }
    g_launch_count++;
}

void launch_dw_ln_gelu(const float* x, const float* wdw, const float* bdw, int k, int dil, const float* gamma,
                       const float* beta, float* out, int C, RowMap map, cudaStream_t st) {
    const unsigned grid = (map.rows + 7) / 8;
    switch (C) {
        case 96: launch_pdl(dw_ln_gelu_kernel<3>, dim3(grid), dim3(256), 0, st, x, wdw, bdw, k, dil, gamma, beta, out, map); break;
        case 192: launch_pdl(dw_ln_gelu_kernel<6>, dim3(grid), dim3(256), 0, st, x, wdw, bdw, k, dil, gamma, beta, out, map); break;
        case 256: launch_pdl(dw_ln_gelu_kernel<8>, dim3(grid), dim3(256), 0, st, x, wdw, bdw, k, dil, gamma, beta, out, map); break;
        default: throw_launch_error("DDSConv: unsupported channel count (96 / 192 / 256)");
    }
    g_launch_count++;
}

static int att_tpad(int max_len) { return (max_len + 3) & ~3; }

size_t attention_smem_bytes(int max_len, int D) {
    const int tpad = att_tpad(max_len);
    return sizeof(float) * ((size_t)ATT_QT * D + (size_t)ATT_QT * tpad + ATT_QT * 32 + ATT_QT + 4 * ATT_QT * D);
}

void launch_attention(const float* qkv, int ldq, const float* relk, const float* relv, int window, float* out,
                      int ldo, int H, int heads, const SegInfo* segs, int nseg, int max_len, cudaStream_t st) {
    const int D = H / heads;
    const int tpad = att_tpad(max_len);
    const size_t smem = attention_smem_bytes(max_len, D);
    dim3 grid((max_len + ATT_QT - 1) / ATT_QT, heads, nseg);
    if (D == 96) {
        set_smem(attention_kernel<96>, smem);
        launch_pdl(attention_kernel<96>, dim3(grid), dim3(192), smem, st, qkv, ldq, relk, relv, window, out, ldo, H, segs, tpad);
    } else if (D == 48) {
        set_smem(attention_kernel<48>, smem);
        launch_pdl(attention_kernel<48>, dim3(grid), dim3(96), smem, st, qkv, ldq, relk, relv, window, out, ldo, H, segs, tpad);
    } else throw_launch_error("attention: unsupported head size (96 / 48)");
    g_launch_count++;
}

void launch_attn_softmax(float* S, int Tp, const float* qkv, int ldq, const float* relk, const float* relv, int window,
                         float* orel, int ldo, int H, int heads, int RX, const SegInfo* segs, const int* seg_of_gran,
                         int gran, int max_len, cudaStream_t st) {
    const int D = H / heads;
    dim3 grid((RX + 7) / 8, heads);
    if (D != 96 || max_len > 1280 || 2 * window + 1 > 32) throw_launch_error("attn_softmax: unsupported head size / length");
    if (max_len <= 640)
        launch_pdl(attn_softmax_kernel<96, 20>, dim3(grid), dim3(256), 0, st, S, Tp, qkv, ldq, relk, relv, window, orel, ldo, RX, segs, seg_of_gran, gran);
    else
        launch_pdl(attn_softmax_kernel<96, 40>, dim3(grid), dim3(256), 0, st, S, Tp, qkv, ldq, relk, relv, window, orel, ldo, RX, segs, seg_of_gran, gran);
    g_launch_count++;
}

void launch_flow_pre(const float* z, int zcol, const float* w, const float* b, const float* g, float* h, int C,
                     RowMap map, cudaStream_t st) {
    const long long n = (long long)map.rows * C;
    launch_pdl(flow_pre_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, z, zcol, w, b, g, h, C, map);
    g_launch_count++;
}

void launch_spline(const float* h29, int ldh, float* z, int tcol, int bins, float inv_sqrt_filter, RowMap map,
                   cudaStream_t st) {
    (void)bins;   // 10 bins is the only configuration Piper ships
    launch_pdl(spline_kernel<10>, dim3((map.rows + 127) / 128), dim3(128), 0, st, h29, ldh, z, tcol, inv_sqrt_filter, map);
    g_launch_count++;
}

void launch_scale_copy2(const float* eps, float s, float* z, RowMap map, cudaStream_t st) {
    launch_pdl(scale_copy2_kernel, dim3((map.rows + 255) / 256), dim3(256), 0, st, eps, s, z, map);
    g_launch_count++;
}

void launch_durations(const float* z, float m0, float logs0, float length_scale, const SegInfo* segs, int nseg,
                      float* logw, int* cum, int* y_len, cudaStream_t st) {
    launch_pdl(durations_kernel, dim3(nseg), dim3(256), 0, st, z, m0, logs0, length_scale, segs, logw, cum, y_len);
    g_launch_count++;
}

void launch_expand(const float* stats, int ldst, int I, const int* cum, const float* eps, float noise_scale,
                   float* zp, const FrameSeg* fsegs, const int* ftile_seg, RowMap ymap, cudaStream_t st) {
    launch_pdl(expand_kernel, dim3((ymap.rows + 7) / 8), dim3(256), 0, st, stats, ldst, I, cum, eps, noise_scale, zp, fsegs, ftile_seg, ymap);
    g_launch_count++;
}

void launch_conv_post(const float* x, int C, const float* w, float* wav, const FrameSeg* fsegs,
                      const int* ftile_seg, int U, RowMap map, cudaStream_t st) {
    if (U % 4 != 0) throw_launch_error("conv_post: samples per frame must be a multiple of 4");
    const unsigned grid = (map.rows + 511) / 512;
    const size_t smem = sizeof(float) * ((size_t)C * 544 + (size_t)C * 8);
    const int vec_ok = (reinterpret_cast<uintptr_t>(wav) & 15) == 0;     // out_off is a multiple of U samples
    switch (C) {
        case 16: set_smem(conv_post_kernel<16>, smem); launch_pdl(conv_post_kernel<16>, dim3(grid), dim3(256), smem, st, x, w, wav, fsegs, ftile_seg, U, map, vec_ok); break;
        case 32: set_smem(conv_post_kernel<32>, smem); launch_pdl(conv_post_kernel<32>, dim3(grid), dim3(256), smem, st, x, w, wav, fsegs, ftile_seg, U, map, vec_ok); break;
        case 64: set_smem(conv_post_kernel<64>, smem); launch_pdl(conv_post_kernel<64>, dim3(grid), dim3(256), smem, st, x, w, wav, fsegs, ftile_seg, U, map, vec_ok); break;
        default: throw_launch_error("conv_post: unsupported channel count (16 / 32 / 64)");
    }
    g_launch_count++;
}

void launch_i16(const float* wav, const FrameSeg* fsegs, int nseg, int hop, long long max_samples, unsigned* maxbits,
                short* out, const PcmPost& post, cudaStream_t st) {
    if (nseg <= 0) return;
    cudaMemsetAsync(maxbits, 0, sizeof(unsigned) * nseg, st);
    int bx = (int)((max_samples + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    dim3 grid(bx, nseg);
    launch_pdl(i16_absmax_kernel, dim3(grid), dim3(256), 0, st, wav, fsegs, hop, maxbits, post);
    launch_pdl(i16_convert_kernel, dim3(grid), dim3(256), 0, st, wav, fsegs, hop, maxbits, out, post);
    g_launch_count += 2;
}

void launch_randn(float* out, long long n, unsigned long long seed, unsigned long long stream_id, cudaStream_t st) {
    const long long n4 = (n + 3) / 4;
    launch_pdl(randn_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, out, n, seed, stream_id);
    g_launch_count++;
}

void launch_cond_bias(const float* w, const float* base, const float* g, int rows, int gin, float* out, cudaStream_t st) {
    launch_pdl(cond_bias_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, w, base, g, rows, gin, out);
    g_launch_count++;
}

void launch_fill_zero(float* p, long long n, cudaStream_t st) {
    const long long n4 = n / 4;
    launch_pdl(fill_zero_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, reinterpret_cast<float4*>(p), n4);
    g_launch_count++;
}

}  // namespace sb200
