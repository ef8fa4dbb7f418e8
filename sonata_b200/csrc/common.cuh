// Shared declarations for libsonata_b200 (sm_100a only).
//
// Activation layout everywhere: TIME-MAJOR fp32 matrices  A[row][channel]  (channel contiguous).
// A "row" is one time step (phoneme id at the X level, frame at the Y level, sample-group at the
// decoder levels).  Utterances of a batch are concatenated along rows as SEGMENTS; segment b
// occupies rows [off_b, off_b + T_b) and is followed by >= HALO all-zero gap rows, so a
// convolution that reads across a segment edge sees the zero padding the reference's B=1
// onnxruntime run would see (piper/src/lib.rs:433-435 runs every utterance alone).
// Every kernel that produces an activation writes ZERO into gap rows (or leaves accumulated
// buffers untouched there), which keeps that invariant without memsets.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <mutex>

#define SB_MAX_TAPS 16

namespace sb200 {

// Row validity: row q is a real time step iff  q < seg_end[q / gran] * seg_mul.
// seg_end is indexed by "granule" (64 ids at the X level, 128 frames at the Y level; a granule
// never straddles two segments) and holds off_b + T_b in granule-level rows; decoder levels
// that run at U x the frame rate pass gran = 128*U, seg_mul = U.
struct RowMap {
    const int* seg_end;
    int gran;
    int seg_mul;
    int rows;   // total rows at this level
};

__device__ __forceinline__ bool row_valid(const RowMap& m, int q) {
    if (q < 0 || q >= m.rows) return false;
    return q < m.seg_end[q / m.gran] * m.seg_mul;
}

enum ConvAct { ACT_NONE = 0, ACT_RELU = 1, ACT_GATE = 2 };

// One implicit-GEMM convolution:  for GEMM row q, output column n
//   acc = bias[n] + sum_t sum_c  f(x[q + tap_off[t]][c]) * w[t][c][n],   f = leaky_relu(in_slope)
//   v   = act(acc) (+ res[orow][n]) ; v *= scale ; y[orow][n] (=|+=) v,   orow = q*orow_mul + orow_add
struct ConvArgs {
    const float* x; int ldx; int rows_in; int cin; float in_slope;
    const float* w; const float* bias; int ldw; int cout;
    const float* wtc; int tc_nt;                                  // tcgen05 weight images (conv_tc.cu) or null
    const float* wcat;                                            // hi/lo-stacked tap-pair images (conv_tc.cu cat mode) or null
    const float* wtf;                                             // tf32 hi/lo images (conv_tf.cu) or null
    int ntaps; int tap_off[SB_MAX_TAPS]; int min_off; int span;   // span = max_off - min_off
    int rows_q; int orow_mul; int orow_add;
    int phase_cols;                                               // >0: fused polyphase ConvTranspose (tcgen05 path only)
    long long* trace;                                             // optional per-role clock64() timeline (debug)
    RowMap map;                                                   // validity of q
    int act; float scale;
    const float* res; int ldres;
    float* y0; int ldy0; int acc0; int split;                     // columns [0, split)
    float* y1; int ldy1; int acc1;                                // columns [split, cout)
    float* yt; int yt_col0; int ldyt;                             // conv_tf.cu: column tiles >= yt_col0 are stored TRANSPOSED,
                                                                  // yt[(n - yt_col0) * ldyt + q] (row index contiguous)
};

// One tile of a grouped GEMM on conv_tf.cu's kernel (the attention contractions): two 128-row m-tiles of A against
// one block of nth B rows over `nkb` 32-column K-blocks;  C = scale * A . B^T (+ res).
struct TfTile {
    int a_row0[2];        // first row of m-tile h in A (rows outside the array read as zeros)
    int a_col0;           // first K column in A
    int b_row0;           // first row of the B block
    int b_col0;           // first K column in B
    int nkb;              // 32-column K-blocks
    int rows_valid[2];    // rows of m-tile h that are stored
    long long out_off[2]; // element offset of (row 0, column 0) of m-tile h's output block in y (and res)
};
struct TfGemm {
    const float* a; int a_rows, a_cols, lda;      // A [a_rows][a_cols], K along the columns
    const float* b; int b_rows, b_cols, ldb;      // B [b_rows][b_cols], rows = output columns
    int nth;                                      // B rows (= output columns) per tile: 96 / 64 / 32
    float* y; int ldy; const float* res; float scale;
    const TfTile* tiles; int ntiles;              // device table
};

// Function attributes (opt-in shared-memory size) are per DEVICE: `run` executes `f` the first time the calling site
// runs on the current device (a process driving several GPUs through the C ABI configures each of them).  The device
// bit is published only AFTER `f` returned, under a mutex, so a second host thread on the same device can never
// launch with more than 48 KB of dynamic shared memory before the opt-in has been applied.
struct PerDeviceOnce {
    std::mutex mu;
    unsigned long long done = 0;
    template <typename F> void run(F&& f) {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        if (__atomic_load_n(&done, __ATOMIC_ACQUIRE) & bit) return;
        std::lock_guard<std::mutex> g(mu);
        if (done & bit) return;
        f();
        __atomic_fetch_or(&done, bit, __ATOMIC_RELEASE);
    }
};

// Programmatic dependent launch.  The step is a chain of ~170-200 dependent kernels; at single-utterance sizes most of
// them run for 5-30 us, so the launch gap and each kernel's prologue (mbarrier init, TMEM allocation, tensor-map fetch)
// are a visible part of the step.  Every kernel of the library is launched with the programmatic-stream-serialization
// attribute and starts with pdl_wait() BEFORE its first access to global memory (tcgen05 kernels: after their prologue),
// after a pdl_trigger() at its very top: the following kernels' CTAs may become resident and run their own prologues (and
// fetch their weights, which no kernel writes) while this kernel is still working; their pdl_wait() returns only when the
// preceding grid has completed and its writes are visible.  Nothing before a pdl_wait() touches memory another kernel
// writes, and a kernel whose CTAs are all resident can always finish, so running ahead cannot deadlock.
// SB200_NO_PDL=1 launches without the attribute (the two instructions are then no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const int allowed = getenv("SB200_NO_PDL") ? 0 : 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = allowed;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Experiment knobs of the launch planners come from the environment.  getenv() scans the whole environment block
// (~1 us), and a planner runs twice per launch with up to seven knobs: 170 launches of a single-utterance call spent
// more host time there than the GPU needed for the kernels.  Each knob is read ONCE per process.
#define SB_ENV_ONCE(name) ([]() -> const char* { static const char* v = getenv(name); return v; }())

void launch_conv_simt(const ConvArgs& a, cudaStream_t st);
int conv_simt_bn_for(int cout);
bool conv_tc_supported(const ConvArgs& a);
bool conv_tc_plan_info(const ConvArgs& a, int* out16);     // planning only, see sb200_debug_plan
void launch_conv_tc(const ConvArgs& a, cudaStream_t st);
bool try_launch_conv_tc(const ConvArgs& a, cudaStream_t st);
size_t conv_tc_weight_floats(int cin, int cout, int ntaps, int nt);
void conv_tc_build_weights(const float* wt, int ldw, int cin, int cout, int ntaps, int nt, float* out);
size_t conv_tc_cat_weight_floats(int cin, int cout, int ntaps, int nt);
void conv_tc_build_weights_cat(const float* wt, int ldw, int cin, int cout, int ntaps, int nt, float* out);
bool conv_tf_supported(const ConvArgs& a);
bool conv_tf_plan_info(const ConvArgs& a, int* out16);
void launch_conv_tf(const ConvArgs& a, cudaStream_t st);
bool try_launch_conv_tf(const ConvArgs& a, cudaStream_t st);
size_t conv_tf_weight_floats(int cin, int cout, int ntaps);
void conv_tf_build_weights(const float* wt, int ldw, int cin, int cout, int ntaps, float* out);
bool gemm_tf_supported(const TfGemm& g);
void launch_gemm_tf(const TfGemm& g, cudaStream_t st);
void throw_launch_error(const char* what);
      // column tile the SIMT kernel will use for this cout (for weight padding)

struct SegInfo { int off; int len; };   // rows

// ---- misc kernels (kernels_misc.cu) ----
void launch_embed(const int* ids_rows, const float* emb, float scale, float* x, int rows, int H, cudaStream_t st);
// out = (res2 ? res2 : 0) + act( LN(x + (res1 ? res1 : 0)) * gamma + beta ),  act: 0 none, 1 exact GELU
void launch_ln(const float* x, const float* res1, const float* res2, const float* gamma, const float* beta,
               float* out, int C, int act, RowMap map, cudaStream_t st);
// out = GELU(LN(depthwise_conv_k(x, dilation)))   (DDSConv first half)
void launch_dw_ln_gelu(const float* x, const float* wdw /*[k][C]*/, const float* bdw, int k, int dil,
                       const float* gamma, const float* beta, float* out, int C, RowMap map, cudaStream_t st);
// Relative-position softmax between the two attention GEMMs (in place on the score rows): see kernels_misc.cu
void launch_attn_softmax(float* S, int Tp, const float* qkv, int ldq, const float* relk, const float* relv, int window,
                         float* orel, int ldo, int H, int heads, int RX, const SegInfo* segs, const int* seg_of_gran,
                         int gran, int max_len, cudaStream_t st);
void launch_attention(const float* qkv, int ldq, const float* relk, const float* relv, int window,
                      float* out, int ldo, int H, int heads, const SegInfo* segs, int nseg, int max_len,
                      cudaStream_t st);
size_t attention_smem_bytes(int max_len, int D);
// h[r][c] = w[c]*z[r][zcol] + b[c] + g[r][c]
void launch_flow_pre(const float* z, int zcol, const float* w, const float* b, const float* g, float* h,
                     int C, RowMap map, cudaStream_t st);
// z[r][tcol] = RQS^-1(z[r][tcol]; params h29[r][0..3*bins-1))
void launch_spline(const float* h29, int ldh, float* z, int tcol, int bins, float inv_sqrt_filter,
                   RowMap map, cudaStream_t st);
// logw = (z[:,0]-m0)*exp(-logs0); w = exp(logw)*length_scale; w_ceil; per-segment inclusive scan
void launch_durations(const float* z, float m0, float logs0, float length_scale, const SegInfo* segs, int nseg,
                      float* logw, int* cum, int* y_len, cudaStream_t st);
struct FrameSeg { int off; int len; int xoff; int xlen; long long out_off; };
// z_p rows: gather m_p/logs_p of the token whose cumulative duration covers the frame, add noise
void launch_expand(const float* stats, int ldst, int I, const int* cum, const float* eps, float noise_scale,
                   float* zp, const FrameSeg* fsegs, const int* ftile_seg, RowMap ymap, cudaStream_t st);
// wav = tanh(conv_k7(lrelu_{0.01}(x)))  ->  compact per-segment output
void launch_conv_post(const float* x, int C, const float* w /*[7][C]*/, float* wav, const FrameSeg* fsegs,
                      const int* ftile_seg, int U, RowMap map, cudaStream_t st);
// per-utterance peak-normalised f32 -> i16 (audio-ops `to_i16_vec`), out indexed like wav
// what the reference applies to a chunk before the 16-bit conversion (see kernels_misc.cu): overlap trim (samples),
// crossfade table of fade_n <= 48 entries, linear gain.  Default = plain to_i16_vec.
struct PcmPost { float gain = 1.f; int fade_n = 0; long long trim_lo = 0, trim_hi = 0; float tab[48] = {0}; };
void launch_i16(const float* wav, const FrameSeg* fsegs, int nseg, int hop, long long max_samples, unsigned* maxbits,
                short* out, const PcmPost& post, cudaStream_t st);
void launch_randn(float* out, long long n, unsigned long long seed, unsigned long long stream_id, cudaStream_t st);
void launch_scale_copy2(const float* eps, float s, float* z, RowMap map, cudaStream_t st);   // z[r][0..1] = eps*s
void launch_fill_zero(float* p, long long n, cudaStream_t st);
// out[r] = base[r] + sum_k w[r][k] * g[k]   (speaker conditioning: effective biases of the conditioned convs)
void launch_cond_bias(const float* w, const float* base, const float* g, int rows, int gin, float* out, cudaStream_t st);

// launch-configuration errors are not sticky and would otherwise be lost: throw immediately
void check_launch(const char* what);

extern unsigned long long g_launch_count;   // kernels launched by this library (host-side counter)

}  // namespace sb200
