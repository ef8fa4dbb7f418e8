/*
 * sonata_b200.h — C ABI of libsonata_b200.so: the B200-native replacement for the one hot path of
 * mush42/sonata, the Piper/VITS phoneme -> waveform synthesis that the reference delegates to
 * onnxruntime (`ort::Session::run`, crates/sonata/models/piper/src/lib.rs:362-379).
 *
 * A Rust `impl SonataModel for B200Vits` (or any FFI host) binds exactly these symbols; each entry
 * point cites the reference interface it replaces.  Plain pointers and sizes only — no torch /
 * CUDA types cross this boundary.  All functions are thread-safe per voice (the reference calls
 * `speak_one_sentence` concurrently from rayon workers on one model, synth/src/lib.rs:316-320).
 *
 * Error convention (mirrors ffi_support's ExternError used by libsonata, capi/libsonata.h:41-49):
 * every fallible call returns 0 on success or an error code and, when `err` is non-NULL, fills
 * `err->code` / `err->message` (heap string, free with sb200_string_free).  Codes reuse
 * libsonata's: 17 FAILED_TO_LOAD_RESOURCE, 18 PHONEMIZATION_ERROR, 19 OPERATION_ERROR
 * (capi/libsonata.h:10-16; SonataError variants at core/src/lib.rs:19-24).
 */
#ifndef SONATA_B200_H
#define SONATA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_OK 0
#define SB200_FAILED_TO_LOAD_RESOURCE 17
#define SB200_PHONEMIZATION_ERROR 18
#define SB200_OPERATION_ERROR 19

typedef struct sb200_voice sb200_voice;     /* = Arc<dyn SonataModel> holding a VitsModel (piper/src/lib.rs:291-297) */
typedef struct sb200_job sb200_job;         /* one batched synthesis in flight */
typedef struct sb200_latent sb200_latent;   /* = EncoderOutputs {z, y_mask} (piper/src/lib.rs:671-677) */

typedef struct sb200_error {
    int32_t code;
    char* message;
} sb200_error;

/* = sonata_core::Audio {samples: Vec<f32>, info.sample_rate, inference_ms} (audio/ops/src/samples.rs:208-214).
 * `data` is library-owned pinned host memory; release with sb200_audio_free. */
typedef struct sb200_audio {
    float* data;
    size_t len;
    float inference_ms;     /* fractional ms (the reference truncates to whole ms, piper/src/lib.rs:380) */
    uint32_t sample_rate;
} sb200_audio;

/* = PiperSynthesisConfig (piper/src/lib.rs:160-166); has_speaker==0 <=> speaker: None */
typedef struct sb200_synth_config {
    int64_t speaker;
    int32_t has_speaker;
    float noise_scale;
    float length_scale;
    float noise_w;
} sb200_synth_config;

/* = AudioInfo (audio/ops/src/samples.rs:9-14; values fixed at piper/src/lib.rs:282-288) */
typedef struct sb200_audio_info {
    uint32_t sample_rate;
    uint32_t num_channels;
    uint32_t sample_width;
} sb200_audio_info;

/* ---- library ---- */
const char* sb200_version(void);
void sb200_string_free(char* s);
void sb200_audio_free(sb200_audio* a);
int32_t sb200_device_count(void);

/* ---- voice: sonata_piper::from_config_path (piper/src/lib.rs:88-110) ----
 * `config_path` is the Piper `<voice>.onnx.json`; weights are read from the sibling `<voice>.svw`
 * (where the reference opens `<voice>.onnx`).  `device` = CUDA ordinal; -1 loads the config only
 * (host-side queries and id mapping work, every synthesis call fails with OPERATION_ERROR). */
int32_t sb200_voice_load(const char* config_path, int32_t device, sb200_voice** out, sb200_error* err);
/* Drops the caller's handle.  Jobs and latents created from the voice share ownership of it (the reference holds the
 * model behind an Arc, capi/src/lib.rs:314,375): they stay valid and may be freed afterwards, in any order. */
void sb200_voice_free(sb200_voice* v);

/* SonataModel::audio_output_info (core/src/lib.rs:83) */
int32_t sb200_audio_output_info(const sb200_voice* v, sb200_audio_info* out, sb200_error* err);
/* SonataModel::get_default_synthesis_config / get_fallback_ / set_fallback_ (core/src/lib.rs:88-90;
 * piper/src/lib.rs:444-462, 215-231: unknown speaker id -> OPERATION_ERROR) */
int32_t sb200_get_default_synthesis_config(const sb200_voice* v, sb200_synth_config* out, sb200_error* err);
int32_t sb200_get_fallback_synthesis_config(const sb200_voice* v, sb200_synth_config* out, sb200_error* err);
int32_t sb200_set_fallback_synthesis_config(sb200_voice* v, const sb200_synth_config* cfg, sb200_error* err);
/* SonataModel::get_language / properties["quality"] / supports_streaming_output (piper/src/lib.rs:180-196, 649-651).
 * Returned strings are heap copies: sb200_string_free. */
int32_t sb200_get_language(const sb200_voice* v, char** out, sb200_error* err);
int32_t sb200_get_quality(const sb200_voice* v, char** out, sb200_error* err);
int32_t sb200_supports_streaming_output(const sb200_voice* v);
/* SonataModel::get_speakers / speaker_name_to_id (piper/src/lib.rs:466-471): returns -1 when unknown */
int32_t sb200_num_speakers(const sb200_voice* v);
int64_t sb200_speaker_name_to_id(const sb200_voice* v, const char* name);

/* VitsModelCommons::phonemes_to_input_ids (piper/src/lib.rs:232-250): [bos] + (id(ch), pad)* + [eos];
 * unknown characters are dropped silently, only the first id of a map entry is used.
 * `*ids` is malloc'ed (free with sb200_ids_free). */
int32_t sb200_phonemes_to_input_ids(const sb200_voice* v, const char* phonemes_utf8, int64_t** ids, size_t* n,
                                    sb200_error* err);
void sb200_ids_free(int64_t* ids);

/* ---- synthesis ---- */
/* SonataModel::speak_one_sentence(phonemes: String) (piper/src/lib.rs:439-443) */
int32_t sb200_speak_one_sentence(sb200_voice* v, const char* phonemes_utf8, sb200_audio* out, sb200_error* err);
/* SonataModel::speak_batch(Vec<String>) (piper/src/lib.rs:425-437).  Same per-utterance result as B
 * sequential calls (the reference loops B=1 runs), computed as one batched pass. */
int32_t sb200_speak_batch(sb200_voice* v, const char* const* phonemes_utf8, size_t batch, sb200_audio* outs,
                          sb200_error* err);
/* VitsModel::infer_with_values(Vec<i64>) (piper/src/lib.rs:342-399) */
int32_t sb200_speak_ids(sb200_voice* v, const int64_t* ids, size_t n, sb200_audio* out, sb200_error* err);
/* batched infer_with_values: utterance b = ids_packed[offsets[b] .. offsets[b+1]) */
int32_t sb200_speak_batch_ids(sb200_voice* v, const int64_t* ids_packed, const size_t* offsets, size_t batch,
                              sb200_audio* outs, sb200_error* err);

/* ---- job API: the same batched pass split into its host<->device steps (bench / multi-GPU plumbing) ----
 * create  : copies ids to the device (H2D).  `eps_w` / `eps_z` optionally inject the graph's two
 *           RandomNormalLike draws (time-major: eps_w[b] = f32[T_x][2], eps_z[b] = f32[T_y][inter]);
 *           NULL -> Philox noise on the device (skipped when the matching scale is 0).
 * run     : all kernels; the waveform stays in HBM (optionally written into caller device memory
 *           `d_out`, capacity in floats, e.g. an NCCL send buffer); returns device time of the pass.
 * fetch   : D2H of the per-utterance waveforms into pinned host memory. */
int32_t sb200_job_create(sb200_voice* v, const int64_t* ids_packed, const size_t* offsets, size_t batch,
                         const float* const* eps_w, const float* const* eps_z, const size_t* eps_z_frames,
                         sb200_job** out, sb200_error* err);
/* keep every intermediate of the next run fetchable through sb200_job_debug_fetch (tests only) */
int32_t sb200_job_set_debug(sb200_job* job, int32_t on);
int32_t sb200_job_run(sb200_job* job, float* d_out, size_t d_out_capacity, float* device_ms, sb200_error* err);
int32_t sb200_job_fetch(sb200_job* job, sb200_audio* outs, sb200_error* err);
size_t sb200_job_batch(const sb200_job* job);
/* per-utterance results available after run: frames (T_y), samples (256*T_y), offset into d_out */
/* Peak-normalised 16-bit PCM of every utterance of a finished job, converted on the device (half the device->host
 * bytes).  Mirrors AudioSamples::to_i16_vec / as_wave_bytes (crates/audio/ops/src/samples.rs:51-78) bit for bit.
 * outs[b] receives a malloc'ed buffer of lens[b] samples: free with sb200_i16_free. */
int32_t sb200_job_fetch_i16(sb200_job* job, int16_t** outs, size_t* lens, sb200_error* err);
void sb200_i16_free(int16_t* p);
int32_t sb200_job_lengths(const sb200_job* job, int64_t* frames, int64_t* samples, int64_t* out_offsets);
/* Copy the result of a finished job into CALLER-OWNED host memory, utterances back to back in batch order.
 * format 0: f32 samples (what infer_with_values returns, piper/src/lib.rs:382-392);
 * format 1: i16 PCM, peak-normalised per utterance on the device (= AudioSamples::to_i16_vec, samples.rs:51-75;
 *           what libsonata hands to its callback, capi/src/lib.rs:416-438) -- half the device->host bytes.
 * `dst` may be ordinary or page-locked memory (see sb200_host_register), e.g. a slice of a segment shared by the
 * per-GPU worker processes of one frontend.  *written = bytes written; fails if `capacity_bytes` is too small. */
int32_t sb200_job_copy_out(sb200_job* job, void* dst, size_t capacity_bytes, int32_t format, size_t* written,
                           sb200_error* err);
/* Page-lock / unlock caller memory for DMA (cudaHostRegister) so that hosts without CUDA bindings (Rust, ctypes)
 * can pin a result segment once and reuse it.  Returns 0 on success. */
int32_t sb200_host_register(void* ptr, size_t bytes, sb200_error* err);
int32_t sb200_host_unregister(void* ptr);
void sb200_job_free(sb200_job* job);

/* ---- streaming: VitsStreamingModel (piper/src/lib.rs:480-669) ----
 * encode = infer_encoder (:537-574): ids -> latent z [T_y][inter] kept on the device.
 * decode_chunk = decoder.onnx on z[:, :, lo:hi] (:793-840) -> 256*(hi-lo) samples (no crossfade here). */
int32_t sb200_encode_ids(sb200_voice* v, const int64_t* ids, size_t n, sb200_latent** out, sb200_error* err);
int64_t sb200_latent_frames(const sb200_latent* z);
int32_t sb200_decode_chunk(sb200_voice* v, const sb200_latent* z, int64_t frame_lo, int64_t frame_hi,
                           sb200_audio* out, sb200_error* err);
void sb200_latent_free(sb200_latent* z);

/* ---- introspection for tests / bench ---- */
/* Copy a named intermediate of the LAST run of `job` to host (time-major fp32, valid rows of
 * utterance b only).  Names: "x","stats","logw","z_p","z","dec.pre","dec.up<i>","dec.mrf<i>".
 * Returns rows via *rows, cols via *cols; data malloc'ed (free with sb200_buffer_free). */
int32_t sb200_job_debug_fetch(sb200_job* job, const char* name, size_t b, float** data, size_t* rows, size_t* cols,
                              sb200_error* err);
void sb200_buffer_free(float* p);
/* cumulative durations (int32 per id) of utterance b */
int32_t sb200_job_debug_durations(sb200_job* job, size_t b, int32_t** cum, size_t* n, sb200_error* err);

typedef struct sb200_region_stat {
    char name[32];
    double ms;            /* device time between the region's CUDA events, last run */
    double flops;         /* algorithmic FLOPs (valid rows only, 2 per MAC) */
    double bytes;         /* layer-wise algorithmic bytes (each operand read once, result written once) */
    int32_t launches;
} sb200_region_stat;
/* region statistics of the last run of `job`; returns the number written (<= cap) */
int32_t sb200_job_profile(const sb200_job* job, sb200_region_stat* out, int32_t cap);
/* one convolution on caller data through either backend (kernel unit tests):
 * y[rows][cout] (=|+=) scale * (act(bias + conv_k,dil(lrelu_slope(x))) + res); w is [cout][cin][k];
 * act 0 none, 1 relu, 2 tanh*sigmoid gate (y is [rows][cout/2]); rows >= valid_rows are masked. */
/* Test hook: the launch configuration the planner of backend 1 (tcgen05 bf16x2 conv) / 2 (tcgen05 3xTF32 conv) would choose
 * for one convolution of `rows` output rows -- nothing is allocated or launched, so it also works without a GPU.
 * out16, backend 1: {nt, image rows, m-tiles, n-tiles, resident, cat, tma_epilogue, pairs, act. stages, weight stages,
 * staging tiles, smem bytes, tma_in, v8, tmem columns, window rows}; backend 2: {nth, image rows, m-pairs, n-tiles, act.
 * stages, weight stages, chunk K-blocks, smem bytes, tmem columns, window rows, 0...}.  Returns 0, or 19 if unsupported. */
int32_t sb200_debug_plan(int32_t backend, int64_t rows, int32_t cin, int32_t cout, int32_t k, int32_t dil, int32_t act,
                         int32_t has_res, int32_t accumulate, int32_t* out16);
int32_t sb200_debug_conv(int32_t device, int32_t backend, const float* x, int32_t rows, int32_t cin, const float* w,
                         const float* bias, int32_t cout, int32_t k, int32_t dil, float in_slope, int32_t act,
                         const float* res, float scale, int32_t accumulate, float* y, int32_t valid_rows,
                         sb200_error* err);
/* kernels launched by this library since load (host-side counter) */
uint64_t sb200_launch_count(void);
/* select the contraction backend: 0 = fp32 CUDA-core implicit GEMM, 1 = tcgen05 (3xTF32) where
 * implemented; returns the previous value */
int32_t sb200_set_backend(sb200_voice* v, int32_t backend);

#ifdef __cplusplus
}
#endif
#endif /* SONATA_B200_H */
