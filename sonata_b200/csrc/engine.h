// Host runtime of libsonata_b200: voice (weights + config), per-call contexts, batched job.
// This is the C++ stand-in for the reference's Rust `VitsModel` (piper/src/lib.rs:291-478): it
// owns what `ort::Session` owns there (weights, execution resources) and mirrors the model-side
// state (`ModelConfig`, `RwLock<PiperSynthesisConfig>`).
#pragma once
#include "common.cuh"
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace sb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SB_CUDA(x)                                                                                       \
    do {                                                                                                 \
        cudaError_t e__ = (x);                                                                           \
        if (e__ != cudaSuccess)                                                                          \
            throw ::sb200::Error(19, std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " +    \
                                         __FILE__ + ":" + std::to_string(__LINE__));                     \
    } while (0)

struct HostTensor {
    std::vector<int> dims;
    std::vector<float> f;
    std::vector<int> i;
    bool is_int = false;
    size_t numel() const { size_t n = 1; for (int d : dims) n *= (size_t)d; return n; }
};

struct Arch {
    int hidden, inter, filter, heads, layers, kernel, window, n_vocab, resblock, up_init, flow_n, wn_layers,
        flow_kernel, dp_kernel, dp_bins, sample_rate;
    std::vector<int> up_rates, up_kernels, res_kernels;
    std::vector<std::vector<int>> res_dils;
    int hop() const { int h = 1; for (int u : up_rates) h *= u; return h; }
};

struct ConvW {
    float* w = nullptr;
    float* bias = nullptr;
    float* wtc = nullptr; int tc_nt = 0;     // tcgen05 hi/lo swizzled weight images
    float* wcat = nullptr;                   // hi/lo-stacked tap-pair images (conv_tc.cu cat mode), column tiles <= 64
    float* wtf = nullptr;                    // tf32 hi/lo images (conv_tf.cu): text-encoder / duration-predictor layers
    int cin = 0, cout = 0, ldw = 0, ntaps = 0;
    int cond_off = -1;                       // multi-speaker voices: offset of this conv's per-call effective bias (Job::d_cond)
    int tap_off[SB_MAX_TAPS] = {0};
    int min_off = 0, span = 0;
};

struct EncLayer { ConvW qkv, o, ffn1, ffn2; float *relk, *relv, *g1, *b1, *g2, *b2; };
struct DDSW { float* wdw[3]; float* bdw[3]; ConvW c1x1[3]; float *g1[3], *b1[3], *g2[3], *b2[3]; };
struct CFlowW { float* pre_w; float* pre_b; DDSW dds; ConvW proj; int ccol, tcol; };
struct CouplingW { ConvW pre; std::vector<ConvW> in, rs; ConvW post; int cond_off, tgt_off; };
struct ResBW { int k; std::vector<int> dils; std::vector<ConvW> c1, c2; };
struct UpStageW { int u, k, cin, cout; std::vector<ConvW> phase; ConvW fused; std::vector<ResBW> res; };   // fused: all phases as one N = u*cout conv (tcgen05 path)

struct SynthConfig { long long speaker = 0; bool has_speaker = false; float noise_scale = 0.667f, length_scale = 1.f, noise_w = 0.8f; };

struct Context;   // stream + arenas for one in-flight call

struct Voice {
    // ---- config (ModelConfig, piper/src/lib.rs:143-158) ----
    std::string config_path, key, quality, language_code, espeak_voice;
    int sample_rate = 22050;
    int num_speakers = 1;
    int num_symbols = 0;
    bool streaming = false;
    std::map<std::string, long long> speaker_id_map;
    std::unordered_map<uint32_t, long long> phoneme_first_id;   // char (code point) -> first id
    SynthConfig factory_cfg;
    mutable std::shared_mutex cfg_mu;     // RwLock<PiperSynthesisConfig> (piper/src/lib.rs:292)
    SynthConfig cfg;

    // ---- weights ----
    int device = 0;
    Arch a;
    std::vector<void*> dev_allocs;
    float* emb = nullptr;
    std::vector<EncLayer> enc;
    ConvW enc_proj;
    ConvW dp_pre, dp_proj;
    DDSW dp_dds;
    std::vector<CFlowW> dp_flows;   // in application order (CF4, CF3, CF2)
    float ea_m0 = 0, ea_logs0 = 0;
    std::vector<CouplingW> flows;   // in application order (f = n-1 .. 0)
    ConvW conv_pre;
    std::vector<UpStageW> ups;
    // multi-speaker conditioning (num_speakers > 1): see the end of load_voice
    float* emb_g = nullptr; int gin = 0, emb_rows = 0;
    float *cond_w = nullptr, *cond_base = nullptr; int cond_rows = 0;
    float* conv_post_w = nullptr;   // [7][C_last]
    int c_last = 0;
    size_t weight_bytes = 0;

    int backend = 1;                // 1 (default): tcgen05 everywhere (bf16x2 split for flow + decoder, chunk-flushed 3xTF32 for the text
                                    // encoder + duration predictor); 2: tcgen05 flow + decoder, fp32 CUDA cores for encoder + predictor;
                                    // 0: fp32 CUDA cores everywhere
    unsigned long long noise_seed = 0x5eed5eedULL;
    std::mutex pool_mu;
    std::vector<Context*> pool;
    std::atomic<unsigned long long> call_counter{0};

    ~Voice();
    Context* acquire();
    void release(Context* c);
    std::vector<long long> phonemes_to_ids(const char* utf8) const;
};

Voice* load_voice(const std::string& config_path, int device);
ConvW debug_make_conv(Voice& v, const float* w, const float* bias, int cout, int cin, int k, int dil);

struct Region { std::string name; cudaEvent_t e0, e1; double flops = 0, bytes = 0; int launches = 0; float ms = 0; };

struct Arena {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    bool dry = false;
    void* alloc(size_t bytes) {
        const size_t a = (used + 255) & ~(size_t)255;
        used = a + bytes;
        if (dry) return nullptr;
        if (used > cap) throw Error(19, "internal: device arena overflow");
        return base + a;
    }
    template <typename T> T* get(size_t n) { return reinterpret_cast<T*>(alloc(n * sizeof(T))); }
};

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    Arena dev;          // device workspace
    char* pin = nullptr; size_t pin_cap = 0;   // pinned staging for small tables
    std::vector<cudaEvent_t> events; size_t events_used = 0;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    void ensure_dev(size_t bytes);
    void ensure_pin(size_t bytes);
    cudaEvent_t next_event();
    ~Context();
};

struct Level {          // one time resolution of the packed batch
    RowMap map;
    long long valid_rows = 0;
};

struct Job {
    Voice* v = nullptr;
    Context* ctx = nullptr;
    SynthConfig cfg;
    size_t B = 0;
    bool debug = false;
    bool encode_only = false;     // stop after the flow (streaming 'encoder.onnx' half)
    float* z_dev = nullptr;
    unsigned long long noise_call = 0;
    // host copies of inputs
    std::vector<long long> ids; std::vector<size_t> offs;
    std::vector<std::vector<float>> eps_w, eps_z; std::vector<size_t> eps_z_frames;
    // X layout
    int RX = 0; std::vector<SegInfo> xsegs; int max_tx = 0;
    // Y layout
    int RY = 0; std::vector<FrameSeg> fsegs; std::vector<int> y_len;
    long long total_samples = 0;
    // device pointers (ctx arena)
    int *d_ids_rows = nullptr, *d_xend = nullptr, *d_cum = nullptr, *d_ylen = nullptr, *d_yend = nullptr, *d_ftile = nullptr;
    SegInfo* d_xsegs = nullptr; FrameSeg* d_fsegs = nullptr;
    // tensor-core attention (conv_tf.cu grouped GEMMs): tile tables built with the X layout, same for every layer
    std::vector<TfTile> tiles_s, tiles_o;      // Q.K^T tiles, P.V tiles
    int att_tp = 0;                            // key columns of a score row (multiple of 96)
    int att_nth_s = 64, att_nth_o = 96;   // column tiles of the two attention GEMMs (narrower when the job is small)
    int* d_xseg_of_gran = nullptr; TfTile *d_tiles_s = nullptr, *d_tiles_o = nullptr;
    float *d_epsw = nullptr, *d_epsz = nullptr;
    float* d_wav = nullptr; bool wav_external = false;
    float* d_cond = nullptr;       // effective biases of the speaker-conditioned convs for this call
    std::map<std::string, std::pair<float*, int>> dbg;   // name -> (device ptr, cols)
    std::map<std::string, int> dbg_level;                // name -> U (rows per frame) or 0 for X level
    std::vector<Region> regions;
    float last_ms = 0;
    bool ran = false;

    ~Job();
    void run(float* d_out, size_t d_out_cap);
};

Job* create_job(Voice* v, const long long* ids, const size_t* offs, size_t B, const float* const* eps_w,
                const float* const* eps_z, const size_t* eps_z_frames, bool debug);

struct Latent {
    Voice* v = nullptr;
    long long sid = 0;    // speaker of the encoder pass (the reference hands `g` from encoder.onnx to decoder.onnx)
    float* z = nullptr;   // device [frames][inter]
    long long frames = 0;
    ~Latent();
};
Latent* encode_latent(Voice* v, const long long* ids, size_t n);
void decode_latent_chunk(Voice* v, const Latent* z, long long lo, long long hi, std::vector<float>& out, float* ms);
// the same chunk as peak-normalised i16 PCM with the reference's post-path done on the DEVICE: drop trim_lo / trim_hi
// overlap frames, crossfade(fade) (samples.rs:144-157), linear gain, to_i16_vec (samples.rs:51-75)
void decode_latent_chunk_pcm(Voice* v, const Latent* z, long long lo, long long hi, long long trim_lo_frames,
                             long long trim_hi_frames, int fade, float gain, std::vector<int16_t>& out, float* ms);
void job_pcm16(Job& j, float gain, std::vector<std::vector<int16_t>>& out);

}  // namespace sb200
