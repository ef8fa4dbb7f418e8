// extern "C" surface of libsonata_b200 (declared in include/sonata_b200.h).
// Error handling mirrors ffi_support::call_with_result used by libsonata (capi/src/lib.rs:187-336):
// no exception crosses the ABI; failures become {code, heap message}.
#include "../../include/sonata_b200.h"
#include "engine.h"
#include "tc_common.cuh"
#include <chrono>
#include <cstring>
#include <deque>

using namespace sb200;

// Handles.  A job returns its context to the voice's pool when it dies and a latent belongs to a voice, so both share
// ownership of the voice: freeing the voice handle first (garbage-collected callers free in any order) is safe.
struct sb200_voice { std::shared_ptr<Voice> v; };
struct sb200_job {
    Job* j; std::shared_ptr<Voice> keep;
    ~sb200_job() { delete j; }
};
struct sb200_latent {
    Latent* l; std::shared_ptr<Voice> keep;
    ~sb200_latent() { delete l; }
};

namespace {

// ---- pinned result blocks, shared by the sb200_audio entries of one batch and recycled ----
struct PinnedBlock { std::atomic<int> refs{0}; float* base = nullptr; size_t bytes = 0; };
std::mutex g_pin_mu;
std::deque<PinnedBlock*> g_pin_free;
std::unordered_map<const float*, PinnedBlock*> g_owner;   // audio.data -> block

// Smallest pooled block that fits (a larger one is fine: an exact-size policy made a caller that alternates between
// big and small batches pay cudaMallocHost + cudaFreeHost, ~1 ms, on every small call).
PinnedBlock* pin_acquire(size_t bytes) {
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        auto best = g_pin_free.end();
        for (auto it = g_pin_free.begin(); it != g_pin_free.end(); ++it)
            if ((*it)->bytes >= bytes && (best == g_pin_free.end() || (*it)->bytes < (*best)->bytes)) best = it;
        if (best != g_pin_free.end()) {
            PinnedBlock* b = *best;
            g_pin_free.erase(best);
            return b;
        }
    }
    PinnedBlock* b = new PinnedBlock();
    b->bytes = bytes + bytes / 8 + 4096;
    void* p = nullptr;
    if (cudaMallocHost(&p, b->bytes) != cudaSuccess) { delete b; throw Error(19, "cudaMallocHost failed for the result buffer"); }
    b->base = (float*)p;
    return b;
}
void pin_release(PinnedBlock* b) {
    PinnedBlock* victim = nullptr;
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        g_pin_free.push_back(b);
        if (g_pin_free.size() > 8) { victim = g_pin_free.front(); g_pin_free.pop_front(); }     // oldest goes
    }
    if (victim) { cudaFreeHost(victim->base); delete victim; }
}

char* dup_cstr(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

template <typename F>
int32_t guarded(sb200_error* err, F&& f) {
    if (err) { err->code = 0; err->message = nullptr; }
    try {
        f();
        return 0;
    } catch (const Error& e) {
        if (err) { err->code = e.code; err->message = dup_cstr(e.what()); }
        return e.code;
    } catch (const std::exception& e) {
        if (err) { err->code = 19; err->message = dup_cstr(e.what()); }
        return 19;
    } catch (...) {
        if (err) { err->code = -1; err->message = dup_cstr("panic"); }
        return -1;
    }
}

void fetch_audio(Job& j, sb200_audio* outs, float wall_ms) {
    if (!j.ran || j.encode_only) throw Error(19, "job has not produced audio");
    Voice& v = *j.v;
    SB_CUDA(cudaSetDevice(v.device));
    PinnedBlock* blk = pin_acquire((size_t)j.total_samples * 4 + 16);
    cudaError_t e = cudaMemcpyAsync(blk->base, j.d_wav, (size_t)j.total_samples * 4, cudaMemcpyDeviceToHost, j.ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(j.ctx->stream);
    if (e != cudaSuccess) { pin_release(blk); throw Error(19, std::string("CUDA error: ") + cudaGetErrorString(e)); }
    blk->refs = (int)j.B;
    const int hop = v.a.hop();
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (size_t b = 0; b < j.B; b++) {
        outs[b].data = blk->base + j.fsegs[b].out_off;
        outs[b].len = (size_t)j.y_len[b] * hop;
        outs[b].sample_rate = (uint32_t)v.sample_rate;
        outs[b].inference_ms = wall_ms * (j.total_samples ? (float)outs[b].len / (float)j.total_samples : 0.f);
        g_owner[outs[b].data] = blk;
    }
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

const char* sb200_version(void) { return "sonata_b200 0.1.0 (sm_100a)"; }
void sb200_string_free(char* s) { free(s); }
void sb200_ids_free(int64_t* ids) { free(ids); }
void sb200_buffer_free(float* p) { free(p); }
int32_t sb200_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }

void sb200_audio_free(sb200_audio* a) {
    if (!a || !a->data) return;
    PinnedBlock* blk = nullptr;
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        auto it = g_owner.find(a->data);
        if (it != g_owner.end()) { blk = it->second; g_owner.erase(it); }
    }
    if (blk) { if (--blk->refs == 0) pin_release(blk); }
    else free(a->data);
    a->data = nullptr; a->len = 0;
}

int32_t sb200_voice_load(const char* config_path, int32_t device, sb200_voice** out, sb200_error* err) {
    return guarded(err, [&] {
        if (!config_path || !out) throw Error(19, "null argument");
        Voice* v = load_voice(config_path, device);
        *out = new sb200_voice{std::shared_ptr<Voice>(v)};
    });
}
void sb200_voice_free(sb200_voice* v) { delete v; }

int32_t sb200_audio_output_info(const sb200_voice* v, sb200_audio_info* out, sb200_error* err) {
    return guarded(err, [&] {
        out->sample_rate = (uint32_t)v->v->sample_rate; out->num_channels = 1; out->sample_width = 2;
    });
}

static void cfg_out(const SynthConfig& c, sb200_synth_config* o) {
    o->speaker = c.speaker; o->has_speaker = c.has_speaker ? 1 : 0;
    o->noise_scale = c.noise_scale; o->length_scale = c.length_scale; o->noise_w = c.noise_w;
}
int32_t sb200_get_default_synthesis_config(const sb200_voice* v, sb200_synth_config* out, sb200_error* err) {
    return guarded(err, [&] {   // always Some(0), piper/src/lib.rs:444-451
        SynthConfig c = v->v->factory_cfg; c.speaker = 0; c.has_speaker = true; cfg_out(c, out);
    });
}
int32_t sb200_get_fallback_synthesis_config(const sb200_voice* v, sb200_synth_config* out, sb200_error* err) {
    return guarded(err, [&] { std::shared_lock<std::shared_mutex> g(v->v->cfg_mu); cfg_out(v->v->cfg, out); });
}
int32_t sb200_set_fallback_synthesis_config(sb200_voice* v, const sb200_synth_config* c, sb200_error* err) {
    return guarded(err, [&] {   // _do_set_default_synth_config, piper/src/lib.rs:215-231
        std::unique_lock<std::shared_mutex> g(v->v->cfg_mu);
        v->v->cfg.length_scale = c->length_scale; v->v->cfg.noise_scale = c->noise_scale; v->v->cfg.noise_w = c->noise_w;
        if (c->has_speaker) {
            bool found = false;
            for (auto& kv : v->v->speaker_id_map) if (kv.second == c->speaker) found = true;
            if (!found) throw Error(19, "No speaker was found with the given id `" + std::to_string(c->speaker) + "`");
            v->v->cfg.speaker = c->speaker; v->v->cfg.has_speaker = true;
        }
    });
}
int32_t sb200_get_language(const sb200_voice* v, char** out, sb200_error* err) {
    return guarded(err, [&] { *out = dup_cstr(v->v->language_code.empty() ? v->v->espeak_voice : v->v->language_code); });
}
int32_t sb200_get_quality(const sb200_voice* v, char** out, sb200_error* err) {
    return guarded(err, [&] { *out = dup_cstr(v->v->quality.empty() ? "unknown" : v->v->quality); });
}
int32_t sb200_supports_streaming_output(const sb200_voice* v) { return v->v->streaming ? 1 : 0; }
int32_t sb200_num_speakers(const sb200_voice* v) { return v->v->num_speakers; }
int64_t sb200_speaker_name_to_id(const sb200_voice* v, const char* name) {
    auto it = v->v->speaker_id_map.find(name ? name : "");
    return it == v->v->speaker_id_map.end() ? -1 : it->second;
}

int32_t sb200_phonemes_to_input_ids(const sb200_voice* v, const char* ph, int64_t** ids, size_t* n, sb200_error* err) {
    return guarded(err, [&] {
        std::vector<long long> r = v->v->phonemes_to_ids(ph);
        *ids = (int64_t*)malloc(r.size() * sizeof(int64_t));
        for (size_t i = 0; i < r.size(); i++) (*ids)[i] = r[i];
        *n = r.size();
    });
}

int32_t sb200_speak_batch_ids(sb200_voice* v, const int64_t* ids, const size_t* offsets, size_t batch,
                              sb200_audio* outs, sb200_error* err) {
    return guarded(err, [&] {
        const double t0 = now_ms();
        static_assert(sizeof(long long) == sizeof(int64_t), "");
        std::unique_ptr<Job> j(create_job(v->v.get(), reinterpret_cast<const long long*>(ids), offsets, batch, nullptr,
                                          nullptr, nullptr, false));
        j->run(nullptr, 0);
        fetch_audio(*j, outs, 0.f);
        const float wall = (float)(now_ms() - t0);
        for (size_t b = 0; b < batch; b++)
            outs[b].inference_ms = wall * (j->total_samples ? (float)outs[b].len / (float)j->total_samples : 0.f);
    });
}
int32_t sb200_speak_ids(sb200_voice* v, const int64_t* ids, size_t n, sb200_audio* out, sb200_error* err) {
    const size_t offs[2] = {0, n};
    return sb200_speak_batch_ids(v, ids, offs, 1, out, err);
}
int32_t sb200_speak_batch(sb200_voice* v, const char* const* ph, size_t batch, sb200_audio* outs, sb200_error* err) {
    std::vector<int64_t> ids; std::vector<size_t> offs{0};
    int32_t rc = guarded(err, [&] {
        for (size_t b = 0; b < batch; b++) {
            std::vector<long long> r = v->v->phonemes_to_ids(ph[b]);
            ids.insert(ids.end(), r.begin(), r.end());
            offs.push_back(ids.size());
        }
    });
    if (rc) return rc;
    return sb200_speak_batch_ids(v, ids.data(), offs.data(), batch, outs, err);
}
int32_t sb200_speak_one_sentence(sb200_voice* v, const char* ph, sb200_audio* out, sb200_error* err) {
    return sb200_speak_batch(v, &ph, 1, out, err);
}

// ---- job API ----
int32_t sb200_job_create(sb200_voice* v, const int64_t* ids, const size_t* offsets, size_t batch,
                         const float* const* eps_w, const float* const* eps_z, const size_t* eps_z_frames,
                         sb200_job** out, sb200_error* err) {
    return guarded(err, [&] {
        Job* j = create_job(v->v.get(), reinterpret_cast<const long long*>(ids), offsets, batch, eps_w, eps_z, eps_z_frames, false);
        *out = new sb200_job{j, v->v};
    });
}
int32_t sb200_job_set_debug(sb200_job* job, int32_t on) { job->j->debug = on != 0; return 0; }
int32_t sb200_job_run(sb200_job* job, float* d_out, size_t cap, float* device_ms, sb200_error* err) {
    return guarded(err, [&] { job->j->run(d_out, cap); if (device_ms) *device_ms = job->j->last_ms; });
}
int32_t sb200_job_fetch(sb200_job* job, sb200_audio* outs, sb200_error* err) {
    return guarded(err, [&] { fetch_audio(*job->j, outs, job->j->last_ms); });
}
int32_t sb200_job_fetch_i16(sb200_job* job, int16_t** outs, size_t* lens, sb200_error* err) {
    return guarded(err, [&] {
        Job& j = *job->j;
        if (!j.ran || j.encode_only) throw Error(19, "job has not produced audio");
        Voice& v = *j.v;
        SB_CUDA(cudaSetDevice(v.device));
        const int hop = v.a.hop();
        long long mx = 0;
        for (size_t b = 0; b < j.B; b++) mx = std::max<long long>(mx, (long long)j.y_len[b] * hop);
        short* d_i16 = nullptr; unsigned* d_max = nullptr;
        SB_CUDA(cudaMallocAsync(&d_i16, (size_t)j.total_samples * 2 + 16, j.ctx->stream));
        SB_CUDA(cudaMallocAsync(&d_max, sizeof(unsigned) * j.B, j.ctx->stream));
        launch_i16(j.d_wav, j.d_fsegs, (int)j.B, hop, mx, d_max, d_i16, PcmPost{}, j.ctx->stream);
        PinnedBlock* blk = pin_acquire((size_t)j.total_samples * 2 + 16);
        cudaError_t e = cudaMemcpyAsync(blk->base, d_i16, (size_t)j.total_samples * 2, cudaMemcpyDeviceToHost, j.ctx->stream);
        cudaFreeAsync(d_i16, j.ctx->stream);
        cudaFreeAsync(d_max, j.ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(j.ctx->stream);
        if (e != cudaSuccess) { pin_release(blk); throw Error(19, std::string("CUDA error: ") + cudaGetErrorString(e)); }
        const int16_t* h = reinterpret_cast<const int16_t*>(blk->base);
        for (size_t b = 0; b < j.B; b++) {
            const size_t n = (size_t)j.y_len[b] * hop;
            outs[b] = (int16_t*)malloc(n * 2 + 2);
            memcpy(outs[b], h + j.fsegs[b].out_off, n * 2);
            lens[b] = n;
        }
        pin_release(blk);
    });
}
void sb200_i16_free(int16_t* p) { free(p); }

int32_t sb200_job_copy_out(sb200_job* job, void* dst, size_t cap, int32_t format, size_t* written, sb200_error* err) {
    return guarded(err, [&] {
        Job& j = *job->j;
        if (!j.ran || j.encode_only) throw Error(19, "job has not produced audio");
        if (!dst) throw Error(19, "null destination");
        Voice& v = *j.v;
        SB_CUDA(cudaSetDevice(v.device));
        cudaStream_t st = j.ctx->stream;
        const size_t n = (size_t)j.total_samples;
        const size_t bytes = n * (format == 1 ? 2 : 4);
        if (bytes > cap) throw Error(19, "destination buffer is too small for the synthesis result");
        if (format == 1) {
            const int hop = v.a.hop();
            long long mx = 0;
            for (size_t b = 0; b < j.B; b++) mx = std::max<long long>(mx, (long long)j.y_len[b] * hop);
            short* d_i16 = nullptr; unsigned* d_max = nullptr;
            SB_CUDA(cudaMallocAsync(&d_i16, n * 2 + 16, st));
            SB_CUDA(cudaMallocAsync(&d_max, sizeof(unsigned) * j.B, st));
            launch_i16(j.d_wav, j.d_fsegs, (int)j.B, hop, mx, d_max, d_i16, PcmPost{}, st);
            cudaError_t e = cudaMemcpyAsync(dst, d_i16, bytes, cudaMemcpyDeviceToHost, st);
            cudaFreeAsync(d_i16, st);
            cudaFreeAsync(d_max, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) throw Error(19, std::string("CUDA error: ") + cudaGetErrorString(e));
        } else {
            SB_CUDA(cudaMemcpyAsync(dst, j.d_wav, bytes, cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
        }
        if (written) *written = bytes;
    });
}
int32_t sb200_host_register(void* ptr, size_t bytes, sb200_error* err) {
    return guarded(err, [&] {
        const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
        if (e != cudaSuccess) { cudaGetLastError(); throw Error(19, std::string("cudaHostRegister failed: ") + cudaGetErrorString(e)); }
    });
}
int32_t sb200_host_unregister(void* ptr) {
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();
    return e == cudaSuccess ? 0 : 19;
}
size_t sb200_job_batch(const sb200_job* job) { return job->j->B; }
int32_t sb200_job_lengths(const sb200_job* job, int64_t* frames, int64_t* samples, int64_t* out_offsets) {
    const Job& j = *job->j;
    if (!j.ran) return 19;
    const int hop = j.v->a.hop();
    for (size_t b = 0; b < j.B; b++) {
        if (frames) frames[b] = j.y_len[b];
        if (samples) samples[b] = (int64_t)j.y_len[b] * hop;
        if (out_offsets) out_offsets[b] = j.fsegs[b].out_off;
    }
    return 0;
}
void sb200_job_free(sb200_job* job) { delete job; }

// ---- streaming halves ----
int32_t sb200_encode_ids(sb200_voice* v, const int64_t* ids, size_t n, sb200_latent** out, sb200_error* err) {
    return guarded(err, [&] { *out = new sb200_latent{encode_latent(v->v.get(), reinterpret_cast<const long long*>(ids), n), v->v}; });
}
int64_t sb200_latent_frames(const sb200_latent* z) { return z->l->frames; }
int32_t sb200_decode_chunk(sb200_voice* v, const sb200_latent* z, int64_t lo, int64_t hi, sb200_audio* out, sb200_error* err) {
    return guarded(err, [&] {
        std::vector<float> w; float ms = 0;
        decode_latent_chunk(v->v.get(), z->l, lo, hi, w, &ms);
        out->data = (float*)malloc(w.size() * 4 + 4);
        memcpy(out->data, w.data(), w.size() * 4);
        out->len = w.size(); out->inference_ms = ms; out->sample_rate = (uint32_t)v->v->sample_rate;
    });
}
void sb200_latent_free(sb200_latent* z) { delete z; }

// ---- introspection ----
int32_t sb200_job_debug_fetch(sb200_job* job, const char* name, size_t b, float** data, size_t* rows, size_t* cols, sb200_error* err) {
    return guarded(err, [&] {
        Job& j = *job->j;
        auto it = j.dbg.find(name);
        if (!j.ran || it == j.dbg.end()) throw Error(19, std::string("no debug buffer named `") + name + "` (was debug enabled before run?)");
        const int U = j.dbg_level[name];
        const int C = it->second.second;
        size_t r0, nr;
        if (U == 0) { r0 = (size_t)j.xsegs[b].off; nr = (size_t)j.xsegs[b].len; }
        else { r0 = (size_t)j.fsegs[b].off * U; nr = (size_t)j.fsegs[b].len * U; }
        *data = (float*)malloc(nr * C * 4 + 4);
        SB_CUDA(cudaSetDevice(j.v->device));
        SB_CUDA(cudaMemcpy(*data, it->second.first + r0 * C, nr * C * 4, cudaMemcpyDeviceToHost));
        *rows = nr; *cols = (size_t)C;
    });
}
int32_t sb200_job_debug_durations(sb200_job* job, size_t b, int32_t** cum, size_t* n, sb200_error* err) {
    return guarded(err, [&] {
        Job& j = *job->j;
        if (!j.ran) throw Error(19, "job not run");
        const size_t len = (size_t)j.xsegs[b].len;
        *cum = (int32_t*)malloc(len * 4 + 4);
        SB_CUDA(cudaSetDevice(j.v->device));
        SB_CUDA(cudaMemcpy(*cum, j.d_cum + j.xsegs[b].off, len * 4, cudaMemcpyDeviceToHost));
        *n = len;
    });
}
int32_t sb200_job_profile(const sb200_job* job, sb200_region_stat* out, int32_t cap) {
    const Job& j = *job->j;
    int32_t n = 0;
    for (const Region& r : j.regions) {
        if (n >= cap) break;
        memset(&out[n], 0, sizeof(out[n]));
        strncpy(out[n].name, r.name.c_str(), sizeof(out[n].name) - 1);
        out[n].ms = r.ms; out[n].flops = r.flops; out[n].bytes = r.bytes; out[n].launches = r.launches;
        n++;
    }
    return n;
}
int32_t sb200_debug_plan(int32_t backend, int64_t rows, int32_t cin, int32_t cout, int32_t k, int32_t dil, int32_t act,
                         int32_t has_res, int32_t accumulate, int32_t* out16) {
    // Planning only: nothing is allocated or launched, so this also runs where there is no GPU (host-logic tests); tensor
    // maps are assumed available, as on any sm_90+ driver.  Buffers are described by stand-in addresses (the planners look
    // at alignment only).  Layer conventions as in sb200_debug_conv / the engine's ResBlock and flow layers.
    if (!out16 || cin <= 0 || cout <= 0 || k <= 0 || k > SB_MAX_TAPS || dil <= 0 || rows <= 0) return 19;
    static float anchor[64];
    float* const stand_in = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(anchor) + 127) & ~(uintptr_t)127);
    ConvArgs p{};
    const int R = (int)((rows + 255) / 256 * 256);
    const int ycols = act == ACT_GATE ? cout / 2 : cout;
    p.x = stand_in; p.ldx = cin; p.rows_in = R; p.cin = cin; p.in_slope = 0.1f;
    p.w = stand_in; p.bias = stand_in; p.ldw = cout; p.cout = cout;
    p.tc_nt = cout <= 128 ? cout : (cout % 128 == 0 ? 128 : (cout % 96 == 0 ? 96 : 0));      // voice.cu tc_tile_for
    p.wtc = p.tc_nt ? stand_in : nullptr; p.wcat = (p.tc_nt && p.tc_nt <= 64) ? stand_in : nullptr; p.wtf = stand_in;
    p.ntaps = k;
    for (int t = 0; t < k; t++) p.tap_off[t] = (t - (k - 1) / 2) * dil;
    p.min_off = p.tap_off[0]; p.span = (k - 1) * dil;
    p.rows_q = R; p.orow_mul = 1; p.orow_add = 0;
    p.act = act; p.scale = 1.f; p.res = has_res ? stand_in : nullptr; p.ldres = cout;
    p.y0 = stand_in; p.ldy0 = ycols; p.acc0 = accumulate; p.split = cout; p.y1 = stand_in; p.ldy1 = ycols; p.acc1 = accumulate;
    plan_assume_tensor_maps() = true;
    const bool ok = backend == 2 ? conv_tf_plan_info(p, out16) : conv_tc_plan_info(p, out16);
    plan_assume_tensor_maps() = false;
    return ok ? 0 : 19;
}

int32_t sb200_debug_conv(int32_t device, int32_t backend, const float* x, int32_t rows, int32_t cin, const float* w,
                         const float* bias, int32_t cout, int32_t k, int32_t dil, float in_slope, int32_t act,
                         const float* res, float scale, int32_t accumulate, float* y, int32_t valid_rows,
                         sb200_error* err) {
    return guarded(err, [&] {
        SB_CUDA(cudaSetDevice(device));
        Voice tmp; tmp.device = device;
        ConvW cw = debug_make_conv(tmp, w, bias, cout, cin, k, dil);
        const int R = (rows + 255) / 256 * 256;
        const int ycols = act == ACT_GATE ? cout / 2 : cout;
        float *dx, *dy, *dres = nullptr; int* dend;
        SB_CUDA(cudaMalloc(&dx, (size_t)R * cin * 4)); SB_CUDA(cudaMemset(dx, 0, (size_t)R * cin * 4));
        SB_CUDA(cudaMemcpy(dx, x, (size_t)rows * cin * 4, cudaMemcpyHostToDevice));
        SB_CUDA(cudaMalloc(&dy, (size_t)R * ycols * 4)); SB_CUDA(cudaMemset(dy, 0, (size_t)R * ycols * 4));
        SB_CUDA(cudaMemcpy(dy, y, (size_t)rows * ycols * 4, cudaMemcpyHostToDevice));
        if (res) { SB_CUDA(cudaMalloc(&dres, (size_t)R * cout * 4)); SB_CUDA(cudaMemset(dres, 0, (size_t)R * cout * 4));
                   SB_CUDA(cudaMemcpy(dres, res, (size_t)rows * cout * 4, cudaMemcpyHostToDevice)); }
        SB_CUDA(cudaMalloc(&dend, 4)); SB_CUDA(cudaMemcpy(dend, &valid_rows, 4, cudaMemcpyHostToDevice));
        ConvArgs p{};
        p.x = dx; p.ldx = cin; p.rows_in = R; p.cin = cin; p.in_slope = in_slope;
        p.w = cw.w; p.bias = cw.bias; p.ldw = cw.ldw; p.cout = cw.cout; p.wtc = cw.wtc; p.tc_nt = cw.tc_nt; p.wcat = cw.wcat;
        p.ntaps = cw.ntaps; memcpy(p.tap_off, cw.tap_off, sizeof(p.tap_off)); p.min_off = cw.min_off; p.span = cw.span;
        p.rows_q = R; p.orow_mul = 1; p.orow_add = 0;
        p.map = RowMap{dend, R, 1, R};
        p.act = act; p.scale = scale; p.res = dres; p.ldres = cout;
        p.y0 = dy; p.ldy0 = ycols; p.acc0 = accumulate; p.split = cout; p.y1 = dy; p.ldy1 = ycols; p.acc1 = accumulate;
        if (res && getenv("SB200_DEBUG_RES_IS_X") && cin == cout) { p.res = dx; p.ldres = cin; }   // ResBlock aliasing (timing only)
        long long* dtrace = nullptr;
        const bool want_trace = backend == 1 && getenv("SB200_TC_TRACE") != nullptr;
        if (want_trace) { SB_CUDA(cudaMalloc(&dtrace, 48 * 8 * 8)); SB_CUDA(cudaMemset(dtrace, 0, 48 * 8 * 8)); }
        p.wtf = cw.wtf;
        if (backend == 2) {
            if (!conv_tf_supported(p)) throw Error(19, "conv shape not supported by the tf32 chunk-flush backend");
            launch_conv_tf(p, 0);
        } else if (backend == 1) {
            if (!conv_tc_supported(p)) throw Error(19, "conv shape not supported by the tcgen05 backend");
            if (want_trace) launch_conv_tc(p, 0); // warm-up before the traced run (timing only: RMW cases run twice)
            p.trace = dtrace;
            launch_conv_tc(p, 0);
        } else launch_conv_simt(p, 0);
        if (want_trace) {
            SB_CUDA(cudaDeviceSynchronize());
            std::vector<long long> t(48 * 8);
            SB_CUDA(cudaMemcpy(t.data(), dtrace, 48 * 8 * 8, cudaMemcpyDeviceToHost));
            const long long t0 = t[0];
            fprintf(stderr, "tile: prod_issue prod_landed prod_conv | mma_afull mma_issued | epi_accfull epi_tmem epi_stored  (cycles rel. to first issue)\n");
            for (int i = 0; i < 48; i++) {
                fprintf(stderr, "%3d:", i);
                for (int k = 0; k < 8; k++) fprintf(stderr, " %9lld", t[i * 8 + k] ? t[i * 8 + k] - t0 : -1LL);
                fprintf(stderr, "\n");
            }
            cudaFree(dtrace);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(y, dy, (size_t)rows * ycols * 4, cudaMemcpyDeviceToHost);
        cudaFree(dx); cudaFree(dy); cudaFree(dend); if (dres) cudaFree(dres);
        if (e != cudaSuccess) throw Error(19, std::string("CUDA error: ") + cudaGetErrorString(e));
    });
}

uint64_t sb200_launch_count(void) { return g_launch_count; }
int32_t sb200_set_backend(sb200_voice* v, int32_t backend) { int32_t p = v->v->backend; v->v->backend = backend; return p; }

}  // extern "C"
