"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the committed
golden vectors.  Two-stage check everywhere: durations EXACT, then waveform
sample-wise max-abs < 1e-3 (the tolerance BASELINE.json states; measured error is ~1e-5)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))      # conv_unit.py (kernel unit cases, torch reference only)
sys.path.insert(0, os.path.join(ROOT, "tests"))      # stage_report.py (uses the oracle: test infrastructure)

import sonata_b200
from oracle import vits_oracle as vo
from sonata_b200 import PiperSynthesisConfig, voicegen, workload
from sonata_b200.job import SynthesisJob

pytestmark = pytest.mark.gpu
TOL_WAV = 1e-3          # BASELINE.json: "waveform max-abs error <1e-3"
TOL_STAGE = 2e-4        # per-stage activations are O(1..5); fp32 path measures ~1e-5
# logw comes out of three inverse rational-quadratic spline flows whose derivative may be as small as 1e-3 (the
# graph's min_derivative), i.e. the INVERSE amplifies its input error by up to 1e3 at a few ids per utterance (the
# fp32 oracle itself is 1.1e-4 away from its fp64 shadow there, profiles/notes_r01.md).  So logw is held to a tight
# MEDIAN and a loose max; what the graph consumes -- ceil(exp(logw)) -- is compared exactly.
TOL_LOGW_MAX = 2e-3
TOL_LOGW_MEDIAN = 1e-5


def _stage_tol(name):
    return TOL_WAV if name == "wav" else TOL_LOGW_MAX if name == "logw" else TOL_STAGE
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.fixture(scope="module")
def models(voice_paths):
    ms = {}

    def get(q):
        if q not in ms:
            ms[q] = sonata_b200.from_config_path(voice_paths[q], device=0)
        return ms[q]
    yield get
    for m in ms.values():
        m.close()


def _det(m):
    m.set_fallback_synthesis_config(PiperSynthesisConfig(None, 0.0, 1.0, 0.0))


@pytest.mark.parametrize("backend", [1, 0], ids=["tcgen05", "fp32simt"])
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_cuda_matches_golden_vectors(path, backend, models):
    g = np.load(path)
    q = os.path.basename(path).split("_")[0]
    m = models(q)
    m.set_backend(backend)
    sc = g["scales"]
    m.set_fallback_synthesis_config(PiperSynthesisConfig(None, float(sc[0]), float(sc[1]), float(sc[2])))
    ew = [g["eps_w"]] if "eps_w" in g else None
    ez = [g["eps_z"]] if "eps_z" in g else None
    job = SynthesisJob(m, [g["ids"]], ew, ez, debug=True)
    job.run()
    frames, samples, _ = job.lengths()
    assert np.array_equal(job.durations(0), g["cum"]), "durations must be exact"
    assert frames[0] == int(g["y_len"]) and samples[0] == 256 * int(g["y_len"])
    assert float(np.abs(job.debug_fetch("z", 0) - g["z"]).max()) < TOL_STAGE
    wav = job.fetch()[0].samples.as_slice()
    assert float(np.abs(wav - g["wav"]).max()) < TOL_WAV
    job.close()
    m.set_backend(1)


@pytest.mark.parametrize("backend", [1, 0], ids=["tcgen05", "fp32simt"])
@pytest.mark.parametrize("quality,ns,noise", [("medium", (16, 40, 5, 0, 1), False), ("medium", (9, 21), True),
                                               ("high", (12, 3), False)])
def test_every_stage_against_oracle(quality, ns, noise, backend):
    from stage_report import stage_report
    rep = stage_report(quality, ns, noise, backend=backend, verbose=False)
    for u in rep["utts"]:
        assert u["durations_exact"] and u["y_len_ref"] == u["y_len_got"], u
        names = [s[0] for s in u["stages"]]
        assert "wav" in names and "z" in names and "dec.mrf0" in names
        for name, err, ref_max in u["stages"]:
            assert err != "SHAPE", (name, u)
            assert err < _stage_tol(name), (name, err, u["n_ids"])
        assert u["logw_median_err"] < TOL_LOGW_MEDIAN, u["logw_median_err"]


# Full-size parity (BASELINE.json configs): utterance seeds screened by tests/screen_margin.py so that every duration
# w = exp(logw) stays >= 1e-3 away from an integer in the fp32 oracle AND its fp64 shadow -- there ceil(w) is well
# defined and the frame counts must match EXACTLY.  (seed, frames) pairs as printed by the screening script.
MARGIN = 1e-3
SCREENED = {
    "C1": ("medium", 128, [(2, 854)]),                                   # 1 x 128 phonemes, T_x = 258
    "C2": ("medium", 256, [(0, 1783), (1, 1794), (2, 1742), (5, 1819)]),  # 4 x 256 phonemes in ONE batch, T_x = 514
    "C3": ("high", 512, [(3, 3559)]),                                    # 1 x 512 phonemes, T_x = 1026
}


@pytest.mark.parametrize("backend", [1, 0], ids=["tcgen05", "fp32simt"])
@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])
def test_baseline_sizes_against_oracle(cfg, backend):
    """The CUDA path against the oracle at the BASELINE.json utterance sizes: durations exact, every stage
    (x, stats, logw, z_p, z, dec.pre, every dec.up*/dec.mrf*) < 2e-4, waveform < 1e-3.  Exercises the multi-tile
    attention path (T_x = 514 / 1026), the block scan of the duration kernel over long segments and the
    frame-level gap logic at 1.8k / 3.5k-frame segments."""
    from stage_report import stage_report
    quality, n, seeds = SCREENED[cfg]
    rep = stage_report(quality, [n] * len(seeds), False, backend=backend, verbose=False, utts=[s for s, _ in seeds])
    for u, (seed, frames) in zip(rep["utts"], seeds):
        assert u["n_ids"] == 2 * n + 2
        assert u["ceil_margin"] >= MARGIN, ("screening is stale: re-run tests/screen_margin.py", u["utt"], u["ceil_margin"])
        assert u["y_len_ref"] == frames, ("oracle frame count moved", u)
        assert u["durations_exact"] and u["y_len_got"] == frames, (cfg, seed, u["flipped_ids"], u["y_len_got"], frames)
        names = [s[0] for s in u["stages"]]
        assert {"x", "stats", "logw", "z_p", "z", "dec.pre", "dec.mrf0", "wav"} <= set(names), names
        for name, err, ref_max in u["stages"]:
            assert err != "SHAPE", (name, u)
            assert err < _stage_tol(name), (cfg, seed, name, err)
        assert u["logw_median_err"] < TOL_LOGW_MEDIAN, (cfg, seed, u["logw_median_err"])


@pytest.mark.parametrize("backend", [1, 0], ids=["tcgen05", "fp32simt"])
def test_multi_speaker_voice_against_oracle(backend, voice_paths):
    """a multi-speaker voice (`emb_g`, `dp.cond`, the WaveNets' `cond_layer`, `dec.cond`) -- the graph
    the reference feeds the `sid` tensor to whenever num_speakers > 1 (piper/src/lib.rs:353-358).  Every stage against
    the oracle for two speakers, and the speakers must differ."""
    from stage_report import stage_report
    waves = {}
    for sid in (0, 3):
        rep = stage_report("medium", (21, 40), False, backend=backend, verbose=False, n_speakers=4, sid=sid)
        for u in rep["utts"]:
            assert u["durations_exact"] and u["y_len_ref"] == u["y_len_got"], (sid, u)
            for name, err, ref_max in u["stages"]:
                assert err != "SHAPE" and err < _stage_tol(name), (sid, name, err)
        waves[sid] = [u["y_len_got"] for u in rep["utts"]]
    cfg = voicegen.write_voice(voicegen.default_voice_dir(), "medium", n_speakers=4)
    m = sonata_b200.from_config_path(cfg, device=0)
    assert m.get_speakers() == {i: f"speaker_{i}" for i in range(4)} and m.speaker_name_to_id("speaker_2") == 2
    ids = workload.synthetic_ids(30, utt=9)
    outs = []
    for sid in (None, 0, 2):                                   # speaker: None -> sid 0 (`unwrap_or(0)`)
        m.set_fallback_synthesis_config(PiperSynthesisConfig(sid, 0.0, 1.0, 0.0))
        outs.append(m.infer_with_values(ids).samples.as_slice().copy())
    assert np.array_equal(outs[0], outs[1])
    assert outs[2].shape != outs[1].shape or float(np.abs(outs[2] - outs[1]).max()) > 1e-2
    with pytest.raises(sonata_b200.OperationError):
        m.set_fallback_synthesis_config(PiperSynthesisConfig(9, 0.0, 1.0, 0.0))     # piper/src/lib.rs:215-231
    m.close()


def test_unscreened_duration_flips_are_cliff_cases():
    """Companion of the screened test: 8 x 256-phoneme utterances with ARBITRARY seeds.  A frame count may differ
    from the oracle's only where the oracle's own duration sits on the ceil() cliff (margin < 1e-3, where fp32
    and fp64 disagree among themselves in ~2 % of C2 batches); everywhere else durations are exact and the
    waveform is within tolerance.  The flip counts are reported (gpurun_out/flip_report.json), not failed on."""
    import json
    from stage_report import stage_report
    seeds = list(range(100, 108))
    rep = stage_report("medium", [256] * len(seeds), False, backend=1, verbose=False, utts=seeds)
    out = []
    for u in rep["utts"]:
        out.append({k: u[k] for k in ("utt", "n_ids", "ceil_margin", "durations_exact", "flipped_ids", "y_len_ref", "y_len_got")})
        if u["ceil_margin"] >= MARGIN:
            assert u["durations_exact"], u
        if u["durations_exact"]:
            wav = [s for s in u["stages"] if s[0] == "wav"][0]
            assert wav[1] < TOL_WAV, u
        else:
            assert u["flipped_ids"] <= 2 and abs(u["y_len_ref"] - u["y_len_got"]) <= u["flipped_ids"], u
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "flip_report.json"), "w") as f:
        json.dump({"margin": MARGIN, "utts": out, "flipped_utts": sum(not o["durations_exact"] for o in out)}, f, indent=1)


@pytest.mark.parametrize("backend", [1, 0], ids=["tcgen05", "fp32simt"])
def test_conv_kernels_against_torch(backend, lib_built):
    """Kernel-level unit check of both contraction backends (same ConvArgs contract): 1x1 / dilated k-tap,
    leaky-ReLU prologue, gate / ReLU / residual / scale / accumulate epilogues, masked rows, and the
    persistent multi-tile path of the tcgen05 kernel (several tiles per CTA on both half-pipelines)."""
    from conv_unit import CASES, run_case
    for c in CASES:
        err, msg = run_case(backend, *c)
        assert err is not None, (c, msg)
        assert err < 1e-4, (c, err)


def test_conv_tf_kernel_fp32_class_accuracy(lib_built):
    """conv_tf.cu (tcgen05 3xTF32, chunk-flushed accumulation) on every encoder / duration-predictor shape against an
    fp64 reference: the kernel replaces fp32 CUDA-core GEMMs in front of the ceil() cliff, so it is held to the
    error of an fp32 FMA chain (measured 2e-6 .. 1e-5), not to the 1e-4 of the bf16x2 decoder kernel."""
    from conv_unit import TF_CASES, TF_TOL, run_case
    for c in TF_CASES:
        err, msg = run_case(2, *c)
        assert err is not None, (c, msg)
        assert err < TF_TOL, (c, err)


@pytest.mark.parametrize("knob", ["SB200_TC_NOTMAST", "SB200_TC_NOTMAIN", "SB200_TC_NOV8", "SB200_TC_NOCAT"])
def test_conv_kernel_variants_stay_correct(knob, lib_built):
    """The fallbacks of the default conv path (no TMA-staged epilogue, cp.async window loads, 128-bit epilogue accesses,
    no hi/lo-stacked weight images) implement the same ConvArgs contract.
    The planner reads the knobs from the environment, so each variant runs in its own process."""
    import subprocess
    env = dict(os.environ, **{knob: "1"})
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "conv_unit.py")
    r = subprocess.run([sys.executable, tool, "1", "0,1,2,3,12,17,18,19,20,21,22,23,25,26"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_device_i16_matches_host_to_i16_vec(models):
    """the 16-bit PCM conversion (`to_i16_vec`: per-buffer peak normalisation, clamp, truncating cast,
    audio/ops/src/samples.rs:51-75) done on the device is bit-identical to the host mirror of the reference."""
    m = models("medium"); _det(m)
    batches = [workload.synthetic_ids(n, utt=70 + i) for i, n in enumerate((21, 5, 40))]
    job = SynthesisJob(m, batches)
    job.run()
    f32 = job.fetch()
    i16 = job.fetch_i16()
    assert len(i16) == len(f32)
    for a, q in zip(f32, i16):
        ref = a.samples.to_i16_vec()
        assert q.dtype == np.int16 and q.shape == ref.shape
        assert np.array_equal(q, ref)
        assert int(np.abs(q.astype(np.int32)).max()) >= 32766       # peak-normalised


def test_batched_equals_sequential(models):
    """speak_batch is a sequential B=1 loop in the reference (piper/src/lib.rs:433-435): the packed
    batched pass must give each utterance the result it gets alone."""
    m = models("medium"); _det(m)
    batches = [workload.synthetic_ids(n, utt=50 + i) for i, n in enumerate((30, 7, 64, 18))]
    together = m.infer_batch_with_values(batches)
    for b, ids in enumerate(batches):
        alone = m.infer_with_values(ids)
        assert len(alone) == len(together[b])
        assert float(np.abs(alone.samples.as_slice() - together[b].samples.as_slice()).max()) < 1e-5


def test_speak_api_surface(models, oracle_weights):
    m = models("medium"); _det(m)
    ph = "hɛloʊ wɜːld"
    a1 = m.speak_one_sentence(ph)
    a2 = m.infer_with_values(m.phonemes_to_input_ids(ph))
    assert np.array_equal(a1.samples.as_slice(), a2.samples.as_slice())          # deterministic at scales [0,1,0]
    assert a1.info.sample_rate == 22050 and a1.inference_ms > 0 and 0 < a1.real_time_factor() < 1
    ref = vo.infer(oracle_weights("medium"), m.phonemes_to_input_ids(ph), [0, 1, 0]).numpy()
    assert a1.samples.as_slice().shape == ref.shape and float(np.abs(a1.samples.as_slice() - ref).max()) < TOL_WAV
    outs = m.speak_batch([ph, "", "a"])
    assert len(outs) == 3 and len(outs[1]) % 256 == 0 and len(outs[1]) > 0       # "" -> [bos, eos]
    assert len(a1.as_wave_bytes()) == 2 * len(a1)
    # length_scale stretches durations (w = exp(logw) * length_scale)
    m.set_fallback_synthesis_config(PiperSynthesisConfig(None, 0.0, 1.7, 0.0))
    assert len(m.speak_one_sentence(ph)) > len(a1)
    _det(m)
    with pytest.raises(sonata_b200.OperationError):
        m.infer_with_values([1, 999, 2])                                        # id outside the embedding table


def test_noise_path_statistics(models):
    """Default scales use the on-device Philox source: results differ call to call (like the graph's
    RandomNormalLike) but stay finite, bounded by tanh, and of plausible length."""
    m = models("medium")
    m.set_fallback_synthesis_config(PiperSynthesisConfig(None, 0.667, 1.0, 0.8))
    ids = workload.synthetic_ids(40, utt=3)
    a, b = m.infer_with_values(ids), m.infer_with_values(ids)
    for x in (a, b):
        s = x.samples.as_slice()
        assert np.isfinite(s).all() and np.abs(s).max() <= 1.0 and len(s) % 256 == 0
        assert 1.5 < len(s) / 256 / len(ids) < 6.0
    n = min(len(a), len(b))
    assert not np.array_equal(a.samples.as_slice()[:n], b.samples.as_slice()[:n])
    _det(m)


def test_full_size_properties(models):
    """BASELINE config 2 (32 x 256 phonemes) at full size through size-independent properties."""
    m = models("medium"); _det(m)
    batches = [workload.synthetic_ids(256, utt=u) for u in range(32)]
    job = SynthesisJob(m, batches)
    ms = job.run()
    frames, samples, offs = job.lengths()
    assert all(s == 256 * f for s, f in zip(samples, frames))
    assert offs == list(np.concatenate([[0], np.cumsum(samples)[:-1]]))
    for b in (0, 17, 31):
        assert int(job.durations(b)[-1]) == frames[b]                           # sum of ceil'd durations == frames
    auds = job.fetch()
    for a in auds:
        s = a.samples.as_slice()
        assert np.isfinite(s).all() and np.abs(s).max() <= 1.0
    # utterance 5 of the big batch == the same utterance alone (batch-size independence)
    alone = m.infer_with_values(batches[5]).samples.as_slice()
    assert alone.shape == auds[5].samples.as_slice().shape
    assert float(np.abs(alone - auds[5].samples.as_slice()).max()) < 1e-5
    prof = {p["name"]: p for p in job.profile()}
    assert prof["dec.mrf2"]["launches"] == 6 and prof["dec.mrf2"]["ms"] > 0
    assert prof["dec.up2"]["launches"] == 1          # phase-fused ConvTranspose on the tensor-core backend
    assert ms > 0
    job.close()


def test_streaming_chunks_match_oracle(voice_paths, oracle_weights, tmp_path):
    """VitsStreamingModel: encoder half -> z on device; decoder half on frame slices with the
    reference's chunk schedule, overlap trim and crossfade(42) (piper/src/lib.rs:765-858)."""
    import json
    import shutil
    cfg = json.load(open(voice_paths["medium"], encoding="utf-8"))
    cfg["streaming"] = True
    p = tmp_path / "rt.onnx.json"
    json.dump(cfg, open(p, "w", encoding="utf-8"), ensure_ascii=False)
    os.symlink(voice_paths["medium"].replace(".onnx.json", ".svw"), tmp_path / "rt.svw")
    m = sonata_b200.from_config_path(str(p), device=0)
    assert isinstance(m, sonata_b200.VitsStreamingModel) and m.supports_streaming_output()
    _det(m)
    W = oracle_weights("medium")
    ids = workload.synthetic_ids(60, utt=77)
    st = {}
    full_ref = vo.infer(W, ids, [0, 1, 0], stages=st).numpy()
    z = st["z"]
    enc = m.infer_encoder(ids)
    assert enc.num_frames == st["y_len"]
    full = enc.infer_decoder().as_slice()
    assert float(np.abs(full - full_ref).max()) < TOL_WAV
    chunks = list(sonata_b200.SpeechStreamer(enc, 45, 3))
    assert len(chunks) > 1
    total = 0
    for ((m0, m1), (a0, a1)), got in zip(sonata_b200.AdaptiveMelChunker(enc.num_frames, 45, 3), chunks):
        hi = enc.num_frames if m1 is None else m1
        ref = vo.decode(W, z[:, :, m0:hi]).view(-1).numpy()
        ref = ref[a0:a1] if a1 is not None else ref[a0:]
        exp = sonata_b200.AudioSamples(ref); exp.crossfade(42)
        assert float(np.abs(got.as_slice() - exp.as_slice()).max()) < TOL_WAV
        total += len(got)
    assert total == 256 * enc.num_frames
    # one-shot rule: frames <= 2*chunk + 2*pad -> a single full decode (:785, :848-853)
    one = list(sonata_b200.SpeechStreamer(m.infer_encoder(workload.synthetic_ids(8, utt=1)), 72, 3))
    assert len(one) == 1
    m.close()


def test_concurrent_calls_are_reentrant(models):
    """The reference calls `speak_one_sentence` concurrently from rayon workers on ONE model
    (synth/src/lib.rs:316-320): every call must own its stream / workspace."""
    import threading
    m = models("medium"); _det(m)
    sents = [workload.synthetic_ids(n, utt=200 + i) for i, n in enumerate((20, 33, 9, 41, 15, 28, 37, 12))]
    expect = [m.infer_with_values(s).samples.as_slice().copy() for s in sents]
    got = [None] * len(sents)
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = m.infer_with_values(sents[i]).samples.as_slice().copy()
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(sents))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for e, g in zip(expect, got):
        assert g is not None and np.array_equal(e, g)


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_handles_may_be_freed_in_any_order(voice_paths):
    """A job hands its context back to the voice's pool and a latent belongs to a voice: both share ownership of the
    voice, so a garbage-collected caller that drops the model first does not touch freed memory (seen as
    `std::system_error: Invalid argument` from the pool mutex at interpreter exit)."""
    m = sonata_b200.from_config_path(voice_paths["medium"], device=0)
    _det(m)
    ids = workload.synthetic_ids(64, utt=3)
    job = SynthesisJob(m, [ids]); job.run()
    n = job.lengths()[1][0]
    m.close()                                     # voice handle first ...
    out = job.fetch()[0].samples.as_slice()       # ... the job still owns the weights it reads
    assert out.shape[0] == n and np.isfinite(out).all()
    job.close()
