"""The libsonata C-ABI facade (boundary #2, capi/libsonata.h:78-109): symbol set + struct layouts on CPU,
callback protocol on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from sonata_b200 import _native


class ExternError(C.Structure):
    _fields_ = [("code", C.c_int32), ("message", C.c_char_p)]


class SynthesisEvent(C.Structure):
    _fields_ = [("event_type", C.c_int32), ("error_ptr", C.POINTER(ExternError)), ("len", C.c_int64),
                ("data", C.POINTER(C.c_uint8))]


CALLBACK = C.CFUNCTYPE(C.c_uint8, SynthesisEvent)


class SynthesisParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("rate", C.c_uint8), ("volume", C.c_uint8), ("pitch", C.c_uint8),
                ("appended_silence_ms", C.c_uint32), ("callback", CALLBACK), ("nonblocking", C.c_uint8)]


class PiperSynthConfig(C.Structure):
    _fields_ = [("speaker", C.c_uint32), ("length_scale", C.c_float), ("noise_scale", C.c_float), ("noise_w", C.c_float)]


class AudioInfoC(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("num_channels", C.c_uint32), ("sample_width", C.c_uint32)]


SYMS = ["libsonataFreeString", "libsonataFreePiperSynthConfig", "libsonataFreeSynthesisEvent",
        "libsonataLoadVoiceFromConfigPath", "libsonataUnloadSonataVoice", "libsonataGetAudioInfo",
        "libsonataGetPiperDefaultSynthConfig", "libsonataSetPiperSynthConfig", "libsonataSpeak", "libsonataSpeakToFile"]


def test_facade_symbols_and_layouts(lib_built):
    for s in SYMS:                       # every function declared at capi/libsonata.h:78-109
        assert hasattr(lib_built, s), s
    assert C.sizeof(ExternError) == 16 and C.sizeof(SynthesisEvent) == 32
    assert C.sizeof(SynthesisParams) == 32 and SynthesisParams.callback.offset == 16 and SynthesisParams.nonblocking.offset == 24
    assert C.sizeof(PiperSynthConfig) == 16 and C.sizeof(AudioInfoC) == 12


def test_facade_load_error(lib_built, tmp_path):
    lib_built.libsonataLoadVoiceFromConfigPath.restype = C.c_void_p
    lib_built.libsonataLoadVoiceFromConfigPath.argtypes = [C.c_char_p, C.POINTER(ExternError)]
    err = ExternError()
    v = lib_built.libsonataLoadVoiceFromConfigPath(str(tmp_path / "nope.onnx.json").encode(), C.byref(err))
    assert not v and err.code == 17 and b"Faild to load model config" in err.message


@pytest.mark.gpu
def test_facade_speak_modes(lib_built, voice_paths, tmp_path):
    lib = lib_built
    lib.libsonataLoadVoiceFromConfigPath.restype = C.c_void_p
    lib.libsonataLoadVoiceFromConfigPath.argtypes = [C.c_char_p, C.POINTER(ExternError)]
    lib.libsonataSpeak.argtypes = [C.c_void_p, C.c_char_p, SynthesisParams, C.POINTER(ExternError)]
    lib.libsonataSpeakToFile.argtypes = [C.c_void_p, C.c_char_p, SynthesisParams, C.c_char_p, C.POINTER(ExternError)]
    lib.libsonataSpeakToFile.restype = C.c_uint8
    lib.libsonataSetPiperSynthConfig.argtypes = [C.c_void_p, PiperSynthConfig, C.POINTER(ExternError)]
    lib.libsonataGetPiperDefaultSynthConfig.restype = C.POINTER(PiperSynthConfig)
    lib.libsonataGetPiperDefaultSynthConfig.argtypes = [C.c_void_p, C.POINTER(ExternError)]
    lib.libsonataGetAudioInfo.argtypes = [C.c_void_p, C.POINTER(AudioInfoC), C.POINTER(ExternError)]
    lib.libsonataFreeSynthesisEvent.argtypes = [SynthesisEvent]
    lib.libsonataUnloadSonataVoice.argtypes = [C.c_void_p]
    err = ExternError()
    v = lib.libsonataLoadVoiceFromConfigPath(voice_paths["medium"].encode(), C.byref(err))
    assert v and err.code == 0
    ai = AudioInfoC()
    lib.libsonataGetAudioInfo(v, C.byref(ai), C.byref(err))
    assert (ai.sample_rate, ai.num_channels, ai.sample_width) == (22050, 1, 2)
    d = lib.libsonataGetPiperDefaultSynthConfig(v, C.byref(err)).contents
    assert d.speaker == 0 and abs(d.noise_scale - 0.667) < 1e-6
    # single-speaker voice: the facade always passes Some(speaker) -> unknown id is an OPERATION_ERROR
    lib.libsonataSetPiperSynthConfig(v, PiperSynthConfig(0, 1.0, 0.0, 0.0), C.byref(err))
    assert err.code == 19 and b"No speaker was found" in err.message

    events = []

    def cb(ev):
        pcm = np.ctypeslib.as_array(ev.data, shape=(max(ev.len, 1),))[:ev.len].copy()
        code = ev.error_ptr.contents.code if ev.error_ptr else 0
        events.append((ev.event_type, pcm.view("<i2"), code))
        lib.libsonataFreeSynthesisEvent(ev)
        return 0

    text = "hɛloʊ wɜːld\nðɪs ɪz ə tɛst əv ðə riːəltaɪm moʊd wɪð ə lɔŋɡɚ sɛntəns ðæt niːdz mɔːɹ ðæn wʌn tʃʌŋk".encode("utf-8")
    totals = {}
    for mode in (0, 1, 2):
        events.clear()
        p = SynthesisParams(mode, 10, 100, 50, 0, CALLBACK(cb), 0)
        lib.libsonataSpeak(v, text, p, C.byref(err))
        assert err.code == 0
        assert events[-1][0] == 1 and all(e[0] == 0 for e in events[:-1])          # SPEECH..., FINISHED
        speech = [e[1] for e in events[:-1]]
        assert all(np.abs(s).max() >= 32766 for s in speech if len(s))             # per-chunk peak normalisation (truncating cast)
        totals[mode] = sum(len(s) for s in speech)
        assert totals[mode] % 256 == 0
    assert totals[0] == totals[1] == totals[2]                                     # same frames in every mode
    # non-neutral rate -> error event, not silence
    events.clear()
    lib.libsonataSpeak(v, text, SynthesisParams(0, 50, 100, 50, 0, CALLBACK(cb), 0), C.byref(err))
    assert events and events[-1][0] == 2 and events[-1][2] == 19
    # invalid mode -> INVALID_SYNTHESIS_MODE through out_error
    lib.libsonataSpeak(v, text, SynthesisParams(7, 10, 100, 50, 0, CALLBACK(cb), 0), C.byref(err))
    assert err.code == 16
    # ---- the device post-path (volume gain, crossfade, peak normalisation, i16) is bit-identical to the host mirror of the
    # reference arithmetic (samples.rs:51-78, 144-157; synth/src/lib.rs:84-86, 106-113) applied to the f32 result
    import sonata_b200
    from sonata_b200 import PiperSynthesisConfig
    m = sonata_b200.from_config_path(voice_paths["medium"], device=0)
    m.set_fallback_synthesis_config(PiperSynthesisConfig(None, 0.0, 1.0, 0.0))      # the facade voice got the same scales above
    sents = text.decode("utf-8").split("\n")
    sil = 30 * 22050 // 1000
    events.clear()
    lib.libsonataSpeak(v, text, SynthesisParams(1, 10, 80, 50, 30, CALLBACK(cb), 0), C.byref(err))
    got = [e[1] for e in events[:-1]]
    vol = np.float32(80 / 100.0)
    for g_, a_ in zip(got, m.speak_batch(sents)):
        x = np.concatenate([a_.samples.as_slice() * vol, np.zeros(sil, dtype=np.float32)])
        assert np.array_equal(g_, sonata_b200.AudioSamples(x).to_i16_vec())
    # realtime: the reference chunk schedule (72, 3), overlap trim and crossfade(42), chunk by chunk
    m.close()
    import json
    cfgd = json.load(open(voice_paths["medium"], encoding="utf-8"))
    cfgd["streaming"] = True
    json.dump(cfgd, open(tmp_path / "rt.onnx.json", "w", encoding="utf-8"), ensure_ascii=False)
    os.symlink(voice_paths["medium"].replace(".onnx.json", ".svw"), tmp_path / "rt.svw")
    ms = sonata_b200.from_config_path(str(tmp_path / "rt.onnx.json"), device=0)
    ms.set_fallback_synthesis_config(PiperSynthesisConfig(None, 0.0, 1.0, 0.0))
    events.clear()
    lib.libsonataSpeak(v, text, SynthesisParams(2, 10, 100, 50, 0, CALLBACK(cb), 0), C.byref(err))
    got = [e[1] for e in events[:-1]]
    exp, chunk, produced = [], 72, 0
    for s in sents:
        if produced:
            chunk = chunk * 1 * produced
        chunks = list(ms.stream_synthesis(s, chunk, 3))
        exp += [c.to_i16_vec() for c in chunks]
        produced += len(chunks)
    assert len(got) == len(exp) and len(exp) > len(sents)
    for g_, e_ in zip(got, exp):
        # the 2 x 42 faded samples go through sinf (libm in the library, numpy's float32 sin in the mirror): allow one
        # LSB there, everything else is bit-identical
        assert g_.shape == e_.shape
        d = np.abs(g_.astype(np.int32) - e_.astype(np.int32))
        assert d.max() <= 1 and int((d != 0).sum()) <= 84 and not d[42:-42].any()
    ms.close()
    # ---- a non-blocking speak keeps the voice alive after unload (the reference clones the Arc, capi/src/lib.rs:314,375)
    import threading
    v2 = lib.libsonataLoadVoiceFromConfigPath(voice_paths["medium"].encode(), C.byref(err))
    done, seen = threading.Event(), []

    def cb2(ev):
        seen.append(ev.event_type)
        lib.libsonataFreeSynthesisEvent(ev)
        if ev.event_type != 0:
            done.set()
        return 0
    cb2_c = CALLBACK(cb2)
    lib.libsonataSpeak(v2, text, SynthesisParams(0, 10, 100, 50, 0, cb2_c, 1), C.byref(err))
    lib.libsonataUnloadSonataVoice(v2)                                             # while the worker thread is synthesising
    assert done.wait(60) and seen[-1] == 1 and seen.count(0) == len(sents)
    out = tmp_path / "o.wav"
    ok = lib.libsonataSpeakToFile(v, text, SynthesisParams(1, 10, 100, 50, 50, CALLBACK(cb), 0), str(out).encode(), C.byref(err))
    assert ok == 1 and os.path.getsize(out) == 44 + 2 * (totals[1] + 2 * (50 * 22050 // 1000))
    lib.libsonataUnloadSonataVoice(v)
