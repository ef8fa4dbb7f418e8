"""Host-side mirror of `sonata-piper` (crates/sonata/models/piper/src/lib.rs) over libsonata_b200.

`from_config_path` / `VitsModel` / `VitsStreamingModel` / `PiperSynthesisConfig` keep the reference's
names and semantics; the arithmetic behind `speak_*` is the CUDA library, never Python.  The integer
host logic that the reference keeps in Rust around `session.run` (the streaming chunk scheduler,
crossfade, one-shot rule) is restated here because it lives on the host side of the FFI boundary
in the reference too.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence

import numpy as np

from . import _native as N
from .core import Audio, AudioInfo, AudioSamples, OperationError, Phonemes, PhonemizationError, SonataError

MIN_CHUNK_SIZE = 44      # piper/src/lib.rs:18
MAX_CHUNK_SIZE = 1024    # piper/src/lib.rs:19
HOP = 256                # piper/src/lib.rs:910


@dataclass
class PiperSynthesisConfig:
    """piper/src/lib.rs:160-166"""
    speaker: Optional[int] = None
    noise_scale: float = 0.667
    length_scale: float = 1.0
    noise_w: float = 0.8


def _check(rc: int, err: N.sb200_error):
    if rc != 0:
        msg = ""
        if err.message:
            msg = C.string_at(err.message).decode("utf-8", "replace")
            N.lib().sb200_string_free(err.message)
        raise SonataError.from_code(err.code if err.code else rc, msg)


class _PinnedOwner:
    """Keeps one library-owned (pinned) result buffer alive for as long as a numpy view of it exists."""

    def __init__(self, a: N.sb200_audio):
        self.a = N.sb200_audio(a.data, a.len, a.inference_ms, a.sample_rate)

    def __del__(self):
        try:
            N.lib().sb200_audio_free(C.byref(self.a))
        except Exception:
            pass


def _take_audio(a: N.sb200_audio) -> Audio:
    """Zero-copy: the returned samples are a read-only view of the library's pinned host buffer (the
    reference copies `outputs[0]` into a Vec at piper/src/lib.rs:392)."""
    if not a.len:
        N.lib().sb200_audio_free(C.byref(a))
        return Audio(AudioSamples(np.zeros(0, np.float32)), int(a.sample_rate), float(a.inference_ms))
    owner = _PinnedOwner(a)
    buf = (C.c_float * a.len).from_address(C.addressof(a.data.contents))
    buf._owner = owner
    arr = np.frombuffer(buf, dtype=np.float32)
    arr.flags.writeable = False
    return Audio(AudioSamples._wrap(arr), int(a.sample_rate), float(a.inference_ms))


class AdaptiveMelChunker:
    """piper/src/lib.rs:860-913 — yields ((mel_start, mel_end|None), (audio_start, audio_end|None))."""

    def __init__(self, num_frames: int, chunk_size: int, chunk_padding: int):
        self.num_frames = num_frames
        self.chunk_size = chunk_size
        self.chunk_padding = chunk_padding
        self.last_end_index: Optional[int] = 0
        self.step = 1

    def consume(self):
        self.last_end_index = None

    def __iter__(self):
        return self

    def __next__(self):
        last_index = self.last_end_index
        if last_index is None:
            raise StopIteration
        chunk_size = min(self.chunk_size * self.step, MAX_CHUNK_SIZE)
        if last_index == 0:
            start_index, start_padding = 0, 0
        else:
            start_index = last_index - self.chunk_padding * 2
            start_padding = self.chunk_padding
        chunk_end = last_index + chunk_size + self.chunk_padding
        remaining = self.num_frames - chunk_end
        if remaining <= MIN_CHUNK_SIZE:
            end_index, end_padding = None, None
        else:
            end_index, end_padding = chunk_end, -self.chunk_padding
        self.step += 1
        self.last_end_index = end_index
        return ((start_index, end_index),
                (start_padding * HOP, None if end_padding is None else end_padding * HOP))


class _VitsCommons:
    def __init__(self, config_path: str, device: int = 0):
        lib = N.lib()
        self._lib = lib
        self._h = C.c_void_p()
        err = N.sb200_error()
        _check(lib.sb200_voice_load(str(config_path).encode("utf-8"), device, C.byref(self._h), C.byref(err)), err)
        self.config_path = str(config_path)
        self.device = device
        self._speakers = None

    def get_speakers(self) -> Optional[dict]:
        """SonataModel::get_speakers (piper/src/lib.rs:463-465): {speaker id: name} of a multi-speaker voice, read from
        the `speaker_id_map` of the voice config (None for single-speaker voices)."""
        if self._speakers is None:
            import json
            with open(self.config_path, encoding="utf-8") as f:
                m = json.load(f).get("speaker_id_map") or {}
            self._speakers = {int(v): k for k, v in m.items()}
        return self._speakers or None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.sb200_voice_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- trait SonataModel (core/src/lib.rs:82-131) ----
    def audio_output_info(self) -> AudioInfo:
        ai, err = N.sb200_audio_info(), N.sb200_error()
        _check(self._lib.sb200_audio_output_info(self._h, C.byref(ai), C.byref(err)), err)
        return AudioInfo(int(ai.sample_rate), int(ai.num_channels), int(ai.sample_width))

    def phonemize_text(self, text: str) -> Phonemes:
        # espeak-ng front-end is outside the hot path; callers pass phonemes.
        raise PhonemizationError("Failed to phonemize given text using espeak-ng. Error: "
                                 "the espeak-ng front-end is not part of sonata_b200; pass phonemes")

    def phonemes_to_input_ids(self, phonemes: str) -> List[int]:
        ids = C.POINTER(C.c_int64)()
        n = C.c_size_t()
        err = N.sb200_error()
        _check(self._lib.sb200_phonemes_to_input_ids(self._h, phonemes.encode("utf-8"), C.byref(ids), C.byref(n),
                                                     C.byref(err)), err)
        out = [int(ids[i]) for i in range(n.value)]
        self._lib.sb200_ids_free(ids)
        return out

    def speak_one_sentence(self, phonemes: str) -> Audio:
        a, err = N.sb200_audio(), N.sb200_error()
        _check(self._lib.sb200_speak_one_sentence(self._h, phonemes.encode("utf-8"), C.byref(a), C.byref(err)), err)
        return _take_audio(a)

    def speak_batch(self, phoneme_batches: Sequence[str]) -> List[Audio]:
        n = len(phoneme_batches)
        if n == 0:
            return []
        arr = (C.c_char_p * n)(*[p.encode("utf-8") for p in phoneme_batches])
        outs = (N.sb200_audio * n)()
        err = N.sb200_error()
        _check(self._lib.sb200_speak_batch(self._h, arr, n, outs, C.byref(err)), err)
        return [_take_audio(outs[i]) for i in range(n)]

    def infer_with_values(self, input_phonemes: Sequence[int]) -> Audio:
        """VitsModel::infer_with_values (piper/src/lib.rs:342-399)."""
        ids = np.ascontiguousarray(input_phonemes, dtype=np.int64)
        a, err = N.sb200_audio(), N.sb200_error()
        _check(self._lib.sb200_speak_ids(self._h, ids.ctypes.data_as(C.POINTER(C.c_int64)), ids.size, C.byref(a),
                                         C.byref(err)), err)
        return _take_audio(a)

    def infer_batch_with_values(self, batches: Sequence[Sequence[int]]) -> List[Audio]:
        n = len(batches)
        packed = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.int64) for b in batches]))
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in batches])
        outs = (N.sb200_audio * n)()
        err = N.sb200_error()
        _check(self._lib.sb200_speak_batch_ids(self._h, packed.ctypes.data_as(C.POINTER(C.c_int64)),
                                               offs.ctypes.data_as(C.POINTER(C.c_size_t)), n, outs, C.byref(err)), err)
        return [_take_audio(outs[i]) for i in range(n)]

    def _cfg(self, fn) -> PiperSynthesisConfig:
        c, err = N.sb200_synth_config(), N.sb200_error()
        _check(fn(self._h, C.byref(c), C.byref(err)), err)
        return PiperSynthesisConfig(int(c.speaker) if c.has_speaker else None, float(c.noise_scale),
                                    float(c.length_scale), float(c.noise_w))

    def get_default_synthesis_config(self) -> PiperSynthesisConfig:
        return self._cfg(self._lib.sb200_get_default_synthesis_config)

    def get_fallback_synthesis_config(self) -> PiperSynthesisConfig:
        return self._cfg(self._lib.sb200_get_fallback_synthesis_config)

    def set_fallback_synthesis_config(self, synthesis_config) -> None:
        if not isinstance(synthesis_config, PiperSynthesisConfig):
            raise OperationError("Invalid configuration for Vits Model")
        c = N.sb200_synth_config(synthesis_config.speaker or 0, 0 if synthesis_config.speaker is None else 1,
                                 synthesis_config.noise_scale, synthesis_config.length_scale, synthesis_config.noise_w)
        err = N.sb200_error()
        _check(self._lib.sb200_set_fallback_synthesis_config(self._h, C.byref(c), C.byref(err)), err)

    def _str(self, fn) -> str:
        p, err = C.c_void_p(), N.sb200_error()
        _check(fn(self._h, C.byref(p), C.byref(err)), err)
        s = C.string_at(p).decode("utf-8")
        self._lib.sb200_string_free(p)
        return s

    def get_language(self) -> Optional[str]:
        return self._str(self._lib.sb200_get_language)

    def properties(self) -> dict:
        return {"quality": self._str(self._lib.sb200_get_quality)}

    def speaker_name_to_id(self, name: str) -> Optional[int]:
        r = int(self._lib.sb200_speaker_name_to_id(self._h, name.encode("utf-8")))
        return None if r < 0 else r

    def supports_streaming_output(self) -> bool:
        return False

    def stream_synthesis(self, phonemes: str, chunk_size: int, chunk_padding: int):
        raise OperationError("Streaming synthesis is not supported for this model")

    def set_backend(self, backend: int) -> int:
        return int(self._lib.sb200_set_backend(self._h, backend))


class VitsModel(_VitsCommons):
    """piper/src/lib.rs:291-478"""


class EncoderOutputs:
    """piper/src/lib.rs:671-763 — `z` stays on the device."""

    def __init__(self, model: "_VitsCommons", handle: C.c_void_p):
        self._m, self._h = model, handle
        self.num_frames = int(model._lib.sb200_latent_frames(handle))

    def infer_decoder(self, lo: int = 0, hi: Optional[int] = None) -> AudioSamples:
        hi = self.num_frames if hi is None else hi
        a, err = N.sb200_audio(), N.sb200_error()
        _check(self._m._lib.sb200_decode_chunk(self._m._h, self._h, lo, hi, C.byref(a), C.byref(err)), err)
        return _take_audio(a).samples

    def __del__(self):
        try:
            if self._h:
                self._m._lib.sb200_latent_free(self._h)
                self._h = None
        except Exception:
            pass


class SpeechStreamer:
    """piper/src/lib.rs:765-858: chunked decoder runs with overlap trimming + crossfade(42)."""

    def __init__(self, enc: EncoderOutputs, chunk_size: int, chunk_padding: int):
        self.enc = enc
        self.chunker = AdaptiveMelChunker(enc.num_frames, chunk_size, chunk_padding)
        self.one_shot = enc.num_frames <= (chunk_size * 2 + chunk_padding * 2)

    def __iter__(self) -> Iterator[AudioSamples]:
        return self

    def __next__(self) -> AudioSamples:
        (m0, m1), (a0, a1) = next(self.chunker)
        if self.one_shot:
            self.chunker.consume()
            return self.enc.infer_decoder()
        hi = self.enc.num_frames if m1 is None else m1
        audio = self.enc.infer_decoder(m0, hi).as_slice()
        audio = audio[a0:a1] if a1 is not None else audio[a0:]
        out = AudioSamples(audio)
        out.crossfade(42)
        return out


class VitsStreamingModel(_VitsCommons):
    """piper/src/lib.rs:480-669"""

    def infer_encoder(self, input_phonemes: Sequence[int]) -> EncoderOutputs:
        ids = np.ascontiguousarray(input_phonemes, dtype=np.int64)
        h, err = C.c_void_p(), N.sb200_error()
        _check(self._lib.sb200_encode_ids(self._h, ids.ctypes.data_as(C.POINTER(C.c_int64)), ids.size, C.byref(h),
                                          C.byref(err)), err)
        return EncoderOutputs(self, h)

    def supports_streaming_output(self) -> bool:
        return True

    def stream_synthesis(self, phonemes: str, chunk_size: int, chunk_padding: int) -> SpeechStreamer:
        ids = self.phonemes_to_input_ids(phonemes)
        return SpeechStreamer(self.infer_encoder(ids), chunk_size, chunk_padding)


def from_config_path(config_path, device: int = 0):
    """sonata_piper::from_config_path (piper/src/lib.rs:88-110): `streaming: true` selects the
    two-stage model."""
    import json
    try:
        with open(config_path, "r", encoding="utf-8") as f:
            streaming = bool(json.load(f).get("streaming") or False)
    except OSError:
        streaming = False   # let the library produce the reference's FailedToLoadResource error
    except ValueError:
        streaming = False
    return (VitsStreamingModel if streaming else VitsModel)(config_path, device)
