"""Host-side mirror of `sonata-core` + the `audio-ops` output types the hot path returns.

Reference: crates/sonata/core/src/lib.rs (SonataError :19-24, Phonemes :53-79, trait SonataModel
:82-131) and crates/audio/ops/src/samples.rs (AudioInfo :9-14, AudioSamples :16-18, to_i16_vec
:51-75, as_wave_bytes :76-78, crossfade :144-157, Audio :208-271).  Same names, argument meaning
and error behaviour, so the parity tests read like the reference's own.
"""
from __future__ import annotations

import math
import struct
import wave
from dataclasses import dataclass
from typing import Optional

import numpy as np


class SonataError(Exception):
    """enum SonataError (core/src/lib.rs:19-24); FFI codes from capi/libsonata.h:10-14."""
    code = 19

    @staticmethod
    def from_code(code: int, message: str) -> "SonataError":
        cls = {17: FailedToLoadResource, 18: PhonemizationError, 19: OperationError}.get(code, OperationError)
        e = cls(message)
        e.code = code
        return e


class FailedToLoadResource(SonataError):
    code = 17

    def __str__(self):  # Display impl, core/src/lib.rs:35-37
        return f"Failed to load resource from. Error `{self.args[0]}`"


class PhonemizationError(SonataError):
    code = 18


class OperationError(SonataError):
    code = 19


class Phonemes:
    """struct Phonemes(Vec<String>) (core/src/lib.rs:53-79)."""

    def __init__(self, sentences):
        self._s = list(sentences)

    def sentences(self):
        return self._s

    def to_vec(self):
        return list(self._s)

    def num_sentences(self):
        return len(self._s)

    def __str__(self):
        return " ".join(self._s)


@dataclass
class AudioInfo:
    sample_rate: int
    num_channels: int = 1
    sample_width: int = 2


_F32_EPS = float(np.finfo(np.float32).eps)


class AudioSamples:
    """struct AudioSamples(Vec<f32>) (audio/ops/src/samples.rs:16-18)."""

    def __init__(self, samples=()):
        self._v = np.asarray(samples, dtype=np.float32).reshape(-1).copy()

    @classmethod
    def _wrap(cls, array: np.ndarray) -> "AudioSamples":
        """Adopt `array` without copying (used for the library's pinned result buffers)."""
        self = cls.__new__(cls)
        self._v = array
        return self

    def as_slice(self) -> np.ndarray:
        return self._v

    def into_vec(self) -> np.ndarray:
        return self._v

    def __len__(self):
        return int(self._v.shape[0])

    def is_empty(self) -> bool:
        return len(self) == 0

    def to_i16_vec(self) -> np.ndarray:
        """samples.rs:51-75: per-buffer peak normalisation to +-32767, clamp, truncating cast."""
        if self.is_empty():
            return np.zeros(0, dtype=np.int16)
        v = self._v
        abs_max = np.float32(max(abs(float(v.max())), abs(float(v.min())), _F32_EPS))
        scale = np.float32(32767.0) / abs_max
        y = np.clip(v * scale, np.float32(-32768.0), np.float32(32767.0))
        return np.trunc(y).astype(np.int16)

    def as_wave_bytes(self) -> bytes:
        return self.to_i16_vec().astype("<i2").tobytes()

    def merge(self, other: "AudioSamples") -> None:
        self._v = np.concatenate([self._v, other._v])

    def _own(self) -> None:
        if not self._v.flags.writeable or self._v.base is not None:
            self._v = self._v.copy()

    def crossfade(self, fade_samples: int) -> None:
        """samples.rs:144-157: quarter-sine fade on both ends, f(i) = sin(i/(n-1) * pi/2)."""
        length = len(self)
        n = min(fade_samples, length // 2)
        if n <= 0:
            return
        self._own()
        att = np.float32(n - 1)
        i = np.arange(n, dtype=np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            f = np.sin((i / att) * np.float32(math.pi) / np.float32(2.0)).astype(np.float32)
        self._v[:n] *= f
        self._v[length - 1 - np.arange(n)] *= f


class Audio:
    """struct Audio {samples, info, inference_ms} (audio/ops/src/samples.rs:208-271)."""

    def __init__(self, samples, sample_rate: int, inference_ms: Optional[float] = None):
        self.samples = samples if isinstance(samples, AudioSamples) else AudioSamples(samples)
        self.info = AudioInfo(sample_rate=sample_rate, num_channels=1, sample_width=2)
        self.inference_ms = inference_ms

    def into_vec(self):
        return self.samples.into_vec()

    def as_wave_bytes(self) -> bytes:
        return self.samples.as_wave_bytes()

    def __len__(self):
        return len(self.samples)

    def is_empty(self):
        return self.samples.is_empty()

    def duration_ms(self) -> float:
        return (len(self) / self.info.sample_rate) * 1000.0

    def real_time_factor(self) -> Optional[float]:
        """inference_ms / duration_ms — lower is better (samples.rs:253-260)."""
        if self.inference_ms is None:
            return None
        d = self.duration_ms()
        return 0.0 if d == 0.0 else self.inference_ms / d

    def save_to_file(self, filename) -> None:
        with wave.open(str(filename), "wb") as w:
            w.setnchannels(self.info.num_channels)
            w.setsampwidth(self.info.sample_width)
            w.setframerate(self.info.sample_rate)
            w.writeframes(self.as_wave_bytes())
