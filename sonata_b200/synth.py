"""Host-side mirror of `sonata-synth` (crates/sonata/synth/src/lib.rs): the callers of the hot path.

`SonataSpeechSynthesizer` keeps the reference's three scheduling modes:

* `synthesize_lazy`      — one sentence per `next()`                                  (synth :297-307)
* `synthesize_parallel`  — the reference fans sentences out over rayon and collects     (synth :314-325);
                           here the fan-out IS the batch: all sentences go through ONE
                           `speak_batch` pass (the packed-segment kernels), same results.
* `synthesize_streamed`  — realtime mode: per sentence `stream_synthesis(chunk, pad)`
                           with the reference's chunk-size growth rule                  (synth :337-382)

`AudioOutputConfig` (rate / volume / pitch through Sonic, appended silence) is CPU post-processing after
the path (out of scope): appended silence and volume are honoured (pure sample
arithmetic); a non-neutral rate or pitch raises OperationError instead of silently being ignored.
The model argument is anything with the `SonataModel` surface (`sonata_b200.VitsModel`, or a fake in
the CPU tests) — like the reference's `Arc<dyn SonataModel + Send + Sync>`.
"""
from __future__ import annotations

import queue
import threading
from dataclasses import dataclass
from typing import Iterator, List, Optional

import numpy as np

from .core import Audio, AudioSamples, OperationError, SonataError

RATE_RANGE = (0.5, 5.5)      # synth/src/lib.rs:13
VOLUME_RANGE = (0.0, 1.0)    # :14
PITCH_RANGE = (0.5, 1.5)     # :15


def percent_to_param(value: int, lo: float, hi: float) -> float:
    """synth/src/utils.rs:6-8 — linear over the range (rate=50 means 3.0x, not 1.0x)."""
    return (value / 100.0) * (hi - lo) + lo


def param_to_percent(value: float, lo: float, hi: float) -> int:
    return int(round((value - lo) / (hi - lo) * 100.0))


@dataclass
class AudioOutputConfig:
    """synth/src/lib.rs:28-34"""
    rate: Optional[int] = None
    volume: Optional[int] = None
    pitch: Optional[int] = None
    appended_silence_ms: Optional[int] = None

    def _check_supported(self):
        if self.rate is not None and abs(percent_to_param(self.rate, *RATE_RANGE) - 1.0) > 1e-6:
            raise OperationError("Sonic Error: time-scale modification (rate) is CPU post-processing outside "
                                 "sonata_b200; use length_scale or rate=10 (1.0x)")
        if self.pitch is not None and abs(percent_to_param(self.pitch, *PITCH_RANGE) - 1.0) > 1e-6:
            raise OperationError("Sonic Error: pitch modification is CPU post-processing outside sonata_b200; "
                                 "use pitch=50 (1.0x)")

    def apply_to_raw_samples(self, samples: AudioSamples) -> AudioSamples:
        self._check_supported()
        v = samples.as_slice()
        if self.volume is not None:
            v = v * np.float32(percent_to_param(self.volume, *VOLUME_RANGE))
        return AudioSamples(v)

    def generate_silence(self, time_ms: int, sample_rate: int) -> AudioSamples:
        return AudioSamples(np.zeros((time_ms * sample_rate) // 1000, dtype=np.float32))   # :107-116

    def apply(self, audio: Audio) -> Audio:
        """synth/src/lib.rs:37-54: silence is appended first, then the whole buffer is processed."""
        s = audio.samples
        if self.appended_silence_ms is not None:
            s = AudioSamples(np.concatenate([s.as_slice(), self.generate_silence(self.appended_silence_ms,
                                                                                  audio.info.sample_rate).as_slice()]))
        return Audio(self.apply_to_raw_samples(s), audio.info.sample_rate, audio.inference_ms)


class SonataSpeechSynthesizer:
    """synth/src/lib.rs:119-203.  `text` is a phoneme string; sentences are separated by newlines when the
    model has no phonemizer (the espeak-ng front-end is outside this repo)."""

    def __init__(self, model):
        self.model = model

    # -- SpeechSynthesisTaskProvider::get_phonemes (:256-258)
    def _phonemes(self, text: str) -> List[str]:
        try:
            return self.model.phonemize_text(text).to_vec()
        except SonataError:
            return [s for s in text.split("\n") if s.strip()]

    def _process(self, audio: Audio, cfg: Optional[AudioOutputConfig]) -> Audio:
        return cfg.apply(audio) if cfg is not None else audio

    def synthesize_lazy(self, text: str, output_config: Optional[AudioOutputConfig] = None) -> Iterator[Audio]:
        for ph in self._phonemes(text):
            yield self._process(self.model.speak_one_sentence(ph), output_config)

    def synthesize_parallel(self, text: str, output_config: Optional[AudioOutputConfig] = None) -> Iterator[Audio]:
        ph = self._phonemes(text)
        results = self.model.speak_batch(ph) if ph else []        # one batched pass == the rayon fan-out + collect
        return iter([self._process(a, output_config) for a in results])

    def synthesize_streamed(self, text: str, output_config: Optional[AudioOutputConfig] = None,
                            chunk_size: int = 72, chunk_padding: int = 3) -> Iterator[AudioSamples]:
        """RealtimeSpeechStream (:337-382): a background producer pushes chunks into an unbounded channel;
        chunk_size is multiplied by the number of chunks already produced for every following sentence."""
        sr = self.model.audio_output_info().sample_rate
        q: "queue.Queue" = queue.Queue()
        done = object()

        def producer():
            cs, produced = chunk_size, 0
            try:
                for ph in self._phonemes(text):
                    if produced != 0:
                        cs = cs * 1 * produced                      # chunk_factor = 1 (:348-356)
                    n = 0
                    for chunk in self.model.stream_synthesis(ph, cs, chunk_padding):
                        q.put(output_config.apply_to_raw_samples(chunk) if output_config else chunk)
                        n += 1
                    produced += n
                    if output_config and output_config.appended_silence_ms:
                        q.put(output_config.generate_silence(output_config.appended_silence_ms, sr))
            except Exception as e:                                  # errors travel through the channel (:368-371)
                q.put(e)
            q.put(done)

        threading.Thread(target=producer, daemon=True).start()
        while True:
            item = q.get()
            if item is done:
                return
            if isinstance(item, Exception):
                raise item
            yield item

    def synthesize_to_file(self, filename, text: str, output_config: Optional[AudioOutputConfig] = None) -> None:
        """:168-198 — parallel mode, concatenated, peak-normalised i16 WAV."""
        parts = [a.samples.as_slice() for a in self.synthesize_parallel(text, output_config)]
        if not parts or sum(len(p) for p in parts) == 0:
            raise OperationError("No speech data to write")
        Audio(AudioSamples(np.concatenate(parts)), self.model.audio_output_info().sample_rate).save_to_file(filename)

    # passthroughs of the SonataModel surface (:205-253)
    def speak_one_sentence(self, phonemes: str) -> Audio:
        return self.model.speak_one_sentence(phonemes)

    def speak_batch(self, phoneme_batches) -> List[Audio]:
        return self.model.speak_batch(phoneme_batches)

    def audio_output_info(self):
        return self.model.audio_output_info()
