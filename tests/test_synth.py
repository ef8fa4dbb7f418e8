"""CPU: the sonata-synth mirror (schedulers around the model) with a fake SonataModel — the reference has
no mock backend but `SonataModel` is a trait, so a fake is the natural unit test."""
import numpy as np
import pytest

from sonata_b200 import Audio, AudioInfo, AudioSamples, OperationError, PhonemizationError
from sonata_b200.synth import (AudioOutputConfig, SonataSpeechSynthesizer, param_to_percent, percent_to_param,
                               RATE_RANGE, PITCH_RANGE, VOLUME_RANGE)


class FakeModel:
    """4 samples per character, value = index of the sentence call."""

    def __init__(self):
        self.batch_calls, self.single_calls, self.stream_args = 0, 0, []

    def audio_output_info(self):
        return AudioInfo(22050, 1, 2)

    def phonemize_text(self, text):
        raise PhonemizationError("no espeak here")

    def _wave(self, ph):
        return np.full(4 * len(ph), 0.5, np.float32)

    def speak_one_sentence(self, ph):
        self.single_calls += 1
        return Audio(self._wave(ph), 22050, 1.0)

    def speak_batch(self, phs):
        self.batch_calls += 1
        return [Audio(self._wave(p), 22050, 1.0) for p in phs]

    def stream_synthesis(self, ph, chunk_size, chunk_padding):
        self.stream_args.append((chunk_size, chunk_padding))
        w = self._wave(ph)
        return iter([AudioSamples(w[:len(w) // 2]), AudioSamples(w[len(w) // 2:])])


def test_percent_param_mapping():
    # synth/src/lib.rs:13-15 + utils.rs:6-8: rate 50 % is 3.0x, 1.0x is rate 10; pitch 50 -> 1.0; volume 50 -> 0.5
    assert percent_to_param(50, *RATE_RANGE) == pytest.approx(3.0)
    assert percent_to_param(10, *RATE_RANGE) == pytest.approx(1.0)
    assert percent_to_param(50, *PITCH_RANGE) == pytest.approx(1.0)
    assert percent_to_param(50, *VOLUME_RANGE) == pytest.approx(0.5)
    assert param_to_percent(1.0, *RATE_RANGE) == 10


def test_modes_and_batching():
    m = FakeModel()
    s = SonataSpeechSynthesizer(m)
    text = "abc\n\nde\nf"
    lazy = list(s.synthesize_lazy(text))
    assert [len(a) for a in lazy] == [12, 8, 4] and m.single_calls == 3
    par = list(s.synthesize_parallel(text))
    assert [len(a) for a in par] == [12, 8, 4] and m.batch_calls == 1       # ONE batched pass for all sentences
    chunks = list(s.synthesize_streamed(text, None, 72, 3))
    assert sum(len(c) for c in chunks) == 24
    # chunk growth: second sentence uses chunk_size * chunks_so_far (2), third * 4 of that (synth :348-356)
    assert m.stream_args == [(72, 3), (144, 3), (576, 3)]


def test_output_config():
    m = FakeModel()
    s = SonataSpeechSynthesizer(m)
    cfg = AudioOutputConfig(rate=10, volume=50, pitch=50, appended_silence_ms=100)
    a = next(s.synthesize_lazy("abcd", cfg))
    assert len(a) == 16 + 2205
    assert np.allclose(a.samples.as_slice()[:16], 0.25) and np.all(a.samples.as_slice()[16:] == 0)
    with pytest.raises(OperationError, match="Sonic"):
        next(s.synthesize_lazy("abcd", AudioOutputConfig(rate=50, volume=50, pitch=50)))
    with pytest.raises(OperationError, match="Sonic"):
        next(s.synthesize_lazy("abcd", AudioOutputConfig(rate=10, volume=50, pitch=70)))


def test_to_file(tmp_path):
    import wave
    s = SonataSpeechSynthesizer(FakeModel())
    p = tmp_path / "o.wav"
    s.synthesize_to_file(p, "abc\nde")
    with wave.open(str(p)) as w:
        assert w.getframerate() == 22050 and w.getnchannels() == 1 and w.getsampwidth() == 2 and w.getnframes() == 20
    with pytest.raises(OperationError, match="No speech data"):
        s.synthesize_to_file(p, "\n\n")
