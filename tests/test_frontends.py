"""Frontends: the `pysonata`-shaped module and the CLI / JSON-lines protocol.
CPU tests drive them with a fake SonataModel (the schedulers and the protocol are host logic); the GPU test runs the
same calls on the synthetic voice."""
import io
import json
import os
import struct

import numpy as np
import pytest

from sonata_b200 import Audio, AudioInfo, AudioSamples, PhonemizationError, PiperSynthesisConfig, SonataError
from sonata_b200 import cli, pysonata
from sonata_b200.synth import SonataSpeechSynthesizer


class FakeModel:
    """4 samples per character; remembers the last synthesis config"""

    def __init__(self):
        self.cfg = PiperSynthesisConfig(None, 0.667, 1.0, 0.8)
        self.stream_args = []

    def audio_output_info(self):
        return AudioInfo(22050, 1, 2)

    def phonemize_text(self, text):
        raise PhonemizationError("no espeak here")

    def _wave(self, ph):
        return (np.arange(4 * len(ph), dtype=np.float32) % 7 - 3) / 4

    def speak_one_sentence(self, ph):
        return Audio(self._wave(ph), 22050, 1.0)

    def speak_batch(self, phs):
        return [Audio(self._wave(p), 22050, 1.0) for p in phs]

    def stream_synthesis(self, ph, chunk_size, chunk_padding):
        self.stream_args.append((chunk_size, chunk_padding))
        w = self._wave(ph)
        return iter([AudioSamples(w[:len(w) // 2]), AudioSamples(w[len(w) // 2:])])

    def get_default_synthesis_config(self):
        return PiperSynthesisConfig(0, 0.667, 1.0, 0.8)

    def get_fallback_synthesis_config(self):
        return self.cfg

    def set_fallback_synthesis_config(self, c):
        self.cfg = c

    def get_language(self):
        return "en-us"

    def get_speakers(self):
        return {0: "a", 1: "b"}

    def close(self):
        pass


def test_pysonata_surface_over_fake_model():
    s = pysonata.Sonata(SonataSpeechSynthesizer(FakeModel()))
    text = "abc\nde"
    lazy = list(s.synthesize(text))
    assert [type(w) for w in lazy] == [pysonata.WaveSamples] * 2
    assert len(lazy[0].get_wave_bytes()) == 2 * 12 and lazy[0].sample_rate == 22050 and lazy[0].sample_width == 2
    assert lazy[0].duration_ms == pytest.approx(12 / 22.05) and lazy[0].real_time_factor is not None
    par = list(s.synthesize_parallel(text, pysonata.AudioOutputConfig(volume=50, appended_silence_ms=10)))
    assert len(par[1].get_wave_bytes()) == 2 * (8 + 220)
    rt = list(s.synthesize_streamed(text))
    assert all(isinstance(c, bytes) for c in rt) and sum(len(c) for c in rt) == 2 * 20
    assert s._s.model.stream_args[0] == (45, 3)                       # pysonata defaults (python/src/lib.rs:379-380)
    assert s.language == "en-us" and s.speakers == {0: "a", 1: "b"}
    ai = s.get_audio_output_info()
    assert (ai.sample_rate, ai.num_channels, ai.sample_width) == (22050, 1, 2)
    with pytest.raises(pysonata.SonataException):
        pysonata.phonemize_text("hello", "en-us")
    with pytest.raises(SonataError):
        list(s.synthesize("abc", pysonata.AudioOutputConfig(rate=50)))  # rate 50 % = 3.0x needs Sonic


def test_cli_json_lines_protocol(tmp_path, monkeypatch):
    fake = FakeModel()
    monkeypatch.setattr(cli, "from_config_path", lambda path, device=0: fake)
    # one JSON request per stdin line -> raw i16 PCM on stdout (main.rs:100-165, 257-260)
    reqs = [{"text": "abc\nde", "mode": "parallel", "volume": 100, "noise_w": 0.1},
            {"text": "xy", "mode": "realtime", "chunk_size": 50},
            {"text": "q", "length_scale": 1.5, "speaker_id": 1}]
    stdin = io.StringIO("\n".join(json.dumps(r) for r in reqs) + "\n")
    out = io.BytesIO()

    class Out:
        buffer = out
    monkeypatch.setattr("sys.stdin", stdin)
    monkeypatch.setattr("sys.stdout", Out())
    assert cli.main(["voice.onnx.json"]) == 0
    assert len(out.getvalue()) == 2 * 4 * (3 + 2 + 2 + 1)
    assert fake.stream_args == [(50, 3)]
    # the last request's config: unspecified scales fall back to the model defaults, speaker id is taken
    assert (fake.cfg.speaker, fake.cfg.length_scale, fake.cfg.noise_scale, fake.cfg.noise_w) == (1, 1.5, 0.667, 0.8)
    # -o with stdin requests: enumerated WAV files (main.rs:243-256)
    monkeypatch.setattr("sys.stdin", io.StringIO(json.dumps({"text": "abc"}) + "\n" + json.dumps({"text": "de\nfgh"}) + "\n"))
    assert cli.main(["voice.onnx.json", "-o", str(tmp_path / "out.wav")]) == 0
    for name, n in (("out-1.wav", 12), ("out-2.wav", 20)):
        raw = open(tmp_path / name, "rb").read()
        assert raw[:4] == b"RIFF" and struct.unpack("<I", raw[40:44])[0] == 2 * n
    # -f: the flags form the request
    (tmp_path / "in.txt").write_text("abcd\n", encoding="utf-8")
    assert cli.main(["voice.onnx.json", "-f", str(tmp_path / "in.txt"), "-o", str(tmp_path / "f.wav"), "--volume", "80"]) == 0
    assert os.path.getsize(tmp_path / "f.wav") == 44 + 2 * 16


@pytest.mark.gpu
def test_frontends_on_the_device(voice_paths, tmp_path):
    m = pysonata.PiperModel(voice_paths["medium"], device=0)
    assert m.speaker is None
    m.set_scales(1.0, 0.0, 0.0)
    sc = m.get_scales()
    assert (sc.length_scale, sc.noise_scale, sc.noise_w) == (1.0, 0.0, 0.0)
    s = pysonata.Sonata.with_piper(m)
    text = "hɛloʊ wɜːld\nðɪs ɪz ə tɛst"
    lazy = [w.get_wave_bytes() for w in s.synthesize(text)]
    par = [w.get_wave_bytes() for w in s.synthesize_parallel(text)]
    assert lazy == par and all(len(b) % 512 == 0 and len(b) > 0 for b in lazy)     # batched == sequential, deterministic scales
    s.synthesize_to_file(str(tmp_path / "o.wav"), text)
    assert os.path.getsize(tmp_path / "o.wav") == 44 + sum(len(b) for b in par)
    out = io.BytesIO()
    cli.process_request(s._s, s._s.model.get_default_synthesis_config(),
                        {"text": text, "mode": "parallel", "noise_scale": 0.0, "noise_w": 0.0}, None, out=out)
    assert out.getvalue() == b"".join(par)
    assert s.speakers is None and s.language
