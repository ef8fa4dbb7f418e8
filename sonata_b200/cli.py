"""`sonata` command-line frontend: the argument set and the JSON-lines protocol of the reference
CLI (crates/frontends/cli/src/main.rs:32-260) over the B200 engine.

    python -m sonata_b200.cli voice.onnx.json -f phonemes.txt -o out.wav --mode parallel
    echo '{"text": "hɛloʊ", "mode": "realtime", "chunk_size": 100}' | python -m sonata_b200.cli voice.onnx.json > pcm.raw

Without `-f`, one JSON request per stdin line (fields of `SynthesisRequest`, main.rs:78-92); without `-o`, raw 16-bit
LE PCM (peak-normalised per sentence / chunk like `as_wave_bytes`) goes to stdout; with `-o` and stdin requests the
files are numbered `<stem>-<n>.<ext>` (main.rs:243-256).  `text` is phonemes, one sentence per line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from typing import Optional

from . import from_config_path
from .piper import PiperSynthesisConfig
from .synth import AudioOutputConfig, SonataSpeechSynthesizer

MODES = ("lazy", "parallel", "realtime")


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="sonata", description="B200-native Piper/VITS synthesis (phoneme input)")
    ap.add_argument("config", help="Model config (<voice>.onnx.json)")
    ap.add_argument("-f", "--input-file", help="Input text file (default stdin: one JSON request per line)")
    ap.add_argument("-o", "--output-file", help="Output WAV file (default stdout: raw i16 PCM)")
    ap.add_argument("--mode", choices=MODES, help="Synthesis mode (default lazy)")
    ap.add_argument("--speaker-id", type=int)
    ap.add_argument("--length-scale", type=float)
    ap.add_argument("--noise-scale", type=float)
    ap.add_argument("--noise-w", type=float)
    ap.add_argument("--rate", type=int, help="Speaking rate [0 - 100] (10 = 1.0x; other values need Sonic, not part of this path)")
    ap.add_argument("--pitch", type=int, help="Speech pitch [0 - 100] (50 = 1.0x)")
    ap.add_argument("--volume", type=int, help="Speech volume [0 - 100]")
    ap.add_argument("--silence", type=int, help="Extra silence (ms) appended to each sentence")
    ap.add_argument("--chunk-size", type=int)
    ap.add_argument("--chunk-padding", type=int)
    ap.add_argument("--device", type=int, default=int(os.environ.get("SONATA_B200_DEVICE", "0")))
    return ap


def process_request(synth: SonataSpeechSynthesizer, default_cfg: PiperSynthesisConfig, req: dict,
                    output_file: Optional[str], out=None) -> None:
    """process_synthesis_request (main.rs:126-165)"""
    out = out or sys.stdout.buffer
    synth.model.set_fallback_synthesis_config(PiperSynthesisConfig(
        req.get("speaker_id"),
        req["noise_scale"] if req.get("noise_scale") is not None else default_cfg.noise_scale,
        req["length_scale"] if req.get("length_scale") is not None else default_cfg.length_scale,
        req["noise_w"] if req.get("noise_w") is not None else default_cfg.noise_w))
    oc = AudioOutputConfig(req.get("rate"), req.get("volume"), req.get("pitch"), req.get("appended_silence_ms"))
    text = req["text"]
    if output_file:
        synth.synthesize_to_file(output_file, text, oc)
        return
    mode = (req.get("mode") or "lazy").lower()
    if mode == "lazy":
        stream = (a.samples for a in synth.synthesize_lazy(text, oc))
    elif mode == "parallel":
        stream = (a.samples for a in synth.synthesize_parallel(text, oc))
    elif mode == "realtime":
        stream = synth.synthesize_streamed(text, oc, req.get("chunk_size") or 100, req.get("chunk_padding") or 3)
    else:
        raise ValueError(f"unknown synthesis mode `{mode}`")
    for samples in stream:
        out.write(samples.as_wave_bytes())
        out.flush()


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    model = from_config_path(args.config, device=args.device)
    synth = SonataSpeechSynthesizer(model)
    default_cfg = model.get_default_synthesis_config()
    if args.input_file:
        with open(args.input_file, encoding="utf-8") as f:
            text = f.read()
        req = {"text": text, "mode": args.mode, "speaker_id": args.speaker_id, "length_scale": args.length_scale,
               "noise_scale": args.noise_scale, "noise_w": args.noise_w, "rate": args.rate, "volume": args.volume,
               "pitch": args.pitch, "appended_silence_ms": args.silence, "chunk_size": args.chunk_size,
               "chunk_padding": args.chunk_padding}
        process_request(synth, default_cfg, req, args.output_file)
    else:
        for i, line in enumerate(sys.stdin):
            if not line.strip():
                continue
            req = json.loads(line)
            out_file = None
            if args.output_file:
                stem, ext = os.path.splitext(args.output_file)
                out_file = f"{stem}-{i + 1}{ext or '.wav'}"
            process_request(synth, default_cfg, req, out_file)
    model.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
