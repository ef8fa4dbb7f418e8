"""Synthetic phoneme-id workloads: the inputs BASELINE.json's configs are quoted on.

ids for utterance `u`: rng = PCG64(20260921 + u); N phoneme slots ~ U{3..num_symbols-1}, interleaved
with pad (0) and wrapped in bos (1) / eos (2) exactly like `phonemes_to_input_ids`
(piper/src/lib.rs:232-250) -> 2N+2 ids.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (quality, batch per GPU, phonemes per utterance)
    "C1": ("medium", 1, 128),
    "C2": ("medium", 32, 256),
    "C3": ("high", 16, 512),
}


def synthetic_ids(n_phonemes: int, utt: int = 0, num_symbols: int = 256, seed: int = 20260921) -> np.ndarray:
    r = np.random.Generator(np.random.PCG64(seed + utt))
    ph = r.integers(3, num_symbols, size=n_phonemes)
    ids = np.zeros(2 * n_phonemes + 2, dtype=np.int64)
    ids[0] = 1
    ids[1:-1:2] = ph
    ids[-1] = 2
    return ids


def mixed_lengths(n_utts: int, lo: int = 64, hi: int = 512, seed: int = 7) -> np.ndarray:
    """C5: N ~ U{lo..hi}."""
    r = np.random.Generator(np.random.PCG64(seed))
    return r.integers(lo, hi + 1, size=n_utts)


def length_buckets(n_ids, batch: int):
    """C5 scheduling ("arrival all-at-once, length-bucketed"): utterance indices sorted by id count,
    longest first, cut into consecutive batches of `batch` -- each batch then holds similar lengths, so the packed
    segment layout wastes little padding and long utterances do not wait behind many short ones."""
    order = sorted(range(len(n_ids)), key=lambda i: (-int(n_ids[i]), i))
    return [order[i:i + batch] for i in range(0, len(order), batch)]


def completion_stats(buckets, bucket_done_s, audio_s):
    """Per-utterance completion latency when every request arrives at t = 0 and bucket k's audio is delivered at
    `bucket_done_s[k]` (cumulative seconds): returns (p50, p99, aggregate audio-s/s)."""
    lat = np.concatenate([np.full(len(b), float(t)) for b, t in zip(buckets, bucket_done_s)]) if buckets else np.zeros(0)
    if lat.size == 0:
        return 0.0, 0.0, 0.0
    total = float(bucket_done_s[-1])
    return float(np.percentile(lat, 50)), float(np.percentile(lat, 99)), (float(np.sum(audio_s)) / total if total > 0 else 0.0)
