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
