"""Seed screening for the full-size parity tests (test infrastructure: imports the oracle).

`ceil(w)` is a cliff: a duration whose fractional part lies within rounding noise of an integer can
legitimately flip between two correct fp32 implementations.  The full-size parity tests therefore run on utterance
seeds whose smallest distance to an integer, over every id of the utterance, is >= MARGIN in the fp32 oracle AND in
its fp64 shadow -- there the frame counts are well defined and must match exactly.  This script searches the seeds
and prints the table committed in tests/test_gpu_parity.py (SCREENED).

  python tests/screen_margin.py medium 128 1      # quality, phonemes, how many seeds
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import vits_oracle as vo  # noqa: E402
from sonata_b200 import voicegen  # noqa: E402

MARGIN = 1e-3


def ceil_margin(W, a, ids, length_scale=1.0):
    """min over ids of the distance of w = exp(logw)*length_scale to the nearest integer (scales [0, ls, 0])"""
    with torch.inference_mode():
        x, _, _ = vo.text_encoder(W, torch.as_tensor(ids).view(1, -1), a)
        eps = torch.zeros(1, 2, len(ids), dtype=x.dtype)
        logw = vo.sdp_reverse(W, x, eps, 0.0, a)
        w, w_ceil, y_len = vo.durations(logw, length_scale)
    w = w.view(-1).double()
    fr = w - torch.floor(w)
    return float(torch.minimum(fr, 1 - fr).min()), int(y_len)


def screen(quality, n_phonemes, count, first=0, limit=400):
    W32 = vo.to_torch(voicegen.make_tensors(quality))
    W64 = vo.to_torch(voicegen.make_tensors(quality), dtype=torch.float64)
    a = vo.arch_of(W32)
    found = []
    for utt in range(first, first + limit):
        ids = vo.synthetic_ids(n_phonemes, utt=utt)
        m32, y32 = ceil_margin(W32, a, ids)
        if m32 < MARGIN:
            continue
        m64, y64 = ceil_margin(W64, a, ids)
        if m64 < MARGIN or y32 != y64:
            continue
        found.append((utt, min(m32, m64), y32))
        print(f"  {quality} n={n_phonemes} utt={utt}: margin {min(m32, m64):.2e}, frames {y32}", flush=True)
        if len(found) >= count:
            break
    return found


if __name__ == "__main__":
    q = sys.argv[1] if len(sys.argv) > 1 else "medium"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    c = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    print(screen(q, n, c, first))
