"""Round-2 planning experiment (CPU only; lives under tests/ because it uses the oracle): would the text encoder and the
duration predictor keep their frame counts if their dense convolutions ran on tensor cores?

Every dense Conv1d of `text_encoder` + `sdp_reverse` is replaced by the tensor-core accumulation model of
tools/emu_tc_accuracy.py (exact K-step partial sums added to an fp32 accumulator that rounds toward zero, K index =
(tap, channel) like the CUDA kernels).  Reported per scheme: max |logw - logw_fp64| and the number of ids whose
`ceil(exp(logw))` differs from the fp64 result, next to the plain fp32 oracle.

  python tests/emu_encoder_durations.py [n_utterances=8] [phonemes=96]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from emu_tc_accuracy import emulate  # noqa: E402
from oracle import vits_oracle as vo  # noqa: E402
from sonata_b200 import voicegen, workload  # noqa: E402

_real_conv1d = torch.nn.functional.conv1d


def make_conv(fmt, acc, chunk):
    def conv1d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
        O, C, k = w.shape
        if groups != 1 or C < 32 or x.dtype != torch.float32:
            return _real_conv1d(x, w, b, stride, padding, dilation, groups)
        assert x.shape[0] == 1 and stride == 1
        if isinstance(padding, (tuple, list)):
            padding = padding[0]
        if isinstance(dilation, (tuple, list)):
            dilation = dilation[0]
        xp = torch.nn.functional.pad(x[0], (padding, padding)).numpy()            # [C, T + 2p]
        T = xp.shape[1] - dilation * (k - 1)                                      # output length
        cols = [xp[:, t * dilation:t * dilation + T].T for t in range(k)]         # tap-major K index
        X = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.float32)  # [T, k*C]
        Wm = np.ascontiguousarray(w.numpy().transpose(2, 1, 0).reshape(k * C, O), dtype=np.float32)
        y = emulate(X, Wm, fmt, acc, chunk)
        if b is not None:
            y = (y + b.numpy()[None, :]).astype(np.float32)
        return torch.from_numpy(np.ascontiguousarray(y.T))[None]
    return conv1d


def logw_of(W, ids, a):
    x, _, _ = vo.text_encoder(W, ids, a)
    eps = torch.zeros(1, 2, ids.shape[1], dtype=W["enc_p.emb.weight"].dtype)
    return vo.sdp_reverse(W, x, eps, 0.0, a)[0, 0]


def main():
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    nph = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    tensors = voicegen.make_tensors("medium", 1234)
    W32, W64 = vo.to_torch(tensors), vo.to_torch(tensors, dtype=torch.float64)
    a = vo.arch_of(W32)
    schemes = [("fp32 (oracle as is)", None), ("bf16x2, truncating acc", ("bf16", "rz", 0)),
               ("3xTF32, truncating acc", ("tf32", "rz", 0)), ("3xTF32, flush / 8 K-steps", ("tf32", "rz", 8)),
               ("3xTF32, flush / 4 K-steps", ("tf32", "rz", 4))]
    res = {name: [0.0, 0, 0, 0.0, []] for name, _ in schemes}     # max err, flips vs fp64, flips vs fp32, sum |dw|, errs
    total = 0
    with torch.inference_mode():
        for u in range(n_utts):
            ids = torch.from_numpy(workload.synthetic_ids(nph, utt=500 + u)).view(1, -1)
            ref = logw_of(W64, ids, a).numpy()
            d_ref = np.ceil(np.exp(ref))
            base = logw_of(W32, ids, a).numpy()
            d_base = np.ceil(np.exp(base.astype(np.float64)))
            total += ids.shape[1]
            for name, cfg in schemes:
                if cfg is None:
                    lw = base
                else:
                    vo.F.conv1d = make_conv(*cfg)
                    try:
                        lw = logw_of(W32, ids, a).numpy()
                    finally:
                        vo.F.conv1d = _real_conv1d
                d = np.ceil(np.exp(lw.astype(np.float64)))
                r = res[name]
                r[0] = max(r[0], float(np.abs(lw - ref).max()))
                r[1] += int((d != d_ref).sum()); r[2] += int((d != d_base).sum())
                # a frame count flips when w = exp(logw) and its fp64 value straddle an integer: for a uniformly
                # distributed fractional part that happens with probability |dw|
                r[3] += float(np.abs(np.exp(lw.astype(np.float64)) - np.exp(ref)).sum())
                r[4].append(np.abs(lw - ref))
            print(f"utterance {u + 1}/{n_utts} done", flush=True)
    print(f"\n{total} ids ({n_utts} x {nph} phonemes), medium voice, noise_w = 0")
    print(f"{'scheme':32s} {'max|logw-fp64|':>15s} {'median':>10s} {'flips vs fp64':>14s} {'flips vs fp32':>14s} {'E[flips] / 16448 ids':>22s}")
    for name, _ in schemes:
        r = res[name]
        med = float(np.median(np.concatenate(r[4])))
        print(f"{name:32s} {r[0]:15.2e} {med:10.2e} {r[1]:14d} {r[2]:14d} {r[3] / total * 16448:22.3f}")


if __name__ == "__main__":
    main()
