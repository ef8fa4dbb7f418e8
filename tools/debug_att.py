"""Isolates the tensor-core attention path (two grouped GEMMs around the relative-position softmax) against the fp32
CUDA-core attention kernel and a numpy restatement, on the first encoder layer.  python tools/debug_att.py [n ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sonata_b200  # noqa: E402
from sonata_b200 import voicegen, workload  # noqa: E402
from sonata_b200.job import SynthesisJob  # noqa: E402
from sonata_b200.piper import PiperSynthesisConfig  # noqa: E402


def run(model, ids, simt):
    if simt:
        os.environ["SB200_ATT_SIMT"] = "1"
    else:
        os.environ.pop("SB200_ATT_SIMT", None)
    job = SynthesisJob(model, ids, debug=True)
    job.run()
    out = []
    for b in range(len(ids)):
        d = {k: job.debug_fetch(k, b) for k in ("qkv0", "att0")}
        if not simt:
            d["p0"] = job.debug_fetch("p0", b)
        d["cum"] = job.durations(b)
        out.append(d)
    job.close()
    return out


def main():
    ns = [int(x) for x in sys.argv[1:]] or [5, 40, 128, 256]
    cfg = voicegen.write_voice(voicegen.default_voice_dir(), "medium")
    m = sonata_b200.from_config_path(cfg, device=0)
    m.set_fallback_synthesis_config(PiperSynthesisConfig(None, 0.0, 1.0, 0.0))
    T = voicegen.make_tensors("medium")
    relk = np.asarray(T["enc_p.encoder.attn_layers.0.emb_rel_k"])[0]      # [9][96]
    relv = np.asarray(T["enc_p.encoder.attn_layers.0.emb_rel_v"])[0]
    ids = [workload.synthetic_ids(n, utt=i) for i, n in enumerate(ns)]
    A = run(m, ids, simt=True)
    B = run(m, ids, simt=False)
    for b, (a, t) in enumerate(zip(A, B)):
        Tn = a["qkv0"].shape[0]
        qk_err = np.abs(a["qkv0"][:, :384] - t["qkv0"][:, :384]).max()
        print(f"utt {b}: T={Tn}  q/k diff {qk_err:.2e}  att0 simt-vs-tc max {np.abs(a['att0'] - t['att0']).max():.3e} "
              f"(|att0| max {np.abs(a['att0']).max():.2f})  durations equal {np.array_equal(a['cum'], t['cum'])}")
        q = a["qkv0"][:, 0:96].astype(np.float64) / np.sqrt(96.0)
        k = a["qkv0"][:, 192:288].astype(np.float64)
        v = a["qkv0"][:, 384:480].astype(np.float64)
        S = q @ k.T
        for i in range(Tn):
            for d in range(9):
                j = i + d - 4
                if 0 <= j < Tn:
                    S[i, j] += q[i] @ relk[d]
        P = np.exp(S - S.max(1, keepdims=True)); P /= P.sum(1, keepdims=True)
        O = P @ v
        for i in range(Tn):
            for d in range(9):
                j = i + d - 4
                if 0 <= j < Tn:
                    O[i] += P[i, j] * relv[d]
        print(f"        numpy vs simt att0[:, :96] {np.abs(O - a['att0'][:, :96]).max():.3e}   numpy vs tc {np.abs(O - t['att0'][:, :96]).max():.3e}")
        p0 = t["p0"][:, :Tn]
        print(f"        P (head 0) tc vs numpy {np.abs(p0 - P).max():.3e}   row sums tc min/max {p0.sum(1).min():.6f}/{p0.sum(1).max():.6f}   "
              f"pad zeros {np.abs(t['p0'][:, Tn:(Tn + 31) // 32 * 32]).max() if Tn % 32 else 0.0:.1e}")
        bad = np.argwhere(np.abs(p0 - P) > 1e-4)
        if len(bad):
            print("        first bad P entries (row, key):", bad[:8].tolist())
    m.close()


if __name__ == "__main__":
    main()
