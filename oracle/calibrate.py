"""ORACLE-SIDE TOOL (test infrastructure): fit the per-tensor gains of the synthetic voice.

Runs the oracle forward once on a calibration utterance and rescales every conv so its
output standard deviation hits a target (LSUV-style), then sets the duration predictor's
ElementwiseAffine so durations average ~3 frames per id (SURVEY §8 calibration target).
Writes ``sonata_b200/data/gains_<quality>.json``; ``sonata_b200/voicegen.py`` multiplies its
seeded base tensors by these gains.  Run:  python -m oracle.calibrate medium high
"""
from __future__ import annotations

import json
import math
import os
import re
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import vits_oracle as vo  # noqa: E402
from sonata_b200 import voicegen  # noqa: E402

SEED = 1234


def make_targets(a):
    nb = a["dp_bins"]
    H = a["hidden"]

    def targets(name):
        if re.search(r"attn_layers\.\d+\.conv_[qkv]$", name):
            return 1.0
        if name.endswith("conv_o") or name.endswith("conv_2"):
            return 0.5
        if name.endswith("conv_1"):
            return 1.0
        if name == "enc_p.proj":
            return 0.6
        if name in ("dp.pre", "dp.proj"):
            return 1.0
        if "convs_1x1" in name:
            return 1.0
        if re.match(r"dp\.flows\.\d+\.pre$", name):
            return 1.0
        if re.match(r"dp\.flows\.\d+\.proj$", name):
            t = torch.ones(3 * nb - 1)
            t[:2 * nb] = math.sqrt(H) * 1.0   # widths / heights are divided by sqrt(H)
            t[2 * nb:] = 1.0
            return t
        if re.match(r"flow\.flows\.\d+\.pre$", name):
            return 1.0
        if "in_layers" in name:
            return 1.0
        if "res_skip_layers" in name:
            return 0.5
        if re.match(r"flow\.flows\.\d+\.post$", name):
            return 0.5
        if name == "dec.conv_pre":
            return 1.0
        if name.startswith("dec.ups."):
            return 1.0
        if "convs1." in name:
            return 1.0
        if "convs2." in name or re.search(r"resblocks\.\d+\.convs\.", name):
            return 0.4
        if name == "dec.conv_post":
            return 0.5
        return None

    return targets


def calibrate(quality: str, n_phonemes: int = 48, mean_log_w: float = math.log(2.35),
              std_log_w: float = 0.35):
    a = voicegen.ARCH[quality]
    W = vo.to_torch(voicegen.make_tensors(quality, SEED, gains={}))
    calib = vo.Calib(make_targets(a))
    ids = vo.synthetic_ids(n_phonemes, utt=999)
    g = torch.Generator().manual_seed(7)
    T = len(ids)
    eps_w = torch.randn(1, 2, T, generator=g)
    arch = vo.arch_of(W)
    # pass 1: encoder + duration predictor, then fit the ElementwiseAffine
    st = {}
    idt = torch.as_tensor(ids).view(1, -1)
    x, m_p, logs_p = vo.text_encoder(W, idt, arch, calib, st)
    W["dp.flows.0.m"] = torch.zeros(2, 1)
    W["dp.flows.0.logs"] = torch.zeros(2, 1)
    vo.sdp_reverse(W, x, eps_w, 0.8, arch, calib, st)
    z0 = st["dp.pre_ea"][0, 0]
    mu, sd = float(z0.mean()), float(z0.std())
    logs0 = math.log(sd / std_log_w)
    m0 = mu - mean_log_w * sd / std_log_w
    W["dp.flows.0.m"] = torch.tensor([[m0], [0.0]])
    W["dp.flows.0.logs"] = torch.tensor([[logs0], [0.0]])
    calib.gains["dp.flows.0.m"] = {"value": [m0, 0.0]}
    calib.gains["dp.flows.0.logs"] = {"value": [logs0, 0.0]}
    # pass 2: full path (already-calibrated layers are re-fitted to ~1.0 -> no-op)
    logw = vo.sdp_reverse(W, x, eps_w, 0.8, arch, None, None)
    _, w_ceil, y_len = vo.durations(logw, 1.0)
    eps_z = torch.randn(1, a["inter"], y_len, generator=g)
    z_p, _ = vo.expand(m_p, logs_p, w_ceil, y_len, eps_z, 0.667)
    z = vo.flow_reverse(W, z_p, arch, calib)
    wav = vo.decoder(W, z, arch, calib)
    print(f"[{quality}] T_x={T} y_len={y_len} frames/id={y_len / T:.3f} "
          f"wav std={float(wav.std()):.3f} absmax={float(wav.abs().max()):.3f}")
    return calib.gains


def main(argv):
    qs = argv or ["medium", "high"]
    os.makedirs(voicegen._DATA_DIR, exist_ok=True)
    for q in qs:
        gains = calibrate(q)
        p = os.path.join(voicegen._DATA_DIR, f"gains_{q}.json")
        with open(p, "w") as f:
            json.dump(gains, f, indent=0, sort_keys=True)
        print("wrote", p, len(gains), "gains")
        # verify through the product-side generator
        W = vo.to_torch(voicegen.make_tensors(q, SEED))
        for n, utt in ((16, 0), (64, 1)):
            ids = vo.synthetic_ids(n, utt)
            for sc in ([0.0, 1.0, 0.0],):
                st = {}
                wav = vo.infer(W, ids, sc, stages=st)
                w = st["w"].view(-1)
                frac = (w - torch.floor(w))
                margin = float(torch.minimum(frac, 1 - frac).min())
                print(f"  N={n} scales={sc} T_x={len(ids)} y_len={st['y_len']} "
                      f"frames/id={st['y_len'] / len(ids):.3f} w[min,max]=({float(w.min()):.2f},"
                      f"{float(w.max()):.2f}) ceil-margin={margin:.2e} wav absmax={float(wav.abs().max()):.3f} "
                      f"std={float(wav.std()):.3f}")


if __name__ == "__main__":
    main(sys.argv[1:])
