"""Per-stage parity report: CUDA path (debug job) vs the torch oracle on the same ids / noise.
Test infrastructure (imports the oracle, so it lives under tests/).  Usage: python tests/stage_report.py [quality] [n_phonemes] [noise: 0|1] [backend]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sonata_b200  # noqa: E402
from oracle import vits_oracle as vo  # noqa: E402
from sonata_b200 import voicegen  # noqa: E402
from sonata_b200.job import SynthesisJob  # noqa: E402
from sonata_b200.piper import PiperSynthesisConfig  # noqa: E402


def stage_report(quality="medium", n_list=(16,), noise=False, backend=0, seed=1234, verbose=True, utts=None,
                 n_speakers=1, sid=None):
    """utts: utterance seeds (default 0, 1, ...: the index in n_list); tests/screen_margin.py picks seeds whose
    durations are well clear of the ceil() cliff for the full-size cases."""
    cfg_path = voicegen.write_voice(voicegen.default_voice_dir(), quality, seed, n_speakers=n_speakers)
    W = vo.to_torch(voicegen.make_tensors(quality, seed, n_speakers=n_speakers))
    a = vo.arch_of(W)
    model = sonata_b200.from_config_path(cfg_path)
    model.set_backend(backend)
    scales = [0.667, 1.0, 0.8] if noise else [0.0, 1.0, 0.0]
    model.set_fallback_synthesis_config(PiperSynthesisConfig(sid, scales[0], scales[1], scales[2]))
    utts = list(range(len(n_list))) if utts is None else list(utts)
    batches = [vo.synthetic_ids(n, utt=u) for u, n in zip(utts, n_list)]
    g = torch.Generator().manual_seed(99)
    refs, eps_w, eps_z = [], [], []
    for ids in batches:
        st = {}
        ew = torch.randn(1, 2, len(ids), generator=g) if noise else None
        # first pass to learn y_len for eps_z
        ez = None
        if noise:
            st0 = {}
            vo.encode(W, ids, scales, eps_w=ew, eps_z=None, stages=st0, sid=sid)
            ez = torch.randn(1, a["inter"], st0["y_len"], generator=g)
        wav = vo.infer(W, ids, scales, eps_w=ew, eps_z=ez, stages=st, sid=sid)
        refs.append(st)
        eps_w.append(None if ew is None else ew[0].T.contiguous().numpy())
        eps_z.append(None if ez is None else ez[0].T.contiguous().numpy())
    job = SynthesisJob(model, batches, eps_w if noise else None, eps_z if noise else None, debug=True)
    ms = job.run()
    audios = job.fetch()
    frames, samples, _ = job.lengths()
    rows = []

    def cmp(name, got, ref):
        ref = np.asarray(ref, dtype=np.float64)
        got = np.asarray(got, dtype=np.float64)
        if got.shape != ref.shape:
            rows.append((name, "SHAPE", str(got.shape), str(ref.shape)))
            return
        err = float(np.abs(got - ref).max()) if got.size else 0.0
        rows.append((name, err, float(np.abs(ref).max()) if ref.size else 0.0, got.shape))

    report = {"device_ms": ms, "utts": []}
    for b, ids in enumerate(batches):
        st = refs[b]
        rows.clear()
        tm = lambda t: t[0].T.numpy()   # [1,C,T] -> [T,C]
        cmp("x", job.debug_fetch("x", b), tm(st["x"]))
        cmp("stats", job.debug_fetch("stats", b), np.concatenate([tm(st["m_p"]), tm(st["logs_p"])], 1))
        logw_got = job.debug_fetch("logw", b)
        cmp("logw", logw_got, tm(st["logw"]))
        logw_med = float(np.median(np.abs(np.asarray(logw_got, dtype=np.float64) - tm(st["logw"]).astype(np.float64))))
        cum = job.durations(b)
        ref_cum = np.cumsum(st["w_ceil"].view(-1).numpy()).astype(np.int64)
        dur_ok = bool(np.array_equal(cum.astype(np.int64), ref_cum))
        y_ok = frames[b] == st["y_len"]
        if dur_ok and y_ok:
            cmp("z_p", job.debug_fetch("z_p", b), tm(st["z_p"]))
            cmp("z", job.debug_fetch("z", b), tm(st["z"]))
            cmp("dec.pre", job.debug_fetch("dec.pre", b), tm(st["dec.pre"]))
            for i in range(len(a["up_rates"])):
                cmp(f"dec.up{i}", job.debug_fetch(f"dec.up{i}", b), tm(st[f"dec.up{i}"]))
                cmp(f"dec.mrf{i}", job.debug_fetch(f"dec.mrf{i}", b), tm(st[f"dec.mrf{i}"]))
            cmp("wav", audios[b].samples.as_slice(), st["wav"].view(-1).numpy())
        w = st["w"].view(-1)
        fr = w - torch.floor(w)
        margin = float(torch.minimum(fr, 1 - fr).min())
        n_flip = int((np.diff(np.concatenate([[0], cum.astype(np.int64)])) != np.diff(np.concatenate([[0], ref_cum]))).sum())
        u = {"n_ids": len(ids), "utt": utts[b], "y_len_ref": st["y_len"], "y_len_got": frames[b], "durations_exact": dur_ok,
             "ceil_margin": margin, "flipped_ids": n_flip, "logw_median_err": logw_med, "stages": [(r[0], r[1], r[2]) for r in rows]}
        report["utts"].append(u)
        if verbose:
            print(f"--- {quality} utt {b}: T_x={len(ids)} y_len ref/got={st['y_len']}/{frames[b]} "
                  f"durations_exact={dur_ok} ceil_margin={margin:.2e} backend={backend} noise={noise}")
            for r in rows:
                if r[1] == "SHAPE":
                    print(f"   {r[0]:10s} SHAPE MISMATCH got {r[2]} ref {r[3]}")
                else:
                    print(f"   {r[0]:10s} max|err| {r[1]:.3e}   ref absmax {r[2]:.3f}   shape {r[3]}")
    report["profile"] = job.profile()
    job.close()
    model.close()
    return report


if __name__ == "__main__":
    q = sys.argv[1] if len(sys.argv) > 1 else "medium"
    ns = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16").split(",")]
    noise = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
    backend = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rep = stage_report(q, ns, noise, backend)
    for p in rep["profile"]:
        print(f"   region {p['name']:10s} {p['ms']:8.3f} ms  {p['launches']:4d} launches  "
              f"{p['flops'] / 1e9:9.3f} GFLOP  {p['flops'] / max(p['ms'], 1e-9) / 1e9:9.2f} TFLOP/s")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"stage_report_{q}_{'noise' if noise else 'det'}_b{backend}.json"), "w") as f:
        json.dump(rep, f, indent=1, default=str)
