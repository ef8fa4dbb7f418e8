"""ORACLE — TEST INFRASTRUCTURE ONLY.  **Parity unpinned against the reference itself** (see below); checked against an
independent third-party implementation of the same published algorithm instead (last paragraph).

CPU restatement (PyTorch fp32, optional fp64 shadow) of the arithmetic the reference runs
inside ``ort::Session::run`` at ``crates/sonata/models/piper/src/lib.rs:362-379``: the Piper
VITS graph ``SynthesizerTrn.infer`` (text encoder -> stochastic duration predictor (reverse)
-> ceil/cumsum/path -> residual-coupling flow (reverse) -> HiFi-GAN generator).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  Nothing under ``sonata_b200/`` does.

Why "parity unpinned": the graph lives in a third-party, un-vendored artefact (a Piper voice
``.onnx`` from rhasspy/piper-voices, executed by onnxruntime 1.20.x through
``ort = 2.0.0-rc.9`` / ``ort-sys = 2.0.0-rc.9``, ``Cargo.lock:1203-1226``).  Neither a voice
file, nor onnxruntime, nor a Rust toolchain exists in this sandbox, and the reference's own
tests assert only "no error" (``crates/sonata/synth/src/tests.rs:5-28``).  This file therefore
restates the *published* algorithm (rhasspy/piper ``src/python/piper_train/vits/{models,
modules,attentions,commons,transforms}.py`` as exported by ``export_onnx.py``) and is anchored
on what the reference does pin at its call sites:

* positional inputs ``input:i64[1,T]``, ``input_lengths:i64[1]``, ``scales:f32[3] =
  [noise_scale, length_scale, noise_w]`` (``piper/src/lib.rs:345-352``) -> ``infer`` below;
* batch == sequential B=1 runs (``piper/src/lib.rs:433-435``) -> every function here is B=1;
* streaming split: encoder outputs named ``z``, ``y_mask`` and a decoder on frame slices of
  axis 2 (``piper/src/lib.rs:681-735, 793-840``) -> ``encode`` / ``decode``;
* hop = 256 samples per frame (``piper/src/lib.rs:910``).

Second opinion (what this restatement IS checked against): Hugging Face ``transformers`` 5.5.0 ``VitsModel`` -- an
independent implementation of the published VITS network, not derived from Piper and not written here -- loaded with
the same synthetic high-quality (ResBlock1, en_US-ryan-high architecture) voice reproduces this oracle's waveform to
2e-6 .. 4e-6, frame counts identical, on the deterministic AND the stochastic path (``tests/hf_reference.py``,
``tests/test_hf_pin.py``, fixtures + generator under ``tests/golden/hf/``), multi-speaker conditioning included (3-speaker
voice, ``speaker_id``); the medium voice agrees with it up to the
vocoder input (``z`` within 5e-5).  Not covered by it: the ResBlock2 wiring of the medium voices' vocoder (transformers
implements ResBlock1 only) and anything Piper's ONNX export may do differently from the published model code.

Every function takes ``W``: dict name -> torch tensor with Piper state-dict names.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1


# --------------------------------------------------------------------------- utilities
def to_torch(tensors, dtype=torch.float32):
    W = {}
    for k, v in tensors.items():
        if k.startswith("hp."):
            W[k] = torch.from_numpy(np.asarray(v).astype(np.int64))
        else:
            W[k] = torch.from_numpy(np.asarray(v)).to(dtype)
    return W


def arch_of(W):
    a = [int(x) for x in W["hp.arch"]]
    keys = ["hidden", "inter", "filter", "heads", "layers", "kernel", "window", "n_vocab",
            "resblock", "up_init", "flow_n", "wn_layers", "flow_kernel", "dp_kernel", "dp_bins",
            "sample_rate"]
    d = dict(zip(keys, a))
    d["up_rates"] = [int(x) for x in W["hp.up_rates"]]
    d["up_kernels"] = [int(x) for x in W["hp.up_kernels"]]
    d["res_kernels"] = [int(x) for x in W["hp.res_kernels"]]
    d["res_dils"] = [[int(y) for y in r] for r in W["hp.res_dils"]]
    return d


class Calib:
    """Optional LSUV-style calibration hook used only by oracle/calibrate.py: when a conv's
    name has a target, its weight+bias are rescaled in place so the output std hits it."""

    def __init__(self, targets):
        self.targets = targets
        self.gains = {}


def _conv(W, name, x, calib=None, **kw):
    w = W[name + ".weight"]
    b = W.get(name + ".bias")
    y = F.conv1d(x, w, b, **kw)
    if calib is not None:
        tgt = calib.targets(name)
        if tgt is not None:
            if torch.is_tensor(tgt):   # per-output-channel targets
                s = tgt / (y.std(dim=(0, 2)) + 1e-12)
                W[name + ".weight"] = w * s.view(-1, 1, 1)
                calib.gains[name + ".weight"] = {"rows": [float(v) for v in s]}
                if b is not None:
                    W[name + ".bias"] = b * s
                    calib.gains[name + ".bias"] = {"rows": [float(v) for v in s]}
            else:
                s = float(tgt / (y.std() + 1e-12))
                W[name + ".weight"] = w * s
                calib.gains[name + ".weight"] = s
                if b is not None:
                    W[name + ".bias"] = b * s
                    calib.gains[name + ".bias"] = s
            y = F.conv1d(x, W[name + ".weight"], W.get(name + ".bias"), **kw)
    return y


def _layer_norm(W, name, x):
    # modules.LayerNorm: normalise over the channel dim of [B, C, T], eps = 1e-5
    C = x.shape[1]
    y = F.layer_norm(x.transpose(1, -1), (C,), W[name + ".gamma"], W[name + ".beta"], 1e-5)
    return y.transpose(1, -1)


# --------------------------------------------------------------------------- text encoder
def _rel_to_abs(x):
    # attentions.MultiHeadAttention._relative_position_to_absolute_position
    b, h, l, _ = x.size()
    x = F.pad(x, [0, 1])
    x_flat = x.view([b, h, l * 2 * l])
    x_flat = F.pad(x_flat, [0, l - 1])
    return x_flat.view([b, h, l + 1, 2 * l - 1])[:, :, :l, l - 1:]


def _abs_to_rel(x):
    b, h, l, _ = x.size()
    x = F.pad(x, [0, l - 1])
    x_flat = x.view([b, h, l ** 2 + l * (l - 1)])
    x_flat = F.pad(x_flat, [l, 0])
    return x_flat.view([b, h, l, 2 * l])[:, :, :, 1:]


def _get_rel_emb(rel, length, window):
    pad_length = max(length - (window + 1), 0)
    start = max((window + 1) - length, 0)
    end = start + 2 * length - 1
    if pad_length > 0:
        rel = F.pad(rel, [0, 0, pad_length, pad_length])
    return rel[:, start:end]


def _mha(W, p, x, a, calib=None):
    H, nh, win = a["hidden"], a["heads"], a["window"]
    kc = H // nh
    q = _conv(W, p + "conv_q", x, calib)
    k = _conv(W, p + "conv_k", x, calib)
    v = _conv(W, p + "conv_v", x, calib)
    b, d, t = k.shape
    q = q.view(b, nh, kc, t).transpose(2, 3)
    k = k.view(b, nh, kc, t).transpose(2, 3)
    v = v.view(b, nh, kc, t).transpose(2, 3)
    qs = q / math.sqrt(kc)
    scores = torch.matmul(qs, k.transpose(-2, -1))
    rel_k = _get_rel_emb(W[p + "emb_rel_k"], t, win)
    rel_logits = torch.matmul(qs, rel_k.unsqueeze(0).transpose(-2, -1))
    scores = scores + _rel_to_abs(rel_logits)
    # x_mask is all ones for B=1 (piper/src/lib.rs:346-347: length == T) -> no masked_fill
    p_attn = F.softmax(scores, dim=-1)
    out = torch.matmul(p_attn, v)
    rel_w = _abs_to_rel(p_attn)
    rel_v = _get_rel_emb(W[p + "emb_rel_v"], t, win)
    out = out + torch.matmul(rel_w, rel_v.unsqueeze(0))
    out = out.transpose(2, 3).contiguous().view(b, d, t)
    return _conv(W, p + "conv_o", out, calib)


def _ffn(W, p, x, a, calib=None):
    k = a["kernel"]
    pl, pr = (k - 1) // 2, k // 2
    y = _conv(W, p + "conv_1", F.pad(x, [pl, pr]), calib)
    y = torch.relu(y)
    y = _conv(W, p + "conv_2", F.pad(y, [pl, pr]), calib)
    return y


def text_encoder(W, ids, a, calib=None, stages=None):
    """enc_p: ids i64[1,T] -> x[1,H,T], m_p[1,I,T], logs_p[1,I,T]."""
    H = a["hidden"]
    x = F.embedding(ids, W["enc_p.emb.weight"]) * math.sqrt(H)
    x = x.transpose(1, -1)
    if stages is not None:
        stages["enc.emb"] = x
    for i in range(a["layers"]):
        y = _mha(W, f"enc_p.encoder.attn_layers.{i}.", x, a, calib)
        x = _layer_norm(W, f"enc_p.encoder.norm_layers_1.{i}", x + y)
        y = _ffn(W, f"enc_p.encoder.ffn_layers.{i}.", x, a, calib)
        x = _layer_norm(W, f"enc_p.encoder.norm_layers_2.{i}", x + y)
        if stages is not None:
            stages[f"enc.layer{i}"] = x
    stats = _conv(W, "enc_p.proj", x, calib)
    m_p, logs_p = torch.split(stats, a["inter"], dim=1)
    return x, m_p, logs_p


# --------------------------------------------------------------------------- duration predictor
def _dds(W, p, x, a, g=None, calib=None):
    k = a["dp_kernel"]
    if g is not None:
        x = x + g
    for i in range(3):
        dil = k ** i
        pad = (k * dil - dil) // 2
        y = F.conv1d(x, W[f"{p}convs_sep.{i}.weight"], W[f"{p}convs_sep.{i}.bias"],
                     groups=x.shape[1], dilation=dil, padding=pad)
        y = _layer_norm(W, f"{p}norms_1.{i}", y)
        y = F.gelu(y)
        y = _conv(W, f"{p}convs_1x1.{i}", y, calib)
        y = _layer_norm(W, f"{p}norms_2.{i}", y)
        y = F.gelu(y)
        x = x + y
    return x


def _searchsorted(bin_locations, inputs, eps=1e-6):
    bl = bin_locations.clone()
    bl[..., -1] += eps
    return torch.sum(inputs[..., None] >= bl, dim=-1) - 1


def _rqs_inverse(inputs, uw, uh, ud, tail_bound=5.0, min_bin_width=1e-3, min_bin_height=1e-3,
                 min_derivative=1e-3):
    """transforms.unconstrained_rational_quadratic_spline(inverse=True, tails='linear')."""
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
    outputs = inputs.clone()
    ud = F.pad(ud, pad=(1, 1))
    constant = math.log(math.exp(1 - min_derivative) - 1)
    ud[..., 0] = constant
    ud[..., -1] = constant
    if not inside.any():
        return outputs
    x = inputs[inside]
    uw, uh, ud = uw[inside, :], uh[inside, :], ud[inside, :]
    left = bottom = -tail_bound
    right = top = tail_bound
    nb = uw.shape[-1]
    widths = F.softmax(uw, dim=-1)
    widths = min_bin_width + (1 - min_bin_width * nb) * widths
    cumwidths = torch.cumsum(widths, dim=-1)
    cumwidths = F.pad(cumwidths, pad=(1, 0), mode="constant", value=0.0)
    cumwidths = (right - left) * cumwidths + left
    cumwidths[..., 0] = left
    cumwidths[..., -1] = right
    widths = cumwidths[..., 1:] - cumwidths[..., :-1]
    derivatives = min_derivative + F.softplus(ud)
    heights = F.softmax(uh, dim=-1)
    heights = min_bin_height + (1 - min_bin_height * nb) * heights
    cumheights = torch.cumsum(heights, dim=-1)
    cumheights = F.pad(cumheights, pad=(1, 0), mode="constant", value=0.0)
    cumheights = (top - bottom) * cumheights + bottom
    cumheights[..., 0] = bottom
    cumheights[..., -1] = top
    heights = cumheights[..., 1:] - cumheights[..., :-1]
    bin_idx = _searchsorted(cumheights, x)[..., None]
    in_cumw = cumwidths.gather(-1, bin_idx)[..., 0]
    in_w = widths.gather(-1, bin_idx)[..., 0]
    in_cumh = cumheights.gather(-1, bin_idx)[..., 0]
    delta = heights / widths
    in_delta = delta.gather(-1, bin_idx)[..., 0]
    in_d = derivatives.gather(-1, bin_idx)[..., 0]
    in_d1 = derivatives[..., 1:].gather(-1, bin_idx)[..., 0]
    in_h = heights.gather(-1, bin_idx)[..., 0]
    aa = (x - in_cumh) * (in_d + in_d1 - 2 * in_delta) + in_h * (in_delta - in_d)
    bb = in_h * in_d - (x - in_cumh) * (in_d + in_d1 - 2 * in_delta)
    cc = -in_delta * (x - in_cumh)
    disc = bb.pow(2) - 4 * aa * cc
    root = (2 * cc) / (-bb - torch.sqrt(disc))
    outputs[inside] = root * in_w + in_cumw
    return outputs


def _conv_flow_reverse(W, p, z, g, a, calib=None, stages=None):
    H, nb = a["hidden"], a["dp_bins"]
    z0, z1 = torch.split(z, [1, 1], 1)
    h = _conv(W, p + "pre", z0, calib)
    h = _dds(W, p + "convs.", h, a, g=g, calib=calib)
    h = _conv(W, p + "proj", h, calib)
    b, c, t = z0.shape
    h = h.reshape(b, c, -1, t).permute(0, 1, 3, 2)
    uw = h[..., :nb] / math.sqrt(H)
    uh = h[..., nb:2 * nb] / math.sqrt(H)
    ud = h[..., 2 * nb:]
    z1 = _rqs_inverse(z1, uw, uh, ud)
    return torch.cat([z0, z1], 1)


def speaker_embedding(W, sid):
    """g = emb_g(sid).unsqueeze(-1) for multi-speaker voices (SynthesizerTrn.infer; the reference feeds `sid` only when
    num_speakers > 1, piper/src/lib.rs:353-358); None for single-speaker voices."""
    if "emb_g.weight" not in W:
        return None
    return W["emb_g.weight"][int(sid or 0)].view(1, -1, 1)


def sdp_reverse(W, x, eps_w, noise_w, a, calib=None, stages=None, g=None):
    """dp(x, reverse=True): x[1,H,T], eps_w[1,2,T] ~ N(0,1) -> logw[1,1,T].  g: speaker embedding [1,gin,1] or None
    (StochasticDurationPredictor: x = pre(x) + cond(g))."""
    h = _conv(W, "dp.pre", x, calib)
    if g is not None:
        h = h + _conv(W, "dp.cond", g, calib)
    h = _dds(W, "dp.convs.", h, a, calib=calib)
    h = _conv(W, "dp.proj", h, calib)
    z = eps_w * noise_w
    # reversed(flows)[:-2] + [EA]  ==  Flip, CF4^-1, Flip, CF3^-1, Flip, CF2^-1, Flip, EA^-1
    for fi in (7, 5, 3):
        z = torch.flip(z, [1])
        z = _conv_flow_reverse(W, f"dp.flows.{fi}.", z, h, a, calib)
        if stages is not None:
            stages[f"dp.flow{fi}"] = z
    z = torch.flip(z, [1])
    if stages is not None:
        stages["dp.pre_ea"] = z
    z = (z - W["dp.flows.0.m"]) * torch.exp(-W["dp.flows.0.logs"])
    return z[:, 0:1]


# --------------------------------------------------------------------------- alignment
def durations(logw, length_scale):
    w = torch.exp(logw) * length_scale
    w_ceil = torch.ceil(w)
    y_len = int(torch.clamp_min(torch.sum(w_ceil), 1).item())
    return w, w_ceil, y_len


def expand(m_p, logs_p, w_ceil, y_len, eps_z, noise_scale):
    """generate_path + attn^T matmul, restated as the gather it is: frame j takes token i
    iff cum[i-1] <= j < cum[i]."""
    cum = torch.cumsum(w_ceil.view(-1), 0)
    j = torch.arange(y_len, dtype=cum.dtype)
    tok = torch.searchsorted(cum, j, right=True)
    T = m_p.shape[2]
    ok = tok < T
    tok_c = tok.clamp_max(T - 1)
    m = m_p[:, :, tok_c] * ok
    lg = logs_p[:, :, tok_c] * ok
    if eps_z is None or noise_scale == 0.0:
        return m, tok
    return m + eps_z * torch.exp(lg) * noise_scale, tok


# --------------------------------------------------------------------------- flow
def _wn(W, p, x, a, calib=None, g=None):
    H, k = a["hidden"], a["flow_kernel"]
    out = torch.zeros_like(x)
    n = a["wn_layers"]
    gc = _conv(W, p + "cond_layer", g, calib) if g is not None else None      # modules.WN: all layers' conditioning at once
    for l in range(n):
        x_in = _conv(W, p + f"in_layers.{l}", x, calib, padding=(k - 1) // 2)
        if gc is not None:
            x_in = x_in + gc[:, 2 * H * l:2 * H * (l + 1)]
        acts = torch.tanh(x_in[:, :H]) * torch.sigmoid(x_in[:, H:])
        rs = _conv(W, p + f"res_skip_layers.{l}", acts, calib)
        if l < n - 1:
            x = x + rs[:, :H]
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out


def flow_reverse(W, z, a, calib=None, stages=None, g=None):
    half = a["inter"] // 2
    for f in reversed(range(a["flow_n"])):
        z = torch.flip(z, [1])
        p = f"flow.flows.{2 * f}."
        x0, x1 = torch.split(z, [half, half], 1)
        h = _conv(W, p + "pre", x0, calib)
        h = _wn(W, p + "enc.", h, a, calib, g=g)
        m = _conv(W, p + "post", h, calib)
        x1 = x1 - m
        z = torch.cat([x0, x1], 1)
        if stages is not None:
            stages[f"flow.{f}"] = z
    return z


# --------------------------------------------------------------------------- HiFi-GAN
def decoder(W, z, a, calib=None, stages=None, g=None):
    x = _conv(W, "dec.conv_pre", z, calib, padding=3)
    if g is not None:
        x = x + _conv(W, "dec.cond", g, calib)
    if stages is not None:
        stages["dec.pre"] = x
    nk = len(a["res_kernels"])
    for i, (u, k) in enumerate(zip(a["up_rates"], a["up_kernels"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        w, b = W[f"dec.ups.{i}.weight"], W[f"dec.ups.{i}.bias"]
        y = F.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
        if calib is not None:
            tgt = calib.targets(f"dec.ups.{i}")
            if tgt is not None:
                s = float(tgt / y.std())
                W[f"dec.ups.{i}.weight"], W[f"dec.ups.{i}.bias"] = w * s, b * s
                calib.gains[f"dec.ups.{i}.weight"] = s
                calib.gains[f"dec.ups.{i}.bias"] = s
                y = y * s
        x = y
        if stages is not None:
            stages[f"dec.up{i}"] = x
        xs = None
        for j, (rk, rd) in enumerate(zip(a["res_kernels"], a["res_dils"])):
            p = f"dec.resblocks.{i * nk + j}."
            xb = x
            if a["resblock"] == 2:
                for m, d in enumerate(rd):
                    xt = F.leaky_relu(xb, LRELU_SLOPE)
                    xt = _conv(W, p + f"convs.{m}", xt, calib, dilation=d, padding=d * (rk - 1) // 2)
                    xb = xt + xb
            else:
                for m, d in enumerate(rd):
                    xt = F.leaky_relu(xb, LRELU_SLOPE)
                    xt = _conv(W, p + f"convs1.{m}", xt, calib, dilation=d, padding=d * (rk - 1) // 2)
                    xt = F.leaky_relu(xt, LRELU_SLOPE)
                    xt = _conv(W, p + f"convs2.{m}", xt, calib, padding=(rk - 1) // 2)
                    xb = xt + xb
            xs = xb if xs is None else xs + xb
        x = xs / nk
        if stages is not None:
            stages[f"dec.mrf{i}"] = x
    x = F.leaky_relu(x)  # default slope 0.01
    x = _conv(W, "dec.conv_post", x, calib, padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------- whole path
def encode(W, ids, scales, eps_w=None, eps_z=None, calib=None, stages=None, sid=None):
    """Streaming 'encoder.onnx' half: ids -> z[1,I,T_y] (piper/src/lib.rs:537-574).
    scales = [noise_scale, length_scale, noise_w] (piper/src/lib.rs:348-352)."""
    a = arch_of(W)
    noise_scale, length_scale, noise_w = (float(s) for s in scales)
    ids = torch.as_tensor(np.asarray(ids, dtype=np.int64)).view(1, -1)
    T = ids.shape[1]
    dt = W["enc_p.emb.weight"].dtype
    x, m_p, logs_p = text_encoder(W, ids, a, calib, stages)
    if eps_w is None:
        eps_w = torch.zeros(1, 2, T, dtype=dt)
    else:
        eps_w = torch.as_tensor(eps_w).to(dt).view(1, 2, T)
    g = speaker_embedding(W, sid)
    logw = sdp_reverse(W, x, eps_w, noise_w, a, calib, stages, g=g)
    w, w_ceil, y_len = durations(logw, length_scale)
    if eps_z is not None:
        eps_z = torch.as_tensor(eps_z).to(dt).view(1, a["inter"], y_len)
    z_p, tok = expand(m_p, logs_p, w_ceil, y_len, eps_z, noise_scale)
    z = flow_reverse(W, z_p, a, calib, stages, g=g)
    if stages is not None:
        stages.update({"x": x, "m_p": m_p, "logs_p": logs_p, "logw": logw, "w": w,
                       "w_ceil": w_ceil, "y_len": y_len, "z_p": z_p, "z": z, "tok": tok})
    return z


def decode(W, z, calib=None, stages=None, sid=None):
    """Streaming 'decoder.onnx' half: z[1,I,T] -> wav[1,1,256*T] (piper/src/lib.rs:736-762; multi-speaker voices also
    pass the encoder's `g` output, :706-735)."""
    return decoder(W, z, arch_of(W), calib, stages, g=speaker_embedding(W, sid))


def infer(W, ids, scales, eps_w=None, eps_z=None, calib=None, stages=None, sid=None):
    """ids (list of i64), scales f32[3] -> waveform float array, as read from ``outputs[0]``
    at piper/src/lib.rs:382-392."""
    with torch.inference_mode(calib is None):
        z = encode(W, ids, scales, eps_w, eps_z, calib, stages, sid=sid)
        wav = decode(W, z, calib, stages, sid=sid)
        if stages is not None:
            stages["wav"] = wav
    return wav.reshape(-1)


def synthetic_ids(n_phonemes: int, utt: int = 0, num_symbols: int = 256, seed: int = 20260921):
    """rng = PCG64(seed + utt); ids ~ U{3..num_symbols-1}, interleaved with pad 0,
    wrapped in bos 1 / eos 2 (the layout of piper/src/lib.rs:232-250) -> 2N+2 ids."""
    r = np.random.Generator(np.random.PCG64(seed + utt))
    ph = r.integers(3, num_symbols, size=n_phonemes)
    ids = np.zeros(2 * n_phonemes + 2, dtype=np.int64)
    ids[0] = 1
    ids[1:-1:2] = ph
    ids[-1] = 2
    return ids
