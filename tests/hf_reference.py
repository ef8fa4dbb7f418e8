"""An INDEPENDENT implementation of the published VITS algorithm as a second opinion on the oracle.

The reference's graph (a Piper voice executed by onnxruntime) cannot run offline, so `oracle/vits_oracle.py` is a
restatement of the published algorithm.  `transformers.models.vits.VitsModel` (Hugging Face transformers, the version
installed in this image; Apache-2.0, written from the same paper / original repository, NOT from Piper and not by us)
implements the same network: text encoder with windowed relative attention, stochastic duration predictor run in
reverse, ceil / cumsum alignment, residual-coupling flow in reverse, HiFi-GAN with ResBlock1.  That is exactly the
architecture of the Piper *high* quality voices (en_US-ryan-high: BASELINE config 3).  This module loads the SAME
synthetic voice tensors (Piper state-dict names, `sonata_b200.voicegen`) into a `VitsModel` and runs its own forward,
so that the oracle -- and the CUDA path -- can be compared against code we did not write.

What this pins: every stage of the high-quality voice.  What it cannot pin: the ResBlock2 wiring of the medium voices
(HF implements ResBlock1 only) and Piper-specific export details (both are covered by the oracle's own tests).

Test infrastructure only (imports torch + transformers).
"""
from __future__ import annotations

import math

import numpy as np
import torch

GIN_CHANNELS = 512      # Piper multi-speaker voices (sonata_b200.voicegen.GIN_CHANNELS)


def build_hf_model(a: dict, decoder: bool = True, n_speakers: int = 1):
    """`a`: an architecture dict of sonata_b200.voicegen.ARCH.  decoder=False builds a throw-away one-stage vocoder so that
    voices whose HiFi-GAN transformers cannot express (ResBlock2: the medium voices) can still be compared on everything
    in front of it through `VitsModelOutput.spectrogram` (= the flow output z)."""
    from transformers import VitsConfig, VitsModel
    if decoder:
        assert a["resblock"] == 1, "transformers' VITS implements ResBlock1 only"
    else:
        a = dict(a, up_init=16, up_rates=(2,), up_kernels=(4,), res_kernels=(3,), res_dils=((1,),))
    cfg = VitsConfig(
        vocab_size=a["n_vocab"], hidden_size=a["hidden"], num_hidden_layers=a["layers"],
        num_attention_heads=a["heads"], window_size=a["window"], use_bias=True, ffn_dim=a["filter"],
        ffn_kernel_size=a["kernel"], flow_size=a["inter"], hidden_act="relu", layerdrop=0.0,
        use_stochastic_duration_prediction=True, num_speakers=n_speakers,
        speaker_embedding_size=GIN_CHANNELS if n_speakers > 1 else 0,
        upsample_initial_channel=a["up_init"], upsample_rates=list(a["up_rates"]),
        upsample_kernel_sizes=list(a["up_kernels"]), resblock_kernel_sizes=list(a["res_kernels"]),
        resblock_dilation_sizes=[list(d) for d in a["res_dils"]], leaky_relu_slope=0.1,
        depth_separable_channels=2, depth_separable_num_layers=3, duration_predictor_flow_bins=a["dp_bins"],
        duration_predictor_tail_bound=5.0, duration_predictor_kernel_size=a["dp_kernel"],
        duration_predictor_num_flows=4, duration_predictor_filter_channels=a["hidden"],
        prior_encoder_num_flows=a["flow_n"], prior_encoder_num_wavenet_layers=a["wn_layers"],
        posterior_encoder_num_wavenet_layers=1, spectrogram_bins=4,      # training-only branch: kept tiny
        wavenet_kernel_size=a["flow_kernel"], wavenet_dilation_rate=1, sampling_rate=a["sample_rate"],
    )
    torch.manual_seed(0)
    return VitsModel(cfg).eval()


def _set(p: torch.nn.Parameter, v):
    v = torch.as_tensor(np.asarray(v), dtype=p.dtype)
    assert tuple(p.shape) == tuple(v.shape), (tuple(p.shape), tuple(v.shape))
    with torch.no_grad():
        p.copy_(v)


def _set_weight_normed(conv, w):
    """torch.nn.utils.parametrizations.weight_norm: weight = g * v / ||v|| (norm over all dims but 0)."""
    w = torch.as_tensor(np.asarray(w), dtype=torch.float32)
    par = conv.parametrizations.weight
    _set(par.original0, w.flatten(1).norm(dim=1).view(-1, 1, 1))
    _set(par.original1, w)


def load_piper_tensors(model, T: dict, a: dict, decoder: bool = True):
    """Copy a Piper-named tensor dict (sonata_b200.voicegen.make_tensors) into the HF model, name by name."""
    te = model.text_encoder
    _set(te.embed_tokens.weight, T["enc_p.emb.weight"])
    for i, lay in enumerate(te.encoder.layers):
        p = f"enc_p.encoder.attn_layers.{i}."
        for ours, theirs in (("conv_q", lay.attention.q_proj), ("conv_k", lay.attention.k_proj),
                             ("conv_v", lay.attention.v_proj), ("conv_o", lay.attention.out_proj)):
            _set(theirs.weight, np.asarray(T[p + ours + ".weight"])[:, :, 0])
            _set(theirs.bias, T[p + ours + ".bias"])
        _set(lay.attention.emb_rel_k, T[p + "emb_rel_k"])
        _set(lay.attention.emb_rel_v, T[p + "emb_rel_v"])
        _set(lay.layer_norm.weight, T[f"enc_p.encoder.norm_layers_1.{i}.gamma"])
        _set(lay.layer_norm.bias, T[f"enc_p.encoder.norm_layers_1.{i}.beta"])
        f = f"enc_p.encoder.ffn_layers.{i}."
        _set(lay.feed_forward.conv_1.weight, T[f + "conv_1.weight"]); _set(lay.feed_forward.conv_1.bias, T[f + "conv_1.bias"])
        _set(lay.feed_forward.conv_2.weight, T[f + "conv_2.weight"]); _set(lay.feed_forward.conv_2.bias, T[f + "conv_2.bias"])
        _set(lay.final_layer_norm.weight, T[f"enc_p.encoder.norm_layers_2.{i}.gamma"])
        _set(lay.final_layer_norm.bias, T[f"enc_p.encoder.norm_layers_2.{i}.beta"])
    _set(te.project.weight, T["enc_p.proj.weight"]); _set(te.project.bias, T["enc_p.proj.bias"])

    def dds(mod, p):
        for j in range(3):
            _set(mod.convs_dilated[j].weight, T[f"{p}convs_sep.{j}.weight"]); _set(mod.convs_dilated[j].bias, T[f"{p}convs_sep.{j}.bias"])
            _set(mod.convs_pointwise[j].weight, T[f"{p}convs_1x1.{j}.weight"]); _set(mod.convs_pointwise[j].bias, T[f"{p}convs_1x1.{j}.bias"])
            _set(mod.norms_1[j].weight, T[f"{p}norms_1.{j}.gamma"]); _set(mod.norms_1[j].bias, T[f"{p}norms_1.{j}.beta"])
            _set(mod.norms_2[j].weight, T[f"{p}norms_2.{j}.gamma"]); _set(mod.norms_2[j].bias, T[f"{p}norms_2.{j}.beta"])

    dp = model.duration_predictor
    _set(dp.conv_pre.weight, T["dp.pre.weight"]); _set(dp.conv_pre.bias, T["dp.pre.bias"])
    dds(dp.conv_dds, "dp.convs.")
    _set(dp.conv_proj.weight, T["dp.proj.weight"]); _set(dp.conv_proj.bias, T["dp.proj.bias"])
    _set(dp.flows[0].translate, T["dp.flows.0.m"]); _set(dp.flows[0].log_scale, T["dp.flows.0.logs"])
    # Piper's flow list interleaves Flip modules: ConvFlow j (1-based) sits at index 2j - 1; the first ConvFlow is
    # dropped at inference by both implementations (`flows[:-2] + [flows[-1]]`) and has no tensors in the voice
    for j in (2, 3, 4):
        p = f"dp.flows.{2 * j - 1}."
        fl = dp.flows[j]
        _set(fl.conv_pre.weight, T[p + "pre.weight"]); _set(fl.conv_pre.bias, T[p + "pre.bias"])
        dds(fl.conv_dds, p + "convs.")
        _set(fl.conv_proj.weight, T[p + "proj.weight"]); _set(fl.conv_proj.bias, T[p + "proj.bias"])

    for f, fl in enumerate(model.flow.flows):
        p = f"flow.flows.{2 * f}."
        _set(fl.conv_pre.weight, T[p + "pre.weight"]); _set(fl.conv_pre.bias, T[p + "pre.bias"])
        for l in range(a["wn_layers"]):
            _set_weight_normed(fl.wavenet.in_layers[l], T[p + f"enc.in_layers.{l}.weight"])
            _set(fl.wavenet.in_layers[l].bias, T[p + f"enc.in_layers.{l}.bias"])
            _set_weight_normed(fl.wavenet.res_skip_layers[l], T[p + f"enc.res_skip_layers.{l}.weight"])
            _set(fl.wavenet.res_skip_layers[l].bias, T[p + f"enc.res_skip_layers.{l}.bias"])
        _set(fl.conv_post.weight, T[p + "post.weight"]); _set(fl.conv_post.bias, T[p + "post.bias"])

    if "emb_g.weight" in T:
        # multi-speaker voices: g = emb_g(sid) enters through 1x1 convs in the duration predictor, every WaveNet and the vocoder
        _set(model.embed_speaker.weight, T["emb_g.weight"])
        _set(dp.cond.weight, T["dp.cond.weight"]); _set(dp.cond.bias, T["dp.cond.bias"])
        for f, fl in enumerate(model.flow.flows):
            p = f"flow.flows.{2 * f}.enc.cond_layer."
            _set_weight_normed(fl.wavenet.cond_layer, T[p + "weight"]); _set(fl.wavenet.cond_layer.bias, T[p + "bias"])
        if decoder:
            _set(model.decoder.cond.weight, T["dec.cond.weight"]); _set(model.decoder.cond.bias, T["dec.cond.bias"])
    if not decoder:
        return model
    dec = model.decoder
    _set(dec.conv_pre.weight, T["dec.conv_pre.weight"]); _set(dec.conv_pre.bias, T["dec.conv_pre.bias"])
    nk = len(a["res_kernels"])
    for i in range(len(a["up_rates"])):
        _set(dec.upsampler[i].weight, T[f"dec.ups.{i}.weight"]); _set(dec.upsampler[i].bias, T[f"dec.ups.{i}.bias"])
        for j in range(nk):
            rb = dec.resblocks[i * nk + j]
            p = f"dec.resblocks.{i * nk + j}."
            for m in range(len(a["res_dils"][j])):
                _set(rb.convs1[m].weight, T[p + f"convs1.{m}.weight"]); _set(rb.convs1[m].bias, T[p + f"convs1.{m}.bias"])
                _set(rb.convs2[m].weight, T[p + f"convs2.{m}.weight"]); _set(rb.convs2[m].bias, T[p + f"convs2.{m}.bias"])
    _set(dec.conv_post.weight, T["dec.conv_post.weight"])
    return model


def hf_infer(model, ids, noise_scale: float, length_scale: float, noise_w: float, seed: int = 0, speaker_id=None):
    """Runs `VitsModel.forward` itself.  Returns (waveform f32[n], eps_w f32[2, T], eps_z f32[inter, frames]): the two
    Gaussian draws the forward made (duration-predictor latents first, then `randn_like(prior_means)`), re-drawn from
    the same generator state so that another implementation can be fed the identical noise."""
    ids = torch.as_tensor(np.asarray(ids, dtype=np.int64)).view(1, -1)
    model.noise_scale = float(noise_scale)
    model.noise_scale_duration = float(noise_w)
    model.speaking_rate = 1.0 / float(length_scale)
    torch.manual_seed(seed)
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=torch.ones_like(ids), speaker_id=speaker_id)
    wav = out.waveform[0].float().numpy().copy()
    hf_infer.last_spectrogram = out.spectrogram[0].float().numpy().copy()      # [inter, frames]: the flow output z
    frames = wav.shape[0] // int(np.prod(model.config.upsample_rates))
    torch.manual_seed(seed)
    eps_w = torch.randn(1, 2, ids.shape[1])
    # `randn_like(prior_means)` fills a TRANSPOSED view ([1, frames, inter] in memory), which also takes torch's
    # scalar (non-vectorised) normal path: draw through the same kind of view to get the same values
    eps_z = torch.randn_like(torch.empty(1, frames, model.config.flow_size).transpose(1, 2))
    return wav, eps_w[0].numpy().copy(), eps_z[0].contiguous().numpy().copy()
