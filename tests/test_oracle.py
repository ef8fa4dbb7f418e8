"""CPU: the oracle against its committed golden vectors and against independent restatements of
the pieces whose upstream definition is a different formula (dense generate_path matmul, naive
relative attention, forward spline, dense ConvTranspose)."""
import glob
import math
import os
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vits_oracle as vo
from sonata_b200 import voicegen

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _crc(t):
    c = 0
    for k in sorted(t):
        c = zlib.crc32(np.ascontiguousarray(t[k]).tobytes(), c)
    return c


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_reproduces_golden(path, oracle_weights):
    g = np.load(path)
    q = os.path.basename(path).split("_")[0]
    assert _crc(voicegen.make_tensors(q)) == int(g["weights_crc"]), "synthetic voice generator drifted"
    W = oracle_weights(q)
    st = {}
    ew = g["eps_w"].T[None] if "eps_w" in g else None
    ez = g["eps_z"].T[None] if "eps_z" in g else None
    wav = vo.infer(W, g["ids"], g["scales"], eps_w=None if ew is None else torch.from_numpy(ew.copy()),
                   eps_z=None if ez is None else torch.from_numpy(ez.copy()), stages=st)
    assert st["y_len"] == int(g["y_len"])
    assert np.array_equal(np.cumsum(st["w_ceil"].view(-1).numpy()).astype(np.int32), g["cum"])
    assert wav.numel() == 256 * int(g["y_len"])          # hop pinned at piper/src/lib.rs:910
    assert float(np.abs(wav.numpy() - g["wav"]).max()) < 2e-5


def test_fp64_shadow_close(oracle_weights):
    t = voicegen.make_tensors("medium")
    W32, W64 = oracle_weights("medium"), vo.to_torch(t, torch.float64)
    ids = vo.synthetic_ids(10, utt=11)
    s32, s64 = {}, {}
    w32 = vo.infer(W32, ids, [0, 1, 0], stages=s32)
    w64 = vo.infer(W64, ids, [0, 1, 0], stages=s64)
    assert torch.equal(s32["w_ceil"].double(), s64["w_ceil"])
    assert float((w32.double() - w64).abs().max()) < 2e-5


def test_expand_equals_dense_generate_path():
    """commons.generate_path + attn^T matmul (the reference graph's dense form) == our gather."""
    torch.manual_seed(0)
    T, C = 13, 6
    w_ceil = torch.randint(0, 5, (1, 1, T)).float()
    y_len = int(max(w_ceil.sum().item(), 1))
    m_p, logs_p = torch.randn(1, C, T), torch.randn(1, C, T)
    cum = torch.cumsum(w_ceil.view(-1), 0)
    j = torch.arange(y_len)
    path = (j[None, :] < cum[:, None]).float()            # sequence_mask(cum_duration, t_y)  [T, y]
    path = path - F.pad(path, [0, 0, 1, 0])[:-1]          # path - shifted path
    dense = torch.matmul(path.T, m_p[0].T).T[None]        # attn^T . m_p
    got, _ = vo.expand(m_p, logs_p, w_ceil, y_len, None, 0.0)
    assert torch.allclose(got, dense, atol=1e-6)


def test_relative_attention_matches_naive():
    a = dict(hidden=8, heads=2, window=2)
    torch.manual_seed(1)
    W = {}
    p = "x."
    for c in ("conv_q", "conv_k", "conv_v", "conv_o"):
        W[p + c + ".weight"] = torch.randn(8, 8, 1) * 0.3
        W[p + c + ".bias"] = torch.randn(8) * 0.1
    W[p + "emb_rel_k"] = torch.randn(1, 5, 4)
    W[p + "emb_rel_v"] = torch.randn(1, 5, 4)
    for T in (1, 2, 3, 7):
        x = torch.randn(1, 8, T)
        got = vo._mha(W, p, x, a)
        q = F.conv1d(x, W[p + "conv_q.weight"], W[p + "conv_q.bias"])[0].view(2, 4, T)
        k = F.conv1d(x, W[p + "conv_k.weight"], W[p + "conv_k.bias"])[0].view(2, 4, T)
        v = F.conv1d(x, W[p + "conv_v.weight"], W[p + "conv_v.bias"])[0].view(2, 4, T)
        out = torch.zeros(2, 4, T)
        for h in range(2):
            for i in range(T):
                qi = q[h, :, i] / 2.0
                s = torch.stack([qi @ k[h, :, j] + (qi @ W[p + "emb_rel_k"][0, j - i + 2] if abs(j - i) <= 2 else 0.0)
                                 for j in range(T)])
                pr = torch.softmax(s, 0)
                o = sum(pr[j] * v[h, :, j] for j in range(T))
                o = o + sum(pr[j] * W[p + "emb_rel_v"][0, j - i + 2] for j in range(T) if abs(j - i) <= 2)
                out[h, :, i] = o
        ref = F.conv1d(out.view(1, 8, T), W[p + "conv_o.weight"], W[p + "conv_o.bias"])
        assert torch.allclose(got, ref, atol=1e-5), T


def _rqs_forward(x, uw, uh, ud, B=5.0):
    """transforms.rational_quadratic_spline(inverse=False), scalar restatement."""
    nb = len(uw)
    w = torch.softmax(uw, 0); w = 1e-3 + (1 - 1e-3 * nb) * w
    cw = F.pad(torch.cumsum(w, 0), (1, 0)); cw = 2 * B * cw - B; cw[0], cw[-1] = -B, B
    h = torch.softmax(uh, 0); h = 1e-3 + (1 - 1e-3 * nb) * h
    ch = F.pad(torch.cumsum(h, 0), (1, 0)); ch = 2 * B * ch - B; ch[0], ch[-1] = -B, B
    c = math.log(math.exp(1 - 1e-3) - 1)
    d = 1e-3 + F.softplus(torch.cat([torch.tensor([c]), ud, torch.tensor([c])]))
    k = int(torch.sum(x >= cw[:-1]).item()) - 1
    k = min(max(k, 0), nb - 1)
    W_, H_ = cw[k + 1] - cw[k], ch[k + 1] - ch[k]
    delta = H_ / W_
    th = (x - cw[k]) / W_
    num = H_ * (delta * th ** 2 + d[k] * th * (1 - th))
    den = delta + (d[k] + d[k + 1] - 2 * delta) * th * (1 - th)
    return ch[k] + num / den


def test_spline_inverse_inverts_forward():
    torch.manual_seed(3)
    for _ in range(50):
        uw, uh, ud = torch.randn(10) * 1.5, torch.randn(10) * 1.5, torch.randn(9)
        x = torch.rand(()) * 9.0 - 4.5
        y = _rqs_forward(x.double(), uw.double(), uh.double(), ud.double())
        xr = vo._rqs_inverse(y.view(1, 1, 1), uw.double().view(1, 1, 1, 10), uh.double().view(1, 1, 1, 10),
                             ud.double().view(1, 1, 1, 9))
        assert abs(float(xr) - float(x)) < 1e-6
    # identity outside the tails
    out = vo._rqs_inverse(torch.tensor([[[7.5]]]), torch.zeros(1, 1, 1, 10), torch.zeros(1, 1, 1, 10), torch.zeros(1, 1, 1, 9))
    assert float(out) == 7.5


def test_flow_flip_folding_identity(oracle_weights):
    """The CUDA path folds the channel flips of the coupling block into its weights (even number of
    flips).  Check the algebra on the oracle: flips + plain weights == no flips + permuted weights."""
    W = oracle_weights("medium")
    a = vo.arch_of(W)
    torch.manual_seed(5)
    z = torch.randn(1, a["inter"], 9)
    ref = vo.flow_reverse(W, z.clone(), a)
    half = a["inter"] // 2
    s = z.clone()
    for step in range(a["flow_n"]):
        f = a["flow_n"] - 1 - step
        p = f"flow.flows.{2 * f}."
        rev = step % 2 == 0
        Wl = dict(W)
        if rev:
            Wl[p + "pre.weight"] = W[p + "pre.weight"].flip(1)
            Wl[p + "post.weight"] = W[p + "post.weight"].flip(0)
            Wl[p + "post.bias"] = W[p + "post.bias"].flip(0)
        cond = s[:, half:] if rev else s[:, :half]
        h = vo._conv(Wl, p + "pre", cond)
        h = vo._wn(Wl, p + "enc.", h, a)
        m = vo._conv(Wl, p + "post", h)
        if rev:
            s = torch.cat([s[:, :half] - m, s[:, half:]], 1)
        else:
            s = torch.cat([s[:, :half], s[:, half:] - m], 1)
    assert torch.allclose(s, ref, atol=1e-5)


def test_polyphase_equals_conv_transpose():
    """The CUDA path runs ConvTranspose1d(k, stride u, pad (k-u)/2) as u phase convolutions."""
    torch.manual_seed(7)
    for (u, k) in ((8, 16), (4, 8), (2, 4)):
        cin, cout, T = 4, 3, 11
        x = torch.randn(1, cin, T); w = torch.randn(cin, cout, k); b = torch.randn(cout)
        ref = F.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
        out = torch.zeros(1, cout, T * u)
        pad = (k - u) // 2
        for p in range(u):
            pp = p + pad
            for d in range(-k, k + 1):
                kk = d * u + pp
                if 0 <= kk < k:
                    for q in range(T):
                        i = q - d
                        if 0 <= i < T:
                            out[0, :, q * u + p] += w[:, :, kk].T @ x[0, :, i]
        out += b[None, :, None]
        assert torch.allclose(out, ref, atol=1e-5)


def test_synthetic_ids_layout():
    ids = vo.synthetic_ids(5, utt=0)
    assert len(ids) == 12 and ids[0] == 1 and ids[-1] == 2
    assert all(ids[2:-1:2] == 0) and all(ids[1:-1:2] >= 3)
    from sonata_b200 import workload
    assert np.array_equal(ids, workload.synthetic_ids(5, utt=0))
