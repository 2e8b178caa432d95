"""ORACLE (test infrastructure, not product): numpy restatement of the reference's shifted-window
attention path, file /root/reference/dl_cs/models/video_swin_transformer_mri_downsample.py:

  window_partition / window_reverse        :41-52 / :55-67
  get_window_size                          :72-85
  WindowAttention3D.__init__ (index)       :115-129   relative_position_index
  WindowAttention3D.forward                :139-170   softmax(q*scale @ k^T + bias (+ mask)) @ v, proj
  SwinTransformerBlock3D.forward_part1     :215-249   LN -> pad -> roll -> partition -> attn -> reverse -> roll -> crop
  compute_mask                             :342-355   region ids -> {0, -100} mask per window
  Mlp / forward_part2 / block forward      :20-38 / :251-252 / :254-273   x + part1, then + fc2(GELU(fc1(LN(x))))

Third-party arithmetic delegated by the reference and absent from /root/reference: torch matmul / softmax /
LayerNorm (eps 1e-5, biased variance) and timm's DropPath (identity in eval) — restated here from their
published definitions in float64.

Parity pin: tests/golden/swin_*.npz are outputs of the unmodified reference modules (oracle/make_golden_swin.py,
run in the build container with a 2-symbol timm shim); tests/test_oracle_golden.py checks this file against them.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def window_partition(x, ws):
    """(B, D, H, W, C) -> (B*nW, ws0*ws1*ws2, C)   (:41-52)"""
    B, D, H, W, C = x.shape
    x = x.reshape(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.transpose(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, ws[0] * ws[1] * ws[2], C)


def window_reverse(windows, ws, B, D, H, W):
    """(B*nW, ws0, ws1, ws2, C) -> (B, D, H, W, C)   (:55-67)"""
    x = windows.reshape(B, D // ws[0], H // ws[1], W // ws[2], ws[0], ws[1], ws[2], -1)
    return x.transpose(0, 1, 4, 2, 5, 3, 6, 7).reshape(B, D, H, W, -1)


def get_window_size(x_size, window_size, shift_size=None):
    """(:72-85) clamp window (and zero the shift) on axes not larger than the window."""
    use_ws = list(window_size)
    use_ss = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            use_ws[i] = x_size[i]
            if use_ss is not None:
                use_ss[i] = 0
    if shift_size is None:
        return tuple(use_ws)
    return tuple(use_ws), tuple(use_ss)


def relative_position_index(ws):
    """(:115-129) pair-wise index into the (2Wd-1)(2Wh-1)(2Ww-1) bias table."""
    coords = np.stack(np.meshgrid(np.arange(ws[0]), np.arange(ws[1]), np.arange(ws[2]), indexing="ij"))
    flat = coords.reshape(3, -1)
    rel = (flat[:, :, None] - flat[:, None, :]).transpose(1, 2, 0).copy()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 2] += ws[2] - 1
    rel[:, :, 0] *= (2 * ws[1] - 1) * (2 * ws[2] - 1)
    rel[:, :, 1] *= (2 * ws[2] - 1)
    return rel.sum(-1)


def compute_mask(D, H, W, ws, ss):
    """(:342-355) (nW, N, N) of {0, -100}; note slice(-0, None) covers the whole axis when the shift is 0."""
    img = np.zeros((1, D, H, W, 1))
    cnt = 0
    for d in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for h in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for w in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, d, h, w, :] = cnt
                cnt += 1
    mw = window_partition(img, ws)[..., 0]
    diff = mw[:, None, :] - mw[:, :, None]
    return np.where(diff != 0, -100.0, 0.0)


def softmax(a, axis=-1):
    a = a - a.max(axis=axis, keepdims=True)
    e = np.exp(a)
    return e / e.sum(axis=axis, keepdims=True)


def layer_norm(x, weight, bias, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * weight + bias


def window_attention(x, p, num_heads, window_size, mask=None, qk_scale=None):
    """WindowAttention3D.forward (:139-170).  x (B_, N, C); p: dict of the module's state_dict arrays."""
    B_, N, C = x.shape
    hd = C // num_heads
    scale = qk_scale or hd ** -0.5
    qkv = x @ p["qkv.weight"].T
    if "qkv.bias" in p and p["qkv.bias"] is not None:
        qkv = qkv + p["qkv.bias"]
    qkv = qkv.reshape(B_, N, 3, num_heads, hd).transpose(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    attn = q @ k.transpose(0, 1, 3, 2)
    idx = p.get("relative_position_index")
    if idx is None:
        idx = relative_position_index(window_size)
    bias = p["relative_position_bias_table"][idx[:N, :N].reshape(-1)].reshape(N, N, -1).transpose(2, 0, 1)
    attn = attn + bias[None]
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.reshape(B_ // nW, nW, num_heads, N, N) + mask[None, :, None]
        attn = attn.reshape(-1, num_heads, N, N)
    attn = softmax(attn, -1)
    out = (attn @ v).transpose(0, 2, 1, 3).reshape(B_, N, C)
    return out @ p["proj.weight"].T + p["proj.bias"]


def block_part1(x, p, num_heads, window_size, shift_size, mask_matrix=None):
    """SwinTransformerBlock3D.forward_part1 (:215-249).  x (B, D, H, W, C); p holds norm1.* and attn.*."""
    B, D, H, W, C = x.shape
    ws, ss = get_window_size((D, H, W), window_size, shift_size)
    x = layer_norm(x, p["norm1.weight"], p["norm1.bias"])
    pd = (ws[0] - D % ws[0]) % ws[0]
    pb = (ws[1] - H % ws[1]) % ws[1]
    pr = (ws[2] - W % ws[2]) % ws[2]
    x = np.pad(x, ((0, 0), (0, pd), (0, pb), (0, pr), (0, 0)))      # zero pad AFTER the norm (:219-225)
    _, Dp, Hp, Wp, _ = x.shape
    shifted = any(i > 0 for i in ss)
    if shifted:
        x = np.roll(x, shift=(-ss[0], -ss[1], -ss[2]), axis=(1, 2, 3))
        if mask_matrix is None:
            mask_matrix = compute_mask(Dp, Hp, Wp, ws, ss)
        mask = mask_matrix
    else:
        mask = None
    xw = window_partition(x, ws)
    ap = {k[len("attn."):]: v for k, v in p.items() if k.startswith("attn.")}
    # the bias index is the MODULE's (un-clamped) window index sliced to [:N, :N] (:152), even when the window was clamped
    aw = window_attention(xw, ap, num_heads, window_size, mask)
    aw = aw.reshape((-1,) + tuple(ws) + (C,))
    x = window_reverse(aw, ws, B, Dp, Hp, Wp)
    if shifted:
        x = np.roll(x, shift=(ss[0], ss[1], ss[2]), axis=(1, 2, 3))
    return x[:, :D, :H, :W, :]


def gelu(x):
    """nn.GELU default (exact): 0.5 x (1 + erf(x / sqrt 2))."""
    from math import erf
    return 0.5 * x * (1.0 + np.vectorize(erf)(x / np.sqrt(2.0)))


def block_part2(x, p):
    """SwinTransformerBlock3D.forward_part2 (:251-252): Mlp(norm2(x)), Mlp = fc2(GELU(fc1(.))) (:20-38); dropouts are p = 0."""
    h = layer_norm(x, p["norm2.weight"], p["norm2.bias"])
    h = gelu(h @ p["mlp.fc1.weight"].T + p["mlp.fc1.bias"])
    return h @ p["mlp.fc2.weight"].T + p["mlp.fc2.bias"]


def block_forward(x, p, num_heads, window_size, shift_size, mask_matrix=None):
    """SwinTransformerBlock3D.forward (:254-273) in eval mode (DropPath is the identity)."""
    x = x + block_part1(x, p, num_heads, window_size, shift_size, mask_matrix)
    return x + block_part2(x, p)
