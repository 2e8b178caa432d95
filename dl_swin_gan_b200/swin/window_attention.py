"""Drop-in mirror of the shifted-window attention half of
dl_cs/models/video_swin_transformer_mri_downsample.py on the B200 kernels.

Same names and semantics as the reference:
    window_partition :41-52, window_reverse :55-67, get_window_size :72-85, compute_mask :342-355,
    WindowAttention3D :88-170 (identical parameter / buffer names and shapes, so `load_state_dict` of a
    reference checkpoint works), SwinTransformerBlock3D :173-273 (forward_part1 fused), BasicLayer :358-437.
The score matrix, the relative-position-bias gather and the shift mask never reach HBM (csrc/swin_kernels.cuh);
the qkv / proj / MLP Linear layers stay library GEMMs.  CUDA only: CPU tensors raise.

Training: the attention core has a hand-written forward; its backward re-materialises the probabilities with
torch ops one call at a time (correct, not yet fast — SURVEY.md §7 step 5 "backward second").
"""
from __future__ import annotations

import ctypes
from functools import reduce
from operator import mul
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint
from torch import nn

from .. import _lib

MASK_NONE, MASK_REGION, MASK_DENSE = 0, 1, 2


def _i3(t):
    return (ctypes.c_int * 3)(*[int(v) for v in t])


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: dl_swin_gan_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32 (got {t.dtype})")


# ----------------------------------------------------------------------------- layout helpers (pure views / indexing)
def window_partition(x, window_size):
    """(B, D, H, W, C) -> (B*num_windows, Wd*Wh*Ww, C)"""
    B, D, H, W, C = x.shape
    x = x.view(B, D // window_size[0], window_size[0], H // window_size[1], window_size[1],
               W // window_size[2], window_size[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, reduce(mul, window_size), C)


def window_reverse(windows, window_size, B, D, H, W):
    """(B*num_windows, Wd, Wh, Ww, C) -> (B, D, H, W, C)"""
    x = windows.view(B, D // window_size[0], H // window_size[1], W // window_size[2],
                     window_size[0], window_size[1], window_size[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(B, D, H, W, -1)


def get_window_size(x_size, window_size, shift_size=None):
    use_window_size = list(window_size)
    if shift_size is not None:
        use_shift_size = list(shift_size)
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            use_window_size[i] = x_size[i]
            if shift_size is not None:
                use_shift_size[i] = 0
    if shift_size is None:
        return tuple(use_window_size)
    return tuple(use_window_size), tuple(use_shift_size)


_MASK_CACHE = {}


def _region_image(D, H, W, window_size, shift_size):
    """Region-id image of compute_mask (:343-350); slice(-0, None) covers the whole axis when a shift is 0."""
    img = np.zeros((D, H, W), dtype=np.int64)
    cnt = 0
    for d in (slice(-window_size[0]), slice(-window_size[0], -shift_size[0]), slice(-shift_size[0], None)):
        for h in (slice(-window_size[1]), slice(-window_size[1], -shift_size[1]), slice(-shift_size[1], None)):
            for w in (slice(-window_size[2]), slice(-window_size[2], -shift_size[2]), slice(-shift_size[2], None)):
                img[d, h, w] = cnt
                cnt += 1
    return img


def compute_mask(D, H, W, window_size, shift_size, device):
    """Same result as the reference ((nW, N, N) of {0, -100}), cached per geometry like its lru_cache.
    The per-token region ids the kernels use instead of this 24 MB tensor are cached alongside."""
    key = (D, H, W, tuple(window_size), tuple(shift_size), str(device))
    hit = _MASK_CACHE.get(key)
    if hit is None:
        img = torch.from_numpy(_region_image(D, H, W, window_size, shift_size)).to(device)
        mw = window_partition(img.view(1, D, H, W, 1), window_size).squeeze(-1)          # nW, N
        attn_mask = (mw.unsqueeze(1) - mw.unsqueeze(2)).to(torch.float32)
        attn_mask = attn_mask.masked_fill(attn_mask != 0, float(-100.0)).masked_fill(attn_mask == 0, float(0.0))
        hit = (attn_mask, mw.to(torch.uint8).contiguous())
        _MASK_CACHE[key] = hit
    return hit[0]


def _region_ids_for(mask_matrix, D, H, W, window_size, shift_size, device):
    """The cached region ids iff mask_matrix is the tensor compute_mask handed out for this geometry."""
    hit = _MASK_CACHE.get((D, H, W, tuple(window_size), tuple(shift_size), str(device)))
    if hit is not None and hit[0] is mask_matrix:
        return hit[1]
    return None


# ----------------------------------------------------------------------------- tensor-core linear (tcgen05 + TMEM)
def linear_tc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """torch.nn.functional.linear on the tcgen05 kernel: fp16 operands, fp32 accumulation.  x [..., K] fp32 CUDA."""
    _require_cuda_f32(x, "linear_tc input")
    lib = _lib.load()
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.contiguous().view(-1, K)
    y = torch.empty((x2.shape[0], N), dtype=torch.float32, device=x.device)
    if y.numel() == 0:
        return y.view(*x.shape[:-1], N)
    w = weight.detach().contiguous()
    b = bias.detach().contiguous() if bias is not None else None
    with torch.cuda.device(x.device):
        st = lib.dlsg_swin_linear_f32(x2.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                      y.data_ptr(), x2.shape[0], N, K, _stream())
    _lib.check(st, "dlsg_swin_linear")
    return y.view(*x.shape[:-1], N)


# ----------------------------------------------------------------------------- attention core
def _attention_core_raw(qkv, table, num_heads, module_window, scale, mask_mode, mask, nW):
    lib = _lib.load()
    B_, N, C3 = qkv.shape
    C = C3 // 3
    out = torch.empty((B_, N, C), dtype=torch.float32, device=qkv.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(qkv.device):
        st = lib.dlsg_swin_window_attention_f32(
            qkv.data_ptr(), table.data_ptr(), out.data_ptr(), B_, N, num_heads, C // num_heads,
            _i3(module_window), float(scale), int(mask_mode), mask.data_ptr() if mask is not None else None,
            int(nW), _stream())
    _lib.check(st, "dlsg_swin_window_attention")
    return out


def _attention_core_torch(qkv, table, index, num_heads, scale, mask):
    """Reference arithmetic (:147-167) in torch ops; used only to differentiate the core in backward."""
    B_, N, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv.reshape(B_, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    attn = (q * scale) @ k.transpose(-2, -1)
    bias = table[index[:N, :N].reshape(-1)].reshape(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, num_heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    return (attn @ v).transpose(1, 2).reshape(B_, N, C)


class _AttnCoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, table, index, num_heads, module_window, scale, mask_mode, mask, dense_mask):
        ctx.cfg = (num_heads, scale)
        ctx.save_for_backward(qkv, table, index, dense_mask if dense_mask is not None else torch.empty(0))
        ctx.has_mask = dense_mask is not None
        nW = mask.shape[0] if mask is not None else 1
        return _attention_core_raw(qkv.contiguous(), table.detach().contiguous(), num_heads, module_window, scale,
                                   mask_mode, mask, nW)

    @staticmethod
    def backward(ctx, grad):
        qkv, table, index, dense_mask = ctx.saved_tensors
        num_heads, scale = ctx.cfg
        with torch.enable_grad():
            q_ = qkv.detach().requires_grad_(True)
            t_ = table.detach().requires_grad_(True)
            out = _attention_core_torch(q_, t_, index, num_heads, scale, dense_mask if ctx.has_mask else None)
            gq, gt = torch.autograd.grad(out, (q_, t_), grad)
        return gq, gt, None, None, None, None, None, None, None


class WindowAttention3D(nn.Module):
    """Window based multi-head self attention (W-MSA) with relative position bias; shifted and non-shifted windows.

    Constructor and forward signature as the reference (:101, :139); `attn_drop` / `proj_drop` must be 0 at
    inference time (they are 0 in every reference config; non-zero values are honoured in training through torch ops).
    """

    def __init__(self, dim, window_size, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.window_size = window_size  # Wd, Wh, Ww
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5

        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1), num_heads))

        coords = torch.stack(torch.meshgrid(torch.arange(window_size[0]), torch.arange(window_size[1]),
                                            torch.arange(window_size[2]), indexing="ij"))
        coords_flatten = torch.flatten(coords, 1)
        relative_coords = (coords_flatten[:, :, None] - coords_flatten[:, None, :]).permute(1, 2, 0).contiguous()
        relative_coords[:, :, 0] += window_size[0] - 1
        relative_coords[:, :, 1] += window_size[1] - 1
        relative_coords[:, :, 2] += window_size[2] - 1
        relative_coords[:, :, 0] *= (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
        relative_coords[:, :, 1] *= (2 * window_size[2] - 1)
        self.register_buffer("relative_position_index", relative_coords.sum(-1))

        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x, mask=None, _region_ids=None):
        """x: (num_windows*B, N, C); mask: (num_windows, N, N) additive or None.
        `_region_ids` (uint8 [nW, N]) is the compact equivalent the fused block passes for compute_mask masks."""
        _require_cuda_f32(x, "WindowAttention3D input")
        if self.training and self.attn_drop.p > 0:
            raise NotImplementedError("attn_drop > 0 in training is not supported by the fused attention core")
        qkv = self._linear(self.qkv, x)                     # (B_, N, 3C): the kernel reads q/k/v in place
        if qkv.dtype != torch.float32:                      # under torch.autocast the Linear returns bf16 / fp16: the core computes in fp32
            qkv = qkv.float()
        if _region_ids is not None:
            mode, m, dense = MASK_REGION, _region_ids, mask
        elif mask is not None:
            _require_cuda_f32(mask, "attention mask")
            mode, m, dense = MASK_DENSE, mask.contiguous(), mask
        else:
            mode, m, dense = MASK_NONE, None, None
        out = _AttnCoreFn.apply(qkv, self.relative_position_bias_table, self.relative_position_index,
                                self.num_heads, tuple(self.window_size), self.scale, mode, m, dense)
        out = self._linear(self.proj, out)
        return self.proj_drop(out)

    tc_linear = True    # qkv / proj on the tcgen05 kernel when no gradient is needed (fp16 operands, fp32 accumulate)

    def _linear(self, layer: nn.Linear, x: torch.Tensor) -> torch.Tensor:
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or layer.weight.requires_grad)
        K, N = layer.in_features, layer.out_features
        if (self.tc_linear and not needs_grad and K % 16 == 0 and N % 32 == 0 and N <= 512
                and (N <= 256 or (N // 2) % 16 == 0) and N * K * 2 + max(256 * K, 33792) + 4 * N + 3328 <= 227 * 1024):
            return linear_tc(x, layer.weight, layer.bias)
        return layer(x)


# ----------------------------------------------------------------------------- fused first half of the block
def _ln_partition(x, norm, window_size, shift_size):
    lib = _lib.load()
    B, D, H, W, C = x.shape
    Dp, Hp, Wp = (-(-n // w) * w for n, w in zip((D, H, W), window_size))
    nW = (Dp // window_size[0]) * (Hp // window_size[1]) * (Wp // window_size[2])
    N = window_size[0] * window_size[1] * window_size[2]
    xw = torch.empty((B * nW, N, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = lib.dlsg_swin_ln_partition_f32(
            x.data_ptr(), norm.weight.data_ptr() if norm.weight is not None else None,
            norm.bias.data_ptr() if norm.bias is not None else None, float(norm.eps), xw.data_ptr(),
            B, D, H, W, C, _i3(window_size), _i3(shift_size), _stream())
    _lib.check(st, "dlsg_swin_ln_partition")
    return xw


def _window_reverse_fused(yw, B, D, H, W, window_size, shift_size):
    lib = _lib.load()
    C = yw.shape[-1]
    y = torch.empty((B, D, H, W, C), dtype=torch.float32, device=yw.device)
    with torch.cuda.device(yw.device):
        st = lib.dlsg_swin_window_reverse_f32(yw.data_ptr(), y.data_ptr(), B, D, H, W, C,
                                              _i3(window_size), _i3(shift_size), _stream())
    _lib.check(st, "dlsg_swin_window_reverse")
    return y


class DropPath(nn.Module):
    """Stochastic depth per sample (what the reference imports from timm, :11); identity in eval."""

    def __init__(self, drop_prob: float = 0.):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class SwinTransformerBlock3D(nn.Module):
    """Swin Transformer block (:173-273) with forward_part1 on the fused kernels."""

    def __init__(self, dim, num_heads, window_size=(2, 7, 7), shift_size=(0, 0, 0), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 use_checkpoint=False):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.window_size, self.shift_size = window_size, shift_size
        self.mlp_ratio, self.use_checkpoint = mlp_ratio, use_checkpoint
        for i in range(3):
            assert 0 <= self.shift_size[i] < self.window_size[i], "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention3D(dim, window_size=self.window_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                      qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward_part1(self, x, mask_matrix):
        B, D, H, W, C = x.shape
        window_size, shift_size = get_window_size((D, H, W), self.window_size, self.shift_size)
        shifted = any(i > 0 for i in shift_size)
        Dp, Hp, Wp = (-(-n // w) * w for n, w in zip((D, H, W), window_size))
        fused = isinstance(self.norm1, nn.LayerNorm) and self._inference(x)
        if fused:
            _require_cuda_f32(x, "SwinTransformerBlock3D input")
            rid = _region_ids_for(mask_matrix, Dp, Hp, Wp, window_size, shift_size, x.device) if shifted else None
            y = self._part1_tensor_core(x.contiguous(), window_size, shift_size,
                                        mask_matrix if shifted else None, rid)
            if y is not None:
                return y
            x_windows = _ln_partition(x.contiguous(), self.norm1, window_size, shift_size)
        else:                                                  # differentiable route: same ops as the reference
            xn = self.norm1(x)
            xn = F.pad(xn, (0, 0, 0, Wp - W, 0, Hp - H, 0, Dp - D))
            if shifted:
                xn = torch.roll(xn, shifts=(-shift_size[0], -shift_size[1], -shift_size[2]), dims=(1, 2, 3))
            x_windows = window_partition(xn, window_size)
        if shifted:
            rid = _region_ids_for(mask_matrix, Dp, Hp, Wp, window_size, shift_size, x.device)
            attn_windows = self.attn(x_windows, mask=mask_matrix, _region_ids=rid)
        else:
            attn_windows = self.attn(x_windows, mask=None)
        if fused:
            return _window_reverse_fused(attn_windows.contiguous(), B, D, H, W, window_size, shift_size)
        attn_windows = attn_windows.view(-1, *(window_size + (C,)))
        shifted_x = window_reverse(attn_windows, window_size, B, Dp, Hp, Wp)
        if shifted:
            shifted_x = torch.roll(shifted_x, shifts=(shift_size[0], shift_size[1], shift_size[2]), dims=(1, 2, 3))
        return shifted_x[:, :D, :H, :W, :].contiguous()

    def _part1_tensor_core(self, x, window_size, shift_size, mask, rid, residual=None):
        """Three launches for the whole of forward_part1 (inference):
        [LN + pad + roll + partition + qkv] tcgen05 GEMM -> attention core on fp16 q/k/v^T -> [proj + reverse + roll + crop]
        tcgen05 GEMM.  Returns None when the shape is outside what those kernels cover (the caller then takes the
        un-fused kernels)."""
        attn = self.attn
        B, D, H, W, C = x.shape
        heads = attn.num_heads
        hd = C // heads
        N = window_size[0] * window_size[1] * window_size[2]
        smem_ok = lambda n_out: n_out * C * 2 + max(256 * C, 33792) + 4 * n_out + 3328 <= 227 * 1024
        if not (attn.tc_linear and C % 16 == 0 and C % 32 == 0 and 3 * C <= 512 and ((3 * C) // 2) % 16 == 0
                and hd <= 32 and hd % 4 == 0 and N % 8 == 0 and smem_ok(3 * C) and smem_ok(C)
                and attn.qkv.in_features == C and attn.proj.out_features == C):
            return None
        if mask is not None and rid is None:
            _require_cuda_f32(mask, "attention mask")
        lib = _lib.load()
        Dp, Hp, Wp = (-(-n // w) * w for n, w in zip((D, H, W), window_size))
        nW = (Dp // window_size[0]) * (Hp // window_size[1]) * (Wp // window_size[2])
        B_ = B * nW
        dev = x.device
        q16 = torch.empty((B_, heads, N, 32), dtype=torch.float16, device=dev)
        k16 = torch.empty_like(q16)
        vt16 = torch.empty_like(q16)                      # v, same row layout (+ a column of ones at index head_dim)
        core = torch.empty((B_, N, C), dtype=torch.float32, device=dev)
        y = torch.empty((B, D, H, W, C), dtype=torch.float32, device=dev)
        n1 = self.norm1
        opt = lambda t: t.detach().contiguous().data_ptr() if t is not None else None
        stream_qkv = C <= 256 and n1.weight is not None and n1.bias is not None      # weight streamed in q / k / v chunks, 2 CTAs/SM
        with torch.cuda.device(dev):
            if stream_qkv:
                st = lib.dlsg_swin_qkv_stream_f32(
                    x.data_ptr(), opt(n1.weight), opt(n1.bias), float(n1.eps), self._qkv_f16_weight().data_ptr(), opt(attn.qkv.bias),
                    q16.data_ptr(), k16.data_ptr(), vt16.data_ptr(), B, D, H, W, C, _i3(window_size), _i3(shift_size),
                    heads, float(attn.scale), _stream())
            else:
                st = lib.dlsg_swin_qkv_fused_f32(
                    x.data_ptr(), opt(n1.weight), opt(n1.bias), float(n1.eps), opt(attn.qkv.weight), opt(attn.qkv.bias),
                    q16.data_ptr(), k16.data_ptr(), vt16.data_ptr(), B, D, H, W, C, _i3(window_size), _i3(shift_size),
                    heads, float(attn.scale), _stream())
            _lib.check(st, "dlsg_swin_qkv")
            if rid is not None:
                mode, mptr = MASK_REGION, rid.data_ptr()
            elif mask is not None:
                mode, mptr = MASK_DENSE, mask.contiguous().data_ptr()
            else:
                mode, mptr = MASK_NONE, None
            st = lib.dlsg_swin_window_attention_f16(
                q16.data_ptr(), k16.data_ptr(), vt16.data_ptr(), opt(attn.relative_position_bias_table), core.data_ptr(),
                B_, N, heads, hd, _i3(attn.window_size), mode, mptr, nW, _stream())
            _lib.check(st, "dlsg_swin_window_attention_f16")
            st = lib.dlsg_swin_proj_scatter_residual_f32(
                core.data_ptr(), opt(attn.proj.weight), opt(attn.proj.bias),
                residual.data_ptr() if residual is not None else None, y.data_ptr(),
                B, D, H, W, C, _i3(window_size), _i3(shift_size), _stream())
            _lib.check(st, "dlsg_swin_proj_scatter")
        return y

    def _inference(self, x) -> bool:
        """True when the fused kernels may run: no gradient is needed and no dropout would fire (a module in train()
        mode under no_grad still applies attn_drop / proj_drop in the reference)."""
        if self.training and (self.attn.attn_drop.p > 0 or self.attn.proj_drop.p > 0):
            return False
        return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))

    def _qkv_f16_weight(self):
        """qkv.weight [3C, C] as fp16 in chunked core-matrix order ([3][k/8][n/8][n%8][k%8]); refreshed when the parameter changes."""
        w = self.attn.qkv.weight
        key = (w.data_ptr(), w._version, w.device)
        cache = getattr(self, "_qkv_f16", None)
        if cache is None or cache[0] != key:
            C = w.shape[1]
            wt = w.detach().to(torch.float16).view(3, C // 8, 8, C // 8, 8).permute(0, 3, 1, 2, 4).contiguous()
            cache = (key, wt)
            self._qkv_f16 = cache
        return cache[1]

    def _mlp_f16_weights(self):
        """fc1 / fc2 weights as fp16 in the kernels' chunked core-matrix order ([H/C][k/8][n/8][n%8][k%8], see
        csrc/swin_mlp_tc.cuh), converted once and refreshed when the parameters change (constants at inference)."""
        w1, w2 = self.mlp.fc1.weight, self.mlp.fc2.weight
        key = (w1.data_ptr(), w1._version, w2.data_ptr(), w2._version, w1.device)
        cache = getattr(self, "_mlp_f16", None)
        if cache is None or cache[0] != key:
            Hd, C = w1.shape
            nk = Hd // C
            # fc1 [H, C]: chunk = C output rows n;  [nc, n/8, n%8, k/8, k%8] -> [nc, k/8, n/8, n%8, k%8]
            w1t = w1.detach().to(torch.float16).view(nk, C // 8, 8, C // 8, 8).permute(0, 3, 1, 2, 4).contiguous()
            # fc2 [C, H]: chunk = C input columns k;  [n/8, n%8, kc, k/8, k%8] -> [kc, k/8, n/8, n%8, k%8]
            w2t = w2.detach().to(torch.float16).view(C // 8, 8, nk, C // 8, 8).permute(2, 3, 0, 1, 4).contiguous()
            cache = (key, w1t, w2t)
            self._mlp_f16 = cache
        return cache[1], cache[2]

    def _part2_tensor_core(self, x, residual: bool):
        """[x +] fc2(GELU(fc1(norm2(x)))) in two tcgen05 GEMM launches (inference).  None when outside what the kernels cover."""
        mlp = self.mlp
        C, Hd = mlp.fc1.in_features, mlp.fc1.out_features
        ok = (self.attn.tc_linear and isinstance(self.norm2, nn.LayerNorm) and isinstance(mlp.act, nn.GELU)
              and getattr(mlp.act, "approximate", "none") == "none"
              and (mlp.drop.p == 0 or not self.training)
              and (isinstance(self.drop_path, nn.Identity) or not self.training))
        if not (ok and x.shape[-1] == C and mlp.fc2.in_features == Hd and mlp.fc2.out_features == C
                and C % 32 == 0 and C <= 256 and Hd % C == 0 and self.norm2.weight is not None and self.norm2.bias is not None
                and max(256 * C, 33792) + 2 * C * C + 4 * Hd + 2200 <= 113 * 1024):
            return None
        _require_cuda_f32(x, "SwinTransformerBlock3D input")
        x = x.contiguous()
        lib = _lib.load()
        M = x.numel() // C
        w1h, w2h = self._mlp_f16_weights()
        hidden = torch.empty((-(-M // 128) * 128, Hd), dtype=torch.float16, device=x.device)    # tile-ordered workspace
        y = torch.empty_like(x)
        if M == 0:
            return y
        opt = lambda t: t.detach().contiguous().data_ptr() if t is not None else None
        with torch.cuda.device(x.device):
            st = lib.dlsg_swin_mlp_f32(x.data_ptr(), opt(self.norm2.weight), opt(self.norm2.bias), float(self.norm2.eps),
                                       w1h.data_ptr(), opt(mlp.fc1.bias), w2h.data_ptr(), opt(mlp.fc2.bias),
                                       hidden.data_ptr(), y.data_ptr(), M, C, Hd, 1 if residual else 0, _stream())
        _lib.check(st, "dlsg_swin_mlp")
        return y

    def forward_part2(self, x):
        if self._inference(x):
            y = self._part2_tensor_core(x, residual=False)
            if y is not None:
                return y
        return self.drop_path(self.mlp(self.norm2(x)))

    def _forward_inference(self, x, mask_matrix):
        """Whole block in five launches: qkv GEMM, attention core, proj GEMM (+ shortcut), fc1, fc2 (+ residual)."""
        if not (isinstance(self.norm1, nn.LayerNorm) and (isinstance(self.drop_path, nn.Identity) or not self.training)):
            return None
        B, D, H, W, C = x.shape
        window_size, shift_size = get_window_size((D, H, W), self.window_size, self.shift_size)
        shifted = any(i > 0 for i in shift_size)
        Dp, Hp, Wp = (-(-n // w) * w for n, w in zip((D, H, W), window_size))
        _require_cuda_f32(x, "SwinTransformerBlock3D input")
        x = x.contiguous()
        rid = _region_ids_for(mask_matrix, Dp, Hp, Wp, window_size, shift_size, x.device) if shifted else None
        y = self._part1_tensor_core(x, window_size, shift_size, mask_matrix if shifted else None, rid, residual=x)
        if y is None:
            return None
        return self._part2_tensor_core(y, residual=True)

    def forward(self, x, mask_matrix):
        if self._inference(x) and not self.use_checkpoint:
            y = self._forward_inference(x, mask_matrix)
            if y is not None:
                return y
        shortcut = x
        if self.use_checkpoint:
            x = checkpoint.checkpoint(self.forward_part1, x, mask_matrix, use_reentrant=False)
        else:
            x = self.forward_part1(x, mask_matrix)
        x = shortcut + self.drop_path(x)
        if self._inference(x) and not self.use_checkpoint:
            y = self._part2_tensor_core(x, residual=True)          # the residual add rides in the fc2 epilogue
            if y is not None:
                return y
        if self.use_checkpoint:
            x = x + checkpoint.checkpoint(self.forward_part2, x, use_reentrant=False)
        else:
            x = x + self.forward_part2(x)
        return x


class BasicLayer(nn.Module):
    """One Swin stage (:358-437): NCDHW -> NDHWC, shared shift mask, `depth` blocks alternating shift."""

    def __init__(self, dim, depth, num_heads, window_size=(1, 7, 7), mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.window_size = window_size
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth, self.use_checkpoint = depth, use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim=dim, num_heads=num_heads, window_size=window_size,
                                   shift_size=(0, 0, 0) if (i % 2 == 0) else self.shift_size, mlp_ratio=mlp_ratio,
                                   qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                   drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                   norm_layer=norm_layer, use_checkpoint=use_checkpoint)
            for i in range(depth)])
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        B, C, D, H, W = x.shape
        window_size, shift_size = get_window_size((D, H, W), self.window_size, self.shift_size)
        x = x.permute(0, 2, 3, 4, 1).contiguous()                         # b c d h w -> b d h w c
        Dp, Hp, Wp = (int(np.ceil(n / w)) * w for n, w in zip((D, H, W), window_size))
        attn_mask = compute_mask(Dp, Hp, Wp, window_size, shift_size, x.device)
        for blk in self.blocks:
            x = blk(x, attn_mask)
        x = x.view(B, D, H, W, -1)
        if self.downsample is not None:
            x = self.downsample(x)
        return x.permute(0, 4, 1, 2, 3).contiguous()                      # b d h w c -> b c d h w
