"""Denoiser convolutions on the B200 tensor cores (csrc/conv3d_tc.cuh): the arithmetic of `set_conv_precision("tf32x3")` -
three TF32 products on operands split into a TF32 head and the remainder - in ONE kernel that splits the operands while it stages
them, instead of three cuDNN convolutions around elementwise passes.  Three layer shapes (all the denoisers have):
  * nn.Conv3d(kernel 3, 'same')                       `dlsg_conv3d_k3_tf32x3_f32`       (Cin % 32 == 0, or the first layer's Cin = 2 / 4)
  * nn.Conv3d(kernel = stride = patch)                `dlsg_conv3d_patch_tf32x3_f32`    (PatchEmbed3D.proj)
  * nn.ConvTranspose3d(kernel = stride = patch)       `dlsg_conv3d_unpatch_tf32x3_f32`  (PatchUnembed3D.proj)
Inference only (no autograd); `layers.conv_forward` routes here when the layer and the tensor qualify.
"""
from __future__ import annotations

import torch

from .. import _lib

K3, PATCH, UNPATCH, FEWOUT = "k3", "patch", "unpatch", "fewout"
_FEWOUT_GEMM = True       # last layer as pointwise GEMM + tap sum (False: the direct CUDA-core kernel)


def _split(t: torch.Tensor):
    """t = hi + lo + O(2^-18 t): bf16 head (round to nearest) and bf16 remainder."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def prepare_weight(weight: torch.Tensor, kc: int = 32, transposed: bool = False) -> torch.Tensor:
    """weight [Cout, Cin, k, k, k] (or [Cin, Cout, k, k, k] of a transposed convolution) -> the kernel's bf16 operand image
    (include/dlsg_b200.h); kc = K elements per pipeline stage, as `dlsg_conv3d_k3_supported(Cin, Cout)` reports it (bit 0x100: stacked).
      Cin % 32 == 0:  [k^3][Cin/kc][hi|lo][kc/8][Cout/8][8][8]                    K walks (tap, channel chunk)
      stacked:        [k^3][Cin/kc][kc/8][hi|lo][Cout/8][8][8]                    one operand of 2 Cout rows, heads first (Cout = 64, 3x3x3)
      Cin = 2 or 4:   [ceil(27 Cin / kc)][hi|lo][kc/8][Cout/8][8][8]              K index = tap * Cin + ci, zero-padded"""
    w = weight.detach().float()
    if transposed:
        w = w.transpose(0, 1)
    w = w.contiguous()
    cout, cin = w.shape[:2]
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    stacked, kc = bool(kc & 0x100), kc & 0xff
    if cin in (2, 4):
        k = taps * cin
        kpad = -(-k // kc) * kc
        wk = torch.zeros(cout, kpad, device=w.device, dtype=torch.float32)
        wk[:, :k] = w.reshape(cout, cin, taps).permute(0, 2, 1).reshape(cout, k)
        hi, lo = _split(wk)
        hl = torch.stack([hi, lo]).reshape(2, cout // 8, 8, kpad // kc, kc // 8, 8)          # [h, co/8, co%8, stage, k8, k%8]
        return hl.permute(3, 0, 4, 1, 2, 5).contiguous()
    hi, lo = _split(w)
    hl = torch.stack([hi, lo]).reshape(2, cout // 8, 8, cin // kc, kc // 8, 8, taps)        # [h, co/8, co%8, chunk, k8, ci%8, tap]
    return hl.permute(*((6, 3, 4, 0, 1, 2, 5) if stacked else (6, 3, 0, 4, 1, 2, 5))).contiguous()


def _cached_weight(weight: torch.Tensor, transposed: bool) -> torch.Tensor:
    # the image lives on the parameter object itself (a dictionary keyed by data_ptr would hand a freed model's image to the next
    # model whose weight lands on the same address); refreshed when the parameter is written (version) or re-homed (data_ptr)
    try:
        ver = weight._version
    except RuntimeError:                   # inference tensors do not track versions: rebuild every call
        ver = None
    tag = (ver, weight.data_ptr(), transposed)
    hit = getattr(weight, "_dlsg_conv_image", None)
    if hit is not None and ver is not None and hit[0] == tag:
        return hit[1]
    cin, cout = (weight.shape[0], weight.shape[1]) if transposed else (weight.shape[1], weight.shape[0])
    kc = _lib.load().dlsg_conv3d_k3_supported(cin, cout)
    if transposed or tuple(weight.shape[2:]) != (3, 3, 3):
        kc &= 0xff                       # the stacked image belongs to the 3x3x3 Cout = 64 kernel only
    prep = prepare_weight(weight, kc, transposed)
    try:
        weight._dlsg_conv_image = (tag, prep)
    except AttributeError:
        pass
    return prep


def kind(conv: torch.nn.Module, x: torch.Tensor):
    """Which of the three kernels takes `conv(x)`, or None (then cuDNN does)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 5) or conv.groups != 1 or tuple(conv.dilation) != (1, 1, 1):
        return None
    if torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad):
        return None
    # the kernels' limits (32-bit element offsets, packed voxel coordinates): larger volumes take the library path
    nb, _, t, y, xx = x.shape
    if t > 2046 or y > 1022 or xx > 1022 or x.numel() >= 2 ** 31 or nb * t * y * xx * max(conv.out_channels, 27 * 4 + 20) >= 2 ** 31:
        return None
    lib = _lib.load()
    k, s, pd = tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding)
    cubic = k[0] == k[1] == k[2]
    if isinstance(conv, torch.nn.ConvTranspose3d):
        if cubic and k == s and pd == (0, 0, 0) and tuple(conv.output_padding) == (0, 0, 0) and k[0] <= 8 \
                and conv.in_channels % 32 == 0 and lib.dlsg_conv3d_k3_supported(conv.in_channels, conv.out_channels):
            return UNPATCH
        return None
    if not isinstance(conv, torch.nn.Conv3d) or conv.padding_mode != "zeros":
        return None
    if k == (3, 3, 3) and s == (1, 1, 1) and pd == (1, 1, 1) and lib.dlsg_conv3d_k3_supported(conv.in_channels, conv.out_channels):
        return K3
    if k == (3, 3, 3) and s == (1, 1, 1) and pd == (1, 1, 1) and conv.out_channels in (2, 4) and conv.in_channels % 32 == 0 \
            and 27 * conv.in_channels * conv.out_channels * 4 <= 100 * 1024:
        return FEWOUT
    if cubic and k == s and pd == (0, 0, 0) and 1 < k[0] <= 8 and conv.in_channels % 32 == 0 \
            and all(d % k[0] == 0 for d in x.shape[2:]) and lib.dlsg_conv3d_k3_supported(conv.in_channels, conv.out_channels):
        return PATCH
    return None


def supported(conv: torch.nn.Module, x: torch.Tensor) -> bool:
    return kind(conv, x) is not None


def _run(fn_name, x, weight, bias, out_dims, last, transposed, residual=None, residual_scale=1.0, t_range=None):
    lib = _lib.load()
    nb, cin, t, y, xx = x.shape
    cout = weight.shape[1] if transposed else weight.shape[0]
    xc = x.permute(0, 2, 3, 4, 1)
    if not xc.is_contiguous():
        xc = xc.contiguous()
    prep = _cached_weight(weight, transposed)
    out = torch.empty((nb,) + tuple(out_dims) + (cout,), device=x.device, dtype=torch.float32)
    b = bias.detach().float().contiguous() if bias is not None else None
    args = [xc.data_ptr(), prep.data_ptr(), b.data_ptr() if b is not None else None]
    if fn_name == "dlsg_conv3d_k3_tf32x3_f32":
        rc = None
        if residual is not None:
            if tuple(residual.shape) != (nb, cout) + tuple(out_dims):
                raise ValueError(f"residual {tuple(residual.shape)} does not match the output {(nb, cout) + tuple(out_dims)}")
            rc = residual.permute(0, 2, 3, 4, 1)
            if not rc.is_contiguous():
                rc = rc.contiguous()
        args += [rc.data_ptr() if rc is not None else None, float(residual_scale)]
    tail = [int(last)]
    if fn_name == "dlsg_conv3d_k3_tf32x3_f32":
        lo, hi = _window(t_range, t)
        tail += [lo, hi - lo]
    st = getattr(lib, fn_name)(*args, out.data_ptr(), nb, t, y, xx, cin, cout, *tail, torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(st, fn_name)
    return out.permute(0, 4, 1, 2, 3)


def _window(t_range, t):
    """(lo, hi) of the output frames to compute, clipped to the volume; None = all."""
    if t_range is None:
        return 0, t
    lo, hi = max(0, int(t_range[0])), min(t, int(t_range[1]))
    if hi <= lo:
        raise ValueError(f"empty frame window {t_range} for {t} frames")
    return lo, hi


def conv3d_k3(x: torch.Tensor, weight: torch.Tensor, bias=None, relu_in: bool = False, residual=None, residual_scale: float = 1.0,
              t_range=None, residual_relu: bool = False) -> torch.Tensor:
    """x [NB, Cin, T, Y, X] (any layout; used channels-last) -> conv(relu?(x)) + bias (+ residual_scale * residual),
    [NB, Cout, T, Y, X] in channels_last_3d.  `t_range = (lo, hi)`: only the output frames lo .. hi - 1 are computed (the others are left
    uninitialised) - for callers that know the outer frames are never read (the denoisers' circular time padding)."""
    if residual_relu:
        if residual_scale <= 0:
            raise ValueError("residual_relu needs a positive residual_scale")
        residual_scale = -residual_scale            # the C ABI's encoding of |scale| * relu(residual)
    return _run("dlsg_conv3d_k3_tf32x3_f32", x, weight, bias, x.shape[2:], relu_in, False, residual, residual_scale, t_range)


def conv3d_patch(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """Conv3d with kernel = stride = patch: [NB, Cin, T, Y, X] -> [NB, Cout, T/p, Y/p, X/p]."""
    p = weight.shape[2]
    return _run("dlsg_conv3d_patch_tf32x3_f32", x, weight, bias, [d // p for d in x.shape[2:]], p, False)


def conv3d_unpatch(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """ConvTranspose3d with kernel = stride = patch (weight [Cin, Cout, p, p, p]): [NB, Cin, T, Y, X] -> [NB, Cout, T p, Y p, X p]."""
    p = weight.shape[2]
    return _run("dlsg_conv3d_unpatch_tf32x3_f32", x, weight, bias, [d * p for d in x.shape[2:]], p, True)


def conv3d_k3_fewout(x: torch.Tensor, weight: torch.Tensor, bias=None, relu_in: bool = False) -> torch.Tensor:
    """The last layer (Cout = 2 or 4) in exact fp32: x [NB, Cin, T, Y, X] -> [NB, Cout, T, Y, X] (channels_last_3d)."""
    lib = _lib.load()
    nb, cin, t, y, xx = x.shape
    cout = weight.shape[0]
    xc = x.permute(0, 2, 3, 4, 1)
    if not xc.is_contiguous():
        xc = xc.contiguous()
    try:
        tag = (weight._version, weight.data_ptr())
    except RuntimeError:
        tag = None
    hit = getattr(weight, "_dlsg_conv_tcc", None)
    if hit is not None and tag is not None and hit[0] == tag:
        wt = hit[1]
    else:
        wt = weight.detach().float().reshape(cout, cin, 27).permute(2, 1, 0).contiguous()
        try:
            weight._dlsg_conv_tcc = (tag, wt)
        except AttributeError:
            pass
    out = torch.empty((nb, t, y, xx, cout), device=x.device, dtype=torch.float32)
    b = bias.detach().float().contiguous() if bias is not None else None
    st = lib.dlsg_conv3d_k3_fewout_f32(xc.data_ptr(), wt.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(),
                                       nb, t, y, xx, cin, cout, int(relu_in), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(st, "dlsg_conv3d_k3_fewout_f32")
    return out.permute(0, 4, 1, 2, 3)


def conv3d_k3_fewout_gemm(x: torch.Tensor, weight: torch.Tensor, bias=None, relu_in: bool = False, t_range=None) -> torch.Tensor:
    """The last layer (Cout = 2 or 4) as a pointwise tensor-core GEMM to 27*Cout tap channels followed by the sum over the 27 shifted
    voxels (csrc/conv3d_fewout.cuh): two passes, each reading its input once, instead of 27 passes over the activations."""
    lib = _lib.load()
    nb, cin, t, y, xx = x.shape
    cout = weight.shape[0]
    zc = -(-27 * cout // 32) * 32
    xc = x.permute(0, 2, 3, 4, 1)
    if not xc.is_contiguous():
        xc = xc.contiguous()
    try:
        tag = (weight._version, weight.data_ptr())
    except RuntimeError:
        tag = None
    hit = getattr(weight, "_dlsg_conv_tapimage", None)
    if hit is not None and tag is not None and hit[0] == tag:
        prep = hit[1]
    else:
        wz = torch.zeros(zc, cin, 1, 1, 1, device=weight.device, dtype=torch.float32)
        wz[:27 * cout, :, 0, 0, 0] = weight.detach().float().reshape(cout, cin, 27).permute(2, 0, 1).reshape(27 * cout, cin)   # row = tap*Cout + co
        prep = prepare_weight(wz, lib.dlsg_conv3d_k3_supported(cin, zc) & 0xff)
        try:
            weight._dlsg_conv_tapimage = (tag, prep)
        except AttributeError:
            pass
    stream = torch.cuda.current_stream(x.device).cuda_stream
    lo, hi = _window(t_range, t)
    zlo, zhi = max(0, lo - 1), min(t, hi + 1)                   # the tap sum reads z one frame beyond the output window
    z = torch.empty((nb, t, y, xx, zc), device=x.device, dtype=torch.float32)
    _lib.check(lib.dlsg_conv3d_pointwise_tf32x3_f32(xc.data_ptr(), prep.data_ptr(), None, z.data_ptr(), nb, t, y, xx, cin, zc, int(relu_in),
                                                    zlo, zhi - zlo, stream), "dlsg_conv3d_pointwise_tf32x3_f32")
    out = torch.empty((nb, t, y, xx, cout), device=x.device, dtype=torch.float32)
    b = bias.detach().float().contiguous() if bias is not None else None
    _lib.check(lib.dlsg_conv3d_tapsum_f32(z.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(), nb, t, y, xx, zc, cout,
                                          lo, hi - lo, stream), "dlsg_conv3d_tapsum_f32")
    return out.permute(0, 4, 1, 2, 3)


def forward(conv: torch.nn.Module, x: torch.Tensor, which: str, relu_in: bool = False, residual=None, residual_scale: float = 1.0,
            t_range=None, residual_relu: bool = False) -> torch.Tensor:
    """`conv(relu?(x)) (+ residual_scale * residual)` on the kernel `kind()` chose; what a kernel cannot fuse is done around it.
    `t_range`: output frames the caller will read (a hint: kernels without a frame window compute everything)."""
    if which == K3:
        return conv3d_k3(x, conv.weight, conv.bias, relu_in, residual, residual_scale, t_range, residual_relu)
    if residual is not None and residual_relu:
        residual = torch.relu(residual)
    if which == FEWOUT:
        gemm = _FEWOUT_GEMM and bool(_lib.load().dlsg_conv3d_k3_supported(conv.in_channels, -(-27 * conv.out_channels // 32) * 32))
        if gemm:
            out = conv3d_k3_fewout_gemm(x, conv.weight, conv.bias, relu_in, t_range)
        else:
            out = conv3d_k3_fewout(x, conv.weight, conv.bias, relu_in)
    else:
        if relu_in:
            x = torch.relu(x)
        out = conv3d_patch(x, conv.weight, conv.bias) if which == PATCH else conv3d_unpatch(x, conv.weight, conv.bias)
    if residual is not None:
        out = out + residual_scale * residual if residual_scale != 1.0 else out + residual
    return out
