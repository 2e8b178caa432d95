"""3 x 3 x 3 'same' convolution of the denoisers on the B200 tensor cores (csrc/conv3d_tc.cuh, `dlsg_conv3d_k3_tf32x3_f32`):
the arithmetic of `set_conv_precision("tf32x3")` - three TF32 products on operands split into a TF32 head and the remainder -
in ONE kernel that splits the operands while it stages them, instead of three cuDNN convolutions around elementwise passes.
Inference only (no autograd); `layers.conv_forward` routes here when the layer and the tensor qualify.
"""
from __future__ import annotations

import torch

from .. import _lib

_WPREP: dict = {}


def _split(t: torch.Tensor):
    hi = ((t.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return hi, t - hi


def prepare_weight(weight: torch.Tensor, kc: int = 32) -> torch.Tensor:
    """weight [Cout, Cin, 3, 3, 3] -> the kernel's operand image (include/dlsg_b200.h); kc = K elements per pipeline stage, as
    `dlsg_conv3d_k3_supported(Cin, Cout)` reports it.
      Cin % 32 == 0:  [27][Cin/kc][hi|lo][kc/4][Cout/8][8][4]                     K walks (tap, channel chunk)
      Cin = 2 or 4:   [ceil(27 Cin / kc)][hi|lo][kc/4][Cout/8][8][4]              K index = tap * Cin + ci, zero-padded"""
    cout, cin = weight.shape[:2]
    w = weight.detach().float().contiguous()
    if cin in (2, 4):
        k = 27 * cin
        kpad = -(-k // kc) * kc
        wk = torch.zeros(cout, kpad, device=w.device, dtype=torch.float32)
        wk[:, :k] = w.reshape(cout, cin, 27).permute(0, 2, 1).reshape(cout, k)
        hi, lo = _split(wk)
        hl = torch.stack([hi, lo]).reshape(2, cout // 8, 8, kpad // kc, kc // 4, 4)          # [h, co/8, co%8, stage, k4, k%4]
        return hl.permute(3, 0, 4, 1, 2, 5).contiguous()
    hi, lo = _split(w)
    hl = torch.stack([hi, lo]).reshape(2, cout // 8, 8, cin // kc, kc // 4, 4, 27)          # [h, co/8, co%8, chunk, k4, ci%4, tap]
    return hl.permute(6, 3, 0, 4, 1, 2, 5).contiguous()


def _cached_weight(weight: torch.Tensor) -> torch.Tensor:
    key = (weight.data_ptr(), weight.device, tuple(weight.shape))
    hit = _WPREP.get(key)
    if hit is not None and hit[0] == weight._version:
        return hit[1]
    kc = _lib.load().dlsg_conv3d_k3_supported(weight.shape[1], weight.shape[0])
    prep = prepare_weight(weight, kc)
    _WPREP[key] = (weight._version, prep)
    return prep


def supported(conv: torch.nn.Module, x: torch.Tensor) -> bool:
    return (isinstance(conv, torch.nn.Conv3d) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 5
            and tuple(conv.kernel_size) == (3, 3, 3) and tuple(conv.stride) == (1, 1, 1) and tuple(conv.padding) == (1, 1, 1)
            and tuple(conv.dilation) == (1, 1, 1) and conv.groups == 1 and conv.padding_mode == "zeros"
            and not (torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad))
            and bool(_lib.load().dlsg_conv3d_k3_supported(conv.in_channels, conv.out_channels)))


def conv3d_k3(x: torch.Tensor, weight: torch.Tensor, bias=None, relu_in: bool = False) -> torch.Tensor:
    """x [NB, Cin, T, Y, X] (any layout; used channels-last) -> [NB, Cout, T, Y, X] in channels_last_3d."""
    lib = _lib.load()
    nb, cin, t, y, xx = x.shape
    cout = weight.shape[0]
    xc = x.permute(0, 2, 3, 4, 1)
    if not xc.is_contiguous():
        xc = xc.contiguous()
    prep = _cached_weight(weight)
    out = torch.empty((nb, t, y, xx, cout), device=x.device, dtype=torch.float32)
    b = bias.detach().float().contiguous() if bias is not None else None
    st = lib.dlsg_conv3d_k3_tf32x3_f32(xc.data_ptr(), prep.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(),
                                       nb, t, y, xx, cin, cout, int(relu_in), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(st, "dlsg_conv3d_k3_tf32x3_f32")
    return out.permute(0, 4, 1, 2, 3)
