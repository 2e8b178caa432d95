"""Convolutional building blocks of the denoisers: mirror of the shared definitions in dl_cs/models/swin3D.py:16-302
(== dl_cs/models/resnet3d.py:14-241): Normalization, Activation, Conv3d, ComplexConv3d, ConvBlock, ResBlock.

These are the cuDNN part of the reconstruction (97 % of the denoiser's FLOPs): library convolutions,
kept in PyTorch with the reference's module tree and parameter names (`layers.2.conv.weight` ...) so state dicts load.
On the device they run channels-last-3d at the precision chosen with `set_conv_precision` (default: 3xTF32, fp32-accurate).
"""
from __future__ import annotations

import torch
from torch import nn

_CONV_PRECISION = "tf32x3"
_TC_CONV = True          # route qualifying 3x3x3 layers to the hand-written tensor-core kernel (models/conv3d.py) in "tf32x3" mode
_CONV_MODES = ("tf32x3", "tf32x2", "tf32r", "fp32", "tf32", "bf16")


def set_conv_precision(mode: str) -> None:
    """Arithmetic of the denoiser convolutions on CUDA tensors (storage stays fp32):
    "tf32x3"  the default, "three split products": every operand is split into a short-mantissa head and the remainder and the
              convolution is  conv(x_hi, w_hi) + conv(x_lo, w_hi) + conv(x_hi, w_lo)  with fp32 accumulation.  At inference the layers
              the denoisers have (3x3x3 'same', first / last layer, patch embed / unembed) run on the repo's own tensor-core kernels
              (models/conv3d.py, csrc/conv3d_tc.cuh: bf16 head + bf16 remainder, 16 mantissa bits, ~1e-5 per layer, ONE pass over the
              activations with the split, the pre-activation ReLU and the skip connection fused); anything else, and training, takes
              three cuDNN TF32 convolutions on TF32 head / remainder (~1e-6, three passes + elementwise kernels).  Keeps a full
              reconstruction within the 1e-3 NRMSE bar of the fp32 reference (tests/test_models.py: 5.1e-4 for five Swin unrolls, of
              which the fp16 attention linears are the larger part);
    "tf32x2" / "tf32r"   two split products / one product on operands rounded to TF32: measured 1.07e-3 / 1.26e-3 on the five-unroll
              Swin fixture, i.e. outside the bar - kept for the precision probe (tools/model_precision_probe.py);
    "fp32"    cuDNN with TF32 switched off (IEEE fp32 FMAs; slow: no tensor cores);
    "tf32"    cuDNN's default for fp32 tensors, which is also what the reference itself runs on a GPU with PyTorch's defaults:
              1.2e-3 NRMSE against the fp32 result after five unrolls of the Swin network;
    "bf16"    autocast to bfloat16 (about 1e-2)."""
    global _CONV_PRECISION
    if mode not in _CONV_MODES:
        raise ValueError("conv precision must be one of %s" % (_CONV_MODES,))
    _CONV_PRECISION = mode


def get_conv_precision() -> str:
    return _CONV_PRECISION


def _tf32_split(t: torch.Tensor):
    """t = hi + lo with hi representable in TF32 (10-bit mantissa, round to nearest): the tensor cores take hi exactly."""
    # same-size dtype views keep the strides, so a channels-last tensor stays channels-last (no layout copies)
    if not (t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last_3d)):
        t = t.contiguous()
    hi = ((t.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return hi, t - hi


class Normalization(nn.Module):
    def __init__(self, in_chans, type):
        super().__init__()
        if type == "none":
            self.norm = nn.Identity()
        elif type == "instance":
            self.norm = nn.InstanceNorm3d(in_chans, affine=False)
        elif type == "batch":
            self.norm = nn.BatchNorm3d(in_chans, affine=False)
        else:
            raise ValueError("Invalid normalization type: %s" % type)

    def forward(self, input):
        if input.is_complex():
            return torch.complex(self.norm(input.real), self.norm(input.imag))
        return self.norm(input)


class Activation(nn.Module):
    def __init__(self, type):
        super().__init__()
        if type == "none":
            self.activ = nn.Identity()
        elif type == "relu":
            # inplace=True is part of the reference's arithmetic, not an optimisation: in `ResBlock.forward`
            # (`self.layers(input) + input`, resnet3d.py:240) the first ConvBlock rectifies `input` itself, so the skip
            # connection adds relu(input).  Kept bit for bit.
            self.activ = nn.ReLU(inplace=True)
        elif type == "leaky_relu":
            self.activ = nn.LeakyReLU(inplace=True)
        elif type == "sigmoid":
            self.activ = nn.Sigmoid()
        else:
            raise ValueError("Invalid activation type: %s" % type)

    def forward(self, input):
        if input.is_complex():
            return torch.complex(self.activ(input.real), self.activ(input.imag))
        return self.activ(input)


def conv_forward(conv: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """`conv(x)` (Conv3d / ConvTranspose3d) in channels-last-3d at the selected precision."""
    if not x.is_cuda:
        return conv(x)
    x = x.contiguous(memory_format=torch.channels_last_3d)
    mode = _CONV_PRECISION
    if mode == "bf16":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return conv(x).float()
    tf32 = mode in ("tf32", "tf32x3", "tf32x2", "tf32r")
    with torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark,
                                    deterministic=torch.backends.cudnn.deterministic, allow_tf32=tf32):
        if mode == "tf32" or not tf32 or x.dtype != torch.float32:
            return conv(x)
        if mode == "tf32x3" and _TC_CONV:
            from . import conv3d
            which = conv3d.kind(conv, x)
            if which is not None:
                return conv3d.forward(conv, x, which)
        if isinstance(conv, nn.ConvTranspose3d):
            def f(a, w, b):
                return torch.nn.functional.conv_transpose3d(a, w, b, conv.stride, conv.padding, conv.output_padding, conv.groups, conv.dilation)
        else:
            def f(a, w, b):
                return torch.nn.functional.conv3d(a, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
        xh, xl = _tf32_split(x)
        wh, wl = _tf32_split(conv.weight)
        xh = xh.contiguous(memory_format=torch.channels_last_3d)
        if mode == "tf32r":
            return f(xh, wh, conv.bias)
        xl = xl.contiguous(memory_format=torch.channels_last_3d)
        if mode == "tf32x2":
            return f(xh, wh, conv.bias) + f(xl, wh, None)
        return f(xh, wh, conv.bias) + f(xl, wh, None) + f(xh, wl, None)


_conv = conv_forward


class Conv3d(nn.Module):
    """nn.Conv3d with 'same' zero padding (swin3D.py:120-134)."""

    def __init__(self, in_chans, out_chans, kernel_size):
        super().__init__()
        self.conv = nn.Conv3d(in_chans, out_chans, kernel_size, padding=(kernel_size - 1) // 2)

    def forward(self, input):
        return _conv(self.conv, input)


class ComplexConv3d(nn.Module):
    """(X + iY) * (a + ib) = (X*a - Y*b) + i (X*b + Y*a) with two real convolutions (swin3D.py:137-168)."""

    def __init__(self, in_chans, out_chans, kernel_size):
        super().__init__()
        padding = (kernel_size - 1) // 2
        self.conv_r = nn.Conv3d(in_chans, out_chans, kernel_size, padding=padding)
        self.conv_i = nn.Conv3d(in_chans, out_chans, kernel_size, padding=padding)

    def forward(self, input):
        re = _conv(self.conv_r, input.real) - _conv(self.conv_i, input.imag)
        im = _conv(self.conv_r, input.imag) + _conv(self.conv_i, input.real)
        return torch.complex(re, im)


class ConvBlock(nn.Module):
    """Pre-activation block Norm -> activation -> Conv (swin3D.py:225-273); children sit in `layers` at 0, 1, 2."""

    def __init__(self, in_chans, out_chans, kernel_size, act_type="relu", norm_type="none", is_complex=False):
        super().__init__()
        self.in_chans, self.out_chans, self.is_complex = in_chans, out_chans, is_complex
        self.name = "ComplexConv3D" if is_complex else "Conv3D"
        conv = (ComplexConv3d if is_complex else Conv3d)(in_chans, out_chans, kernel_size=kernel_size)
        self.layers = nn.Sequential(Normalization(in_chans, norm_type), Activation(act_type), conv)

    def forward(self, input, residual=None, residual_scale=1.0, input_is_temp=False, t_range=None, residual_relu=False):
        """`conv(act(input))` (+ residual_scale * residual).  The extra arguments only let the B200 kernels fuse what surrounds the
        convolution; the arithmetic is the reference's.  `input_is_temp`: nobody reads `input` after this call, so the in-place
        ReLU of the reference (which would rectify `input` itself) may be applied inside the kernel's operand load instead.
        `t_range = (lo, hi)`: the caller reads only the output frames lo .. hi - 1 (the rest may be left uninitialised)."""
        act, conv_mod = self.layers[1], self.layers[2]
        fused = None
        if (_TC_CONV and _CONV_PRECISION == "tf32x3" and input.is_cuda and isinstance(conv_mod, Conv3d)
                and isinstance(self.layers[0].norm, nn.Identity) and isinstance(act.activ, (nn.Identity, nn.ReLU))):
            from . import conv3d
            fused = conv3d.kind(conv_mod.conv, input)
        if fused is None:
            out = self.layers(input)
            if residual is not None:
                if residual_relu:
                    residual = torch.relu(residual)
                out = out + (residual if residual_scale == 1.0 else residual_scale * residual)
            return out
        relu = isinstance(act.activ, nn.ReLU)
        if relu and not input_is_temp:
            input = act(input)                      # in place, visible to the caller, exactly as in the reference
            relu = False
        return conv3d.forward(conv_mod.conv, input, fused, relu_in=relu, residual=residual, residual_scale=residual_scale, t_range=t_range,
                              residual_relu=residual_relu)

    def __repr__(self):
        return f"{self.name}(in_chans={self.in_chans}, out_chans={self.out_chans})"


class ResBlock(nn.Module):
    """Two ConvBlocks and a residual connection (swin3D.py:276-302)."""

    def __init__(self, chans, kernel_size, act_type="relu", is_complex=False):
        super().__init__()
        self.layers = nn.Sequential(ConvBlock(chans, chans, kernel_size, act_type=act_type, is_complex=is_complex),
                                    ConvBlock(chans, chans, kernel_size, act_type=act_type, is_complex=is_complex))

    def forward(self, input, t_range=None, input_is_temp=False):
        # the first block rectifies `input` in place (reference quirk, see Activation), so the skip below adds relu(input).
        # `input_is_temp` (nobody else reads `input` afterwards): the ReLU is folded into both reads instead and `input` is not rewritten
        relu_skip = input_is_temp and isinstance(self.layers[0].layers[1].activ, nn.ReLU)
        h = self.layers[0](input, t_range=grow_frames(t_range, 1), input_is_temp=input_is_temp)
        return self.layers[1](h, residual=input, input_is_temp=True, t_range=t_range, residual_relu=relu_skip)


def grow_frames(t_range, k):
    """The frame window a layer must produce so that a 3-tap consumer can produce `t_range` (None stays None = all frames)."""
    return None if t_range is None else (t_range[0] - k, t_range[1] + k)


def time_pad_circular(x: torch.Tensor, pad: int) -> torch.Tensor:
    """Circular padding of axis 2 (time) by `pad` on both sides (swin3D.py:401-404)."""
    return nn.functional.pad(x, (0, 0, 0, 0, pad, pad), mode="circular")


def complex_to_channels(x: torch.Tensor) -> torch.Tensor:
    """[N, C, T, Y, X] complex -> [N, 2C, T, Y, X] real: real parts first, then imaginary parts (swin3D.py:399)."""
    return torch.cat((x.real, x.imag), dim=1)


def channels_to_complex(x: torch.Tensor) -> torch.Tensor:
    c = x.shape[1]
    return torch.complex(x[:, :(c // 2)].contiguous(), x[:, (c // 2):].contiguous())
