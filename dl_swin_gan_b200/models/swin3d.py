"""Swin denoiser of the unrolled network: mirror of dl_cs/models/swin3D.py
(SwinTransformer3DBlock :304-325, ResSwinTransformer3DBlock :327-340, DeepFeatureExtraction :342-369,
SwinTransformer3DNet :371-435) with the B200 window-attention backbone inside.

Module tree and parameter names equal the reference's (`SFE.layers.2.conv.*`, `DFE.resswin_blocks.N.layers.0.transformer.*`,
`DFE.layers.N.*` - the reference registers the ResSwin blocks twice, in `resswin_blocks` and in `layers`, and so does this
mirror - `final_layer.*`), so `load_state_dict` of a reference checkpoint works.  The Swin hyper-parameters are the
hard-coded ones of the reference (`depths=[6], num_heads=[8], window_size=(7,8,8)`, swin3D.py:315): the constructor
arguments `window_size / num_heads / num_layers` are accepted and ignored exactly as there.
"""
from __future__ import annotations

import torch
from torch import nn

from ..mri.utils import center_crop
from ..swin.backbone import SwinTransformer3D
from .layers import ConvBlock, channels_to_complex, complex_to_channels, grow_frames, time_pad_circular


class SwinTransformer3DBlock(nn.Module):
    def __init__(self, in_chans, chans, window_size, num_heads, num_layers, is_complex=False):
        super().__init__()
        self.is_complex = is_complex
        self.transformer = SwinTransformer3D(in_chans=in_chans, embed_dim=chans, depths=[6], num_heads=[8], window_size=(7, 8, 8))

    def forward(self, input):
        return self.transformer(input)


class ResSwinTransformer3DBlock(nn.Module):
    def __init__(self, in_chans, chans, window_size, num_heads, num_layers, act_type="relu", is_complex=False):
        super().__init__()
        self.layers = nn.Sequential(
            SwinTransformer3DBlock(in_chans, chans, window_size, num_heads, num_layers, is_complex=is_complex),
            ConvBlock(chans, chans, kernel_size=3, act_type=act_type, is_complex=is_complex))

    def forward(self, input, t_range=None):
        return self.layers[1](self.layers[0](input), residual=input, input_is_temp=True, t_range=t_range)


class DeepFeatureExtraction(nn.Module):
    def __init__(self, in_chans, chans, window_size, num_heads, num_layers, num_swinblocks, act_type="relu", is_complex=False):
        super().__init__()
        self.resswin_blocks = nn.ModuleList(
            [ResSwinTransformer3DBlock(in_chans, chans, window_size, num_heads, num_layers, act_type=act_type, is_complex=is_complex)
             for _ in range(num_swinblocks)])
        self.layers = nn.Sequential(*self.resswin_blocks,
                                    ConvBlock(chans, chans, kernel_size=3, act_type=act_type, is_complex=is_complex))

    def forward(self, input, skip_scale=1.0, t_range=None):
        """`layers(input) + skip_scale * input`; the caller's own `x + DFE(x)` (swin3D.py:430) rides in as skip_scale = 2.
        `t_range`: the output frames the caller reads.  Only the last block's convolution can be narrowed with it - the Swin backbone
        of every block attends over all frames."""
        x = input
        last = len(self.resswin_blocks) - 1
        for i, block in enumerate(self.resswin_blocks):
            x = block(x, t_range=grow_frames(t_range, 1) if i == last else None)
        return self.layers[-1](x, residual=input, residual_scale=skip_scale, input_is_temp=x is not input, t_range=t_range)


class SwinTransformer3DNet(nn.Module):
    """SwinMR-style denoiser on [N, emaps, T, Y, X] complex64 series."""

    def __init__(self, num_swinblocks, in_chans, chans, kernel_size, window_size, act_type="relu", num_heads=[4],
                 num_layers=[4], use_complex_layers=False, circular_pad=True):
        super().__init__()
        if use_complex_layers:
            raise NotImplementedError("the Swin backbone is real-valued; the reference's complex-layer variant cannot run either "
                                      "(LayerNorm / Linear on complex tensors) - use CONV_BLOCK.COMPLEX: False (config_swin.yaml:19)")
        self.use_complex_layers = use_complex_layers
        self.circular_pad = circular_pad
        self.pad_size = (2 * num_swinblocks + 2) * (kernel_size - 1) // 2
        self.SFE = ConvBlock(in_chans, chans, kernel_size=3, act_type="none", is_complex=use_complex_layers)      # shallow features
        self.DFE = DeepFeatureExtraction(chans, chans, window_size, num_heads, num_layers, num_swinblocks, act_type=act_type,
                                         is_complex=use_complex_layers)
        self.final_layer = ConvBlock(chans, in_chans, kernel_size=3, act_type=act_type, is_complex=use_complex_layers)

    def _preprocess(self, input):
        if not self.use_complex_layers:
            input = complex_to_channels(input)
        if self.circular_pad:
            input = time_pad_circular(input, self.pad_size)
        return input

    def _postprocess(self, output):
        output = center_crop(output, dims=[2], shapes=[output.shape[2] - 2 * self.pad_size])
        if not self.use_complex_layers:
            output = channels_to_complex(output)
        return output

    def forward(self, x):
        x = self._preprocess(x)
        # _postprocess keeps the frames pad_size .. T - pad_size: a layer's outer frames that no later layer reads are not computed
        # (bit-identical on the frames that are; the reference computes and discards them)
        keep = (self.pad_size, x.shape[2] - self.pad_size)
        x = self.SFE(x)
        x = self.DFE(x, skip_scale=2.0, t_range=grow_frames(keep, 1))      # x + DFE(x) = layers(x) + 2 x
        x = self.final_layer(x, input_is_temp=True, t_range=keep)
        return self._postprocess(x)
