"""ResNet denoiser of the unrolled network: mirror of dl_cs/models/resnet3d.py:243-317 (`ResNet`), the denoiser of
configs/basic/example.yaml and configs/config_recon.yaml (2 residual blocks x 64 features).  cuDNN convolutions in the
reference's module tree (`init_layer`, `res_blocks.N.layers.M`, `final_layer`); note the `+ input` skip around the whole
network (:312) that the Swin denoiser does not have."""
from __future__ import annotations

from torch import nn

from ..mri.utils import center_crop
from .layers import ConvBlock, ResBlock, channels_to_complex, complex_to_channels, grow_frames, time_pad_circular


class ResNet(nn.Module):
    def __init__(self, num_resblocks, in_chans, chans, kernel_size, act_type="relu", use_complex_layers=False, circular_pad=True):
        super().__init__()
        self.use_complex_layers = use_complex_layers
        self.circular_pad = circular_pad
        self.pad_size = (2 * num_resblocks + 2) * (kernel_size - 1) // 2
        chans = int(chans / 1.4142) + 1 if use_complex_layers else chans
        self.init_layer = ConvBlock(in_chans, chans, kernel_size, act_type="none", is_complex=use_complex_layers)
        self.res_blocks = nn.ModuleList([ResBlock(chans, kernel_size, act_type=act_type, is_complex=use_complex_layers)
                                         for _ in range(num_resblocks)])
        self.final_layer = ConvBlock(chans, in_chans, kernel_size, act_type=act_type, is_complex=use_complex_layers)

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

    def forward(self, input):
        input = self._preprocess(input)
        # _postprocess keeps the frames pad_size .. T - pad_size; each 3-tap layer widens what its producer must deliver by one frame,
        # and what nobody reads is not computed (bit-identical on the frames that are; the reference computes and discards them)
        keep = (self.pad_size, input.shape[2] - self.pad_size)
        n = len(self.res_blocks)
        output = self.init_layer(input, t_range=grow_frames(keep, 2 * n + 1))
        for i, res_block in enumerate(self.res_blocks):
            output = res_block(output, t_range=grow_frames(keep, 2 * (n - 1 - i) + 1), input_is_temp=True)
        output = self.final_layer(output, residual=input, input_is_temp=n > 0, t_range=keep)
        return self._postprocess(output)
