"""Single-stage Swin backbone around the B200 window-attention blocks: mirror of
dl_cs/models/video_swin_transformer_mri_downsample.py  PatchEmbed3D :440-479, PatchUnembed3D :481-531,
SwinTransformer3D :534-761 (the "downsample" variant the denoiser imports, dl_cs/models/swin3D.py:14).

The denoiser instantiates it with `depths=[6], num_heads=[8], window_size=(7,8,8)` (swin3D.py:315), i.e. ONE stage:
no PatchMerging / PatchExpand is ever built (`:647`, `:651-654` loop over num_layers-1 = 0).  That is what is mirrored
here; more than one stage raises.  Parameter and buffer names equal the reference's (`patch_embed.proj.*`,
`patch_unembed.proj.*`, `layers.0.blocks.N.*`, `norm.*`), so a reference state dict loads.
The patch (un)embedding convolutions are cuDNN library calls; the six blocks are the hand-written
five-launch blocks of window_attention.py.
"""
from __future__ import annotations

from math import ceil, floor

import torch
import torch.nn.functional as F
from torch import nn

from .window_attention import BasicLayer


def _conv(conv, x):
    from ..models.layers import conv_forward       # late import: models imports this module
    return conv_forward(conv, x)


class PatchEmbed3D(nn.Module):
    """Video to patch embedding: zero-pad each axis up to a multiple of the patch, Conv3d(kernel = stride = patch)."""

    def __init__(self, patch_size=(2, 4, 4), in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.patch_size = patch_size
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        _, _, D, H, W = x.size()
        pd, ph, pw = self.patch_size
        if W % pw or H % ph or D % pd:
            x = F.pad(x, (0, -W % pw, 0, -H % ph, 0, -D % pd))
        x = _conv(self.proj, x)                           # B C D Wh Ww
        if self.norm is not None:
            D, Wh, Ww = x.size(2), x.size(3), x.size(4)
            x = self.norm(x.flatten(2).transpose(1, 2))
            x = x.transpose(1, 2).view(-1, self.embed_dim, D, Wh, Ww)
        return x


class PatchUnembed3D(nn.Module):
    """Patch tokens back to a video: ConvTranspose3d(kernel = stride = patch), then a centred crop to `pre_size`."""

    def __init__(self, patch_size, in_channels: int = 3, embed_dim: int = 96, norm_layer=None) -> None:
        super().__init__()
        self.tuple_patch_size = (patch_size[0], patch_size[1], patch_size[2])
        self.proj = nn.ConvTranspose3d(embed_dim, in_channels, kernel_size=self.tuple_patch_size, stride=self.tuple_patch_size)
        self.norm = norm_layer(in_channels) if norm_layer is not None else nn.Identity()

    def forward(self, x, pre_size):
        x = _conv(self.proj, x)
        cur = x.size()
        lo = [ceil((cur[j] - pre_size[j]) / 2) for j in range(5)]
        hi = [cur[j] - floor((cur[j] - pre_size[j]) / 2) for j in range(5)]
        x = x[:, :, lo[2]:hi[2], lo[3]:hi[3], lo[4]:hi[4]]
        if not isinstance(self.norm, nn.Identity):
            x = self.norm(x.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)
        return x


class SwinTransformer3D(nn.Module):
    def __init__(self, pretrained=None, pretrained2d=True, patch_size=(4, 4, 4), in_chans=3, embed_dim=96,
                 depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=(2, 7, 7), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=nn.LayerNorm,
                 patch_norm=False, frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if len(depths) != 1:
            raise NotImplementedError(
                "the B200 mirror covers the single-stage configuration the denoiser uses (depths=[6], swin3D.py:315); "
                "PatchMerging / PatchExpand stages are not built")
        if pretrained is not None:
            raise NotImplementedError("2-D checkpoint inflation (inflate_weights) is not part of the hot path")
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.frozen_stages = frozen_stages
        self.window_size = window_size
        self.patch_size = patch_size
        self.patch_embed = PatchEmbed3D(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                        norm_layer=norm_layer if patch_norm else None)
        self.patch_unembed = PatchUnembed3D(patch_size=patch_size, embed_dim=embed_dim,
                                            norm_layer=norm_layer if patch_norm else None, in_channels=in_chans)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]      # stochastic depth decay rule (:603)
        self.layers = nn.ModuleList([
            BasicLayer(dim=embed_dim, depth=depths[0], num_heads=num_heads[0], window_size=window_size, mlp_ratio=mlp_ratio,
                       qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                       drop_path=dpr[:depths[0]], norm_layer=norm_layer, downsample=None, use_checkpoint=use_checkpoint)])
        self.layers_up = nn.ModuleList()
        self.num_features = embed_dim
        self.norm = norm_layer(self.num_features)             # present in the reference's state dict, unused in forward (:661)
        # the reference's init_weights (:708-733) is never called by the denoiser: every layer keeps torch's default initialisation
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for p in self.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 1:
            self.pos_drop.eval()
            for i in range(self.frozen_stages):
                self.layers[i].eval()
                for p in self.layers[i].parameters():
                    p.requires_grad = False

    def forward(self, x):
        pre_size = x.size()
        x = self.pos_drop(self.patch_embed(x))
        x = self.layers[0](x.contiguous())
        return self.patch_unembed(x, pre_size)

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        return self
