"""Generate tests/golden/swin_*.npz by running the UNMODIFIED reference Swin modules
(/root/reference/dl_cs/models/video_swin_transformer_mri_downsample.py) in the build container.

`timm` is not installed; the reference imports only DropPath and trunc_normal_ from it (:11), neither
of which does inference arithmetic, so a 2-symbol in-memory shim is registered before the import.
Inputs and weights are regenerated from seeds by oracle/golden_inputs.py::swin_case (not stored).
"""
from __future__ import annotations

import os
import sys
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch
from torch import nn

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

from oracle.golden_inputs import SWIN_CASES, swin_case  # noqa: E402


def install_timm_shim():
    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.trunc_normal_ = nn.init.trunc_normal_
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def main():
    install_timm_shim()
    sys.path.insert(0, REF)
    from dl_cs.models import video_swin_transformer_mri_downsample as ref

    os.makedirs(OUT, exist_ok=True)
    for name in SWIN_CASES:
        c = swin_case(name)
        blk = ref.SwinTransformerBlock3D(dim=c["dim"], num_heads=c["heads"], window_size=c["window"],
                                         shift_size=c["shift"], qkv_bias=True).eval()
        sd = blk.state_dict()
        for k, v in c["params"].items():
            assert sd[k].shape == tuple(v.shape), (k, sd[k].shape, v.shape)
            sd[k] = torch.from_numpy(v)
        blk.load_state_dict(sd)
        x = torch.from_numpy(c["x"])
        B, D, H, W, C = x.shape
        ws, ss = ref.get_window_size((D, H, W), c["window"], c["shift"])
        Dp, Hp, Wp = (int(np.ceil(n / w)) * w for n, w in zip((D, H, W), ws))
        mask = ref.compute_mask(Dp, Hp, Wp, ws, ss, x.device)
        with torch.no_grad():
            part1 = blk.forward_part1(x, mask)
            # the attention module alone on the partitioned (un-normed) input, with and without the mask
            xp = torch.nn.functional.pad(x, (0, 0, 0, Wp - W, 0, Hp - H, 0, Dp - D))
            xw = ref.window_partition(xp, ws)
            attn_nomask = blk.attn(xw, mask=None)
            attn_mask = blk.attn(xw, mask=mask)
            block = blk(x, mask)                 # whole block: x + part1, then + forward_part2 (:254-273)
        out = {"part1": part1.numpy(), "attn_nomask": attn_nomask.numpy(), "attn_mask": attn_mask.numpy(),
               "mask": mask.numpy().astype(np.float32), "ws": np.array(ws), "ss": np.array(ss),
               "relative_position_index": blk.attn.relative_position_index.numpy()}
        np.savez_compressed(os.path.join(OUT, f"swin_{name}.npz"), **out)
        np.savez_compressed(os.path.join(OUT, f"swin_block_{name}.npz"), block=block.numpy())
        print(name, tuple(x.shape), "ws", ws, "ss", ss, "nW", mask.shape[0])


if __name__ == "__main__":
    main()
