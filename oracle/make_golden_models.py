"""Generate tests/golden/model_*.npz by running the UNMODIFIED reference unrolled networks
(/root/reference/dl_cs/models/{unrolledswin,unrolled,swin3D,resnet3d,video_swin_transformer_mri_downsample}.py with the
reference SenseModel, /root/reference/dl_cs/mri/transforms.py) in the build container, torch CPU fp32, .eval():
    python oracle/make_golden_models.py
Weights and problems come from oracle/model_weights.py (functions of names and seeds, not stored); the fixtures hold the
reference outputs: the reconstruction `pgd(y, A)` per case, one denoiser pass, the 6-block BasicLayer and the
SwinTransformer3D backbone.
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

from oracle.make_golden_swin import install_timm_shim, REF, OUT   # noqa: E402
from oracle.model_weights import MODEL_CASES, fill_state_dict, model_config, model_problem, _rng   # noqa: E402


def load_weights(module):
    sd = module.state_dict()
    new = fill_state_dict(sd)
    for k, v in new.items():
        sd[k] = torch.from_numpy(v)
    module.load_state_dict(sd)
    module.eval()            # the reference's SwinTransformer3D.train() override returns None: do not chain
    return module


def main(only=None):
    install_timm_shim()
    sys.path.insert(0, REF)
    from dl_cs.mri.transforms import SenseModel
    from dl_cs.models import unrolledswin, unrolled
    from dl_cs.models import video_swin_transformer_mri_downsample as vs
    torch.manual_seed(0)

    for name in MODEL_CASES:
        if only and name not in only:
            continue
        cfg = model_config(name)
        swin = name.startswith("swin")
        mod = unrolledswin if swin else unrolled
        cls = mod.HalfQuadraticSplitting if cfg.MODEL.META_ARCHITECTURE == "modl" else mod.ProximalGradientDescent
        net = load_weights(cls(cfg))
        maps, mask, y = model_problem(name)
        A = SenseModel(torch.from_numpy(maps), weights=torch.from_numpy(mask))
        t0 = time.time()
        with torch.no_grad():
            aty = A(torch.from_numpy(y), adjoint=True)
            recon = net(torch.from_numpy(y), A)
            den = net.cnn_update[0](aty)                  # one denoiser pass on A^H y
        keys = sorted(net.state_dict().keys())
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), recon=recon.numpy(), denoised=den.numpy(),
                            n_params=np.int64(sum(p.numel() for p in net.parameters())),
                            keys=np.array(keys))
        print(name, tuple(recon.shape), "params", sum(p.numel() for p in net.parameters()), "%.1f s" % (time.time() - t0),
              "|recon| %.3f |aty| %.3f" % (recon.abs().mean(), aty.abs().mean()))

    if not only or "layers" in only:
        # the 6-block stage and the whole backbone at the denoiser's hard-coded configuration (swin3D.py:315)
        layer = load_weights(vs.BasicLayer(dim=160, depth=6, num_heads=8, window_size=(7, 8, 8), qkv_bias=True,
                                           drop_path=[0.0, 0.04, 0.08, 0.12, 0.16, 0.2]))
        x = torch.from_numpy(_rng("basic_layer_x").standard_normal((1, 160, 7, 12, 10)).astype(np.float32))
        with torch.no_grad():
            out = layer(x)
        np.savez_compressed(os.path.join(OUT, "model_basic_layer.npz"), out=out.numpy())
        print("basic_layer", tuple(out.shape))
        bb = load_weights(vs.SwinTransformer3D(in_chans=8, embed_dim=160, depths=[6], num_heads=[8], window_size=(7, 8, 8)))
        x = torch.from_numpy(_rng("backbone_x").standard_normal((1, 8, 27, 46, 40)).astype(np.float32))   # ragged: padded to the patch
        with torch.no_grad():
            out = bb(x)
        np.savez_compressed(os.path.join(OUT, "model_swin_backbone.npz"), out=out.numpy(), keys=np.array(sorted(bb.state_dict().keys())))
        print("swin_backbone", tuple(out.shape))


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
