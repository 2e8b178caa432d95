"""Unrolled networks (dl_swin_gan_b200.models, mirrors of dl_cs/models/{unrolledswin,unrolled,swin3D,resnet3d}.py).

CPU tests: module trees / state-dict keys and parameter counts equal the reference's (recorded in the fixtures that
oracle/make_golden_models.py produced from the unmodified reference), the ResNet solvers run on CPU tensors with the
oracle's SENSE operator injected and reproduce the reference reconstructions (the denoiser convolutions are torch ops).
GPU tests: the same networks with the B200 SenseModel and window-attention kernels against the reference outputs,
NRMSE = ||a - b|| / ||b|| <= 1e-3 (north_star's bar for full unrolled reconstructions).
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle.model_weights import MODEL_CASES, _rng, fill_state_dict, model_config, model_problem

NRMSE_TOL = 1e-3


def golden(name):
    return np.load(os.path.join(GOLDEN, f"model_{name}.npz"))


def build(name):
    from dl_swin_gan_b200.models import build_unrolled
    torch.manual_seed(0)
    net = build_unrolled(model_config(name))
    load_weights(net)
    return net.eval()


def load_weights(module):
    sd = module.state_dict()
    for k, v in fill_state_dict(sd).items():
        assert tuple(sd[k].shape) == v.shape, k
        sd[k] = torch.from_numpy(v)
    module.load_state_dict(sd)
    return module.eval()


class OracleSense:
    """The reference operator restated with torch.fft on CPU (oracle/sense_oracle_torch.py), duck-typed like SenseModel."""

    def __init__(self, maps, mask):
        from oracle import sense_oracle_torch as st
        self.st, self.maps, self.mask = st, maps, mask

    def __call__(self, data, adjoint=False):
        return self.st.adjoint(data, self.maps, self.mask) if adjoint else self.st.forward(data, self.maps, self.mask)


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_state_dict_keys_and_parameter_count_equal_reference(name):
    g = golden(name)
    net = build(name)
    assert sorted(net.state_dict().keys()) == [str(k) for k in g["keys"]]
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"])


def test_backbone_state_dict_keys_equal_reference():
    from dl_swin_gan_b200.swin.backbone import SwinTransformer3D
    bb = SwinTransformer3D(in_chans=8, embed_dim=160, depths=[6], num_heads=[8], window_size=(7, 8, 8))
    assert sorted(bb.state_dict().keys()) == [str(k) for k in golden("swin_backbone")["keys"]]
    with pytest.raises(NotImplementedError):
        SwinTransformer3D(depths=[2, 2])


@pytest.mark.parametrize("name", [n for n in sorted(MODEL_CASES) if n.startswith("resnet")])
def test_resnet_solvers_reproduce_reference_on_cpu(name):
    g = golden(name)
    net = build(name)
    maps, mask, y = (torch.from_numpy(a) for a in model_problem(name))
    A = OracleSense(maps, mask)
    with torch.no_grad():
        recon = net(y, A)
        den = net.cnn_update[0](A(y, adjoint=True))
    assert rel_l2(den.numpy(), g["denoised"]) < 1e-5
    assert rel_l2(recon.numpy(), g["recon"]) < 1e-4


def test_config_presets_and_yaml_loader(tmp_path):
    from dl_swin_gan_b200.config import get_cfg, load_cfg
    c = load_cfg("config_swin")
    p = c.MODEL.PARAMETERS
    assert (p.NUM_UNROLLS, p.NUM_SWINBLOCKS, p.NUM_FEATURES, p.NUM_EMAPS, p.FIX_STEP_SIZE) == (5, 1, 160, 2, True)
    assert p.CONV_BLOCK.KERNEL_SIZE == (3,) and p.CONV_BLOCK.COMPLEX is False and p.MODL.NUM_CG_STEPS == 10
    assert get_cfg().MODEL.PARAMETERS.CONV_BLOCK.COMPLEX is True          # the reference default (defaults.py)
    f = tmp_path / "c.yaml"
    f.write_text("MODEL:\n  PARAMETERS:\n    NUM_UNROLLS: 3\nAUG_TRAIN:\n  UNDERSAMPLE:\n    ACCELERATIONS: (10, 15)\n")
    c = load_cfg(str(f))
    assert c.MODEL.PARAMETERS.NUM_UNROLLS == 3 and c.AUG_TRAIN.UNDERSAMPLE.ACCELERATIONS == (10, 15)
    assert c.MODEL.PARAMETERS.NUM_RESBLOCKS == 2


def test_checkpointed_unrolls_match_plain_on_cpu():
    """GRAD_CHECKPOINT (unrolledswin.py:101-118): same outputs and gradients with and without re-computation."""
    import copy
    from dl_swin_gan_b200.models import build_unrolled
    cfg = model_config("resnet_hqs_example")
    cfg.MODEL.META_ARCHITECTURE = "dlespirit"
    cfg.MODEL.PARAMETERS.FIX_STEP_SIZE = False
    a = load_weights(build_unrolled(cfg)).train()
    cfg2 = copy.deepcopy(cfg)
    cfg2.MODEL.PARAMETERS.GRAD_CHECKPOINT = True
    b = load_weights(build_unrolled(cfg2)).train()
    maps, mask, y = (torch.from_numpy(t) for t in model_problem("resnet_hqs_example"))
    A = OracleSense(maps, mask)
    la = a(y, A).abs().sum()
    lb = b(y, A).abs().sum()
    la.backward(); lb.backward()
    assert torch.allclose(la, lb, rtol=1e-5)
    assert torch.allclose(a.step_size.grad, b.step_size.grad, rtol=1e-4)
    ga, gb = a.cnn_update[0].init_layer.layers[2].conv.weight.grad, b.cnn_update[0].init_layer.layers[2].conv.weight.grad
    assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-6)


# ----------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_unrolled_reconstruction_matches_reference_on_b200(name):
    """Full reconstruction through the B200 SenseModel (fused DC steps) and, for the Swin case, the B200 attention blocks."""
    from dl_swin_gan_b200.mri import SenseModel
    dev = torch.device("cuda:0")
    g = golden(name)
    net = build(name).to(dev)
    maps, mask, y = (torch.from_numpy(a).to(dev) for a in model_problem(name))
    A = SenseModel(maps, weights=mask)
    with torch.no_grad():
        recon = net(y, A)
        den = net.cnn_update[0](A(y, adjoint=True))
    e_den, e_rec = rel_l2(den.cpu().numpy(), g["denoised"]), rel_l2(recon.cpu().numpy(), g["recon"])
    print(f"{name}: NRMSE denoiser pass {e_den:.2e}, reconstruction {e_rec:.2e}")
    assert e_den < NRMSE_TOL and e_rec < NRMSE_TOL


@gpu
def test_basic_layer_six_blocks_matches_reference_on_b200():
    """The 6-block stage (alternating shift, video_swin...:417-437) at the denoiser's configuration vs the reference."""
    from dl_swin_gan_b200.swin.window_attention import BasicLayer
    dev = torch.device("cuda:0")
    layer = load_weights(BasicLayer(dim=160, depth=6, num_heads=8, window_size=(7, 8, 8), qkv_bias=True,
                                    drop_path=[0.0, 0.04, 0.08, 0.12, 0.16, 0.2])).to(dev)
    x = torch.from_numpy(_rng("basic_layer_x").standard_normal((1, 160, 7, 12, 10)).astype(np.float32)).to(dev)
    with torch.no_grad():
        out = layer(x)
    err = rel_l2(out.cpu().numpy(), golden("basic_layer")["out"])
    print(f"BasicLayer x6: rel-L2 {err:.2e}")
    assert err < 1e-3


@gpu
def test_swin_backbone_matches_reference_on_b200():
    from dl_swin_gan_b200.swin.backbone import SwinTransformer3D
    dev = torch.device("cuda:0")
    bb = load_weights(SwinTransformer3D(in_chans=8, embed_dim=160, depths=[6], num_heads=[8], window_size=(7, 8, 8))).to(dev)
    x = torch.from_numpy(_rng("backbone_x").standard_normal((1, 8, 27, 46, 40)).astype(np.float32)).to(dev)
    with torch.no_grad():
        out = bb(x)
    err = rel_l2(out.cpu().numpy(), golden("swin_backbone")["out"])
    print(f"SwinTransformer3D: rel-L2 {err:.2e}")
    assert err < 1e-3
