"""Deterministic model weights and problems for the unrolled-network fixtures (shared by oracle/make_golden_models.py,
which runs the UNMODIFIED reference in the build container, and by tests/).  Test infrastructure only.

Weights are a function of the parameter NAME and shape (a numpy PCG64 stream seeded by a hash of the name), not of the
construction order, so the reference module tree and the mirror get identical tensors as long as their state-dict keys
and shapes agree - which is itself part of what the tests check.
"""
from __future__ import annotations

import zlib

import numpy as np

# name -> (preset, overrides of MODEL.PARAMETERS, (B, C, T, Y, X), acceleration)
MODEL_CASES = {
    "swin_pgd_2unroll": ("config_swin", {"NUM_UNROLLS": 2}, (1, 4, 20, 64, 64), 4),
    "swin_pgd_5unroll": ("config_swin", {}, (1, 4, 12, 48, 48), 4),          # all five unrolls; T = 12 -> depth 5 < 7: clamped window
    "resnet_pgd_example": ("example", {}, (1, 8, 20, 64, 48), 6),
    "resnet_pgd_recon": ("config_recon", {}, (2, 6, 12, 96, 64), 8),
    "resnet_hqs_example": ("example", {"MODL": {"NUM_CG_STEPS": 4, "MU": 0.1, "FIX_PENALTY": False}, "NUM_UNROLLS": 2},
                           (1, 4, 8, 32, 32), 4),
}


def _rng(name: str, salt: int = 0):
    return np.random.default_rng(zlib.crc32(name.encode()) + 7919 * salt)


def fill_state_dict(sd):
    """name -> numpy array for every floating-point entry of a state dict (dict of tensors / arrays with .shape)."""
    out = {}
    for name, t in sd.items():
        shape = tuple(t.shape)
        if name.endswith("relative_position_index"):
            continue                                     # integer buffer, identical by construction
        r = _rng(name)
        if name.endswith("step_size"):
            v = np.full(shape, -2.0)
        elif name.endswith("lamda"):
            v = np.full(shape, 0.1)
        elif name.endswith("relative_position_bias_table"):
            v = 0.2 * r.standard_normal(shape)
        elif ("norm" in name.split(".")[-2]) and name.endswith("weight"):
            v = 1.0 + 0.1 * r.standard_normal(shape)
        elif name.endswith("bias"):
            v = 0.05 * r.standard_normal(shape)
        else:                                            # conv / linear / transposed-conv weights: unit-gain fan-in scaling
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            if "patch_unembed" in name:                  # ConvTranspose3d: [in, out, kd, kh, kw], one input patch per output voxel
                fan_in = shape[0]
            v = r.standard_normal(shape) / np.sqrt(fan_in)
        out[name] = v.astype(np.float32)
    return out


def model_problem(name: str):
    """maps [B,E,C,1,Y,X], mask [B,1,T,Y,X], y [B,C,T,Y,X] = mask * A(x_true), x0 = None; E from the preset."""
    from dl_swin_gan_b200.config import PRESETS
    from dl_swin_gan_b200.mri.subsample import VDktMaskFunc
    preset, _, (B, C, T, Y, X), accel = MODEL_CASES[name]
    E = PRESETS[preset]["MODEL"]["PARAMETERS"]["NUM_EMAPS"]
    r = _rng(name, 1)

    def cn(*s):
        return ((r.standard_normal(s) + 1j * r.standard_normal(s)) / np.sqrt(2)).astype(np.complex64)

    maps = (cn(B, E, C, 1, Y, X) / np.sqrt(E * C)).astype(np.complex64)
    x_true = cn(B, E, T, Y, X)
    f = VDktMaskFunc((accel, accel), sim_partial_kx=0.25, sim_partial_ky=0.0)
    mask = np.concatenate([f((1, 1, T, Y, X), seed=100 + b).numpy() for b in range(B)], axis=0)
    img = (maps[:, :, :, 0][:, :, :, None] * x_true[:, :, None]).sum(1)               # [B,C,T,Y,X]
    y = (mask * np.fft.fftn(img, axes=(-2, -1), norm="ortho")).astype(np.complex64)
    return maps, mask.astype(np.float32), y


def model_config(name: str):
    from dl_swin_gan_b200.config import PRESETS, get_cfg
    import copy
    preset, over, _, _ = MODEL_CASES[name]
    tree = copy.deepcopy(PRESETS[preset])
    tree["MODEL"]["PARAMETERS"].update(over)
    if name.startswith("resnet_hqs"):
        tree["MODEL"]["META_ARCHITECTURE"] = "modl"
    return get_cfg(tree)
