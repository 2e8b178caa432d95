"""Configuration stand-in for dl_cs/config (an fvcore CfgNode tree, dl_cs/config/defaults.py:17-209): attribute access
`cfg.MODEL.PARAMETERS.NUM_UNROLLS` over nested namespaces, the reference's default values for the keys the unrolled
models read (dl_cs/models/unrolledswin.py:23-35, unrolled.py:23-32), yaml overrides merged on top, and the three
configurations BASELINE.json names as presets (their MODEL sections, copied value by value from
configs/config_swin.yaml:2-26, configs/config_recon.yaml:2-20, configs/basic/example.yaml:2-19).
"""
from __future__ import annotations

import copy
from types import SimpleNamespace
from typing import Any, Dict, Optional

_DEFAULTS: Dict[str, Any] = {
    "MODEL": {
        "DEVICE": "cuda", "NAME": "DLESPIRiT3D", "MODEL_TYPE": "RES", "WEIGHTS": "", "META_ARCHITECTURE": "dlespirit",
        "STRATEGY": "standard",
        "PARAMETERS": {
            "NUM_UNROLLS": 5, "NUM_RESBLOCKS": 2, "NUM_SWINBLOCKS": 2, "NUM_HEADS": 6, "RR": 16, "NUM_FEATURES": 256,
            "DROPOUT": 0.0, "NUM_EMAPS": 2, "NUM_COILS": 8, "FIX_STEP_SIZE": False, "SHARE_WEIGHTS": False,
            "SLWIN_INIT": False, "GRAD_CHECKPOINT": False,
            "MODL": {"NUM_CG_STEPS": 10, "MU": 0.1, "FIX_PENALTY": False},
            "WINDOW_SIZE": (4, 4), "NUM_HEAD": 4, "NUM_LAYERS": 4,
            "CONV_BLOCK": {"KERNEL_SIZE": (3,), "CIRCULAR_PAD": True, "ACTIVATION": "relu", "NORM": "none",
                           "SEPARABLE": True, "COMPLEX": True},
        },
        "RECON_LOSS": {"NAME": "complex_l1", "RENORMALIZE_DATA": True, "LOSS_WEIGHT": False},
    },
    "AUG_TRAIN": {"CROP_READOUT": 64, "UNDERSAMPLE": {"NAME": "VDktMaskFunc", "ACCELERATIONS": (10, 15),
                                                     "PARTIAL_KX": 0.25, "PARTIAL_KY": 0.25}},
    "OPTIMIZER": {"NAME": "Adam", "MAX_EPOCHS": 1000, "GRAD_ACCUM_ITERS": 1, "ADAM": {"LR": 0.0001}},
    "DATALOADER": {"TRAIN_BATCH_SIZE": 1, "VAL_BATCH_SIZE": 1},
    "SEED": 1000,
}

_CONV_REAL = {"ACTIVATION": "relu", "NORM": "none", "CIRCULAR_PAD": True, "COMPLEX": False}
PRESETS: Dict[str, Dict[str, Any]] = {
    # configs/config_swin.yaml: Swin-denoiser unrolled network (training config, BASELINE configs[2])
    "config_swin": {"MODEL": {"MODEL_TYPE": "SWIN", "META_ARCHITECTURE": "dlespirit", "STRATEGY": "standard", "PARAMETERS": {
        "NUM_UNROLLS": 5, "NUM_RESBLOCKS": 2, "NUM_SWINBLOCKS": 1, "NUM_FEATURES": 160, "NUM_EMAPS": 2, "SHARE_WEIGHTS": False,
        "FIX_STEP_SIZE": True, "SLWIN_INIT": True, "GRAD_CHECKPOINT": False, "CONV_BLOCK": _CONV_REAL},
        "RECON_LOSS": {"NAME": "complex_l1", "RENORMALIZE_DATA": False, "LOSS_WEIGHT": False}}},
    # configs/config_recon.yaml: ResNet unrolled network, one map set (inference config, BASELINE configs[1])
    "config_recon": {"MODEL": {"META_ARCHITECTURE": "dlespirit", "PARAMETERS": {
        "NUM_UNROLLS": 5, "NUM_RESBLOCKS": 2, "NUM_FEATURES": 64, "NUM_EMAPS": 1, "SHARE_WEIGHTS": False, "FIX_STEP_SIZE": True,
        "SLWIN_INIT": False, "GRAD_CHECKPOINT": False, "CONV_BLOCK": _CONV_REAL},
        "RECON_LOSS": {"NAME": "complex_l1", "RENORMALIZE_DATA": False, "LOSS_WEIGHT": True}}},
    # configs/basic/example.yaml: ResNet unrolled network, two map sets (BASELINE configs[0])
    "example": {"MODEL": {"META_ARCHITECTURE": "dlespirit", "PARAMETERS": {
        "NUM_UNROLLS": 5, "NUM_RESBLOCKS": 2, "NUM_FEATURES": 64, "NUM_EMAPS": 2, "SHARE_WEIGHTS": False, "FIX_STEP_SIZE": True,
        "SLWIN_INIT": True, "GRAD_CHECKPOINT": False, "CONV_BLOCK": _CONV_REAL},
        "RECON_LOSS": {"NAME": "complex_l1", "RENORMALIZE_DATA": False}}},
}


def _merge(base: Dict[str, Any], over: Dict[str, Any]) -> Dict[str, Any]:
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _merge(base[k], v)
        else:
            base[k] = copy.deepcopy(v)
    return base


def _ns(d):
    return SimpleNamespace(**{k: _ns(v) if isinstance(v, dict) else v for k, v in d.items()})


def get_cfg(overrides: Optional[Dict[str, Any]] = None) -> SimpleNamespace:
    """Defaults (defaults.py) with `overrides` merged on top, as a nested namespace."""
    tree = copy.deepcopy(_DEFAULTS)
    if overrides:
        _merge(tree, overrides)
    return _ns(tree)


def load_cfg(name_or_path: str) -> SimpleNamespace:
    """A preset name ('config_swin', 'config_recon', 'example') or the path of a yaml file in the reference's format."""
    if name_or_path in PRESETS:
        return get_cfg(PRESETS[name_or_path])
    import yaml
    with open(name_or_path) as f:
        text = f.read()
    # the reference's yaml files write tuples as "(10, 15)": fvcore evaluates those strings; do the same for that one form
    over = yaml.safe_load(text) or {}

    def fix(node):
        if isinstance(node, dict):
            return {k: fix(v) for k, v in node.items()}
        if isinstance(node, str) and node.startswith("(") and node.endswith(")"):
            import ast
            try:
                return ast.literal_eval(node)
            except (ValueError, SyntaxError):
                return node
        return node

    return get_cfg(fix(over))
