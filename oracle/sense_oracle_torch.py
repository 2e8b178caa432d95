"""ORACLE / CPU BASELINE (test infrastructure, not product): the reference SENSE path restated with
torch CPU ops, i.e. the same arithmetic library (torch.fft, pocketfft/MKL, intra-op threads) the
reference's own CPU path uses (SenseModel inside CinePreprocess / DataTransform workers,
dl_cs/data/preprocess.py:140-164, scripts/reconstruct.py:146-163).

Used by bench.py only as the timed `cpu_baseline` / `--impl reference` arm (kind "port": the Python
reference itself cannot travel to the GPU box) and by tests/ as a second checker.  Restates
    SenseModel._forward_op / _adjoint_op   dl_cs/mri/transforms.py:84-98
    ProximalGradientDescent data-consistency loop   dl_cs/models/unrolled.py:96-117 (denoiser omitted)
Pinned against tests/golden by tests/test_oracle_golden.py::test_torch_port_matches_reference_outputs.
"""
from __future__ import annotations

import torch


def _fft2(data: torch.Tensor, ndims: int, inverse: bool) -> torch.Tensor:
    dims = [i for i in range(-1, -1 - ndims, -1)]
    if inverse:
        return torch.fft.ifftn(data, dim=dims, norm="ortho")
    return torch.fft.fftn(data, dim=dims, norm="ortho")


def ndims_of(maps: torch.Tensor) -> int:
    return len(list(maps[0, 0, 0, ...].squeeze().shape))


def forward(x, maps, weights=None):
    """k = W * F( sum_e maps * x ) ; x [B,E,T,Y,X], maps [B,E,C,1,Y,X]."""
    w = 1.0 if weights is None else weights
    coil = (x.unsqueeze(2) * maps).sum(1)
    return w * _fft2(coil, ndims_of(maps), False)


def adjoint(k, maps, weights=None):
    """x = sum_c conj(maps) * F^H( W * k )."""
    w = 1.0 if weights is None else weights
    img = _fft2(w * k, ndims_of(maps), True)
    return (img.unsqueeze(1) * torch.conj(maps)).sum(2)


def dc_path(y, maps, weights, num_unrolls: int, step: float, x0=None):
    """A^H y once, then num_unrolls data-consistency updates (unrolled.py:96-117 with the CNN omitted)."""
    aty = adjoint(y, maps, weights)
    x = aty if x0 is None else x0
    for _ in range(num_unrolls):
        x = x + step * (adjoint(forward(x, maps, weights), maps, weights) - aty)
    return x
