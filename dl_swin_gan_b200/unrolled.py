"""Host-side mirror of the unrolled solvers' data-consistency loop.

`ProximalGradientDescent.forward(y, A, x0)` of the reference (dl_cs/models/unrolledswin.py:91-122,
dl_cs/models/unrolled.py:86-117) with the operator `A` injected exactly as there.  The per-unroll
update `x = x + step_size * (A(A(x), adjoint=True) - ATy)` (unrolledswin.py:109) goes through the
fused `A.dc_step` entry when the operator offers it and through two operator calls otherwise, so a
reference `SenseModel` and the B200 `SenseModel` are interchangeable here.

The denoiser list `cnn_update` is a constructor argument: the Swin / ResNet denoisers are the callers'
business (cuDNN convolutions); with `None` the loop is the pure data-consistency path
that bench.py measures.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
from torch import nn


class ProximalGradientDescent(nn.Module):
    def __init__(self, num_unrolls: int = 5, cnn_update: Optional[Sequence[Callable]] = None,
                 fix_step_size: bool = True, step_init: float = -2.0) -> None:
        super().__init__()
        self.num_unrolls = num_unrolls
        self.cnn_update = nn.ModuleList(cnn_update) if cnn_update is not None and all(
            isinstance(m, nn.Module) for m in cnn_update) else cnn_update
        # PGD step size (unrolledswin.py:88-89)
        self.step_size = nn.Parameter(torch.tensor([step_init], dtype=torch.float32),
                                      requires_grad=(not fix_step_size))

    def forward(self, y, A, x0=None, ATy=None):
        """Reference signature `forward(y, A, x0)`; `ATy` (extension) takes A^H y when the caller already has it,
        e.g. computed on a transfer stream while the previous batch is reconstructed."""
        if ATy is None:
            ATy = A(y, adjoint=True)                   # unrolledswin.py:96
        xi = ATy if x0 is None else x0
        fused = getattr(A, "dc_step", None)
        for i in range(self.num_unrolls):
            if fused is not None:
                xi = fused(xi, ATy, self.step_size)
            else:
                xi = xi + self.step_size * (A(A(xi), adjoint=True) - ATy)
            if self.cnn_update is not None:
                xi = self.cnn_update[i](xi)
        return xi


class HalfQuadraticSplitting(nn.Module):
    """MoDL / half-quadratic-splitting solver, mirror of dl_cs/models/unrolledswin.py:125-172 (== unrolled.py:120-167):
    x <- CG( A^H A + lamda I,  A^H y + lamda * cnn_i(x) ).  The normal equations go through the fused
    `A.normal_reg` entry when the operator offers it."""

    def __init__(self, num_unrolls: int = 5, num_cg_iter: int = 10, cnn_update: Optional[Sequence[Callable]] = None,
                 fix_step_size: bool = True, lamda_init: float = 0.1) -> None:
        super().__init__()
        from .mri.algorithms import ConjugateGradient
        self._cg_cls = ConjugateGradient
        self.num_unrolls = num_unrolls
        self.num_cg_iter = num_cg_iter
        self.cnn_update = nn.ModuleList(cnn_update) if cnn_update is not None and all(
            isinstance(m, nn.Module) for m in cnn_update) else cnn_update
        self.lamda = nn.Parameter(torch.tensor([lamda_init], dtype=torch.float32), requires_grad=(not fix_step_size))

    def forward(self, y, A, x0=None):
        ATy = A(y, adjoint=True)
        xi = ATy if x0 is None else x0
        fused = getattr(A, "normal_reg", None)
        if fused is not None:
            model_normal = lambda m: fused(m, self.lamda)
        else:
            model_normal = lambda m: A(A(m), adjoint=True) + self.lamda * m
        cg_solve = self._cg_cls(model_normal, self.num_cg_iter)
        for i in range(self.num_unrolls):
            z = self.cnn_update[i](xi) if self.cnn_update is not None else xi
            xi = cg_solve(xi, ATy + self.lamda * z)
        return xi
