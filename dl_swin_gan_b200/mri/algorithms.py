"""Mirror of dl_cs/mri/algorithms.py::ConjugateGradient (:11-73) for operators living on the GPU.

Same constructor and call signature (`ConjugateGradient(A, num_iter, dbprint)(x, y)`), same arithmetic order; the dot
products stay on the device as 0-dim tensors (no host synchronisation), and — like the reference — they reduce over the
WHOLE batch, so slices are coupled through alpha / beta (SURVEY.md §8e caveat): shard only whole CG problems.
`A` is any callable; for the MoDL / HQS normal equations use `SenseModel.normal_reg` (A^H A x + lamda x in one fused chain).
"""
from __future__ import annotations

import torch
from torch import nn


class ConjugateGradient(nn.Module):
    def __init__(self, A, num_iter, dbprint=False):
        super().__init__()
        self.A = A
        self.num_iter = num_iter
        self.dbprint = dbprint

    def zdot(self, x1, x2):
        """Complex dot product between tensors x1 and x2 (:32-36)."""
        return torch.sum(x1.conj() * x2)

    def zdot_single(self, x):
        return self.zdot(x, x).real

    def forward(self, x, y):
        r = y - self.A(x)
        rsold = self.zdot_single(r)
        p = r
        for i in range(self.num_iter):
            Ap = self.A(p)
            pAp = self.zdot(p, Ap)
            alpha = rsold / pAp
            x = x + alpha * p
            r = r - alpha * Ap
            rsnew = self.zdot_single(r)
            beta = rsnew / rsold
            rsold = rsnew
            p = beta * p + r
            if self.dbprint:
                print(f"CG Iteration {i}: {rsnew}")
        return x
