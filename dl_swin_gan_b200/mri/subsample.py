"""Host-side k-t sampling masks: mirror of dl_cs/mri/subsample.py (VDktMaskFunc :65-254, subsample :257-283).

This is the generator of the weights the SENSE operator is given in training and in the synthetic benchmarks
(SURVEY.md §8d): a variable-density ky-t pattern with a golden-ratio shift per cardiac phase, optional partial
readout (the first `sim_partial_kx` share of kx is empty) and partial phase encoding.  It is numpy on the host, like
the reference (it runs inside data-loader workers there); the arithmetic and the order of the random draws follow
the reference exactly, so that the same seed gives the same mask bit for bit (tests/test_subsample.py checks that
against masks produced by the unmodified reference, tests/golden/vdkt_mask_*.npz).
"""
from __future__ import annotations

from math import ceil, floor

import numpy as np
import torch

_EPS = 0.0001          # tolerance added before floor(), as in the reference (subsample.py:180)


class MaskFunc:
    """Base class: holds the acceleration range and a private numpy RandomState (subsample.py:13-33)."""

    def __init__(self, accelerations):
        self.accelerations = accelerations
        self.rng = np.random.RandomState()

    def choose_acceleration(self):
        lo, hi = self.accelerations[0], self.accelerations[1]
        return lo + (hi - lo) * self.rng.rand()          # one draw even when lo == hi


class VDktMaskFunc(MaskFunc):
    """Variable-density k-t mask; `__call__(out_shape=[1, 1, T, Y, X], seed)` -> float32 tensor of that shape."""

    golden_ratio = 0.618034

    def __init__(self, accelerations, sim_partial_kx=0.25, sim_partial_ky=0.0):
        super().__init__(accelerations)
        self.sim_partial_kx = sim_partial_kx
        self.sim_partial_ky = sim_partial_ky

    def __call__(self, out_shape, seed=None):
        self.rng.seed(seed)
        nkx, nky, nphases = out_shape[4], out_shape[3], out_shape[2]
        accel = self.choose_acceleration()
        if self.sim_partial_ky > 0.0:
            kt = self.vdkt_partial_ky(nky, nphases, accel, partialFourierFactor=self.sim_partial_ky)
        else:
            kt = self.vdkt(nky, nphases, accel)
        if self.sim_partial_kx > 0.0:
            full = np.repeat(kt[None], nkx, axis=0)            # [kx, ky, t]
            full[: int(self.sim_partial_kx * nkx)] = 0         # partial echo: the first kx samples are not acquired
            kt = full
        kt = kt.transpose(2, 1, 0) if kt.ndim == 3 else kt.transpose(1, 0)
        return torch.from_numpy(np.ascontiguousarray(kt).reshape(out_shape).astype(np.float32))

    def goldenratio_shift(self, accel, nt):
        return np.round(np.arange(0, nt) * self.golden_ratio * accel) % accel

    # ---- one cardiac phase of the pattern ---------------------------------------------------------------
    def _perturbed_lines(self, start, ny, accel, perturb, adhere):
        """Uniform lines start, start+accel, ... with the reference's random jitter (sequential: a line's own
        test sees the pull of the previous line), subsample.py:149-163."""
        y = np.arange(start, ny, accel)
        if perturb > 0:
            lo, hi = perturb * accel, ny - perturb * accel
            last = y.size - 1
            for i in range(y.size):
                if y[i] < lo or y[i] >= hi:
                    continue
                shift = perturb * accel * (self.rng.rand() - 0.5)
                y[i] += shift
                if i > 0:
                    y[i - 1] += adhere * shift
                if i < last:
                    y[i + 1] += adhere * shift
        return y

    @staticmethod
    def _fit_half(col, pos, radius, ny, upper):
        """Snap the warped positions `pos` (sorted by |pos|) of one half of k-space onto free grid lines of `col`,
        re-stretching the remaining positions whenever a collision consumed a line (subsample.py:173-214)."""
        stretch, origin = 1.0, 0.0
        edge = floor(pos[0] * radius + radius + _EPS)
        if not upper and col[edge] == 1:
            edge -= 1
            origin = pos[0]
            stretch = (radius + float(edge - radius)) / (radius * (1.0 - abs(origin)))
        for p in pos:
            loc = floor((origin + (p - origin) * stretch) * radius + radius + _EPS)
            loc = min(loc, ny - 1) if upper else max(loc, 0)
            if col[loc] == 0:
                col[loc] = 1
                edge = loc + 1
            else:
                col[edge] = 1
                origin = p
                stretch = (radius - float(edge - radius)) / (radius * (1 - abs(origin)))
                edge += 1 if upper else -1

    def vdkt(self, ny, nt, accel, nCal=1, vdDegree=1.5, vdFactor=None, perturbFactor=0.4, adhereFactor=0.33):
        """[ny, nt] float32 pattern (subsample.py:121-221; algorithm by Peng Lai, GE Healthcare, 2018)."""
        vdDegree = max(vdDegree, 0.0)
        perturbFactor = min(max(perturbFactor, 0.0), 1.0)
        adhereFactor = min(max(adhereFactor, 0.0), 1.0)
        nCal = max(nCal, 0)
        if vdFactor is None or vdFactor > accel:
            vdFactor = accel
        centre, radius = floor(ny / 2.0), (ny - 1) / 2.0
        if vdDegree > 0:
            vdFactor = vdFactor ** (1.0 / vdDegree)
        a_coef, b_coef = (vdFactor - 1.0) / vdFactor, 1.0 / vdFactor

        mask = np.zeros([ny, nt], np.float32)
        shifts = self.goldenratio_shift(accel, nt)
        for t in range(nt):
            y = self._perturbed_lines(shifts[t], ny, accel, perturbFactor, adhereFactor)
            y = (np.clip(y, 0, ny - 1) - radius) / radius                      # to [-1, 1]
            y = y * (a_coef * np.abs(y) + b_coef) ** vdDegree                  # denser towards the centre
            y = y[np.argsort(np.abs(y))]
            col = mask[:, t]
            self._fit_half(col, y[y >= 0], radius, ny, upper=True)
            self._fit_half(col, y[y < 0], radius, ny, upper=False)
        # calibration lines (an empty range for the default nCal = 1, as in the reference)
        mask[(centre - ceil(nCal / 2)):(centre + nCal - 1 - ceil(nCal / 2)), :] = 1
        return mask

    def vdkt_partial_ky(self, nky, nphases, tgt_accel, partialFourierFactor=0.25, tol=0.1, max_iter=10):
        """Bisection on the nominal acceleration until the partial-Fourier pattern reaches `tgt_accel` (subsample.py:223-254)."""
        lo, hi = 1.0, tgt_accel
        actual, it = 1.0, 0
        mask = None
        while abs(actual - tgt_accel) > tol and it < max_iter:
            trial = 0.5 * (lo + hi)
            mask = self.vdkt(nky, nphases, trial)
            cut = int(nky * partialFourierFactor)
            mask[(nky - cut):nky, 0::2] = 0                   # alternate halves of ky dropped on even / odd phases
            mask[0:cut, 1::2] = 0
            actual = (nky * nphases) / np.sum(mask)
            if actual < tgt_accel:
                lo = trial
            else:
                hi = trial
            it += 1
        return mask


def subsample(data, mask_func, seed=None, mode="2D"):
    """mask * data and the mask; mask shape (1, 1) + data.shape[2:4] ('2D') or data.shape[2:5] ('3D') (subsample.py:257-283)."""
    shape = tuple(data.shape)
    if mode == "2D":
        mask_shape = (1, 1) + shape[2:4]
    elif mode == "3D":
        mask_shape = (1, 1) + shape[2:5]
    else:
        raise ValueError("Only 2D and 3D undersampling masks are supported.")
    mask = mask_func(mask_shape, seed)
    return mask * data, mask
