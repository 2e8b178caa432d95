"""Generate tests/golden/sense_*.npz by running the UNMODIFIED reference implementation.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):      python oracle/make_golden.py
Each fixture holds the outputs (inputs are regenerated from oracle/golden_inputs.py seeds) of
    dl_cs.mri.transforms.SenseModel / FFT        (/root/reference/dl_cs/mri/transforms.py)
    dl_cs.mri.utils.fftmod / sliding_window      (/root/reference/dl_cs/mri/utils.py)
computed with torch on CPU in complex64 (what the reference runs) and, for an error floor,
in complex128.  tests/test_oracle_golden.py pins oracle/sense_oracle.py to these vectors and
tests/test_sense_gpu.py compares the CUDA path with them on the B200.
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

from oracle.golden_inputs import CASES, make_inputs  # noqa: E402


def main():
    sys.path.insert(0, REF)
    from dl_cs.mri.transforms import SenseModel, FFT   # the reference, unmodified
    from dl_cs.mri import utils as ref_utils

    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    for name, (B, E, C, T, Y, X, wkind) in CASES.items():
        maps, x, k, aty, w = make_inputs(name, B, E, C, T, Y, X, wkind)
        out = {"step": np.float32(-2.0)}
        small = k.size <= 5000
        for tag, cdt in (("c64", torch.complex64), ("c128", torch.complex128)):
            if tag == "c128" and not small:
                continue
            tm = torch.from_numpy(maps).to(cdt)
            tw = None if w is None else torch.from_numpy(w).to(torch.float32 if cdt == torch.complex64 else torch.float64)
            A = SenseModel(tm, weights=tw)
            tx, tk, ta = (torch.from_numpy(a).to(cdt) for a in (x, k, aty))
            fwd = A(tx)
            adj = A(tk, adjoint=True)
            nrm = A(A(tx), adjoint=True)
            dc = tx + float(out["step"]) * (nrm - ta)            # unrolledswin.py:109
            out[f"fwd_{tag}"] = fwd.numpy()
            out[f"adj_{tag}"] = adj.numpy()
            out[f"normal_{tag}"] = nrm.numpy()
            out[f"dc_{tag}"] = dc.numpy()
            out["ndims"] = np.int32(A.fft.ndims)
        np.savez_compressed(os.path.join(OUT, f"sense_{name}.npz"), **out)
        print(name, {k_: v.shape for k_, v in out.items() if hasattr(v, "shape") and v.ndim > 0})

    # FFT module (centered / adjoint) and the helpers either side of the operator
    rng = np.random.default_rng(77)
    img = ((rng.standard_normal((3, 12, 10)) + 1j * rng.standard_normal((3, 12, 10)))).astype(np.complex64)
    img_even = ((rng.standard_normal((2, 8, 6)) + 1j * rng.standard_normal((2, 8, 6)))).astype(np.complex64)
    series = ((rng.standard_normal((1, 2, 6, 4, 5)) + 1j * rng.standard_normal((1, 2, 6, 4, 5)))).astype(np.complex64)
    series[:, :, ::3] = 0   # unsampled frames exercise get_mask / time_average
    f2, f1 = FFT(2), FFT(1)
    timg = torch.from_numpy(img)
    misc = {
        "img": img, "img_even": img_even, "series": series,
        "fft2": f2(timg).numpy(), "ifft2": f2(timg, adjoint=True).numpy(),
        "fft2_centered": f2(timg, centered=True).numpy(),
        "ifft2_centered": f2(timg, adjoint=True, centered=True).numpy(),
        "fft1": f1(timg).numpy(),
        "fftmod": ref_utils.fftmod(torch.from_numpy(img_even.copy())).numpy(),
        "sliding_window5": ref_utils.sliding_window(torch.from_numpy(series), 2, 5).numpy(),
        "get_mask": ref_utils.get_mask(torch.from_numpy(series)).numpy(),
    }
    # conjugate gradient on the regularised normal equations (MoDL / HQS), reference code on the "tiny_odd" and "pow2_32" inputs
    from dl_cs.mri.algorithms import ConjugateGradient
    for name in ("tiny_odd", "pow2_32"):
        B, E, C, T, Y, X, wkind = CASES[name]
        maps, x, k, aty, w = make_inputs(name, B, E, C, T, Y, X, wkind)
        A = SenseModel(torch.from_numpy(maps), weights=None if w is None else torch.from_numpy(w))
        lam = 0.1
        model_normal = lambda m: A(A(m), adjoint=True) + lam * m
        ty = torch.from_numpy(k)
        ATy = A(ty, adjoint=True)
        cg = ConjugateGradient(model_normal, 6)
        misc[f"cg6_{name}"] = cg(torch.from_numpy(x), ATy + lam * torch.from_numpy(aty)).numpy()
        xi = ATy
        for _ in range(2):                                   # HalfQuadraticSplitting.forward with cnn_update = identity
            xi = ConjugateGradient(model_normal, 4)(xi, ATy + lam * xi)
        misc[f"hqs2x4_{name}"] = xi.numpy()
    # DataTransform.__call__ (scripts/reconstruct.py:132-171): the script itself needs pytorch_lightning, so the same
    # reference functions are called here in the same order on an undersampled cine-like series (even sizes: fftmod)
    rng = np.random.default_rng(99)
    C_, T_, Y_, X_ = 3, 7, 12, 10
    ksp = (rng.standard_normal((C_, T_, Y_, X_)) + 1j * rng.standard_normal((C_, T_, Y_, X_))).astype(np.complex64)
    keep = rng.random((1, T_, Y_, 1)) < 0.4
    ksp = ksp * keep
    mp = (rng.standard_normal((2, C_, 1, Y_, X_)) + 1j * rng.standard_normal((2, C_, 1, Y_, X_))).astype(np.complex64) / 2
    misc["dt_kspace_in"], misc["dt_maps_in"] = ksp, mp
    for slwin in (False, True):
        kspace = torch.from_numpy(ksp.copy()).unsqueeze(0)
        maps_t = torch.from_numpy(mp.copy()).unsqueeze(0)
        mask = ref_utils.get_mask(kspace)[:, 0, None]
        kspace = ref_utils.fftmod(kspace)
        maps_t = ref_utils.fftmod(maps_t)
        A = SenseModel(maps_t, weights=None)
        averaged = ref_utils.time_average(kspace, dim=2)
        image = A(averaged, adjoint=True)
        magnitude = torch.abs(image).reshape(-1)
        kk = int(round(0.05 * magnitude.numel()))
        scale = torch.min(torch.topk(magnitude, kk).values)
        kspace /= scale
        init_kspace = ref_utils.sliding_window(kspace, dim=2, window_size=5) if slwin else kspace
        init_image = A(init_kspace, adjoint=True)
        tag = "slwin" if slwin else "plain"
        misc[f"dt_{tag}_kspace"] = kspace.squeeze(0).numpy()
        misc[f"dt_{tag}_maps"] = maps_t.squeeze(0).numpy()
        misc[f"dt_{tag}_mask"] = mask.squeeze(0).numpy()
        misc[f"dt_{tag}_init"] = init_image.squeeze(0).numpy()
        misc[f"dt_{tag}_scale"] = np.float32(scale.item())
        misc["time_average"] = averaged.numpy()
    np.savez_compressed(os.path.join(OUT, "sense_misc.npz"), **misc)
    print("misc", list(misc))


if __name__ == "__main__":
    main()
