"""Generate tests/golden/vdkt_masks.npz by running the UNMODIFIED reference mask generator
(/root/reference/dl_cs/mri/subsample.py: VDktMaskFunc).  Build container only:
    python oracle/make_golden_masks.py
Masks are stored bit-packed along kx-collapsed (ky, t) planes plus the kx profile, so the fixture stays small;
tests/test_subsample.py rebuilds the full masks and compares dl_swin_gan_b200.mri.subsample bit for bit.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name -> (accelerations, sim_partial_kx, sim_partial_ky, shape, seed)
CASES = {
    "cine_r12": ((12, 12), 0.25, 0.0, (1, 1, 20, 192, 160), 1000),
    "cine_r12_pky": ((12, 12), 0.25, 0.25, (1, 1, 20, 192, 160), 1000),
    "cine_r10_15": ((10, 15), 0.25, 0.25, (1, 1, 20, 180, 64), 7),          # config_swin.yaml ACCELERATIONS range, odd-ish ky
    "perf_r8": ((8, 8), 0.25, 0.0, (1, 1, 40, 256, 256), 1001),
    "small_r4": ((4, 4), 0.5, 0.0, (1, 1, 5, 33, 8), 3),
    "seed_none_shape": ((6, 6), 0.25, 0.0, (1, 1, 3, 48, 16), 12345),
}


def main():
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from dl_cs.mri.subsample import VDktMaskFunc        # the reference, unmodified
    out = {}
    for name, (acc, pkx, pky, shape, seed) in CASES.items():
        m = VDktMaskFunc(acc, sim_partial_kx=pkx, sim_partial_ky=pky)(shape, seed=seed).numpy()
        assert m.shape == tuple(shape) and m.dtype == np.float32
        assert set(np.unique(m)) <= {0.0, 1.0}
        kt = m[0, 0].max(axis=-1)                           # [T, Y]: 1 where the line is sampled at any kx
        kx = m[0, 0].max(axis=(0, 1))                       # [X]
        assert np.array_equal(m[0, 0], kt[:, :, None] * kx[None, None, :])     # the mask is separable
        out[f"{name}_kt"] = np.packbits(kt.astype(np.uint8))
        out[f"{name}_kx"] = kx.astype(np.uint8)
        print(name, shape, "sampled fraction %.4f" % m.mean(), "lines per phase", kt.sum(1).astype(int).tolist()[:8])
    np.savez_compressed(os.path.join(OUT, "vdkt_masks.npz"), **out)


if __name__ == "__main__":
    main()
