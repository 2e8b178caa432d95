"""dl_swin_gan_b200.mri.subsample (mirror of dl_cs/mri/subsample.py) against masks produced by the unmodified
reference generator (oracle/make_golden_masks.py -> tests/golden/vdkt_masks.npz): bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.make_golden_masks import CASES
from dl_swin_gan_b200.mri.subsample import VDktMaskFunc, subsample


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "vdkt_masks.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_vdkt_mask_equals_reference(golden, name):
    acc, pkx, pky, shape, seed = CASES[name]
    T, Y, X = shape[2:]
    kt = np.unpackbits(golden[f"{name}_kt"])[: T * Y].reshape(T, Y).astype(np.float32)
    want = kt[:, :, None] * golden[f"{name}_kx"].astype(np.float32)[None, None, :]
    got = VDktMaskFunc(acc, sim_partial_kx=pkx, sim_partial_ky=pky)(shape, seed=seed)
    assert got.dtype == torch.float32 and tuple(got.shape) == tuple(shape)
    assert np.array_equal(got.numpy()[0, 0], want)


def test_same_seed_same_mask_and_subsample_shape():
    f = VDktMaskFunc((12, 12))
    a, b = f((1, 1, 6, 64, 32), seed=5), f((1, 1, 6, 64, 32), seed=5)
    assert torch.equal(a, b)
    assert not torch.equal(a, f((1, 1, 6, 64, 32), seed=6))
    data = torch.ones(2, 3, 6, 64, 32, dtype=torch.complex64)
    masked, mask = subsample(data, f, seed=5, mode="3D")
    assert tuple(mask.shape) == (1, 1, 6, 64, 32) and torch.equal(masked, mask * data)
    with pytest.raises(ValueError):
        subsample(data, f, mode="4D")
