"""Slice-sharded inference on the device(s): the reconstruction of a batch split by `slice_range` over ranks equals the
single-rank result bit for bit (slices never interact and the kernels' results do not depend on the batch composition).
With one GPU the shards run on the same device; with two or more every shard runs on its own device, as bench.py does
(one process per GPU, no data-path collective)."""
import numpy as np
import pytest
import torch

from oracle import sense_oracle as so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pgd_equals_single_rank_bit_for_bit(world):
    from dl_swin_gan_b200.mri import SenseModel
    from dl_swin_gan_b200.sharding import slice_range
    from dl_swin_gan_b200.unrolled import ProximalGradientDescent
    from dl_swin_gan_b200.mri.subsample import VDktMaskFunc
    B, E, C, T, Y, X = 7, 1, 4, 6, 192, 160
    maps, x, _ = so.synthetic_problem(B, E, C, T, Y, X, seed=3)
    f = VDktMaskFunc((12, 12), sim_partial_kx=0.25, sim_partial_ky=0.25)
    mask = np.concatenate([f((1, 1, T, Y, X), seed=40 + b).numpy() for b in range(B)], 0)
    ndev = torch.cuda.device_count()
    d0 = torch.device("cuda:0")
    tm, tw, tx = (torch.from_numpy(a) for a in (maps, mask, x))
    pgd = ProximalGradientDescent(num_unrolls=3, step_init=-2.0)
    with torch.no_grad():
        A = SenseModel(tm.to(d0), weights=tw.to(d0))
        y = A(tx.to(d0))
        whole = pgd.to(d0)(y, A).cpu()
        parts = []
        for rank in range(world):
            b0, b1 = slice_range(rank, world, B)
            dev = torch.device("cuda", rank % ndev)
            A_r = SenseModel(tm[b0:b1].to(dev), weights=tw[b0:b1].to(dev))
            parts.append(pgd.to(dev)(y[b0:b1].to(dev), A_r).cpu())
    assert sum(p.shape[0] for p in parts) == B
    assert torch.equal(torch.cat(parts, 0), whole)
