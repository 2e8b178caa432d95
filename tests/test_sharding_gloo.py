"""World-size-2 gloo test (CPU) of the multi-GPU host logic: slice ranges, max-over-ranks timing reduction and
result gathering.  The per-rank work is the torch-CPU port of the reference path standing in for the CUDA operator:
sharded result == unsharded result bit for bit (no collective touches the data path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dl_swin_gan_b200.sharding import gather_slices, max_over_ranks, slice_range


def test_slice_ranges_partition_everything():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 64, 65):
            ranges = [slice_range(r, world, n) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slice_range(2, 2, 8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sense_oracle as so
        from oracle import sense_oracle_torch as st
        B = 5
        maps, x, mask = so.synthetic_problem(B, 1, 3, 2, 16, 12, seed=1000)
        tm, tw = torch.from_numpy(maps), torch.from_numpy(mask)
        y = st.forward(torch.from_numpy(x), tm, tw)
        b, e = slice_range(rank, world, B)
        local = st.dc_path(y[b:e], tm[b:e], tw[b:e], 3, -0.5)
        ms = max_over_ranks(10.0 + rank)                       # slowest rank defines the step time
        full = gather_slices(local, B)
        if rank == 0:
            ref = st.dc_path(y, tm, tw, 3, -0.5)
            np.save(os.path.join(out_dir, "ok.npy"), np.array([float(torch.equal(full, ref)), ms, float(full.shape[0])]))
        else:
            assert full is None and ms == 10.0 + world - 1
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_dc_path_equals_single_rank(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ok, ms, n = np.load(tmp_path / "ok.npy")
    assert ok == 1.0 and ms == 11.0 and n == 5
