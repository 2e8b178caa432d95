"""Training slice on CPU (gloo, world size 2): the DDP step of dl_swin_gan_b200.train with the ResNet unrolled network and the
oracle's SENSE operator gives the same parameters after one Adam step as a single process on the concatenated batch
(the gradient all-reduce is the only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.model_weights import fill_state_dict, model_config, model_problem


class OracleSense:
    def __init__(self, maps, weights=None):
        from oracle import sense_oracle_torch as st
        self.st, self.maps, self.weights = st, maps, weights

    def __call__(self, data, adjoint=False):
        return (self.st.adjoint if adjoint else self.st.forward)(data, self.maps, self.weights)


def _cfg():
    cfg = model_config("resnet_hqs_example")
    cfg.MODEL.META_ARCHITECTURE = "dlespirit"
    cfg.MODEL.PARAMETERS.FIX_STEP_SIZE = False
    cfg.MODEL.PARAMETERS.NUM_FEATURES = 8
    return cfg


def _batch(b0, b1):
    maps, mask, y = (torch.from_numpy(a) for a in model_problem("resnet_pgd_recon"))      # 2 slices, E = 1
    maps = torch.cat([maps, maps.conj()], 1) / 2 ** 0.5                                    # two map sets for the E = 2 network
    A = OracleSense(maps, mask)
    init = A(y, adjoint=True)
    target = init * 0.5 + 0.1
    scale = torch.ones(2)
    sl = slice(b0, b1)
    return y[sl], mask[sl], maps[sl], init[sl], scale[sl], target[sl]


def _make(ddp):
    from dl_swin_gan_b200.models import build_unrolled
    from dl_swin_gan_b200.train import Trainer
    torch.manual_seed(0)
    net = build_unrolled(_cfg())
    sd = net.state_dict()
    sd.update({k: torch.from_numpy(v) for k, v in fill_state_dict(sd).items()})
    net.load_state_dict(sd)
    return Trainer(_cfg(), torch.device("cpu"), conv_precision="fp32", ddp=ddp, model=net, operator=OracleSense)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    tr = _make(ddp=True)
    loss = tr.step(_batch(rank, rank + 1))
    if rank == 0:
        torch.save({"loss": loss, "sd": {k: v.clone() for k, v in tr.net.state_dict().items()}}, out)
    dist.destroy_process_group()


def test_ddp_step_equals_large_batch_step(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.set_num_threads(2)
    # single process: the mean of the two per-slice losses has the gradient DDP averages
    tr = _make(ddp=False)
    b0, b1 = _batch(0, 1), _batch(1, 2)
    k, m, mp_, i, s, t = b0
    p0 = tr.model(k, OracleSense(mp_, m), i)
    k, m, mp_, i, s, t1 = b1
    p1 = tr.model(k, OracleSense(mp_, m), i)
    loss = 0.5 * (tr.loss(p0, b0[5]) + tr.loss(p1, t1))
    tr.opt.zero_grad()
    loss.backward()
    tr.opt.step()
    for key, v in tr.net.state_dict().items():
        assert torch.allclose(v, got["sd"][key], rtol=2e-4, atol=1e-6), key
    assert np.isfinite(got["loss"])


def test_losses_match_reference_definitions():
    from dl_swin_gan_b200.train import complex_l1, complex_l2
    g = torch.Generator().manual_seed(1)
    a = torch.view_as_complex(torch.randn(2, 2, 5, 6, 4, 2, generator=g))
    b = torch.view_as_complex(torch.randn(2, 2, 5, 6, 4, 2, generator=g))
    assert torch.allclose(complex_l1(a, b), (a - b).abs().mean())
    # the reference's weight (metrics.py:11-17) repeats the temporal std along axis 2 of the REDUCED tensor (the y axis) and
    # reshapes: W[b,c,t,y,x] = std[b,c,(t*Y + y) // T, x] - reproduced literally, not "fixed"
    std = torch.std(a, dim=2).abs()
    T_, Y_ = a.shape[2], a.shape[3]
    idx = (torch.arange(T_)[:, None] * Y_ + torch.arange(Y_)[None, :]) // T_
    W = std[:, :, idx, :]
    assert torch.allclose(complex_l1(a, b, True), (W * (a - b)).abs().mean(), rtol=1e-5)
    assert torch.allclose(complex_l2(a, b), ((a - b).abs() ** 2).mean().sqrt())
