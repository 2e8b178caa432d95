"""Training slice of the unrolled networks (BASELINE.json configs[2], config_swin.yaml): the loop the reference runs inside
PyTorch Lightning - `LitUnrolled.training_step` (scripts/train_swin.py:111-136), Adam(lr 1e-4) + StepLR (`:155-166`), DDP when
several devices are given (`:253-261`) - as a plain torch loop: one process per GPU, `DistributedDataParallel` over NCCL
(NVLink 5 / NVSwitch) for the gradient all-reduce and nothing else on the wire.

Batches follow the contract of `CinePreprocess.__call__` (dl_cs/data/preprocess.py:128-180): (kspace, mask, maps,
initial_guess, scale, target) with target = A^H k_full, the VDkt k-t mask of the config's acceleration range, kspace = mask *
k_full, readout cropped to AUG_TRAIN.CROP_READOUT columns.  There is no dataset offline, so `synthetic_batch` draws the same
tensors from seeds (the HDF5 loader, dl_cs/data/dataset.py, stays out of scope).

    loss = complex_l1(target, model(y=kspace, A=SenseModel(maps, weights=mask), x0=initial_guess))     (utils/metrics.py:31-39)

The SENSE operator inside the loop is the B200 one (autograd: VJP(A) = A^H); the Swin blocks differentiate through their
torch route (the fused inference kernels are forward-only; a hand-written attention backward is the open item of DESIGN.md).
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch
from torch import nn


def complex_l1(ref: torch.Tensor, pred: torch.Tensor, weight: bool = False) -> torch.Tensor:
    """`metrics.l1` (dl_cs/utils/metrics.py:31-39): mean |W (ref - pred)|, W = temporal standard deviation when weighted."""
    if weight:
        nt = ref.shape[2]
        W = torch.reshape(torch.repeat_interleave(torch.abs(torch.std(ref, dim=2)), nt, dim=2), ref.shape)
        return torch.mean(torch.abs(W * (ref - pred)))
    return torch.mean(torch.abs(ref - pred))


def complex_l2(ref: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """`metrics.l2` (:20-28), unweighted: root mean squared magnitude of the difference."""
    return torch.sqrt(torch.mean(torch.abs(ref - pred) ** 2))


def synthetic_batch(config, device, batch: int, seed: int, coils: int = 8, frames: int = 20, ny: int = 192,
                    nx: Optional[int] = None, operator=None):
    """(kspace, mask, maps, initial_guess, scale, target) with the shapes and semantics of CinePreprocess' output.
    `operator(maps, weights)` builds the SENSE operator (default: the B200 SenseModel)."""
    from .mri.subsample import VDktMaskFunc
    if operator is None:
        from .mri import SenseModel as operator
    E = config.MODEL.PARAMETERS.NUM_EMAPS
    nx = nx or config.AUG_TRAIN.CROP_READOUT                     # training crops the readout (config_swin.yaml:44)
    us = config.AUG_TRAIN.UNDERSAMPLE
    g = torch.Generator(device=device).manual_seed(seed)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=device)) * (0.5 ** 0.5)

    maps = cn(batch, E, coils, 1, ny, nx) / (E * coils) ** 0.5
    image = cn(batch, E, frames, ny, nx)
    k_full = operator(maps, weights=None)(image)
    target = operator(maps, weights=None)(k_full, adjoint=True)                  # preprocess.py:143
    f = VDktMaskFunc(tuple(us.ACCELERATIONS), sim_partial_kx=us.PARTIAL_KX, sim_partial_ky=us.PARTIAL_KY)
    mask = torch.cat([f((1, 1, frames, ny, nx), seed=seed * 1000 + b) for b in range(batch)], 0).to(device)
    kspace = mask * k_full
    A = operator(maps, weights=mask)
    init = A(kspace, adjoint=True)
    scale = torch.quantile(init.abs().flatten(1)[:, :: max(1, init[0].numel() // 65536)], 0.95, dim=1)   # 95th percentile (preprocess.py:149-153)
    s = scale.view(-1, 1, 1, 1, 1)
    return kspace / s, mask, maps, init / s, scale, target / s


class Trainer:
    """One optimisation step = forward through the unrolled network, complex-L1 loss, backward (DDP all-reduce), Adam."""

    def __init__(self, config, device, conv_precision: str = "bf16", ddp: bool = False, model: Optional[nn.Module] = None,
                 operator=None, autocast: Optional[bool] = None):
        from .models import build_unrolled, set_conv_precision
        set_conv_precision(conv_precision)
        self.config, self.device = config, device
        # BASELINE configs[2] trains in bf16: with bf16 convolutions the denoiser's Linear layers run under torch.autocast too
        # (the SENSE operator is complex64 and the attention core fp32 either way; parameters and Adam states stay fp32)
        self.autocast = (conv_precision == "bf16" and device.type == "cuda") if autocast is None else bool(autocast)
        self.operator = operator
        if operator is None:
            from .mri import SenseModel
            self.operator = SenseModel
        net = (model if model is not None else build_unrolled(config)).to(device).train()
        # `SwinTransformer3D.norm` exists in the reference's module tree but never takes part in forward (video_swin...:661,
        # 735-756): it would never receive a gradient, which DistributedDataParallel treats as an error - freeze it
        for name, p in net.named_parameters():
            if name.endswith("transformer.norm.weight") or name.endswith("transformer.norm.bias"):
                p.requires_grad_(False)
        self.net = net
        self.model = (nn.parallel.DistributedDataParallel(net, device_ids=[device.index] if device.type == "cuda" else None,
                                                          gradient_as_bucket_view=True) if ddp else net)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=config.OPTIMIZER.ADAM.LR)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=getattr(getattr(config, "LR_SCHEDULER", None), "STEP_SIZE", 10 ** 6),
                                                     gamma=getattr(getattr(config, "LR_SCHEDULER", None), "GAMMA", 1.0))
        self.loss_name = config.MODEL.RECON_LOSS.NAME
        self.loss_weight = bool(getattr(config.MODEL.RECON_LOSS, "LOSS_WEIGHT", False))

    def loss(self, pred, target):
        if self.loss_name == "complex_l2":
            return complex_l2(target, pred)
        return complex_l1(target, pred, self.loss_weight)

    def step(self, batch) -> float:
        kspace, mask, maps, init, scale, target = batch
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=self.autocast):
            pred = self.model(kspace, self.operator(maps, weights=mask), init)
        if self.config.MODEL.RECON_LOSS.RENORMALIZE_DATA:
            s = scale.view(-1, 1, 1, 1, 1)
            pred, target = pred * s, target * s
        loss = self.loss(pred, target)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def grad_bytes(self) -> int:
        return sum(p.numel() * p.element_size() for p in self.net.parameters() if p.requires_grad)


def time_allreduce(nbytes: int, device, iters: int = 5) -> Optional[float]:
    """Milliseconds of one NCCL all-reduce of a gradient-sized fp32 buffer (what DDP overlaps with backward)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=device)
    for _ in range(2):
        dist.all_reduce(buf)
    torch.cuda.synchronize(device)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        dist.all_reduce(buf)
    e.record()
    torch.cuda.synchronize(device)
    return s.elapsed_time(e) / iters
