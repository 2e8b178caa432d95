"""Config-driven unrolled networks: mirror of dl_cs/models/unrolledswin.py (UnrolledSwinNet :15-74, ProximalGradientDescent
:77-122, HalfQuadraticSplitting :125-172) and dl_cs/models/unrolled.py (UnrolledNet :15-80, same solvers on the ResNet
denoiser).  Same constructor (`config`), same attribute and parameter names (`cnn_update.N.*`, `step_size`, `lamda`), same
forward(y, A, x0): a reference state dict loads, and either SENSE operator (the reference's or the B200 one) can be injected.
The data-consistency update goes through the operator's fused `dc_step` / `normal_reg` entries when it has them
(dl_swin_gan_b200/unrolled.py holds that loop); GRAD_CHECKPOINT wraps each unroll in torch.utils.checkpoint as the
reference does (unrolledswin.py:117-118).
"""
from __future__ import annotations

import torch
import torch.utils.checkpoint as cp
from torch import nn

from ..mri.algorithms import ConjugateGradient
from .resnet3d import ResNet
from .swin3d import SwinTransformer3DNet


class _Unrolled(nn.Module):
    def __init__(self, config):
        super().__init__()
        p = config.MODEL.PARAMETERS
        self.num_unrolls = p.NUM_UNROLLS
        self.num_features = p.NUM_FEATURES
        self.kernel_size = p.CONV_BLOCK.KERNEL_SIZE[0]
        self.num_emaps = p.NUM_EMAPS
        self.share_weights = p.SHARE_WEIGHTS
        self.fix_step_size = p.FIX_STEP_SIZE
        self.use_complex_layers = p.CONV_BLOCK.COMPLEX
        self.circular_pad = p.CONV_BLOCK.CIRCULAR_PAD
        self.do_checkpoint = p.GRAD_CHECKPOINT
        self._read(p)
        self.cnn_update = self.init_nets()

    def _read(self, p):
        pass

    def _make_net(self, in_chans):
        raise NotImplementedError

    def init_nets(self):
        in_chans = self.num_emaps if self.use_complex_layers else 2 * self.num_emaps
        if self.share_weights:
            return nn.ModuleList([self._make_net(in_chans)] * self.num_unrolls)       # the same network at every unroll
        return nn.ModuleList([self._make_net(in_chans) for _ in range(self.num_unrolls)])

    def _run(self, update, xi):
        for i in range(self.num_unrolls):
            xi = cp.checkpoint(update(i), xi, use_reentrant=False) if self.do_checkpoint else update(i)(xi)
        return xi


class UnrolledSwinNet(_Unrolled):
    def _read(self, p):
        self.num_swinblocks = p.NUM_SWINBLOCKS
        self.window_size = p.WINDOW_SIZE
        self.num_heads = p.NUM_HEAD
        self.num_layers = p.NUM_LAYERS

    def _make_net(self, in_chans):
        return SwinTransformer3DNet(in_chans=in_chans, chans=self.num_features, num_swinblocks=self.num_swinblocks,
                                    use_complex_layers=self.use_complex_layers, window_size=self.window_size,
                                    kernel_size=self.kernel_size, circular_pad=self.circular_pad, num_heads=self.num_heads,
                                    num_layers=self.num_layers)


class UnrolledNet(_Unrolled):
    def _read(self, p):
        self.num_resblocks = p.NUM_RESBLOCKS

    def _make_net(self, in_chans):
        return ResNet(in_chans=in_chans, chans=self.num_features, num_resblocks=self.num_resblocks,
                      use_complex_layers=self.use_complex_layers, kernel_size=self.kernel_size, circular_pad=self.circular_pad)


class _PGD:
    """x <- cnn_i( x + step_size * (A^H A x - A^H y) )   (unrolledswin.py:105-112)"""

    def _init_pgd(self):
        self.step_size = nn.Parameter(torch.tensor([-2.0], dtype=torch.float32), requires_grad=(not self.fix_step_size))

    def forward(self, y, A, x0=None):
        ATy = A(y, adjoint=True)
        xi = ATy if x0 is None else x0
        if self.training and self.do_checkpoint:
            xi.requires_grad_()
        fused = getattr(A, "dc_step", None)

        def update(i):
            def update_fn(x):
                if fused is not None:
                    x = fused(x, ATy, self.step_size)
                else:
                    x = x + self.step_size * (A(A(x), adjoint=True) - ATy)
                return self.cnn_update[i](x)
            return update_fn

        return self._run(update, xi)


class _HQS:
    """z = cnn_i(x);  x <- CG(A^H A + lamda I, A^H y + lamda z)   (unrolledswin.py:139-172)"""

    def _init_hqs(self, config):
        self.num_cg_iter = config.MODEL.PARAMETERS.MODL.NUM_CG_STEPS
        self.lamda = nn.Parameter(torch.tensor([0.1], dtype=torch.float32), requires_grad=(not self.fix_step_size))

    def forward(self, y, A, x0=None):
        ATy = A(y, adjoint=True)
        xi = ATy if x0 is None else x0
        if self.training and self.do_checkpoint:
            xi.requires_grad_()
        fused = getattr(A, "normal_reg", None)
        if fused is not None:
            model_normal = lambda m: fused(m, self.lamda)
        else:
            model_normal = lambda m: A(A(m), adjoint=True) + self.lamda * m
        cg_solve = ConjugateGradient(model_normal, self.num_cg_iter)

        def update(i):
            def update_fn(x):
                z = self.cnn_update[i](x)
                return cg_solve(x, ATy + self.lamda * z)
            return update_fn

        return self._run(update, xi)


class SwinProximalGradientDescent(_PGD, UnrolledSwinNet):
    """dl_cs.models.unrolledswin.ProximalGradientDescent"""

    def __init__(self, config):
        UnrolledSwinNet.__init__(self, config)
        self._init_pgd()


class SwinHalfQuadraticSplitting(_HQS, UnrolledSwinNet):
    """dl_cs.models.unrolledswin.HalfQuadraticSplitting"""

    def __init__(self, config):
        UnrolledSwinNet.__init__(self, config)
        self._init_hqs(config)


class ProximalGradientDescent(_PGD, UnrolledNet):
    """dl_cs.models.unrolled.ProximalGradientDescent (ResNet denoiser)"""

    def __init__(self, config):
        UnrolledNet.__init__(self, config)
        self._init_pgd()


class HalfQuadraticSplitting(_HQS, UnrolledNet):
    """dl_cs.models.unrolled.HalfQuadraticSplitting (ResNet denoiser)"""

    def __init__(self, config):
        UnrolledNet.__init__(self, config)
        self._init_hqs(config)


def build_unrolled(config) -> nn.Module:
    """The solver the reference's training script picks (scripts/train_swin.py:33-39 / train.py): MODEL_TYPE selects the
    denoiser family, META_ARCHITECTURE the solver ('dlespirit' = PGD, 'modl' = HQS)."""
    swin = getattr(config.MODEL, "MODEL_TYPE", "RES") == "SWIN"
    arch = config.MODEL.META_ARCHITECTURE
    if arch == "dlespirit":
        return (SwinProximalGradientDescent if swin else ProximalGradientDescent)(config)
    if arch == "modl":
        return (SwinHalfQuadraticSplitting if swin else HalfQuadraticSplitting)(config)
    raise ValueError(f"META_ARCHITECTURE {arch!r}: expected 'dlespirit' or 'modl'")
