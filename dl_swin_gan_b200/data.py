"""GPU mirror of the inference pre-processing of the reference: DataTransform.__call__ (scripts/reconstruct.py:132-171),
which is also the tail of CinePreprocess.__call__ (dl_cs/data/preprocess.py:140-164).

mask from the acquired k-space, fftmod of k-space and maps (so the operator can use un-centred FFTs), scale = 95th
percentile of the view-shared image magnitude, network initialisation by (sliding-window) adjoint.  The reference
does this on CPU tensors inside DataLoader workers; here everything stays on the device and goes through the
B200 operator and the k-t kernels of dl_swin_gan_b200.mri.utils.
"""
from __future__ import annotations

import numpy as np
import torch

from .mri import utils
from .mri.transforms import SenseModel


class DataTransform:
    def __init__(self, slwin_init: bool = False, device="cuda"):
        self.slwin_init = slwin_init
        self.device = torch.device(device)

    def __call__(self, kspace, maps):
        """kspace [coils, t, y, x], maps [emaps, coils, 1, y, x] (numpy or tensors, complex64), un-batched like the reference."""
        to_dev = lambda a: (torch.from_numpy(a) if isinstance(a, np.ndarray) else a).to(self.device).unsqueeze(0).contiguous()
        kspace, maps = to_dev(kspace).clone(), to_dev(maps).clone()
        mask = utils.get_mask(kspace)[:, 0, None]                     # no need to store all coils (:139)
        kspace = utils.fftmod(kspace)
        maps = utils.fftmod(maps)
        A = SenseModel(maps, weights=None)
        averaged = utils.time_average(kspace, dim=2)
        image = A(averaged, adjoint=True)
        magnitude = torch.abs(image).reshape(-1)
        k = int(round(0.05 * magnitude.numel()))
        scale = torch.min(torch.topk(magnitude, k).values)
        kspace /= scale
        init_kspace = utils.sliding_window(kspace, dim=2, window_size=5) if self.slwin_init else kspace
        init_image = A(init_kspace, adjoint=True)
        return kspace.squeeze(0), maps.squeeze(0), mask.squeeze(0), init_image.squeeze(0), scale
