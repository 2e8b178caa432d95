"""dl_swin_gan_b200 — B200-native (sm_100a) hot path of dl-cs-dynamic (tjtiger86/dl-swin-gan).

Only what the per-unroll data-consistency path needs lives here:
  csrc/   hand-written CUDA kernels and the C ABI (include/dlsg_b200.h)
  mri/    SenseModel / FFT, mirroring dl_cs/mri/transforms.py
  swin/   WindowAttention3D and friends, mirroring dl_cs/models/video_swin_transformer_mri_downsample.py
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
