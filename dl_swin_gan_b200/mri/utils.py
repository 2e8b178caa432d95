"""Mirror of dl_cs/mri/utils.py on the GPU: fftmod :7-19, get_mask :69-79, time_average :29-34, sliding_window :37-49
(root_sum_of_squares :22 and center_crop :52-66 are one-line torch indexing and are provided for completeness).
Same names and argument meaning; complex64 CUDA tensors only (no CPU fallback)."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from .transforms import _require_cuda_c64


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def fftmod(out: torch.Tensor) -> torch.Tensor:
    """In place, like the reference: multiplies every other line by exp(j*pi) along the last two axes."""
    _require_cuda_c64(out, "fftmod input")
    if not out.is_contiguous():
        raise ValueError("fftmod works in place and needs a contiguous tensor")
    if out.numel() == 0:
        return out
    Y, X = out.shape[-2], out.shape[-1]
    with torch.cuda.device(out.device):
        st = _lib.load().dlsg_fftmod_c64(out.data_ptr(), out.numel() // (Y * X), Y, X, _stream())
    _lib.check(st, "dlsg_fftmod")
    return out


def get_mask(data: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    assert torch.is_complex(data)  # force complex (utils.py:73)
    _require_cuda_c64(data, "get_mask input")
    data = data.contiguous()
    mask = torch.empty(data.shape, dtype=torch.float32, device=data.device)
    if data.numel():
        with torch.cuda.device(data.device):
            st = _lib.load().dlsg_get_mask_c64(data.data_ptr(), mask.data_ptr(), data.numel(), float(eps), _stream())
        _lib.check(st, "dlsg_get_mask")
    return mask


def _window_average(data: torch.Tensor, dim: int, window: int, keep_all: bool, eps: float) -> torch.Tensor:
    _require_cuda_c64(data, "sliding_window input")
    dim = dim % data.dim()
    data = data.contiguous()
    T = data.shape[dim]
    outer = 1
    for s in data.shape[:dim]:
        outer *= s
    inner = 1
    for s in data.shape[dim + 1:]:
        inner *= s
    shape = list(data.shape)
    if not keep_all:
        shape[dim] = 1
    out = torch.empty(shape, dtype=torch.complex64, device=data.device)
    if out.numel():
        with torch.cuda.device(data.device):
            st = _lib.load().dlsg_sliding_window_c64(data.data_ptr(), out.data_ptr(), outer, T, inner, int(window),
                                                     1 if keep_all else 0, 1e-12, float(eps), _stream())
        _lib.check(st, "dlsg_sliding_window")
    return out


def time_average(data, dim, eps=1e-6, keepdim=True):
    out = _window_average(data, dim, data.shape[dim], False, eps)
    return out if keepdim else out.squeeze(dim)


def sliding_window(data, dim, window_size):
    assert 0 < window_size <= data.shape[dim]
    return _window_average(data, dim, window_size, True, 1e-6)


def root_sum_of_squares(x, dim=0):
    return torch.sqrt(torch.sum(torch.abs(x) ** 2, dim=dim))


def center_crop(data, shapes, dims):
    for i, dim in enumerate(dims):
        assert 0 < shapes[i] <= data.shape[dim]
        idx_start = (data.shape[dim] - shapes[i]) // 2
        data = data.narrow(dim, idx_start, shapes[i])
    return data
