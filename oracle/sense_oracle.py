"""ORACLE (test infrastructure, not product): CPU restatement of the reference SENSE path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module, and only as the checker or the timed CPU baseline.  The product
(dl_swin_gan_b200) never imports it and has no CPU fallback.

What is restated (all citations relative to /root/reference):
  * FFT.forward                dl_cs/mri/transforms.py:28-46   n-D (i)FFT over the trailing
                               `ndims` axes, norm='ortho', optional ifftshift/fftshift
  * SenseModel.__init__        dl_cs/mri/transforms.py:60-82   ndims from the squeezed map image,
                               weights=None -> 1.0
  * SenseModel._forward_op     dl_cs/mri/transforms.py:92-98   k = W * F( sum_e S * x )
  * SenseModel._adjoint_op     dl_cs/mri/transforms.py:84-90   x = sum_c conj(S) * F^H( W * k )
  * DC update                  dl_cs/models/unrolledswin.py:109 (== dl_cs/models/unrolled.py:104)
  * ConjugateGradient          dl_cs/mri/algorithms.py:11-73, HalfQuadraticSplitting dl_cs/models/unrolledswin.py:139-172
  * fftmod                     dl_cs/mri/utils.py:7-19
  * sliding_window / time_average / get_mask   dl_cs/mri/utils.py:29-49, 69-79

Third-party arithmetic the reference delegates to and that is not under /root/reference:
torch.fft.fftn / ifftn (pocketfft on CPU, cuFFT on CUDA; torch pinned 1.9.0 / 2.4.0 in
environment*.yaml, 2.11.0 in this image).  Its published definition is the plain DFT
  X[k] = (1/sqrt(N)) sum_n x[n] exp(-2 pi i n k / N)         (norm='ortho', forward)
which is what numpy.fft implements too; `dft_matrix_*` below restates it with an explicit
DFT matrix in float64 so that small cases do not depend on any FFT library at all.

Parity pin: the reference ships no tests or golden vectors.  This oracle
is pinned against outputs of the reference itself: tests/golden/sense_*.npz were produced by
oracle/make_golden.py importing /root/reference/dl_cs/mri/transforms.py in the build
container, and tests/test_oracle_golden.py checks every function here against them.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- FFT
def fft_nd(data: np.ndarray, ndims: int, adjoint: bool = False, centered: bool = False) -> np.ndarray:
    """FFT.forward (transforms.py:28-46): ortho (i)FFT over the last `ndims` axes."""
    assert np.iscomplexobj(data)
    axes = tuple(range(-1, -1 - ndims, -1))
    if ndims == 0:
        return data.copy()
    if centered:
        data = np.fft.ifftshift(data, axes=axes)
    if adjoint:
        data = np.fft.ifftn(data, axes=axes, norm="ortho")
    else:
        data = np.fft.fftn(data, axes=axes, norm="ortho")
    if centered:
        data = np.fft.fftshift(data, axes=axes)
    return data


def dft_matrix(n: int, inverse: bool = False) -> np.ndarray:
    """Explicit ortho DFT matrix in complex128 (the published definition, no FFT library)."""
    idx = np.arange(n)
    sign = 2.0 if inverse else -2.0
    return np.exp(sign * 1j * np.pi * np.outer(idx, idx) / n) / np.sqrt(n)


def dft_matrix_2d(img: np.ndarray, inverse: bool = False) -> np.ndarray:
    """Ortho 2-D DFT of the last two axes by matrix products (small sizes only)."""
    fy = dft_matrix(img.shape[-2], inverse)
    fx = dft_matrix(img.shape[-1], inverse)
    return np.einsum("ky,...yx,lx->...kl", fy, img.astype(np.complex128), fx)


# --------------------------------------------------------------------------- SENSE
def sense_ndims(maps: np.ndarray) -> int:
    """SenseModel.__init__ (transforms.py:76): len(maps[0,0,0].squeeze().shape)."""
    return int(np.squeeze(maps[0, 0, 0]).ndim)


def _weights(weights):
    return 1.0 if weights is None else weights


def sense_forward(x: np.ndarray, maps: np.ndarray, weights=None) -> np.ndarray:
    """SenseModel._forward_op (transforms.py:92-98).

    x [B,E,T,Y,X] complex, maps [B,E,C,1,Y,X] complex, weights broadcastable to [B,C,T,Y,X].
    """
    assert np.iscomplexobj(x) and np.iscomplexobj(maps)
    nd = sense_ndims(maps)
    coil_images = (x[:, :, None] * maps).sum(1)          # [B,C,T,Y,X]
    return _weights(weights) * fft_nd(coil_images, nd, adjoint=False)


def sense_adjoint(k: np.ndarray, maps: np.ndarray, weights=None) -> np.ndarray:
    """SenseModel._adjoint_op (transforms.py:84-90)."""
    assert np.iscomplexobj(k) and np.iscomplexobj(maps)
    nd = sense_ndims(maps)
    img = fft_nd(_weights(weights) * k, nd, adjoint=True)  # [B,C,T,Y,X]
    return (img[:, None] * np.conj(maps)).sum(2)           # [B,E,T,Y,X]


def sense_normal(x: np.ndarray, maps: np.ndarray, weights=None) -> np.ndarray:
    """A(A(x), adjoint=True), the inner term of unrolledswin.py:109."""
    return sense_adjoint(sense_forward(x, maps, weights), maps, weights)


def dc_step(x: np.ndarray, aty: np.ndarray, maps: np.ndarray, weights, step: float) -> np.ndarray:
    """x + step * (A^H A x - A^H y)   (unrolledswin.py:109, unrolled.py:104)."""
    return x + step * (sense_normal(x, maps, weights) - aty)


def conjugate_gradient(A, x, y, num_iter: int):
    """ConjugateGradient.forward (dl_cs/mri/algorithms.py:38-73): dot products over the whole batch."""
    zdot = lambda a, b: np.sum(np.conj(a) * b)
    r = y - A(x)
    rsold = zdot(r, r).real
    p = r
    for _ in range(num_iter):
        Ap = A(p)
        alpha = rsold / zdot(p, Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rsnew = zdot(r, r).real
        p = (rsnew / rsold) * p + r
        rsold = rsnew
    return x


def hqs_path(y, maps, weights, num_unrolls: int, num_cg_iter: int, lamda: float, x0=None):
    """HalfQuadraticSplitting.forward (unrolledswin.py:139-172) with the CNN omitted (z = x)."""
    aty = sense_adjoint(y, maps, weights)
    x = aty if x0 is None else x0
    normal = lambda m: sense_normal(m, maps, weights) + lamda * m
    for _ in range(num_unrolls):
        x = conjugate_gradient(normal, x, aty + lamda * x, num_cg_iter)
    return x


# --------------------------------------------------------------------------- helpers either side of the path
def fftmod(arr: np.ndarray) -> np.ndarray:
    """dl_cs/mri/utils.py:7-19 (in-place sign modulation that centres an even-sized FFT)."""
    arr[..., ::2] *= -1
    arr[..., ::2, :] *= -1
    arr *= -1
    return arr


def get_mask(data: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """dl_cs/mri/utils.py:69-79."""
    mag = np.abs(data)
    return np.where(mag > eps, np.ones_like(mag), np.zeros_like(mag))


def time_average(data: np.ndarray, axis: int, eps: float = 1e-6) -> np.ndarray:
    """dl_cs/mri/utils.py:29-34."""
    mask = get_mask(data)
    return data.sum(axis, keepdims=True) / (mask.sum(axis, keepdims=True) + eps)


def sliding_window(data: np.ndarray, axis: int, window: int) -> np.ndarray:
    """dl_cs/mri/utils.py:37-49 (circular sliding-window time average)."""
    assert 0 < window <= data.shape[axis]
    out = []
    for i in range(data.shape[axis]):
        rolled = np.roll(data, int(window / 2) - i, axis)
        win = np.take(rolled, np.arange(window), axis=axis)
        out.append(time_average(win, axis))
    return np.concatenate(out, axis=axis)


# --------------------------------------------------------------------------- synthetic cine data
def synthetic_problem(B, E, C, T, Y, X, seed=1000, accel=12, dtype=np.complex64):
    """Deterministic maps / image / undersampling weights of the cine shape.

    maps ~ CN(0,1)/sqrt(E*C), x ~ CN(0,1); the weights are a k-t lattice-like 0/1 mask
    with a fully-zero first quarter of kx (the reference's partial-kx, subsample.py:107-109)
    broadcast over coils, shape [B,1,T,Y,X].  Pure numpy so it runs on the GPU box too.
    """
    rng = np.random.default_rng(seed)

    def cn(*shape):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2.0)

    maps = (cn(B, E, C, 1, Y, X) / np.sqrt(E * C)).astype(dtype)
    x = cn(B, E, T, Y, X).astype(dtype)
    mask = np.zeros((B, 1, T, Y, X), dtype=np.float32)
    for b in range(B):
        for t in range(T):
            rows = (np.arange(Y) + 5 * t + 3 * b) % accel == 0
            centre = np.abs(np.arange(Y) - Y // 2) < max(1, Y // 32)
            keep = np.roll(rows | centre, -(Y // 2))        # un-centred k-space (fftmod'ed data)
            mask[b, 0, t, keep, :] = 1.0
    mask[..., : X // 4] = 0.0
    return maps, x, mask
