"""In-register DFT butterflies of csrc/fft_radix.cuh, compiled for the host, against numpy."""
import ctypes

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib(fft_radix_host_lib):
    lib = ctypes.CDLL(fft_radix_host_lib)
    lib.tw_cos.restype = ctypes.c_double
    lib.tw_sin.restype = ctypes.c_double
    lib.tw_cos.argtypes = [ctypes.c_longlong] * 2
    lib.tw_sin.argtypes = [ctypes.c_longlong] * 2
    return lib


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 32])
@pytest.mark.parametrize("direction", [-1, 1])
def test_dft_matches_numpy(lib, n, direction):
    rng = np.random.default_rng(n * 7 + direction)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    buf = x.copy()
    assert lib.dft_run(n, direction, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
    x128 = x.astype(np.complex128)
    ref = np.fft.fft(x128) if direction < 0 else np.fft.ifft(x128) * n
    assert np.linalg.norm(buf - ref) / np.linalg.norm(ref) < 3e-7


def test_dft_of_delta_and_constant(lib):
    for n in (5, 6, 10, 32):
        d = np.zeros(n, np.complex64); d[1] = 1
        assert lib.dft_run(n, -1, d.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
        assert np.allclose(d, np.exp(-2j * np.pi * np.arange(n) / n), atol=2e-7)
        c = np.ones(n, np.complex64)
        lib.dft_run(n, 1, c.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        assert np.allclose(c, np.r_[n, np.zeros(n - 1)], atol=1e-5)


def test_constexpr_trig_is_double_accurate(lib):
    worst = 0.0
    for n in (3, 5, 7, 32, 160, 192, 320, 1000):
        for k in range(-n, 2 * n):
            worst = max(worst, abs(lib.tw_cos(k, n) - np.cos(2 * np.pi * k / n)),
                        abs(lib.tw_sin(k, n) - np.sin(2 * np.pi * k / n)))
    assert worst < 1e-14
    assert lib.tw_cos(1, 4) == 0.0 and lib.tw_sin(1, 4) == 1.0 and lib.tw_cos(1, 2) == -1.0
