"""The product's SENSE kernels + C-ABI dispatch, built by g++ under the thread-per-CUDA-thread
emulation of csrc/cuda_compat.h (TEST ONLY), against the oracle.  Debugs the index math,
shared-memory layouts and chunk walk without a GPU; the GPU parity tests are in test_sense_gpu.py."""
import ctypes

import numpy as np
import pytest

from conftest import rel_l2
from oracle import sense_oracle as so

vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def lib(sense_emu_lib):
    lib = ctypes.CDLL(sense_emu_lib)
    tail = [vp, ctypes.c_size_t] + [ctypes.c_int] * 7 + [vp]
    for name in ("forward", "adjoint", "normal"):
        getattr(lib, f"dlsg_sense_{name}_c64").argtypes = [vp] * 5 + tail
    lib.dlsg_sense_dc_step_c64.argtypes = [vp] * 5 + [ctypes.c_float, vp, vp] + tail
    lib.dlsg_sense_normal_reg_c64.argtypes = [vp] * 4 + [ctypes.c_float, vp, vp] + tail
    lib.dlsg_sense_workspace_bytes.restype = ctypes.c_size_t
    lib.dlsg_sense_workspace_bytes.argtypes = [ctypes.c_int] * 7 + [ctypes.c_size_t]
    return lib


def ptr(a):
    return a.ctypes.data_as(vp)


def run_all(lib, B, E, C, T, Y, X, wkind, budget=0, ndims=2, seed=7):
    rng = np.random.default_rng(seed)
    maps, x, mask = so.synthetic_problem(B, E, C, T, Y, X, seed=seed)
    w = {"none": None, "mask": mask,
         "full": rng.standard_normal((B, C, T, Y, X)).astype(np.float32)}[wkind]
    k = (rng.standard_normal((B, C, T, Y, X)) + 1j * rng.standard_normal((B, C, T, Y, X))).astype(np.complex64)
    aty = (rng.standard_normal((B, E, T, Y, X)) + 1j * rng.standard_normal((B, E, T, Y, X))).astype(np.complex64)
    m5 = np.ascontiguousarray(maps[:, :, :, 0])
    if w is not None:
        wb = np.broadcast_to(w, (B, C, T, Y, X))
        ws, wp = (ctypes.c_longlong * 5)(*[s // 4 for s in wb.strides]), ptr(w)
    else:
        ws, wp = None, None
    nbytes = lib.dlsg_sense_workspace_bytes(B, E, C, T, Y, X, ndims, budget)
    work = np.zeros(nbytes // 8 + 1, np.complex64)
    tail = (ptr(work), nbytes, B, E, C, T, Y, X, ndims, None)
    m128, x128, k128, a128 = (a.astype(np.complex128) for a in (maps, x, k, aty))
    err = {}
    out = np.full((B, C, T, Y, X), np.nan, np.complex64)
    assert lib.dlsg_sense_forward_c64(ptr(x), ptr(m5), wp, ws, ptr(out), *tail) == 0
    err["fwd"] = rel_l2(out, so.sense_forward(x128, m128, w))
    out = np.full((B, E, T, Y, X), np.nan, np.complex64)
    assert lib.dlsg_sense_adjoint_c64(ptr(k), ptr(m5), wp, ws, ptr(out), *tail) == 0
    err["adj"] = rel_l2(out, so.sense_adjoint(k128, m128, w))
    out = np.full((B, E, T, Y, X), np.nan, np.complex64)
    assert lib.dlsg_sense_normal_c64(ptr(x), ptr(m5), wp, ws, ptr(out), *tail) == 0
    err["normal"] = rel_l2(out, so.sense_normal(x128, m128, w))
    out = np.full((B, E, T, Y, X), np.nan, np.complex64)
    assert lib.dlsg_sense_dc_step_c64(ptr(x), ptr(aty), ptr(m5), wp, ws, -2.0, None, ptr(out), *tail) == 0
    err["dc"] = rel_l2(out, so.dc_step(x128, a128, m128, w, -2.0))
    out = np.full((B, E, T, Y, X), np.nan, np.complex64)
    assert lib.dlsg_sense_normal_reg_c64(ptr(x), ptr(m5), wp, ws, 0.25, None, ptr(out), *tail) == 0
    err["reg"] = rel_l2(out, so.sense_normal(x128, m128, w) + 0.25 * x128)
    return err


# (B, E, C, T, Y, X, weights, kwargs) — radix kernels (32*R axes), direct kernels (odd axes), mixes, chunking
CASES = [
    ((1, 1, 2, 2, 32, 32, "mask"), {}),
    ((1, 2, 3, 1, 64, 96, "none"), {}),
    ((1, 2, 2, 1, 12, 10, "mask"), {}),
    ((1, 1, 2, 2, 7, 32, "full"), {}),
    ((1, 3, 2, 1, 32, 15, "full"), {}),
    ((2, 3, 2, 1, 64, 64, "full"), {"budget": 1}),          # one (slice, frame) per chunk, E > 2 groups
    ((1, 2, 2, 2, 1, 64, "mask"), {"ndims": 1}),
    ((1, 2, 6, 2, 32, 320, "mask"), {}),                    # two coil groups: partial sums through `out`
    ((1, 1, 3, 5, 32, 64, "full"), {}),                     # frame chunk shared by 4 warps, ragged batches
]


@pytest.mark.parametrize("case,kw", CASES)
def test_emulated_kernels_match_oracle(lib, case, kw):
    err = run_all(lib, *case, **kw)
    for key, e in err.items():
        assert e < 1e-5, (case, key, e)      # north_star tolerance: 1e-5 relative L2 in fp32


@pytest.mark.slow
def test_emulated_cine_radices_192x160(lib):
    err = run_all(lib, 1, 1, 2, 1, 192, 160, "mask")
    assert max(err.values()) < 1e-5


def test_status_codes(lib):
    z = np.zeros(16, np.complex64)
    tail = (ptr(z), 128, 1, 1, 1, 1, 4, 4, 2, None)
    assert lib.dlsg_sense_forward_c64(None, ptr(z), None, None, ptr(z), *tail) == -1
    assert lib.dlsg_sense_forward_c64(ptr(z), ptr(z), None, None, ptr(z), ptr(z), 8, 1, 1, 1, 1, 4, 4, 2, None) == -2
    assert lib.dlsg_sense_forward_c64(ptr(z), ptr(z), None, None, ptr(z), ptr(z), 128, 1, 1, 1, 1, 4, 4, 3, None) == -1
    assert lib.dlsg_sense_forward_c64(ptr(z), ptr(z), None, None, ptr(z), ptr(z), 128, 1, 1, 1, 1, 4, 5000, 2, None) == -3
    assert lib.dlsg_sense_forward_c64(ptr(z), ptr(z), ptr(z), None, ptr(z), *tail) == -1   # weights without strides


def test_kt_kernels_emulated(sense_emu_lib):
    """fftmod / get_mask / sliding-window kernels under the host emulation vs the oracle helpers."""
    lib = ctypes.CDLL(sense_emu_lib)
    ll = ctypes.c_longlong
    lib.dlsg_fftmod_c64.argtypes = [vp, ll, ctypes.c_int, ctypes.c_int, vp]
    lib.dlsg_get_mask_c64.argtypes = [vp, vp, ll, ctypes.c_float, vp]
    lib.dlsg_sliding_window_c64.argtypes = [vp, vp, ll, ctypes.c_int, ll, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, vp]
    rng = np.random.default_rng(4)
    a = (rng.standard_normal((2, 3, 6, 5)) + 1j * rng.standard_normal((2, 3, 6, 5))).astype(np.complex64)
    b = a.copy()
    assert lib.dlsg_fftmod_c64(ptr(b), 6, 6, 5, None) == 0
    assert np.array_equal(b, so.fftmod(a.copy()))
    s = (rng.standard_normal((2, 7, 4, 3)) + 1j * rng.standard_normal((2, 7, 4, 3))).astype(np.complex64)
    s[:, ::3] = 0
    m = np.empty(s.shape, np.float32)
    assert lib.dlsg_get_mask_c64(ptr(s), ptr(m), s.size, 1e-12, None) == 0
    assert np.array_equal(m, so.get_mask(s))
    out = np.empty_like(s)
    assert lib.dlsg_sliding_window_c64(ptr(s), ptr(out), 2, 7, 12, 5, 1, 1e-12, 1e-6, None) == 0
    assert rel_l2(out, so.sliding_window(s, 1, 5)) < 1e-6
    avg = np.empty((2, 1, 4, 3), np.complex64)
    assert lib.dlsg_sliding_window_c64(ptr(s), ptr(avg), 2, 7, 12, 7, 0, 1e-12, 1e-6, None) == 0
    assert rel_l2(avg, so.time_average(s, 1)) < 1e-6
    assert lib.dlsg_sliding_window_c64(ptr(s), ptr(out), 2, 7, 12, 9, 1, 1e-12, 1e-6, None) == -1


# ---------------------------------------------------------------------------- sampled-line path (sense_lines.cuh)
def _lines_lib(sense_emu_lib):
    lib = ctypes.CDLL(sense_emu_lib)
    i, sz = ctypes.c_int, ctypes.c_size_t
    lib.dlsg_sense_line_plan_bytes.restype = sz
    lib.dlsg_sense_line_plan_bytes.argtypes = [i, i, i]
    lib.dlsg_sense_line_plan_build.argtypes = [vp, vp, i, i, i, i, i, vp, vp]
    lib.dlsg_sense_lines_supported.argtypes = [i, i, i, i]
    lib.dlsg_sense_lines_workspace_bytes.restype = sz
    lib.dlsg_sense_lines_workspace_bytes.argtypes = [i, i, i, i, i, sz]
    tail = [vp, sz] + [i] * 6 + [vp]
    lib.dlsg_sense_forward_lines_c64.argtypes = [vp, vp, vp, vp, vp, i, vp] + tail
    lib.dlsg_sense_adjoint_lines_c64.argtypes = [vp, vp, vp, vp, vp, i, vp] + tail
    lib.dlsg_sense_adjoint_lines_gather_c64.argtypes = [vp, vp, vp, vp, vp, i, vp] + tail
    lib.dlsg_sense_normal_lines_c64.argtypes = [vp, vp, vp, vp, vp, vp, i, ctypes.c_float, vp, i, vp] + tail
    return lib


def _line_weights(rng, B, C, T, Y, X, kind):
    """k-t style weights: a few ky lines per (slice, frame); 'mask' = 0/1 broadcast over coils with an empty kx quarter,
    'coil' = real-valued per-coil weights on those lines, 'ragged' = different line counts per frame incl. an empty one."""
    w = np.zeros((B, C if kind == "coil" else 1, T, Y, X), np.float32)
    for b in range(B):
        for t in range(T):
            nl = 0 if (kind == "ragged" and t == 1) else 3 + (2 * t + b) % 4
            lines = rng.choice(Y, size=nl, replace=False)
            for c in range(w.shape[1]):
                w[b, c, t, lines, :] = rng.standard_normal((nl, X)) if kind == "coil" else 1.0
    if kind == "mask":
        w[..., : X // 4] = 0
    return w


LINE_CASES = [
    (1, 1, 2, 2, 32, 32, "mask", 0),
    (2, 2, 3, 3, 64, 96, "coil", 0),
    (2, 1, 2, 3, 96, 64, "ragged", 1),                      # one slice per workspace chunk, a frame without lines
    (1, 3, 2, 2, 64, 160, "mask", 0),                       # E = 3: maps of e >= 1 read through the cache
    # Y = 16 * R2 with R2 >= 8: the packed, pruned column kernels of sense_lines_p.cuh
    (1, 1, 3, 2, 128, 32, "mask", 0),
    (2, 2, 2, 3, 128, 32, "coil", 1),                       # E = 2, one slice per workspace chunk
    (1, 1, 2, 3, 160, 32, "ragged", 0),                     # R2 = 10, a frame without lines
    (1, 3, 2, 1, 128, 32, "mask", 0),                       # E = 3 stays on the first-generation kernels
    (1, 1, 3, 2, 128, 128, "coil", 0),                      # X = 16 * 8: second-generation row pass of the normal operator, odd row count
]


@pytest.mark.parametrize("B,E,C,T,Y,X,kind,budget", LINE_CASES)
def test_emulated_sampled_line_path_matches_oracle(sense_emu_lib, B, E, C, T, Y, X, kind, budget):
    lib = _lines_lib(sense_emu_lib)
    rng = np.random.default_rng(11)
    maps, x, _ = so.synthetic_problem(B, E, C, T, Y, X, seed=5)
    w = _line_weights(rng, B, C, T, Y, X, kind)
    k = (rng.standard_normal((B, C, T, Y, X)) + 1j * rng.standard_normal((B, C, T, Y, X))).astype(np.complex64)
    aty = (rng.standard_normal((B, E, T, Y, X)) + 1j * rng.standard_normal((B, E, T, Y, X))).astype(np.complex64)
    m5 = np.ascontiguousarray(maps[:, :, :, 0])
    wb = np.broadcast_to(w, (B, C, T, Y, X))
    ws = (ctypes.c_longlong * 5)(*[s // 4 for s in wb.strides])

    plan = np.full(lib.dlsg_sense_line_plan_bytes(B, T, Y) // 4, -7, np.int32)
    assert lib.dlsg_sense_line_plan_build(ptr(w), ws, B, C, T, Y, X, ptr(plan), None) == 0
    counts = plan[: B * T]
    sampled = (wb != 0).any(axis=(1, 4))                    # [B, T, Y]
    assert np.array_equal(counts.reshape(B, T), sampled.sum(-1))
    lst = plan[B * T: B * T * (1 + Y)].reshape(B, T, Y)
    pos = plan[B * T * (1 + Y): B * T * (1 + 2 * Y)].reshape(B, T, Y)
    order = plan[B * T * (1 + 2 * Y): B * T * (1 + 3 * Y)].reshape(B, T, Y)
    cls = plan[B * T * (1 + 3 * Y): B * T * (1 + 3 * Y + 36)].reshape(B, T, 36)
    for b in range(B):
        for t in range(T):
            ky = np.flatnonzero(sampled[b, t])
            assert np.array_equal(lst[b, t, : len(ky)], ky)
            assert np.array_equal(pos[b, t, ky], np.arange(len(ky))) and (pos[b, t, ~sampled[b, t]] == -1).all()
            # class tables of the pruned column kernels: lines grouped by ky mod 16, stable
            want_order = np.argsort(ky % 16, kind="stable")
            assert np.array_equal(order[b, t, : len(ky)], want_order)
            seg = cls[b, t, :17]
            assert seg[0] == 0 and seg[16] == len(ky) and np.array_equal(np.diff(seg), np.bincount(ky % 16, minlength=16))
            nonempty = np.flatnonzero(np.diff(seg))
            assert cls[b, t, 33] == len(nonempty) and np.array_equal(cls[b, t, 17:17 + len(nonempty)], nonempty)
            assert cls[b, t, 34] == sum(1 << int(k) for k in nonempty)
    lmax = max(1, int(counts.max()))
    assert lib.dlsg_sense_lines_supported(E, Y, X, lmax) == 1

    nbytes = lib.dlsg_sense_lines_workspace_bytes(B, C, T, X, lmax, budget)
    work = np.zeros(nbytes // 8 + 1, np.complex64)
    tail = (ptr(work), nbytes, B, E, C, T, Y, X, None)
    m128, x128, k128, a128 = (a.astype(np.complex128) for a in (maps, x, k, aty))
    out = np.full((B, C, T, Y, X), np.nan, np.complex64)
    assert lib.dlsg_sense_forward_lines_c64(ptr(x), ptr(m5), ptr(w), ws, ptr(plan), lmax, ptr(out), *tail) == 0
    assert rel_l2(out, so.sense_forward(x128, m128, w)) < 1e-5
    assert not out[np.broadcast_to(~sampled[:, None, :, :, None], out.shape)].any()      # unsampled lines are exact zeros
    out = np.full((B, E, T, Y, X), np.nan, np.complex64)
    assert lib.dlsg_sense_adjoint_lines_c64(ptr(k), ptr(m5), ptr(w), ws, ptr(plan), lmax, ptr(out), *tail) == 0
    assert rel_l2(out, so.sense_adjoint(k128, m128, w)) < 1e-5
    out2 = np.full((B, E, T, Y, X), np.nan, np.complex64)        # same through the line gather (host-resident k-space entry)
    assert lib.dlsg_sense_adjoint_lines_gather_c64(ptr(k), ptr(m5), ptr(w), ws, ptr(plan), lmax, ptr(out2), *tail) == 0
    assert np.array_equal(out, out2)
    want = {0: so.sense_normal(x128, m128, w), 1: so.dc_step(x128, a128, m128, w, -2.0),
            2: so.sense_normal(x128, m128, w) + (-2.0) * x128}
    for epi in (0, 1, 2):
        out = np.full((B, E, T, Y, X), np.nan, np.complex64)
        assert lib.dlsg_sense_normal_lines_c64(ptr(x), ptr(aty), ptr(m5), ptr(w), ws, ptr(plan), lmax, -2.0, None, epi,
                                               ptr(out), *tail) == 0
        assert rel_l2(out, want[epi]) < 1e-5, epi


def test_sampled_line_status_codes(sense_emu_lib):
    lib = _lines_lib(sense_emu_lib)
    assert lib.dlsg_sense_lines_supported(1, 192, 160, 27) == 1
    assert lib.dlsg_sense_lines_supported(1, 190, 160, 27) == 0       # not 32*R
    assert lib.dlsg_sense_lines_supported(1, 192, 160, 0) == 0
    assert lib.dlsg_sense_lines_supported(8, 320, 320, 160) == 0      # accumulators do not fit in shared memory
    z = np.zeros(64, np.complex64)
    ws = (ctypes.c_longlong * 5)(0, 0, 0, 32, 1)
    args = (ptr(z), ptr(z), ptr(z), ws, ptr(z), 4, ptr(z))
    assert lib.dlsg_sense_forward_lines_c64(*args, ptr(z), 8, 1, 1, 1, 1, 32, 32, None) == -2
    assert lib.dlsg_sense_forward_lines_c64(*args, ptr(z), 1 << 20, 1, 1, 1, 1, 33, 32, None) == -3
    assert lib.dlsg_sense_forward_lines_c64(ptr(z), ptr(z), None, ws, ptr(z), 4, ptr(z), ptr(z), 1 << 20, 1, 1, 1, 1, 32, 32, None) == -1
