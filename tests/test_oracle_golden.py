"""Pins oracle/sense_oracle.py to vectors produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2
from oracle import sense_oracle as so
from oracle.golden_inputs import CASES

NAMES = list(CASES)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_outputs(golden_cases, name):
    inp, ref = golden_cases[name]
    maps, x, k, aty, w = inp["maps"], inp["x"], inp["k"], inp["aty"], inp["weights"]
    step = float(ref["step"])
    assert so.sense_ndims(maps) == int(ref["ndims"])
    got = {
        "fwd": so.sense_forward(x, maps, w),
        "adj": so.sense_adjoint(k, maps, w),
        "normal": so.sense_normal(x, maps, w),
        "dc": so.dc_step(x, aty, maps, w, step),
    }
    for key, val in got.items():
        # fp32 reference run: both sides carry ~1e-7 rounding
        assert rel_l2(val, ref[f"{key}_c64"]) < 2e-6, (name, key)
        if f"{key}_c128" in ref.files:
            val128 = getattr(so, {"fwd": "sense_forward", "adj": "sense_adjoint", "normal": "sense_normal"}.get(key, "dc_step"))
            if key == "dc":
                hi = so.dc_step(x.astype(np.complex128), aty.astype(np.complex128), maps.astype(np.complex128), w, step)
            else:
                arg = (k if key == "adj" else x).astype(np.complex128)
                hi = val128(arg, maps.astype(np.complex128), w)
            assert rel_l2(hi, ref[f"{key}_c128"]) < 1e-12, (name, key)


def test_oracle_fft_module_and_helpers_match_reference():
    ref = np.load(os.path.join(GOLDEN, "sense_misc.npz"))
    img = ref["img"]
    assert rel_l2(so.fft_nd(img, 2), ref["fft2"]) < 2e-6
    assert rel_l2(so.fft_nd(img, 2, adjoint=True), ref["ifft2"]) < 2e-6
    assert rel_l2(so.fft_nd(img, 2, centered=True), ref["fft2_centered"]) < 2e-6
    assert rel_l2(so.fft_nd(img, 2, adjoint=True, centered=True), ref["ifft2_centered"]) < 2e-6
    assert rel_l2(so.fft_nd(img, 1), ref["fft1"]) < 2e-6
    assert np.array_equal(so.fftmod(ref["img_even"].copy()), ref["fftmod"])
    assert rel_l2(so.sliding_window(ref["series"], 2, 5), ref["sliding_window5"]) < 1e-6
    assert np.array_equal(so.get_mask(ref["series"]), ref["get_mask"])


@pytest.mark.parametrize("shape", [(5, 7), (12, 10), (16, 16), (1, 9)])
def test_fft_library_agrees_with_explicit_dft_matrix(shape):
    rng = np.random.default_rng(3)
    img = rng.standard_normal((2,) + shape) + 1j * rng.standard_normal((2,) + shape)
    assert rel_l2(so.fft_nd(img, 2), so.dft_matrix_2d(img)) < 1e-12
    assert rel_l2(so.fft_nd(img, 2, adjoint=True), so.dft_matrix_2d(img, inverse=True)) < 1e-12


def test_adjointness_and_fftmod_centering():
    maps, x, mask = so.synthetic_problem(1, 2, 4, 3, 16, 12, seed=5, dtype=np.complex128)
    rng = np.random.default_rng(9)
    y = rng.standard_normal((1, 4, 3, 16, 12)) + 1j * rng.standard_normal((1, 4, 3, 16, 12))
    lhs = np.vdot(y, so.sense_forward(x, maps, mask))
    rhs = np.vdot(so.sense_adjoint(y, maps, mask), x)
    assert abs(lhs - rhs) / abs(lhs) < 1e-12
    # fftmod on both sides of an un-centred FFT == centred FFT for even sizes (SURVEY gotcha 1)
    img = rng.standard_normal((16, 12)) + 1j * rng.standard_normal((16, 12))
    lhs = so.fftmod(so.fft_nd(so.fftmod(img.copy()), 2))
    assert rel_l2(lhs, so.fft_nd(img, 2, centered=True)) < 1e-12


def test_synthetic_problem_is_deterministic_and_undersampled():
    a = so.synthetic_problem(1, 1, 2, 4, 32, 32)
    b = so.synthetic_problem(1, 1, 2, 4, 32, 32)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    mask = a[2]
    assert mask.shape == (1, 1, 4, 32, 32) and set(np.unique(mask)) <= {0.0, 1.0}
    assert 0.02 < mask.mean() < 0.3 and mask[..., :8].sum() == 0


@pytest.mark.parametrize("name", NAMES)
def test_torch_port_matches_reference_outputs(golden_cases, name):
    import torch
    from oracle import sense_oracle_torch as st
    inp, ref = golden_cases[name]
    maps, x, k = (torch.from_numpy(inp[n]) for n in ("maps", "x", "k"))
    w = None if inp["weights"] is None else torch.from_numpy(inp["weights"])
    assert rel_l2(st.forward(x, maps, w).numpy(), ref["fwd_c64"]) < 2e-6
    assert rel_l2(st.adjoint(k, maps, w).numpy(), ref["adj_c64"]) < 2e-6
    # one DC update from x0 = x reproduces the reference's dc vector when aty is A^H(k)
    aty = st.adjoint(k, maps, w)
    one = x + float(ref["step"]) * (st.adjoint(st.forward(x, maps, w), maps, w) - aty)
    got = st.dc_path(k, maps, w, 1, float(ref["step"]), x0=x)
    assert rel_l2(got.numpy(), one.numpy()) < 1e-6
    assert rel_l2(st.adjoint(st.forward(x, maps, w), maps, w).numpy(), ref["normal_c64"]) < 2e-6


# ----------------------------------------------------------------------------- shifted-window attention oracle
from oracle import swin_oracle as sw  # noqa: E402
from oracle.golden_inputs import SWIN_CASES, swin_case  # noqa: E402


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_swin_oracle_matches_reference_outputs(name):
    c = swin_case(name)
    ref = np.load(os.path.join(GOLDEN, f"swin_{name}.npz"))
    x, p = c["x"].astype(np.float64), c["params"]
    B, D, H, W, C = x.shape
    ws, ss = sw.get_window_size((D, H, W), c["window"], c["shift"])
    assert tuple(ref["ws"]) == ws and tuple(ref["ss"]) == ss
    Dp, Hp, Wp = (int(np.ceil(n / w)) * w for n, w in zip((D, H, W), ws))
    mask = sw.compute_mask(Dp, Hp, Wp, ws, ss)
    assert np.array_equal(mask.astype(np.float32), ref["mask"])
    assert np.array_equal(sw.relative_position_index(c["window"]), ref["relative_position_index"])
    assert rel_l2(sw.block_part1(x, p, c["heads"], c["window"], c["shift"], mask), ref["part1"]) < 5e-6
    xp = np.pad(x, ((0, 0), (0, Dp - D), (0, Hp - H), (0, Wp - W), (0, 0)))
    xw = sw.window_partition(xp, ws)
    ap = {k[5:]: v for k, v in p.items() if k.startswith("attn.")}
    assert rel_l2(sw.window_attention(xw, ap, c["heads"], c["window"], None), ref["attn_nomask"]) < 5e-6
    assert rel_l2(sw.window_attention(xw, ap, c["heads"], c["window"], mask), ref["attn_mask"]) < 5e-6
    # whole block (x + part1, then + fc2(GELU(fc1(norm2)))) against the reference block's output
    blk = np.load(os.path.join(GOLDEN, f"swin_block_{name}.npz"))["block"]
    assert rel_l2(sw.block_forward(x, p, c["heads"], c["window"], c["shift"], mask), blk) < 5e-6


def test_swin_host_helpers_match_oracle():
    """The product's host-side helpers (pure torch indexing, CPU-runnable) against the oracle."""
    import torch
    from dl_swin_gan_b200.swin import window_attention as wa
    for (D, H, W, ws, ss) in [(4, 8, 8, (2, 4, 4), (1, 2, 2)), (7, 16, 16, (7, 8, 8), (0, 4, 4)), (6, 8, 12, (2, 4, 4), (0, 0, 0))]:
        m = wa.compute_mask(D, H, W, ws, ss, torch.device("cpu"))
        assert np.array_equal(m.numpy(), sw.compute_mask(D, H, W, ws, ss).astype(np.float32))
        assert wa.compute_mask(D, H, W, ws, ss, torch.device("cpu")) is m          # cached like the reference's lru_cache
        rid = wa._region_ids_for(m, D, H, W, ws, ss, torch.device("cpu")).numpy().astype(np.int64)
        assert np.array_equal(np.where(rid[:, None, :] != rid[:, :, None], -100.0, 0.0), m.numpy())
    x = torch.arange(2 * 4 * 8 * 8 * 3, dtype=torch.float32).view(2, 4, 8, 8, 3)
    xw = wa.window_partition(x, (2, 4, 4))
    assert np.array_equal(xw.numpy(), sw.window_partition(x.numpy(), (2, 4, 4)))
    back = wa.window_reverse(xw.view(-1, 2, 4, 4, 3), (2, 4, 4), 2, 4, 8, 8)
    assert torch.equal(back, x)
    assert wa.get_window_size((2, 8, 8), (4, 4, 4), (2, 2, 2)) == sw.get_window_size((2, 8, 8), (4, 4, 4), (2, 2, 2))
    mod = wa.WindowAttention3D(160, (7, 8, 8), 8, qkv_bias=True)
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    assert shapes == {"relative_position_bias_table": (2925, 8), "relative_position_index": (448, 448),
                      "qkv.weight": (480, 160), "qkv.bias": (480,), "proj.weight": (160, 160), "proj.bias": (160,)}
    assert np.array_equal(mod.relative_position_index.numpy(), sw.relative_position_index((7, 8, 8)))


@pytest.mark.parametrize("name", ["tiny_odd", "pow2_32"])
def test_oracle_cg_and_hqs_match_reference(golden_cases, name):
    ref = np.load(os.path.join(GOLDEN, "sense_misc.npz"))
    inp, _ = golden_cases[name]
    maps, x, k, aty, w = inp["maps"], inp["x"], inp["k"], inp["aty"], inp["weights"]
    lam = 0.1
    normal = lambda m: so.sense_normal(m, maps, w) + lam * m
    ATy = so.sense_adjoint(k, maps, w)
    got = so.conjugate_gradient(normal, x, ATy + lam * aty, 6)
    assert rel_l2(got, ref[f"cg6_{name}"]) < 2e-5
    assert rel_l2(so.hqs_path(k, maps, w, 2, 4, lam), ref[f"hqs2x4_{name}"]) < 2e-5


def test_oracle_preprocessing_chain_matches_reference():
    """DataTransform.__call__ (scripts/reconstruct.py:132-171) restated with the oracle helpers."""
    ref = np.load(os.path.join(GOLDEN, "sense_misc.npz"))
    for slwin, tag in ((False, "plain"), (True, "slwin")):
        k = ref["dt_kspace_in"][None].copy()
        maps = ref["dt_maps_in"][None].copy()
        mask = so.get_mask(k)[:, 0, None]
        k, maps = so.fftmod(k), so.fftmod(maps)
        img = so.sense_adjoint(so.time_average(k, 2), maps, None)
        mag = np.sort(np.abs(img).reshape(-1))[::-1]
        scale = mag[int(round(0.05 * mag.size)) - 1]
        k = k / scale
        init = so.sense_adjoint(so.sliding_window(k, 2, 5) if slwin else k, maps, None)
        assert np.array_equal(mask[0], ref[f"dt_{tag}_mask"])
        assert abs(scale - ref[f"dt_{tag}_scale"]) / ref[f"dt_{tag}_scale"] < 1e-5
        assert rel_l2(k[0], ref[f"dt_{tag}_kspace"]) < 2e-6 and rel_l2(init[0], ref[f"dt_{tag}_init"]) < 2e-6
