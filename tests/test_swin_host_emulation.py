"""Shifted-window attention kernels + C-ABI dispatch under the g++ thread emulation (TEST ONLY) vs the numpy oracle.
The emulation includes a software model of mma.sync.m16n8k16 with the PTX fragment layouts, so fragment indexing,
shared-memory staging, the a(i)-a(j) bias addressing and the region-id mask are debugged without a GPU."""
import ctypes

import numpy as np
import pytest

from conftest import rel_l2
from oracle import swin_oracle as sw
from oracle.golden_inputs import SWIN_CASES, swin_case

vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


@pytest.fixture(scope="module")
def lib(sense_emu_lib):
    lib = ctypes.CDLL(sense_emu_lib)
    lib.dlsg_swin_ln_partition_f32.argtypes = [vp, vp, vp, cf, vp] + [ci] * 5 + [vp, vp, vp]
    lib.dlsg_swin_window_reverse_f32.argtypes = [vp, vp] + [ci] * 5 + [vp, vp, vp]
    lib.dlsg_swin_window_attention_f32.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, cf, ci, vp, ci, vp]
    lib.dlsg_swin_window_attention_f16.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, ci, vp, ci, vp]
    return lib


def ptr(a):
    return a.ctypes.data_as(vp)


def i3(t):
    return (ctypes.c_int * 3)(*t)


def region_ids(Dp, Hp, Wp, ws, ss):
    """per-window per-token region id equivalent to compute_mask's cnt image (:342-355)."""
    img = np.zeros((1, Dp, Hp, Wp, 1))
    cnt = 0
    for d in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for h in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for w in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, d, h, w, :] = cnt
                cnt += 1
    return sw.window_partition(img, ws)[..., 0].astype(np.uint8)


SMALL = [n for n in SWIN_CASES if not n.startswith("cine")]


@pytest.mark.parametrize("name", SMALL)
def test_ln_partition_and_reverse(lib, name):
    c = swin_case(name)
    x, p = c["x"], c["params"]
    B, D, H, W, C = x.shape
    ws, ss = sw.get_window_size((D, H, W), c["window"], c["shift"])
    Dp, Hp, Wp = (int(np.ceil(n / w)) * w for n, w in zip((D, H, W), ws))
    nW, N = (Dp // ws[0]) * (Hp // ws[1]) * (Wp // ws[2]), ws[0] * ws[1] * ws[2]
    xw = np.full((B * nW, N, C), np.nan, np.float32)
    assert lib.dlsg_swin_ln_partition_f32(ptr(x), ptr(p["norm1.weight"]), ptr(p["norm1.bias"]), 1e-5, ptr(xw),
                                          B, D, H, W, C, i3(ws), i3(ss), None) == 0
    ref = sw.layer_norm(x.astype(np.float64), p["norm1.weight"], p["norm1.bias"])
    ref = np.pad(ref, ((0, 0), (0, Dp - D), (0, Hp - H), (0, Wp - W), (0, 0)))
    if any(ss):
        ref = np.roll(ref, shift=tuple(-s for s in ss), axis=(1, 2, 3))
    ref = sw.window_partition(ref, ws)
    assert rel_l2(xw, ref) < 1e-6
    # reverse is the exact inverse on the un-padded region
    yw = np.random.default_rng(0).standard_normal(xw.shape).astype(np.float32)
    y = np.full((B, D, H, W, C), np.nan, np.float32)
    assert lib.dlsg_swin_window_reverse_f32(ptr(yw), ptr(y), B, D, H, W, C, i3(ws), i3(ss), None) == 0
    r = sw.window_reverse(yw.reshape((-1,) + tuple(ws) + (C,)), ws, B, Dp, Hp, Wp)
    if any(ss):
        r = np.roll(r, shift=tuple(ss), axis=(1, 2, 3))
    assert np.array_equal(y, r[:, :D, :H, :W, :])


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("mode", ["none", "region", "dense"])
def test_attention_core(lib, name, mode):
    c = swin_case(name)
    x, p = c["x"], c["params"]
    B, D, H, W, C = x.shape
    heads, hd = c["heads"], C // c["heads"]
    ws, ss = sw.get_window_size((D, H, W), c["window"], c["shift"])
    Dp, Hp, Wp = (int(np.ceil(n / w)) * w for n, w in zip((D, H, W), ws))
    nW, N = (Dp // ws[0]) * (Hp // ws[1]) * (Wp // ws[2]), ws[0] * ws[1] * ws[2]
    rng = np.random.default_rng(5)
    xw = rng.standard_normal((B * nW, N, C)).astype(np.float32)
    qkv = (xw @ p["attn.qkv.weight"].T + p["attn.qkv.bias"]).astype(np.float32)
    mask = sw.compute_mask(Dp, Hp, Wp, ws, ss).astype(np.float32)
    ap = {k[5:]: v for k, v in p.items() if k.startswith("attn.")}
    ap["proj.weight"], ap["proj.bias"] = np.eye(C), np.zeros(C)
    ref = sw.window_attention(xw.astype(np.float64), ap, heads, c["window"], None if mode == "none" else mask)
    out = np.full((B * nW, N, C), np.nan, np.float32)
    table = np.ascontiguousarray(p["attn.relative_position_bias_table"])
    if mode == "none":
        args = (0, None, 1)
    elif mode == "region":
        rid = np.ascontiguousarray(region_ids(Dp, Hp, Wp, ws, ss))
        args = (1, ptr(rid), nW)
    else:
        args = (2, ptr(mask), nW)
    st = lib.dlsg_swin_window_attention_f32(ptr(qkv), ptr(table), ptr(out), B * nW, N, heads, hd, i3(c["window"]),
                                            hd ** -0.5, *args, None)
    assert st == 0
    assert rel_l2(out, ref) < 1e-3          # north_star: 1e-3 relative for the tensor-core attention


@pytest.mark.parametrize("name", ["small_shift", "small_pad_shift"])
def test_attention_core_fp16_inputs(lib, name):
    """The layout the fused qkv GEMM writes: q16 (pre-scaled by scale*log2e), k16, v16, each [B_,heads,N,32]."""
    c = swin_case(name)
    x, p = c["x"], c["params"]
    B, D, H, W, C = x.shape
    heads, hd = c["heads"], C // c["heads"]
    ws, ss = sw.get_window_size((D, H, W), c["window"], c["shift"])
    Dp, Hp, Wp = (int(np.ceil(n / w)) * w for n, w in zip((D, H, W), ws))
    nW, N = (Dp // ws[0]) * (Hp // ws[1]) * (Wp // ws[2]), ws[0] * ws[1] * ws[2]
    rng = np.random.default_rng(6)
    xw = rng.standard_normal((B * nW, N, C)).astype(np.float32)
    qkv = (xw @ p["attn.qkv.weight"].T + p["attn.qkv.bias"]).astype(np.float32).reshape(B * nW, N, 3, heads, hd)
    scale = hd ** -0.5
    q16 = np.zeros((B * nW, heads, N, 32), np.float16)
    k16 = np.zeros((B * nW, heads, N, 32), np.float16)
    vt16 = np.zeros((B * nW, heads, N, 32), np.float16)
    q16[..., :hd] = (qkv[:, :, 0].transpose(0, 2, 1, 3) * scale * np.log2(np.e)).astype(np.float16)
    k16[..., :hd] = qkv[:, :, 1].transpose(0, 2, 1, 3).astype(np.float16)
    vt16[..., :hd] = qkv[:, :, 2].transpose(0, 2, 1, 3).astype(np.float16)
    vt16[..., hd] = 1.0                                      # the column of ones the fused GEMM appends (head_dim < 32)
    mask = sw.compute_mask(Dp, Hp, Wp, ws, ss).astype(np.float32)
    rid = np.ascontiguousarray(region_ids(Dp, Hp, Wp, ws, ss))
    ap = {k[5:]: v for k, v in p.items() if k.startswith("attn.")}
    ap["proj.weight"], ap["proj.bias"] = np.eye(C), np.zeros(C)
    ref = sw.window_attention(xw.astype(np.float64), ap, heads, c["window"], mask)
    out = np.full((B * nW, N, C), np.nan, np.float32)
    table = np.ascontiguousarray(p["attn.relative_position_bias_table"])
    st = lib.dlsg_swin_window_attention_f16(ptr(q16), ptr(k16), ptr(vt16), ptr(table), ptr(out), B * nW, N, heads, hd,
                                            i3(c["window"]), 1, ptr(rid), nW, None)
    assert st == 0
    assert rel_l2(out, ref) < 1e-3


def test_attention_core_generic_path_odd_window_width(lib):
    """Odd window width: no (tab[i], tab[i-1]) pairing, explicit row sums (the non-FAST template instance)."""
    rng = np.random.default_rng(8)
    window, heads, C = (2, 3, 3), 2, 16
    N, hd, B_ = 18, 8, 3
    table = (0.5 * rng.standard_normal(((2 * 2 - 1) * 5 * 5, heads))).astype(np.float32)
    p = {"relative_position_bias_table": table, "qkv.weight": (rng.standard_normal((3 * C, C)) / 4).astype(np.float32),
         "qkv.bias": np.zeros(3 * C, np.float32), "proj.weight": np.eye(C), "proj.bias": np.zeros(C)}
    xw = rng.standard_normal((B_, N, C)).astype(np.float32)
    qkv = (xw @ p["qkv.weight"].T).astype(np.float32)
    ref = sw.window_attention(xw.astype(np.float64), p, heads, window, None)
    out = np.full((B_, N, C), np.nan, np.float32)
    st = lib.dlsg_swin_window_attention_f32(ptr(qkv), ptr(table), ptr(out), B_, N, heads, hd, i3(window), hd ** -0.5, 0, None, 1, None)
    assert st == 0 and rel_l2(out, ref) < 1e-3


def test_status_codes(lib):
    z = np.zeros(64, np.float32)
    w = i3((2, 4, 4))
    assert lib.dlsg_swin_window_attention_f32(None, ptr(z), ptr(z), 1, 32, 1, 8, w, 1.0, 0, None, 1, None) == -1
    assert lib.dlsg_swin_window_attention_f32(ptr(z), ptr(z), ptr(z), 1, 32, 1, 64, w, 1.0, 0, None, 1, None) == -3
    assert lib.dlsg_swin_window_attention_f32(ptr(z), ptr(z), ptr(z), 1, 33, 1, 8, w, 1.0, 0, None, 1, None) == -1
    assert lib.dlsg_swin_window_attention_f32(ptr(z), ptr(z), ptr(z), 3, 32, 1, 8, w, 1.0, 1, ptr(z), 2, None) == -1
    assert lib.dlsg_swin_ln_partition_f32(ptr(z), None, None, 1e-5, ptr(z), 1, 1, 1, 1, 1, i3((2, 2, 2)), i3((2, 0, 0)), None) == -1
