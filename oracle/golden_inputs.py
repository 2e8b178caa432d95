"""Seeded inputs of the golden SENSE fixtures (shared by oracle/make_golden.py and tests/).

The fixtures store only what the reference produced; the inputs are regenerated from these
seeds (numpy PCG64 streams are stable across numpy versions), which keeps tests/golden small.
Test infrastructure — never imported by the product.
"""
import numpy as np

# name: (B, E, C, T, Y, X, weights kind)
CASES = {
    "tiny_odd": (1, 2, 3, 2, 12, 10, "mask"),
    "pow2_32": (2, 1, 2, 2, 32, 32, "none"),
    "mixed_64x96": (1, 2, 2, 3, 64, 96, "full"),
    "prime_7x13": (1, 1, 2, 2, 7, 13, "full"),
    "cine_tile_192x160": (1, 1, 2, 1, 192, 160, "mask"),
    "line_y1": (1, 2, 2, 2, 1, 64, "mask"),
    "line_x1": (1, 1, 2, 2, 16, 1, "full"),
}


def make_inputs(name, B, E, C, T, Y, X, wkind):
    seed = 1000 + sum(ord(ch) for ch in name)
    rng = np.random.default_rng(seed)

    def cn(*s):
        return ((rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(2)).astype(np.complex64)

    maps = (cn(B, E, C, 1, Y, X) / np.sqrt(E * C)).astype(np.complex64)
    x = cn(B, E, T, Y, X)
    k = cn(B, C, T, Y, X)
    aty = cn(B, E, T, Y, X)
    if wkind == "none":
        w = None
    elif wkind == "mask":
        w = (rng.random((B, 1, T, Y, X)) < 0.3).astype(np.float32)
    else:
        w = rng.standard_normal((B, C, T, Y, X)).astype(np.float32)
    return maps, x, k, aty, w




# ------------------------------------------------------------------ shifted-window attention fixtures
# name: (B, D, H, W, dim, heads, window, shift)
SWIN_CASES = {
    "small_shift": (2, 4, 8, 8, 16, 2, (2, 4, 4), (1, 2, 2)),
    "small_pad_noshift": (1, 3, 6, 5, 16, 2, (2, 4, 4), (0, 0, 0)),
    "small_pad_shift": (1, 5, 7, 9, 24, 3, (2, 4, 4), (1, 2, 2)),
    "clamped_depth": (1, 2, 8, 8, 16, 4, (4, 4, 4), (2, 2, 2)),          # D <= window: window and shift clamp (:72-85)
    "cine_cfg_shift": (1, 7, 16, 16, 160, 8, (7, 8, 8), (3, 4, 4)),       # hard-coded Swin config (swin3D.py:315), 4 windows
    "cine_cfg_noshift": (1, 7, 8, 16, 160, 8, (7, 8, 8), (0, 0, 0)),
}


def swin_case(name):
    B, D, H, W, dim, heads, window, shift = SWIN_CASES[name]
    rng = np.random.default_rng(2000 + sum(ord(ch) for ch in name))
    f32 = np.float32
    ntab = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
    params = {
        "norm1.weight": (1 + 0.1 * rng.standard_normal(dim)).astype(f32),
        "norm1.bias": (0.1 * rng.standard_normal(dim)).astype(f32),
        "attn.relative_position_bias_table": (0.5 * rng.standard_normal((ntab, heads))).astype(f32),
        "attn.qkv.weight": (rng.standard_normal((3 * dim, dim)) / np.sqrt(dim)).astype(f32),
        "attn.qkv.bias": (0.1 * rng.standard_normal(3 * dim)).astype(f32),
        "attn.proj.weight": (rng.standard_normal((dim, dim)) / np.sqrt(dim)).astype(f32),
        "attn.proj.bias": (0.1 * rng.standard_normal(dim)).astype(f32),
    }
    x = rng.standard_normal((B, D, H, W, dim)).astype(f32)
    # second half of the block (norm2 + Mlp, :208, :20-38): a separate stream so that the arrays above keep their values
    rng2 = np.random.default_rng(3000 + sum(ord(ch) for ch in name))
    hidden = 4 * dim
    params.update({
        "norm2.weight": (1 + 0.1 * rng2.standard_normal(dim)).astype(f32),
        "norm2.bias": (0.1 * rng2.standard_normal(dim)).astype(f32),
        "mlp.fc1.weight": (rng2.standard_normal((hidden, dim)) / np.sqrt(dim)).astype(f32),
        "mlp.fc1.bias": (0.1 * rng2.standard_normal(hidden)).astype(f32),
        "mlp.fc2.weight": (rng2.standard_normal((dim, hidden)) / np.sqrt(hidden)).astype(f32),
        "mlp.fc2.bias": (0.1 * rng2.standard_normal(dim)).astype(f32),
    })
    return dict(x=x, params=params, dim=dim, heads=heads, window=window, shift=shift)
