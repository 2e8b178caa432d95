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


