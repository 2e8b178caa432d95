import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.fixture(scope="session")
def golden_cases():
    """name -> (inputs dict, reference outputs npz) for every committed SENSE fixture."""
    from oracle.golden_inputs import CASES, make_inputs
    out = {}
    for name, (B, E, C, T, Y, X, wkind) in CASES.items():
        maps, x, k, aty, w = make_inputs(name, B, E, C, T, Y, X, wkind)
        ref = np.load(os.path.join(GOLDEN, f"sense_{name}.npz"))
        out[name] = (dict(maps=maps, x=x, k=k, aty=aty, weights=w, dims=(B, E, C, T, Y, X)), ref)
    return out


def _build_host_lib(src_rel, out_name, extra=()):
    build_dir = os.path.join(ROOT, "tests", "host", "_build")
    os.makedirs(build_dir, exist_ok=True)
    out = os.path.join(build_dir, out_name)
    src = os.path.join(ROOT, src_rel)
    deps = [src] + [os.path.join(ROOT, "dl_swin_gan_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "dl_swin_gan_b200", "csrc"))
                    if f.endswith((".cu", ".cuh", ".h"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cmd = ["g++", "-std=c++17", "-O1", "-pthread", "-shared", "-fPIC", "-x", "c++", *extra, "-o", out, src]
        subprocess.run(cmd, check=True, cwd=ROOT)
    return out


@pytest.fixture(scope="session")
def fft_radix_host_lib():
    return _build_host_lib("tests/host/fft_radix_host.cpp", "fft_radix_host.so")


@pytest.fixture(scope="session")
def sense_emu_lib():
    """The product's kernel + dispatch sources compiled by g++ under the thread emulation (test-only)."""
    return _build_host_lib("tests/host/sense_emu.cpp", "libdlsg_emu.so")
