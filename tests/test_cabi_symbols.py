"""The nvcc-built C-ABI library loads on a GPU-less box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(dlsg_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    from dl_swin_gan_b200 import _lib
    return _lib


def test_header_declares_the_hot_path_entry_points():
    names = declared_functions()
    for must in ("dlsg_sense_forward_c64", "dlsg_sense_adjoint_c64", "dlsg_sense_normal_c64",
                 "dlsg_sense_dc_step_c64", "dlsg_sense_workspace_bytes", "dlsg_status_string", "dlsg_version"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by {built_lib.LIB_PATH}"
    assert set(built_lib.SIGNATURES) == set(declared_functions())


def test_argument_validation_returns_before_any_cuda_call(built_lib):
    lib = built_lib.load()
    assert lib.dlsg_version() >= 100
    assert b"ok" in lib.dlsg_status_string(0)
    # invalid arguments are rejected on the host: no launch, so this is safe without a GPU
    assert lib.dlsg_sense_forward_c64(None, None, None, None, None, None, 0, 1, 1, 1, 1, 32, 32, 2, None) == built_lib.DLSG_EINVAL
    assert lib.dlsg_sense_adjoint_c64(None, None, None, None, None, None, 0, 1, 1, 1, 1, 32, 0, 2, None) == built_lib.DLSG_EINVAL
    assert lib.dlsg_sense_workspace_bytes(64, 1, 8, 20, 192, 160, 2, 0) == 8 * 192 * 160 * 8
    assert lib.dlsg_sense_workspace_bytes(64, 1, 8, 20, 192, 160, 2, 64 << 20) == 8 * 20 * 192 * 160 * 8


def test_cpu_tensors_fail_loudly(built_lib):
    import torch
    from dl_swin_gan_b200.mri import SenseModel
    A = SenseModel(torch.randn(1, 1, 2, 1, 8, 8, dtype=torch.complex64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        A(torch.randn(1, 1, 2, 8, 8, dtype=torch.complex64))
    with pytest.raises(AssertionError):
        A(torch.randn(1, 1, 2, 8, 8))
