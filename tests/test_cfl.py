"""BART .cfl reader / writer (dl_swin_gan_b200.fileio.cfl, mirror of dl_cs/fileio/cfl.py): byte-exact files against a fixture
written by the unmodified reference (tests/golden/cfl_ref.*, made by oracle/make_golden_cfl.py) and round trips."""
import os

import numpy as np

from conftest import GOLDEN
from dl_swin_gan_b200.fileio import cfl


def _array():
    rng = np.random.default_rng(5)
    return (rng.standard_normal((2, 3, 4, 5)) + 1j * rng.standard_normal((2, 3, 4, 5))).astype(np.complex64)


def test_reads_files_written_by_the_reference():
    a = _array()
    assert np.array_equal(cfl.read(os.path.join(GOLDEN, "cfl_ref_c")), a)
    assert np.array_equal(cfl.readcfl(os.path.join(GOLDEN, "cfl_ref_f")), a)
    assert cfl.read_hdr(os.path.join(GOLDEN, "cfl_ref_c")) == [2, 3, 4, 5]
    assert cfl.read_hdr(os.path.join(GOLDEN, "cfl_ref_f"), order="F") == [2, 3, 4, 5]


def test_writes_the_same_bytes_as_the_reference(tmp_path):
    a = _array()
    cfl.write(str(tmp_path / "c"), a)
    cfl.writecfl(str(tmp_path / "f"), a)
    for ours, ref in (("c", "cfl_ref_c"), ("f", "cfl_ref_f")):
        for ext in (".hdr", ".cfl"):
            assert open(str(tmp_path / ours) + ext, "rb").read() == open(os.path.join(GOLDEN, ref) + ext, "rb").read()


def test_round_trip_and_pinned_reader(tmp_path):
    a = _array()[0]
    cfl.write(str(tmp_path / "x"), a)
    assert np.array_equal(cfl.read(str(tmp_path / "x")), a)
    t = cfl.read_pinned(str(tmp_path / "x"))
    assert np.array_equal(t.numpy(), a)
