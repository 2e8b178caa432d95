"""BART `.cfl` / `.hdr` complex-float files: mirror of dl_cs/fileio/cfl.py:12-67 (the format the reference's inference script
reads k-space and ESPIRiT maps from and writes images to, scripts/reconstruct.py:59-120, 245).

A dataset is two files: `<name>.hdr` - a comment line, then the dimensions in BART's (column-major, fastest first) order - and
`<name>.cfl` - the raw complex64 samples.  `order='C'` (the default of `read` / `write`) presents the array with the
dimensions reversed, so that the last numpy axis is BART's first; `readcfl` / `writecfl` keep BART's order with Fortran layout.
Host-side I/O (numpy), as in the reference; `read_pinned` additionally returns the samples in pinned host memory so that
`SenseModel(...)(kspace, adjoint=True)` can read the sampled lines in place.
"""
from __future__ import annotations

import numpy as np


def read_hdr(name, order="C"):
    """Dimensions from `<name>.hdr` (second line), reversed for C order."""
    with open(name + ".hdr", "r") as h:
        h.readline()                                   # "# Dimensions"
        dims = [int(tok) for tok in h.readline().split()]
    if order == "C":
        dims.reverse()
    return dims


def read(name, order="C"):
    dims = read_hdr(name, order)
    n = int(np.prod(dims))
    with open(name + ".cfl", "rb") as d:
        a = np.fromfile(d, dtype=np.complex64, count=n)
    return a.reshape(dims, order=order)


def readcfl(name):
    return read(name, order="F")


def write(name, array, order="C"):
    array = np.asarray(array)
    shape = array.shape[::-1] if order == "C" else array.shape
    with open(name + ".hdr", "w") as h:
        h.write("# Dimensions\n")
        h.write("".join("%d " % i for i in shape))
        h.write("\n")
    with open(name + ".cfl", "wb") as d:
        (array if order == "C" else array.T).astype(np.complex64).tofile(d)      # .T: column-major sample order


def writecfl(name, array):
    write(name, array, order="F")


def read_pinned(name, order="C"):
    """`read` into page-locked host memory (a torch tensor), ready for the host-resident k-space adjoint."""
    import torch
    a = read(name, order)
    t = torch.empty(a.shape, dtype=torch.complex64, pin_memory=torch.cuda.is_available())
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return t
