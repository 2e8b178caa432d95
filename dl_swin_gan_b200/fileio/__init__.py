"""Disk formats either side of the path: BART .cfl/.hdr (mirror of dl_cs/fileio/cfl.py)."""
from . import cfl  # noqa: F401
