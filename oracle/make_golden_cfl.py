"""Write tests/golden/cfl_ref_{c,f}.{hdr,cfl} with the UNMODIFIED reference writer (/root/reference/dl_cs/fileio/cfl.py).
Build container only:  python oracle/make_golden_cfl.py"""
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, "/root/reference")
from dl_cs.fileio import cfl as ref   # noqa: E402

rng = np.random.default_rng(5)
a = (rng.standard_normal((2, 3, 4, 5)) + 1j * rng.standard_normal((2, 3, 4, 5))).astype(np.complex64)
ref.write(os.path.join(OUT, "cfl_ref_c"), a)
ref.writecfl(os.path.join(OUT, "cfl_ref_f"), a)
print("written", [f for f in os.listdir(OUT) if f.startswith("cfl_ref")])
