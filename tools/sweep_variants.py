"""Time the SENSE ops for each tuning variant of the library (dl_swin_gan_b200/csrc/variants/libdlsg_*.so)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import torch, bench
from dl_swin_gan_b200.mri import SenseModel
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOAD); cfg["B"] = int(os.environ.get("SWEEP_B", 64)); cfg["E"] = int(os.environ.get("SWEEP_E", 1))
maps, mask, y = bench.device_problem(cfg, dev, 1000)
A = SenseModel(maps, weights=mask)
x = A(y, adjoint=True)
B, E, C, T, Y, X = (cfg[k] for k in "BECTYX")
for mb in [int(v) for v in os.environ.get("SWEEP_MB", "64,8192").split(",")]:
    os.environ["DLSG_WORKSPACE_MB"] = str(mb)
    r = {"lib": os.path.basename(os.environ.get("DLSG_LIB_PATH", "default")), "E": E, "workspace_MB": mb}
    with torch.no_grad():
        for name, fn in (("forward", lambda: A(x)), ("adjoint", lambda: A(y, adjoint=True)),
                         ("normal", lambda: A.normal(x)), ("dc_step", lambda: A.dc_step(x, x, -2.0))):
            ms = bench.time_op(fn, 5, dev)
            nbytes = bench.algorithmic_bytes(name, B, E, C, T, Y, X, mask.numel())
            r[name] = {"ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1)}
    print(json.dumps(r), flush=True)
''' % ROOT

libs = sorted(glob.glob(os.path.join(ROOT, "dl_swin_gan_b200", "csrc", "variants", "libdlsg_*.so")))
only = os.environ.get("SWEEP_LIBS")
for lib in libs:
    if only and not any(("libdlsg_%s.so" % o) == os.path.basename(lib) for o in only.split(",")):
        continue
    env = dict(os.environ, DLSG_LIB_PATH=lib)
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
