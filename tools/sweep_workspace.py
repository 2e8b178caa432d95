"""Sweep the L2-resident intermediate size (DLSG_WORKSPACE_MB) and the SENSE op timings at the bench shape."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench

import __graft_entry__ as ge
ge.build()
from dl_swin_gan_b200.mri import SenseModel

dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOAD)
cfg["B"] = int(os.environ.get("SWEEP_B", 64))
cfg["E"] = int(os.environ.get("SWEEP_E", 1))
maps, mask, y = bench.device_problem(cfg, dev, 1000)
A = SenseModel(maps, weights=mask)
x = A(y, adjoint=True)
B, E, C, T, Y, X = (cfg[k] for k in "BECTYX")
rows = []
for mb in [int(v) for v in os.environ.get("SWEEP_MB", "8,20,40,64,80,120,200,400,8192").split(",")]:
    os.environ["DLSG_WORKSPACE_MB"] = str(mb)
    r = {"workspace_MB": mb}
    with torch.no_grad():
        for name, fn in (("forward", lambda: A(x)), ("adjoint", lambda: A(y, adjoint=True)),
                         ("normal", lambda: A.normal(x)), ("dc_step", lambda: A.dc_step(x, x, -2.0))):
            ms = bench.time_op(fn, 5, dev)
            nbytes = bench.algorithmic_bytes(name, B, E, C, T, Y, X, mask.numel())
            r[name] = {"ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1)}
    rows.append(r)
    print(json.dumps(r), flush=True)
# torch eager reference of the same ops on this GPU (cuFFT + ATen), for context
import oracle.sense_oracle_torch as st
with torch.no_grad():
    for name, fn in (("forward", lambda: st.forward(x, maps, mask)), ("adjoint", lambda: st.adjoint(y, maps, mask)),
                     ("normal", lambda: st.adjoint(st.forward(x, maps, mask), maps, mask))):
        ms = bench.time_op(fn, 3, dev)
        nbytes = bench.algorithmic_bytes(name, B, E, C, T, Y, X, mask.numel())
        print(json.dumps({"torch_eager_gpu": name, "ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
