"""Run each SENSE op once (plus warm-up) at a given batch for ncu captures."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from dl_swin_gan_b200.mri import SenseModel
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOAD); cfg["B"] = int(os.environ.get("PROF_B", 16)); cfg["E"] = int(os.environ.get("PROF_E", 1))
maps, mask, y = bench.device_problem(cfg, dev, 1000)
A = SenseModel(maps, weights=mask)
with torch.no_grad():
    for _ in range(int(os.environ.get("PROF_REPS", 2))):
        x = A(y, adjoint=True)
        k = A(x)
        n = A.normal(x)
        d = A.dc_step(x, x, -2.0)
torch.cuda.synchronize()
print("ok", float(d.abs().mean()))
