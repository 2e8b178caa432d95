"""Per-op CUDA-event times of the SENSE operator at the bench workload (default batch 64), with a parity check of the
sampled-line path against the dense kernels on a small batch.  QUICK_E selects the number of map sets."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from dl_swin_gan_b200.mri import SenseModel
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOAD); cfg["B"] = int(os.environ.get("QUICK_B", 64)); cfg["E"] = int(os.environ.get("QUICK_E", 1))
if cfg["E"] > 1:
    os.environ["DLSG_SAMPLED_LINES"] = "force"
maps, mask, y = bench.device_problem(cfg, dev, 1000)
A = SenseModel(maps, weights=mask)
with torch.no_grad():
    aty = A(y, adjoint=True)
    x = aty
    # parity on 2 slices: sampled-line kernels vs dense kernels
    A2 = SenseModel(maps[:2], weights=mask[:2])
    got = {"fwd": A2(x[:2]), "adj": A2(y[:2], adjoint=True), "dc": A2.dc_step(x[:2], aty[:2], -2.0)}
    os.environ["DLSG_SAMPLED_LINES"] = "0"
    A3 = SenseModel(maps[:2], weights=mask[:2])
    want = {"fwd": A3(x[:2]), "adj": A3(y[:2], adjoint=True), "dc": A3.dc_step(x[:2], aty[:2], -2.0)}
    os.environ["DLSG_SAMPLED_LINES"] = "force" if cfg["E"] > 1 else "1"
    err = {k: float((got[k] - want[k]).norm() / want[k].norm()) for k in got}
    t = {"fwd": bench.time_op(lambda: A(x), 5, dev), "adj": bench.time_op(lambda: A(y, adjoint=True), 5, dev),
         "normal": bench.time_op(lambda: A.normal(x), 5, dev), "dc": bench.time_op(lambda: A.dc_step(x, aty, -2.0), 5, dev)}
plan = A._line_plans.get(tuple(cfg[k] for k in "BECTYX"))
nb = bench.algorithmic_bytes("dc_step", *(cfg[k] for k in "BECTYX"), mask.numel())
print("lib %s E=%d lmax=%s | " % (os.path.basename(os.environ.get("DLSG_LIB_PATH", "shipped")), cfg["E"], plan[1] if plan else None)
      + " ".join("%s %.3f" % kv for kv in t.items()) + " | dc frac %.3f | err vs dense " % (nb / (t["dc"] * 1e-3) / 1e9 / 6451.2)
      + " ".join("%s %.1e" % kv for kv in err.items()))
