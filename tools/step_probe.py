"""Per-op CUDA-event times inside one bench step (A^H y + 5 DC updates), to compare with the isolated op timings."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from dl_swin_gan_b200.mri import SenseModel
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOAD)
maps, mask, y = bench.device_problem(cfg, dev, 1000)
A = SenseModel(maps, weights=mask)
step = torch.tensor([-2.0], device=dev)
def one_step(record):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    with torch.no_grad():
        ev[0].record()
        aty = A(y, adjoint=True); ev[1].record()
        x = aty
        for i in range(5):
            x = A.dc_step(x, aty, step); ev[2 + i].record()
    torch.cuda.synchronize()
    if record:
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(6)]
        print("adj %.3f | dc %s | total %.3f | max|x| %.3e finite %s" % (ts[0], " ".join("%.3f" % v for v in ts[1:]), ev[0].elapsed_time(ev[6]),
              float(x.abs().max()), bool(torch.isfinite(x.abs().max()))))
for k in range(5):
    one_step(k >= 2)
