"""Is the bench step time stable?  30 x (10 steps) in one process, per-kernel-chain breakdown of the slowest repeat."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from dl_swin_gan_b200.mri import SenseModel
from dl_swin_gan_b200.unrolled import ProximalGradientDescent
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOAD)
maps, mask, y = bench.device_problem(cfg, dev, 1000)
A = SenseModel(maps, weights=mask)
pgd = ProximalGradientDescent(num_unrolls=5, step_init=-2.0).to(dev)
def step():
    with torch.no_grad():
        return pgd(y, A)
for _ in range(20): step()
torch.cuda.synchronize()
res = []
for rep in range(30):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    evs[0].record()
    for i in range(10):
        out = step(); evs[i + 1].record()
    torch.cuda.synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(10)]
    res.append((sum(per) / 10, max(per), min(per)))
print("ms/step per repeat:", " ".join("%.2f" % r[0] for r in res))
print("worst single step %.2f, best %.2f" % (max(r[1] for r in res), min(r[2] for r in res)))
print("address of y %x, out %x" % (y.data_ptr(), out.data_ptr()))

# same with the step captured once in a CUDA graph (host work per step = one launch)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out_g = step()
res = []
for rep in range(30):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    evs[0].record()
    for i in range(10):
        g.replay(); evs[i + 1].record()
    torch.cuda.synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(10)]
    res.append((sum(per) / 10, max(per), min(per)))
print("graph ms/step per repeat:", " ".join("%.2f" % r[0] for r in res))
print("graph worst single step %.2f, best %.2f; equal to eager: %s" % (max(r[1] for r in res), min(r[2] for r in res), bool(torch.equal(out_g, out))))
