"""Timeline of one end-to-end step (bench.py's pipeline): per chunk, when copies / A^H y / DC steps / D2H finish (ms from start)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from dl_swin_gan_b200.mri import SenseModel
from dl_swin_gan_b200.unrolled import ProximalGradientDescent
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOAD)
B, E, C, T, Y, X = (cfg[k] for k in "BECTYX")
maps, mask, y = bench.device_problem(cfg, dev, 1000)
pgd = ProximalGradientDescent(num_unrolls=5, step_init=-2.0).to(dev)
hy = torch.empty(y.shape, dtype=y.dtype, pin_memory=True).copy_(y)
hm = torch.empty(maps.shape, dtype=maps.dtype, pin_memory=True).copy_(maps)
hw = torch.empty(mask.shape, dtype=mask.dtype, pin_memory=True).copy_(mask)
hx = torch.empty((B, E, T, Y, X), dtype=torch.complex64, pin_memory=True)
n_chunk = int(os.environ.get("CHUNKS", 8)); cb = B // n_chunk
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
xfer = torch.cuda.Stream(device=dev, priority=-1)
def ev(): return torch.cuda.Event(enable_timing=True)
def step(record):
    marks = []
    t_host = []
    start = ev(); start.record(xfer)
    for st in streams: st.wait_stream(xfer)
    h0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n_chunk):
            sl = slice(i * cb, (i + 1) * cb); st = streams[i % 2]
            with torch.cuda.stream(xfer):
                dw = hw[sl].to(dev, non_blocking=True); dm = hm[sl].to(dev, non_blocking=True)
                e_cp = ev(); e_cp.record(xfer)
                A_i = SenseModel(dm, weights=dw)
                aty_i = A_i(hy[sl], adjoint=True)
                e_aty = ev(); e_aty.record(xfer)
            with torch.cuda.stream(st):
                st.wait_event(e_aty)
                for tns in (dm, dw, aty_i): tns.record_stream(st)
                res = pgd(None, A_i, ATy=aty_i)
                e_dc = ev(); e_dc.record(st)
                hx[sl].copy_(res, non_blocking=True)
                e_out = ev(); e_out.record(st)
            marks.append((e_cp, e_aty, e_dc, e_out)); t_host.append((time.perf_counter() - h0) * 1e3)
    torch.cuda.synchronize()
    if record:
        for i, (a, b_, c, d) in enumerate(marks):
            print(f"chunk {i}: copies done {start.elapsed_time(a):6.2f}  A^H y done {start.elapsed_time(b_):6.2f}  DC done {start.elapsed_time(c):6.2f}  D2H done {start.elapsed_time(d):6.2f}   host enqueued at {t_host[i]:6.2f}")
for k in range(3): step(k == 2)
