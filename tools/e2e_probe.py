"""Where does the end-to-end step go?  Times the host-k-space adjoint, the copies and pipeline shapes (chunks x streams)."""
import os, sys, time, itertools
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

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

with torch.no_grad():
    cb = 16
    A = SenseModel(maps[:cb], weights=mask[:cb]); A(y[:cb], adjoint=True)
    plan = A._line_plans[(cb, E, C, T, Y, X)]
    rows = int(plan[0][: cb * T].sum().item()); yb = rows * C * X * 8
    t_dev = timed(lambda: A(y[:cb], adjoint=True))
    t_host = timed(lambda: A(hy[:cb], adjoint=True))
    print(f"adjoint 16 slices: device k {t_dev:.3f} ms, pinned-host k {t_host:.3f} ms -> gather of {yb/1e6:.1f} MB at {yb/1e6/(t_host-t_dev+1e-9):.1f} GB/s (incremental)")
    t_cp = timed(lambda: hy[:cb].to(dev, non_blocking=True))
    print(f"dense H2D of 16 slices ({hy[:cb].numel()*8/1e6:.0f} MB): {t_cp:.3f} ms = {hy[:cb].numel()*8/1e6/t_cp:.1f} GB/s")
    t_mm = timed(lambda: (hm[:cb].to(dev, non_blocking=True), hw[:cb].to(dev, non_blocking=True)))
    print(f"maps+mask H2D of 16 slices: {t_mm:.3f} ms")
    t_pl = timed(lambda: SenseModel(maps[:cb], weights=mask[:cb])._line_plan((cb, E, C, T, Y, X), mask[:cb].expand(cb, C, T, Y, X), 2))
    print(f"plan build + read-back 16 slices: {t_pl:.3f} ms")
    t_d2h = timed(lambda: hx[:cb].copy_(y[:cb, 0:1], non_blocking=True))
    print(f"D2H of 16 slices: {t_d2h:.3f} ms")
    t_comp = timed(lambda: pgd(y[:cb], A))
    print(f"compute 16 slices (device y): {t_comp:.3f} ms")

    for n_chunk, n_streams in itertools.product((4, 8, 16), (2, 3, 4)):
        cb = B // n_chunk
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        def step():
            for i in range(n_chunk):
                st = streams[i % n_streams]; sl = slice(i * cb, (i + 1) * cb)
                with torch.cuda.stream(st):
                    dm = hm[sl].to(dev, non_blocking=True); dw = hw[sl].to(dev, non_blocking=True)
                    res = pgd(hy[sl], SenseModel(dm, weights=dw))
                    hx[sl].copy_(res, non_blocking=True)
            for st in streams:
                torch.cuda.current_stream(dev).wait_stream(st)
        ms = timed(step, 4)
        print(f"e2e chunks={n_chunk:2d} streams={n_streams}: {ms:.2f} ms/step = {B/ms*1e3:.0f} slices/s")
