"""BASELINE.json config 5: SENSE operator microbenchmark sweep (coils x frames x matrix), forward / adjoint GB/s."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
import __graft_entry__ as ge
ge.build()
from dl_swin_gan_b200.mri import SenseModel

dev = torch.device("cuda:0")
peak = bench.measured_peaks()[0]
g = torch.Generator(device=dev).manual_seed(1000)


def cn(*s):
    return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))


E = int(os.environ.get("SWEEP_E", 2))
for (Y, X) in [(128, 128), (160, 160), (192, 160), (192, 192), (256, 256), (320, 320)]:
    for C in (8, 16, 32):
        for T in (10, 20, 40, 60):
            per_slice = 8 * T * Y * X * (E + C)
            B = max(1, min(64, int(1.5e9 // per_slice)))          # ~1.5 GB of k-space per call, >> L2
            maps = cn(B, E, C, 1, Y, X)
            one = torch.from_numpy(bench.make_mask(min(B, 4), T, Y, X, 12, 1000)).to(dev)      # reference k-t masks (VDktMaskFunc, R = 12)
            mask = one.repeat((B + one.shape[0] - 1) // one.shape[0], 1, 1, 1, 1)[:B].contiguous()
            A = SenseModel(maps, weights=mask)
            x, k = cn(B, E, T, Y, X), cn(B, C, T, Y, X)
            with torch.no_grad():
                tf = bench.time_op(lambda: A(x), 3, dev)
                ta = bench.time_op(lambda: A(k, adjoint=True), 3, dev)
                tn = bench.time_op(lambda: A.normal(x), 3, dev)
                td = bench.time_op(lambda: A.dc_step(x, x, -2.0), 3, dev)
            nb = bench.algorithmic_bytes("forward", B, E, C, T, Y, X, mask.numel())
            nn_ = bench.algorithmic_bytes("normal", B, E, C, T, Y, X, mask.numel())
            nd = bench.algorithmic_bytes("dc_step", B, E, C, T, Y, X, mask.numel())
            print(json.dumps({"Y": Y, "X": X, "C": C, "T": T, "E": E, "B": B,
                              "fwd_GBps": round(nb / tf / 1e6), "adj_GBps": round(nb / ta / 1e6), "normal_GBps": round(nn_ / tn / 1e6),
                              "fwd_frac": round(nb / tf / 1e6 / peak, 3), "adj_frac": round(nb / ta / 1e6 / peak, 3),
                              "dc_frac": round(nd / td / 1e6 / peak, 3), "dc_us_per_slice": round(1e3 * td / B, 1),
                              "path": ("lines" if A._line_plans.get((B, E, C, T, Y, X)) is not None else "dense"),
                              "fwd_us_per_slice": round(1e3 * tf / B, 1), "adj_us_per_slice": round(1e3 * ta / B, 1),
                              "normal_us_per_slice": round(1e3 * tn / B, 1)}), flush=True)
            del A, maps, mask, x, k
            torch.cuda.empty_cache()
