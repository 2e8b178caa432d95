"""Sampled-line vs dense kernels over matrix sizes / map sets (10 % random ky lines): ms per call for forward / adjoint / normal."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from dl_swin_gan_b200.mri import SenseModel
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1000)
def cn(*s): return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))
frac = float(os.environ.get("LINE_FRACTION", 0.1))
for E in (1, 2):
    for (Y, X) in [(128, 128), (192, 160), (256, 256), (320, 320)]:
        for C in (8, 32):
            T = 20
            B = max(1, min(32, int(1.0e9 // (8 * T * Y * X * (E + C)))))
            maps = cn(B, E, C, 1, Y, X)
            mask = (torch.rand(B, 1, T, Y, 1, generator=g, device=dev) < frac).float().expand(B, 1, T, Y, X).contiguous()
            x, k = cn(B, E, T, Y, X), cn(B, C, T, Y, X)
            row = {"E": E, "Y": Y, "X": X, "C": C, "B": B}
            for mode in ("1", "0"):
                os.environ["DLSG_SAMPLED_LINES"] = mode
                A = SenseModel(maps, weights=mask)
                with torch.no_grad():
                    t = [bench.time_op(f, 3, dev) for f in (lambda: A(x), lambda: A(k, adjoint=True), lambda: A.normal(x))]
                key = "lines" if A._line_plans.get((B, E, C, T, Y, X)) is not None else "dense"
                row[key + ("" if mode == "1" else "_forced")] = [round(v, 3) for v in t]
            print(json.dumps(row), flush=True)
            del maps, mask, x, k
            torch.cuda.empty_cache()
