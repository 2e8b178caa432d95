import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import __graft_entry__ as ge
ge.build()
from dl_swin_gan_b200.swin import window_attention as wa
import bench
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (M, N, K) in [(128, 32, 16), (128, 160, 160), (300, 480, 160), (107520, 480, 160), (107520, 160, 160)]:
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    y = wa.linear_tc(x, w, b)
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    err = float((y.double() - ref).norm() / ref.norm())
    mx = float((y.double() - ref).abs().max())
    t = bench.time_op(lambda: wa.linear_tc(x, w, b), 5, dev) if M > 1000 else 0.0
    t_ref = bench.time_op(lambda: torch.nn.functional.linear(x, w, b), 5, dev) if M > 1000 else 0.0
    print(json.dumps({"M": M, "N": N, "K": K, "rel_err": err, "max_abs": mx, "tc_ms": round(t, 4), "torch_fp32_ms": round(t_ref, 4),
                      "TFLOPs": round(2.0 * M * N * K / max(t, 1e-9) / 1e9, 1)}), flush=True)
