"""Kernel-time breakdown of one denoiser pass (torch.profiler, CUDA activities) - which kernels the time of a reconstruction goes to."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from dl_swin_gan_b200.config import load_cfg
from dl_swin_gan_b200.models import build_unrolled, set_conv_precision
dev = torch.device("cuda:0")
cfg = os.environ.get("CFG", "config_swin")
net = build_unrolled(load_cfg(cfg)).to(dev).eval()
den = net.cnn_update[0]
E = 2 if cfg == "config_swin" else 1
x = torch.view_as_complex(torch.randn(int(os.environ.get("NB", "2")), E, 20, 192, 160, 2, device=dev))
set_conv_precision(os.environ.get("PREC", "tf32x3"))
with torch.no_grad():
    den(x); den(x); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        den(x); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"{cfg} NB={x.shape[0]} total {tot / 1e3:.2f} ms")
for e in rows[:int(os.environ.get("TOP", "18"))]:
    print(f"{e.device_time_total / 1e3:8.3f} ms {100 * e.device_time_total / tot:5.1f}% x{e.count:<4d} {e.key[:110]}")
