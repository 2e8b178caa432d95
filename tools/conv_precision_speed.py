"""Time one Swin denoiser pass (config_swin size, 1 slice) per convolution precision."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dl_swin_gan_b200.config import load_cfg
from dl_swin_gan_b200.models import build_unrolled, set_conv_precision
dev = torch.device("cuda:0")
net = build_unrolled(load_cfg(os.environ.get("CFG", "config_swin"))).to(dev).eval()
den = net.cnn_update[0]
E = 2 if os.environ.get("CFG", "config_swin") == "config_swin" else 1
x = torch.view_as_complex(torch.randn(int(os.environ.get("NB", "2")), E, 20, 192, 160, 2, device=dev))
from dl_swin_gan_b200.models import layers
for mode in os.environ.get("PROBE_MODES", "tf32x3,tf32x3-cudnn,tf32,bf16").split(","):
    layers._TC_CONV = not mode.endswith("-cudnn")
    set_conv_precision(mode.replace("-cudnn", ""))
    with torch.no_grad():
        for _ in range(2):
            den(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            den(x)
        e1.record(); torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1) / 3:.2f} ms per denoiser pass, {x.shape[0]} slices")
