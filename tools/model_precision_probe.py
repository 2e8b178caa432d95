"""NRMSE of the Swin unrolled network on the B200 against the reference fixture for the precision choices available:
cuDNN TF32 on/off for the denoiser convolutions x fp16-operand tensor-core linears / attention on/off."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from test_models import build, golden, model_problem, rel_l2
from dl_swin_gan_b200.mri import SenseModel
from dl_swin_gan_b200.swin import window_attention as wa
dev = torch.device("cuda:0")
name = os.environ.get("PROBE_CASE", "swin_pgd_2unroll")
g = golden(name)
net = build(name).to(dev)
maps, mask, y = (torch.from_numpy(a).to(dev) for a in model_problem(name))
A = SenseModel(maps, weights=mask)
for tf32 in os.environ.get("PROBE_MODES", "tf32x3,tf32x2,tf32r,tf32,fp32").split(","):
    for tc in (True, False):
        from dl_swin_gan_b200.models import set_conv_precision
        set_conv_precision(tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        wa.WindowAttention3D.tc_linear = tc
        with torch.no_grad():
            recon = net(y, A)
            den = net.cnn_update[0](A(y, adjoint=True))
        print(f"conv {tf32} tc_linear(fp16)={tc}: denoiser {rel_l2(den.cpu().numpy(), g['denoised']):.2e} recon {rel_l2(recon.cpu().numpy(), g['recon']):.2e}")
