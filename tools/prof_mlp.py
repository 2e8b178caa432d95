"""forward_part2 (tcgen05 MLP kernels) a few times at the cine size, for ncu."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dl_swin_gan_b200.swin import window_attention as wa
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = wa.SwinTransformerBlock3D(160, 8, (7, 8, 8), (3, 4, 4), qkv_bias=True).to(dev).eval()
x = torch.randn(int(os.environ.get("SWIN_B", 8)), 7, 48, 40, 160, device=dev)
with torch.no_grad():
    for _ in range(3):
        y = blk._part2_tensor_core(x, residual=True)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
