"""One launch of the tensor-core convolution at the Swin denoiser's layer shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dl_swin_gan_b200.models.conv3d import conv3d_k3
dev = torch.device("cuda:0")
c = int(os.environ.get("CONV_C", "160"))
inp = torch.randn(1, c, 28, 192, 160, device=dev).contiguous(memory_format=torch.channels_last_3d)
w = torch.randn(c, c, 3, 3, 3, device=dev) / (27 * c) ** 0.5
for _ in range(2):
    conv3d_k3(inp, w, None)
torch.cuda.synchronize()
