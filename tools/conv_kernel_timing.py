"""Time the tensor-core 3x3x3 convolution against cuDNN (TF32, one pass) at the denoisers' layer shapes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dl_swin_gan_b200.models.conv3d import conv3d_k3
dev = torch.device("cuda:0")
def timeit(f, n=5):
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (nb, c, t, y, x) in [(1, 160, 28, 192, 160), (2, 160, 28, 192, 160), (8, 64, 32, 192, 160)]:
    inp = torch.randn(nb, c, t, y, x, device=dev).contiguous(memory_format=torch.channels_last_3d)
    w = torch.randn(c, c, 3, 3, 3, device=dev) / (27 * c) ** 0.5
    b = torch.randn(c, device=dev)
    wc = w.contiguous(memory_format=torch.channels_last_3d)
    flop = 2.0 * 27 * c * c * nb * t * y * x
    ms = timeit(lambda: conv3d_k3(inp, w, b))
    torch.backends.cudnn.allow_tf32 = True
    ms_c = timeit(lambda: torch.nn.functional.conv3d(inp, wc, b, padding=1))
    print(f"[{nb},{c},{t},{y},{x}] tensor-core 3xTF32 kernel {ms:.3f} ms = {flop / ms / 1e9:.0f} TFLOP/s useful ({3 * flop / ms / 1e9:.0f} issued) | "
          f"cuDNN TF32 one pass {ms_c:.3f} ms = {flop / ms_c / 1e9:.0f} TFLOP/s (x3 = {3 * ms_c:.3f} ms)")
