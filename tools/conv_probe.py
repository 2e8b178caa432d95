"""cuDNN Conv3d 160->160, 3x3x3 at the Swin denoiser's size (2 slices: [2,160,28,192,160]) and 64->64 at the ResNet's:
layout (NCDHW / channels-last-3d) x precision (TF32 / bf16) x cudnn.benchmark, ms per call and TFLOP/s."""
import itertools, sys, time
import torch
dev = torch.device("cuda:0")
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for (B, C, D, H, W) in [(2, 160, 28, 192, 160), (8, 64, 32, 192, 160)]:
    conv = torch.nn.Conv3d(C, C, 3, padding=1).to(dev)
    x = torch.randn(B, C, D, H, W, device=dev)
    flops = 2.0 * B * D * H * W * C * C * 27
    for bench, cl, prec in itertools.product((False, True), (False, True), ("tf32", "bf16")):
        torch.backends.cudnn.benchmark = bench
        xi = x.contiguous(memory_format=torch.channels_last_3d) if cl else x
        cv = conv.to(memory_format=torch.channels_last_3d) if cl else conv.to(memory_format=torch.contiguous_format)
        with torch.no_grad():
            if prec == "bf16":
                def fn():
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return cv(xi)
            else:
                torch.backends.cudnn.allow_tf32 = True
                fn = lambda: cv(xi)
            ms = t(fn)
        print(f"C={C} benchmark={bench} channels_last={cl} {prec}: {ms:.2f} ms  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
