"""PCIe: one-directional vs simultaneous H2D + D2H copy rates, and the zero-copy gather alongside a D2H copy."""
import time, torch
dev = torch.device("cuda:0")
n = 1 << 28                                     # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_()
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.empty(n, dtype=torch.float32, device=dev).normal_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); el = (time.perf_counter() - t0) / reps
    return el
for name, a, b in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
    run(a, b, 1); el = run(a, b)
    gb = 4 * n / 1e9
    print(f"{name:9s}: {el*1e3:7.2f} ms per GiB-pair -> H2D {gb/el if a else 0:5.1f} GB/s, D2H {gb/el if b else 0:5.1f} GB/s, total {gb*(a+b)/el:5.1f} GB/s")
