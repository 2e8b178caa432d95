"""The tensor-core 3 x 3 x 3 convolution (csrc/conv3d_tc.cuh) against fp64 torch on the B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # NB, Cin, Cout, T, Y, X, bias
    (1, 32, 32, 3, 5, 7, True),          # a single partial tile
    (2, 64, 64, 4, 9, 16, True),         # several tiles, batch boundary inside a tile
    (1, 160, 160, 3, 16, 24, True),      # the Swin denoiser's channel count
    (1, 64, 32, 2, 33, 21, False),
    (1, 96, 160, 5, 12, 40, True),
    (2, 4, 160, 4, 10, 24, True),        # first layer of the Swin denoiser (re / im of two map sets)
    (1, 2, 64, 5, 17, 20, True),         # first layer of the ResNet denoiser at E = 1
    (1, 4, 32, 2, 3, 5, False),
]


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


@pytest.mark.parametrize("case", CASES)
def test_conv3d_k3_matches_fp64(case):
    from dl_swin_gan_b200.models.conv3d import conv3d_k3
    nb, cin, cout, t, y, x, has_bias = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(cin * 1000 + cout + t)
    inp = torch.randn(nb, cin, t, y, x, generator=g, device=dev)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, device=dev) / (27 * cin) ** 0.5
    b = torch.randn(cout, generator=g, device=dev) if has_bias else None
    ref = torch.nn.functional.conv3d(inp.double(), w.double(), b.double() if has_bias else None, padding=1)
    out = conv3d_k3(inp, w, b)
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last_3d)
    # 3xTF32 with the tensor core's fp32 accumulator: 2e-6 at K = 27*32, 1.5e-5 at K = 27*160 (measured); TF32 alone is 5e-4
    assert rel(out, ref) < 3e-5, rel(out, ref)
    # fused pre-activation and skip connection: conv(relu(x)) + b + 2 r
    res = torch.randn(nb, cout, t, y, x, generator=g, device=dev)
    out_r = conv3d_k3(inp, w, b, relu_in=True, residual=res, residual_scale=2.0)
    ref_r = torch.nn.functional.conv3d(inp.double().relu(), w.double(), b.double() if has_bias else None, padding=1) + 2.0 * res.double()
    assert rel(out_r, ref_r) < 3e-5
    # skip connection through a ReLU (ResBlock's in-place pre-activation on the tensor its skip re-reads)
    out_rr = conv3d_k3(inp, w, b, relu_in=True, residual=res, residual_scale=1.0, residual_relu=True)
    ref_rr = torch.nn.functional.conv3d(inp.double().relu(), w.double(), b.double() if has_bias else None, padding=1) + res.double().relu()
    assert rel(out_rr, ref_rr) < 3e-5


@pytest.mark.parametrize("case", [(1, 32, 64, 2, 3, 5, 2), (2, 160, 160, 2, 6, 5, 4), (1, 64, 32, 3, 9, 7, 1)])
def test_patch_embed_and_unembed_match_fp64(case):
    from dl_swin_gan_b200.models.conv3d import conv3d_patch, conv3d_unpatch
    nb, cin, cout, t, y, x, p = case                      # token grid t, y, x; patch p
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(p * 100 + cin)
    vol = torch.randn(nb, cin, t * p, y * p, x * p, generator=g, device=dev)
    w = torch.randn(cout, cin, p, p, p, generator=g, device=dev) / (p ** 3 * cin) ** 0.5
    b = torch.randn(cout, generator=g, device=dev)
    ref = torch.nn.functional.conv3d(vol.double(), w.double(), b.double(), stride=p)
    out = conv3d_patch(vol, w, b)
    # K = p^3 * Cin = 10 240 accumulations in the tensor core's fp32 accumulator at the Swin size: 5e-5 (measured); one TF32 pass is 5e-4
    assert out.shape == ref.shape and rel(out, ref) < 1e-4, rel(out, ref)
    tok = torch.randn(nb, cin, t, y, x, generator=g, device=dev)
    wt = torch.randn(cin, cout, p, p, p, generator=g, device=dev) / cin ** 0.5
    ref_t = torch.nn.functional.conv_transpose3d(tok.double(), wt.double(), b.double(), stride=p)
    out_t = conv3d_unpatch(tok, wt, b)
    assert out_t.shape == ref_t.shape and rel(out_t, ref_t) < 3e-5, rel(out_t, ref_t)


@pytest.mark.parametrize("case", [(1, 160, 4, 3, 9, 20), (2, 64, 2, 4, 7, 13), (1, 32, 4, 1, 1, 3)])
def test_last_layer_kernel_matches_fp64(case):
    from dl_swin_gan_b200.models.conv3d import conv3d_k3_fewout, conv3d_k3_fewout_gemm
    nb, cin, cout, t, y, x = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(cin + cout)
    inp = torch.randn(nb, cin, t, y, x, generator=g, device=dev)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, device=dev) / (27 * cin) ** 0.5
    b = torch.randn(cout, generator=g, device=dev)
    ref = torch.nn.functional.conv3d(inp.double(), w.double(), b.double(), padding=1)
    assert rel(conv3d_k3_fewout(inp, w, b), ref) < 2e-6                      # plain fp32 FMAs
    ref_r = torch.nn.functional.conv3d(inp.double().relu(), w.double(), None, padding=1)
    assert rel(conv3d_k3_fewout(inp, w, None, relu_in=True), ref_r) < 2e-6
    # the same layer as pointwise tensor-core GEMM + sum over the shifted taps (split bf16 operands)
    assert rel(conv3d_k3_fewout_gemm(inp, w, b), ref) < 3e-5
    assert rel(conv3d_k3_fewout_gemm(inp, w, None, relu_in=True), ref_r) < 3e-5


@pytest.mark.parametrize("case", [(2, 96, 160, 7, 9, 12, 2, 6), (3, 64, 64, 8, 10, 16, 1, 7), (2, 4, 32, 6, 5, 9, 0, 4), (1, 64, 64, 5, 6, 7, 3, 5)])
def test_frame_window_computes_exactly_the_requested_frames(case):
    """t_range: the kernels write only the requested output frames of every sample, bit-identical to the full convolution there
    (the denoisers use it to skip the outer frames of the circular time padding that no later layer reads)."""
    from dl_swin_gan_b200.models.conv3d import conv3d_k3, conv3d_k3_fewout_gemm
    nb, cin, cout, t, y, x, lo, hi = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(cin + 7 * cout + lo)
    inp = torch.randn(nb, cin, t, y, x, generator=g, device=dev)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, device=dev) / (27 * cin) ** 0.5
    b = torch.randn(cout, generator=g, device=dev)
    res = torch.randn(nb, cout, t, y, x, generator=g, device=dev)
    full = conv3d_k3(inp, w, b, relu_in=True, residual=res)
    part = conv3d_k3(inp, w, b, relu_in=True, residual=res, t_range=(lo, hi))
    assert torch.equal(part[:, :, lo:hi], full[:, :, lo:hi])
    # a window reaching beyond the volume is clipped
    part = conv3d_k3(inp, w, b, t_range=(lo - 5, t + 3))
    assert torch.equal(part, conv3d_k3(inp, w, b))
    if cin % 32 == 0:
        w4 = torch.randn(4, cin, 3, 3, 3, generator=g, device=dev) / (27 * cin) ** 0.5
        full4 = conv3d_k3_fewout_gemm(inp, w4, b[:4])
        part4 = conv3d_k3_fewout_gemm(inp, w4, b[:4], t_range=(lo, hi))
        assert torch.equal(part4[:, :, lo:hi], full4[:, :, lo:hi])


def test_volumes_beyond_the_kernels_limits_take_the_library_path():
    from dl_swin_gan_b200.models import conv3d
    conv = torch.nn.Conv3d(64, 64, 3, padding=1)
    big = torch.empty(0, device="cuda:0").new_empty((1, 64, 1, 8, 8)).expand(64, 64, 2048, 8, 8)      # a view: shape only, no memory
    with torch.no_grad():
        assert conv3d.kind(conv.cuda(), big) is None
        assert conv3d.kind(conv.cuda(), torch.zeros(1, 64, 4, 8, 8, device="cuda:0")) == conv3d.K3


def test_conv_forward_routes_to_the_kernel_and_matches_cudnn_path():
    from dl_swin_gan_b200 import _lib
    from dl_swin_gan_b200.models import layers
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    conv = torch.nn.Conv3d(64, 64, 3, padding=1).to(dev)
    x = torch.randn(2, 64, 6, 20, 24, device=dev)
    keep = layers.get_conv_precision()
    try:
        layers.set_conv_precision("tf32x3")
        n0 = _lib.load().dlsg_launch_count()
        with torch.no_grad():
            y = layers.conv_forward(conv, x)
        assert _lib.load().dlsg_launch_count() == n0 + 1
        layers.set_conv_precision("fp32")
        with torch.no_grad():
            y32 = layers.conv_forward(conv, x)
    finally:
        layers.set_conv_precision(keep)
    assert rel(y, y32.double()) < 2e-5
    # training keeps the differentiable cuDNN path
    xg = x.clone().requires_grad_(True)
    n0 = _lib.load().dlsg_launch_count()
    layers.conv_forward(conv, xg).sum().backward()
    assert _lib.load().dlsg_launch_count() == n0 and xg.grad is not None
