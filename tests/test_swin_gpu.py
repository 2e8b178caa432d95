"""Parity of the CUDA shifted-window attention path on a B200 against outputs of the unmodified reference
modules (tests/golden/swin_*.npz) and the numpy oracle.  Bar (north_star): 1e-3 relative error."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import swin_oracle as sw
from oracle.golden_inputs import SWIN_CASES, swin_case

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__ as ge
    ge.build()
    return torch.device("cuda:0")


def make_block(c, dev):
    from dl_swin_gan_b200.swin.window_attention import SwinTransformerBlock3D
    blk = SwinTransformerBlock3D(dim=c["dim"], num_heads=c["heads"], window_size=c["window"], shift_size=c["shift"],
                                 qkv_bias=True).eval()
    sd = blk.state_dict()
    for k, v in c["params"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
        sd[k] = torch.from_numpy(v)
    blk.load_state_dict(sd)                      # reference parameter names load unchanged
    return blk.to(dev)


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_block_part1_and_attention_match_reference(dev, name):
    from dl_swin_gan_b200.swin import window_attention as wa
    c = swin_case(name)
    ref = np.load(os.path.join(GOLDEN, f"swin_{name}.npz"))
    blk = make_block(c, dev)
    x = torch.from_numpy(c["x"]).to(dev)
    B, D, H, W, C = x.shape
    ws, ss = wa.get_window_size((D, H, W), c["window"], c["shift"])
    Dp, Hp, Wp = (-(-n // w) * w for n, w in zip((D, H, W), ws))
    mask = wa.compute_mask(Dp, Hp, Wp, ws, ss, dev)
    assert np.array_equal(mask.cpu().numpy(), ref["mask"])
    with torch.no_grad():
        part1 = blk.forward_part1(x, mask)                                   # fused LN/partition, region-id mask
        assert rel_l2(part1.cpu().numpy(), ref["part1"]) < TOL
        part1_dense = blk.forward_part1(x, mask.clone())                     # foreign mask tensor -> dense path
        assert rel_l2(part1_dense.cpu().numpy(), ref["part1"]) < TOL
        xp = torch.nn.functional.pad(x, (0, 0, 0, Wp - W, 0, Hp - H, 0, Dp - D))
        xw = wa.window_partition(xp, ws)
        assert rel_l2(blk.attn(xw, mask=None).cpu().numpy(), ref["attn_nomask"]) < TOL
        assert rel_l2(blk.attn(xw, mask=mask.clone()).cpu().numpy(), ref["attn_mask"]) < TOL
    # differentiable route gives the same forward
    xg = x.clone().requires_grad_()
    assert rel_l2(blk.forward_part1(xg, mask).detach().cpu().numpy(), ref["part1"]) < TOL
    # whole block and its MLP half (tcgen05 fc1 / fc2 kernels when dim % 32 == 0, i.e. the cine configuration)
    ref_block = np.load(os.path.join(GOLDEN, f"swin_block_{name}.npz"))["block"]
    p64 = {k: v.astype(np.float64) for k, v in c["params"].items()}
    with torch.no_grad():
        assert rel_l2(blk(x, mask).cpu().numpy(), ref_block) < TOL
        assert rel_l2(blk.forward_part2(x).cpu().numpy(), sw.block_part2(c["x"].astype(np.float64), p64)) < TOL
    assert rel_l2(blk(xg, mask).detach().cpu().numpy(), ref_block) < TOL          # differentiable route


def test_fused_partition_and_reverse_are_exact(dev):
    from dl_swin_gan_b200.swin import window_attention as wa
    for (B, D, H, W, C, ws0, ss0) in [(2, 5, 7, 9, 24, (2, 4, 4), (1, 2, 2)), (1, 7, 48, 40, 160, (7, 8, 8), (3, 4, 4))]:
        x = torch.randn(B, D, H, W, C, device=dev)
        ws, ss = wa.get_window_size((D, H, W), ws0, ss0)
        Dp, Hp, Wp = (-(-n // w) * w for n, w in zip((D, H, W), ws))
        norm = torch.nn.LayerNorm(C).to(dev)
        with torch.no_grad():
            norm.weight.normal_(1, 0.1); norm.bias.normal_(0, 0.1)
            got = wa._ln_partition(x, norm, ws, ss)
            xn = torch.nn.functional.pad(norm(x), (0, 0, 0, Wp - W, 0, Hp - H, 0, Dp - D))
            xn = torch.roll(xn, shifts=tuple(-s for s in ss), dims=(1, 2, 3))
            want = wa.window_partition(xn, ws)
            assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < 1e-6
            yw = torch.randn_like(want)
            back = wa._window_reverse_fused(yw, B, D, H, W, ws, ss)
            r = wa.window_reverse(yw.view(-1, *ws, C), ws, B, Dp, Hp, Wp)
            r = torch.roll(r, shifts=tuple(ss), dims=(1, 2, 3))[:, :D, :H, :W]
            assert torch.equal(back, r)


def test_cine_size_attention_against_torch_reference_arithmetic(dev):
    """One cine slice worth of windows (30 x 448 tokens, 8 heads, shift mask) vs the reference arithmetic in fp32 torch."""
    from dl_swin_gan_b200.swin import window_attention as wa
    torch.manual_seed(0)
    mod = wa.WindowAttention3D(160, (7, 8, 8), 8, qkv_bias=True).to(dev).eval()
    with torch.no_grad():
        mod.relative_position_bias_table.normal_(0, 0.5)
    ws, ss = (7, 8, 8), (0, 4, 4)
    mask = wa.compute_mask(7, 48, 40, ws, ss, dev)
    rid = wa._region_ids_for(mask, 7, 48, 40, ws, ss, dev)
    x = torch.randn(30, 448, 160, device=dev)
    with torch.no_grad():
        qkv = mod.qkv(x)
        want = mod.proj(wa._attention_core_torch(qkv, mod.relative_position_bias_table, mod.relative_position_index, 8, mod.scale, mask))
        got_region = mod(x, mask=mask, _region_ids=rid)
        got_dense = mod(x, mask=mask)
        got_none = mod(x)
        want_none = mod.proj(wa._attention_core_torch(qkv, mod.relative_position_bias_table, mod.relative_position_index, 8, mod.scale, None))
    assert rel_l2(got_region.cpu().numpy(), want.cpu().numpy()) < TOL
    assert rel_l2(got_dense.cpu().numpy(), want.cpu().numpy()) < TOL
    assert rel_l2(got_none.cpu().numpy(), want_none.cpu().numpy()) < TOL
    assert float((got_region - want).abs().max() / want.abs().max()) < 5e-3


def test_attention_gradients(dev):
    from dl_swin_gan_b200.swin import window_attention as wa
    torch.manual_seed(1)
    mod = wa.WindowAttention3D(16, (2, 4, 4), 2, qkv_bias=True).to(dev)
    with torch.no_grad():
        mod.relative_position_bias_table.normal_(0, 0.5)
    mask = wa.compute_mask(4, 8, 8, (2, 4, 4), (1, 2, 2), dev)
    x = torch.randn(16, 32, 16, device=dev, requires_grad=True)
    out = mod(x, mask=mask)
    go = torch.randn_like(out)
    got = torch.autograd.grad(out, [x, mod.relative_position_bias_table, mod.qkv.weight], go)
    x2 = x.detach().clone().requires_grad_()
    ref = mod.proj(wa._attention_core_torch(mod.qkv(x2), mod.relative_position_bias_table, mod.relative_position_index, 2, mod.scale, mask))
    want = torch.autograd.grad(ref, [x2, mod.relative_position_bias_table, mod.qkv.weight], go)
    for a, b in zip(got, want):
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < TOL


def test_basic_layer_runs_and_matches_blockwise_oracle(dev):
    from dl_swin_gan_b200.swin import window_attention as wa
    torch.manual_seed(2)
    layer = wa.BasicLayer(dim=16, depth=2, num_heads=2, window_size=(2, 4, 4), qkv_bias=True).to(dev).eval()
    x = torch.randn(1, 16, 4, 8, 8, device=dev)
    with torch.no_grad():
        y = layer(x)
    assert y.shape == x.shape and torch.isfinite(y).all()
    # oracle: block 0 (no shift) part1 on channels-last input
    p = {k: (v.detach().cpu().numpy() if "index" in k else v.detach().cpu().numpy().astype(np.float64))
         for k, v in layer.blocks[1].state_dict().items()}
    xin = torch.randn(1, 4, 8, 8, 16, device=dev)
    with torch.no_grad():
        got = layer.blocks[1].forward_part1(xin, wa.compute_mask(4, 8, 8, (2, 4, 4), (1, 2, 2), dev))
    want = sw.block_part1(xin.cpu().numpy().astype(np.float64), p, 2, (2, 4, 4), (1, 2, 2))
    assert rel_l2(got.cpu().numpy(), want) < TOL


def test_tcgen05_linear_matches_torch(dev):
    from dl_swin_gan_b200.swin import window_attention as wa
    torch.manual_seed(3)
    for (M, N, K) in [(128, 32, 16), (1, 160, 160), (300, 480, 160), (13440, 480, 160), (13440, 160, 160), (257, 512, 64)]:
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev)
        want = torch.nn.functional.linear(x.double(), w.double(), b.double())
        got = wa.linear_tc(x, w, b)
        assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < 5e-4, (M, N, K)       # fp16 operand rounding ~2e-4
        got_nb = wa.linear_tc(x, w, None)
        assert rel_l2(got_nb.cpu().numpy(), (want - b.double()).cpu().numpy()) < 5e-4
    with pytest.raises(ValueError):
        wa.linear_tc(torch.randn(8, 24, device=dev), torch.randn(32, 24, device=dev), None)       # K % 16 != 0


def test_error_behaviour(dev):
    from dl_swin_gan_b200.swin import window_attention as wa
    mod = wa.WindowAttention3D(16, (2, 4, 4), 2).to(dev)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.randn(2, 32, 16))
    big = wa.WindowAttention3D(128, (2, 4, 4), 2).to(dev)          # head_dim 64 > 32
    with pytest.raises(ValueError):
        big(torch.randn(2, 32, 128, device=dev))


def test_mlp_kernels_at_cine_size_and_ragged_tiles(dev):
    """forward_part2 on the tensor-core kernels vs fp32 torch: one cine slice of tokens, a token count that is not a
    multiple of the 128-row tile, residual on/off, and the weight cache following a parameter update."""
    from dl_swin_gan_b200.swin import window_attention as wa
    torch.manual_seed(3)
    blk = wa.SwinTransformerBlock3D(160, 8, (7, 8, 8), (0, 0, 0), qkv_bias=True).to(dev).eval()
    with torch.no_grad():
        blk.norm2.weight.normal_(1, 0.1); blk.norm2.bias.normal_(0, 0.1)
        for shape in [(1, 7, 48, 40, 160), (1, 3, 5, 7, 160), (2, 1, 1, 1, 160)]:
            x = torch.randn(*shape, device=dev)
            want = blk.mlp(blk.norm2(x))
            got = blk._part2_tensor_core(x, residual=False)
            assert got is not None
            assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < TOL
            assert float((got - want).abs().max() / want.abs().max()) < 5e-3
            got_r = blk._part2_tensor_core(x, residual=True)
            assert rel_l2(got_r.cpu().numpy(), (x + want).cpu().numpy()) < TOL
        blk.mlp.fc1.weight.mul_(0.5)                       # in-place update bumps the version: the fp16 copies are refreshed
        x = torch.randn(1, 2, 8, 8, 160, device=dev)
        assert rel_l2(blk.forward_part2(x).cpu().numpy(), blk.mlp(blk.norm2(x)).cpu().numpy()) < TOL


def test_block_launches_are_bit_identical(dev):
    """The five launches of the block (tcgen05 GEMMs with cp.async rings, attention core) are deterministic: 10 repeats, same bits."""
    from dl_swin_gan_b200.swin import window_attention as wa
    torch.manual_seed(5)
    blk = wa.SwinTransformerBlock3D(160, 8, (7, 8, 8), (3, 4, 4), qkv_bias=True).to(dev).eval()
    x = torch.randn(2, 7, 48, 40, 160, device=dev)
    mask = wa.compute_mask(7, 48, 40, (7, 8, 8), (0, 4, 4), dev)
    with torch.no_grad():
        first = blk(x, mask)
        for _ in range(10):
            assert torch.equal(blk(x, mask), first)
