"""Timing of the shifted-window attention path at the cine size: fused kernels vs the reference arithmetic in torch eager."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import __graft_entry__ as ge
ge.build()
from dl_swin_gan_b200.swin import window_attention as wa
import bench

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("SWIN_B", 8))
D, H, W, C, heads = 7, 48, 40, 160, 8
ws0, ss0 = (7, 8, 8), (3, 4, 4)
for shift in (False, True):
    blk = wa.SwinTransformerBlock3D(C, heads, ws0, ss0 if shift else (0, 0, 0), qkv_bias=True).to(dev).eval()
    with torch.no_grad():
        blk.attn.relative_position_bias_table.normal_(0, 0.5)
    x = torch.randn(B, D, H, W, C, device=dev)
    ws, ss = wa.get_window_size((D, H, W), ws0, ss0 if shift else (0, 0, 0))
    mask = wa.compute_mask(7, 48, 40, ws, ss, dev)
    rid = wa._region_ids_for(mask, 7, 48, 40, ws, ss, dev) if shift else None

    def eager():
        xn = blk.norm1(x)
        if shift:
            xn = torch.roll(xn, shifts=tuple(-s for s in ss), dims=(1, 2, 3))
        xw = wa.window_partition(xn, ws)
        qkv = blk.attn.qkv(xw)
        o = wa._attention_core_torch(qkv, blk.attn.relative_position_bias_table, blk.attn.relative_position_index,
                                     heads, blk.attn.scale, mask if shift else None)
        o = blk.attn.proj(o)
        r = wa.window_reverse(o.view(-1, *ws, C), ws, B, D, H, W)
        if shift:
            r = torch.roll(r, shifts=tuple(ss), dims=(1, 2, 3))
        return r

    with torch.no_grad():
        xw = wa._ln_partition(x, blk.norm1, ws, ss)
        qkv = blk.attn.qkv(xw).contiguous()
        tab = blk.attn.relative_position_bias_table.detach().contiguous()
        core = lambda: wa._attention_core_raw(qkv, tab, heads, ws0, blk.attn.scale, 1 if shift else 0, rid, mask.shape[0])
        t_fused = bench.time_op(lambda: blk.forward_part1(x, mask), 10, dev)
        t_core = bench.time_op(core, 10, dev)
        t_eager = bench.time_op(eager, 5, dev)
        err = float((blk.forward_part1(x, mask) - eager()).norm() / eager().norm())
    # the three fused launches one by one
    import ctypes
    from dl_swin_gan_b200 import _lib
    lib = _lib.load()
    heads_, hd = heads, C // heads
    N = 448
    B_ = B * 30
    q16 = torch.empty((B_, heads, N, 32), dtype=torch.float16, device=dev); k16 = torch.empty_like(q16)
    vt16 = torch.empty_like(q16)
    core_o = torch.empty((B_, N, C), device=dev); yv = torch.empty_like(x)
    i3 = wa._i3
    att = blk.attn
    def f_qkv():
        lib.dlsg_swin_qkv_fused_f32(x.data_ptr(), blk.norm1.weight.data_ptr(), blk.norm1.bias.data_ptr(), 1e-5, att.qkv.weight.data_ptr(),
                                    att.qkv.bias.data_ptr(), q16.data_ptr(), k16.data_ptr(), vt16.data_ptr(), B, D, H, W, C, i3(ws), i3(ss),
                                    heads, float(att.scale), wa._stream())
    def f_core():
        lib.dlsg_swin_window_attention_f16(q16.data_ptr(), k16.data_ptr(), vt16.data_ptr(), tab.data_ptr(), core_o.data_ptr(), B_, N, heads, hd,
                                           i3(ws0), 1 if shift else 0, rid.data_ptr() if shift else None, 30, wa._stream())
    def f_proj():
        lib.dlsg_swin_proj_scatter_f32(core_o.data_ptr(), att.proj.weight.data_ptr(), att.proj.bias.data_ptr(), yv.data_ptr(), B, D, H, W, C,
                                       i3(ws), i3(ss), wa._stream())
    f_qkv(); f_core(); f_proj()
    parts = {"qkv_fused_ms": round(bench.time_op(f_qkv, 10, dev), 4), "core16_ms": round(bench.time_op(f_core, 10, dev), 4),
             "proj_scatter_ms": round(bench.time_op(f_proj, 10, dev), 4)}
    print(json.dumps(parts), flush=True)
    nwin = B * mask.shape[0] if shift else B * 30
    flops_core = 4.0 * 448 * 448 * 160 * nwin
    print(json.dumps({"shift": shift, "slices": B, "windows": nwin, "part1_fused_ms": round(t_fused, 3), "core_ms": round(t_core, 3),
                      "part1_eager_ms": round(t_eager, 3), "core_useful_TFLOPs": round(flops_core / t_core / 1e9, 2),
                      "core_us_per_slice": round(1e3 * t_core / B, 1), "rel_err_vs_eager": err}), flush=True)

# ---- second half of the block (forward_part2) and the whole block: tcgen05 MLP kernels vs torch eager (fp32 and TF32 cuBLAS)
blk = wa.SwinTransformerBlock3D(C, heads, ws0, ss0, qkv_bias=True).to(dev).eval()
x = torch.randn(B, D, H, W, C, device=dev)
ws, ss = wa.get_window_size((D, H, W), ws0, ss0)
mask = wa.compute_mask(7, 48, 40, ws, ss, dev)
with torch.no_grad():
    eager2 = lambda: x + blk.mlp(blk.norm2(x))
    t_k = bench.time_op(lambda: blk._part2_tensor_core(x, residual=True), 10, dev)
    t_e = bench.time_op(eager2, 10, dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    t_e_tf32 = bench.time_op(eager2, 10, dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    err2 = float((blk._part2_tensor_core(x, residual=True) - eager2()).norm() / eager2().norm())
    t_blk = bench.time_op(lambda: blk(x, mask), 10, dev)
tokens = x.numel() // C
print(json.dumps({"what": "x + forward_part2(x)", "slices": B, "tokens": tokens, "kernels_ms": round(t_k, 4), "eager_fp32_ms": round(t_e, 4),
                  "eager_tf32_ms": round(t_e_tf32, 4), "useful_TFLOPs": round(4.0 * tokens * C * 4 * C / t_k / 1e9, 1),
                  "rel_err_vs_eager": err2, "whole_block_ms": round(t_blk, 4)}), flush=True)
