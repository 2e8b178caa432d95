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
    nwin = B * mask.shape[0] if shift else B * 30
    flops_core = 4.0 * 448 * 448 * 160 * nwin
    print(json.dumps({"shift": shift, "slices": B, "windows": nwin, "part1_fused_ms": round(t_fused, 3), "core_ms": round(t_core, 3),
                      "part1_eager_ms": round(t_eager, 3), "core_useful_TFLOPs": round(flops_core / t_core / 1e9, 2),
                      "core_us_per_slice": round(1e3 * t_core / B, 1), "rel_err_vs_eager": err}), flush=True)
