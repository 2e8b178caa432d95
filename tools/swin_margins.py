import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge; ge.build()
from dl_swin_gan_b200.swin import window_attention as wa
from oracle.golden_inputs import SWIN_CASES, swin_case
dev = torch.device("cuda:0")
for name in SWIN_CASES:
    c = swin_case(name)
    ref = np.load(os.path.join("tests", "golden", f"swin_{name}.npz"))
    blk = wa.SwinTransformerBlock3D(dim=c["dim"], num_heads=c["heads"], window_size=c["window"], shift_size=c["shift"], qkv_bias=True).eval()
    sd = blk.state_dict(); sd.update({k: torch.from_numpy(v) for k, v in c["params"].items()}); blk.load_state_dict(sd); blk = blk.to(dev)
    x = torch.from_numpy(c["x"]).to(dev)
    B, D, H, W, C = x.shape
    ws, ss = wa.get_window_size((D, H, W), c["window"], c["shift"])
    Dp, Hp, Wp = (-(-n // w) * w for n, w in zip((D, H, W), ws))
    mask = wa.compute_mask(Dp, Hp, Wp, ws, ss, dev)
    with torch.no_grad():
        got = blk.forward_part1(x, mask).cpu().numpy()
        wa.WindowAttention3D.tc_linear = False
        got_fp32lin = blk.forward_part1(x, mask).cpu().numpy()
        wa.WindowAttention3D.tc_linear = True
    r = ref["part1"]
    print(name, "fused tc path: %.2e" % (np.linalg.norm(got - r) / np.linalg.norm(r)), " fp32 linears + fp16 core: %.2e" % (np.linalg.norm(got_fp32lin - r) / np.linalg.norm(r)))
