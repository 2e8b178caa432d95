"""Where a training step of the Swin unrolled network goes (config_swin.yaml, batch 1, readout 64): top CUDA ops by time."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from dl_swin_gan_b200.config import load_cfg
from dl_swin_gan_b200.train import Trainer, synthetic_batch
dev = torch.device("cuda:0")
cfg = load_cfg("config_swin")
torch.manual_seed(0)
tr = Trainer(cfg, dev, conv_precision="bf16")
batch = synthetic_batch(cfg, dev, 1, seed=1)
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step(batch)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
