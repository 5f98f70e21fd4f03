#!/usr/bin/env python3
"""Where does the encoder training step spend its time?  torch.profiler over two S2 steps; prints the
CUDA-time table of the top kernels.   python scripts/prof_train_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from auto_avsr_b200 import ConformerEncoder  # noqa: E402
from auto_avsr_b200.synthetic import SHAPES, encoder_input, encoder_state_dict  # noqa: E402

dev = torch.device("cuda:0")
lengths = list(SHAPES["S2"])
enc = ConformerEncoder()
enc.load_state_dict(encoder_state_dict(0))
enc = enc.to(dev).train()
xs = encoder_input(lengths).to(dev)
mask = (torch.arange(max(lengths))[None, :] < torch.tensor(lengths)[:, None]).unsqueeze(1).to(dev)
for _ in range(2):
    enc.zero_grad(set_to_none=True)
    enc(xs, mask)[0].pow(2).mean().backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        enc.zero_grad(set_to_none=True)
        enc(xs, mask)[0].pow(2).mean().backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
