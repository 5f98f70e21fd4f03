#!/usr/bin/env python3
"""Run a few encoder forwards WITHOUT the CUDA graph so that ncu sees every kernel launch (S2 workload)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from auto_avsr_b200 import ConformerEncoder
from auto_avsr_b200.synthetic import SHAPES, encoder_input, encoder_state_dict

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
dev = torch.device("cuda:0")
lengths = list(SHAPES["S2"])
enc = ConformerEncoder()
enc.load_state_dict(encoder_state_dict(0))
enc = enc.to(dev).eval()
enc.precision = prec
enc.use_graph = False
enc.assume_frozen = True
xs = encoder_input(lengths).to(dev)
mask = (torch.arange(max(lengths))[None, :] < torch.tensor(lengths)[:, None]).unsqueeze(1).to(dev)
with torch.no_grad():
    for _ in range(steps):
        out, _ = enc(xs, mask)
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
