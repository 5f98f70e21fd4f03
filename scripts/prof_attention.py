#!/usr/bin/env python3
"""ncu target: the fused rel-pos attention kernel at the S2 shape (B=4, T=400, H=12), a few launches.
    ncu --set full --clock-control none --import-source on -k regex:attention_f16 -s 2 -c 1 -o gpurun_out/attn python scripts/prof_attention.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from auto_avsr_b200 import ops  # noqa: E402

B, T, H = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4, 400, 12)))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
D = H * 64
q, k, v = ((torch.randn(B, T, D, generator=g) * 0.5).to(dev) for _ in range(3))      # moderate logits: few reference moves
p = torch.randn(2 * T - 1, D, generator=g).to(dev)
u, vb = (torch.randn(H, 64, generator=g) * 0.3).to(dev), (torch.randn(H, 64, generator=g) * 0.3).to(dev)
for _ in range(4):
    ops.relpos_attention(q, k, v, p, u, vb, None, H, precision="f16")
torch.cuda.synchronize()
