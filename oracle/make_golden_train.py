#!/usr/bin/env python3
"""Generate tests/golden/train_*.npz: forward outputs AND gradients of the UNMODIFIED reference modules in train mode
(float64, dropout p = 0 so that the result is deterministic) for the training slice of SURVEY.md 8f #2.

TEST INFRASTRUCTURE ONLY.  Run in the build container:   PYTHONPATH=/root/reference python oracle/make_golden_train.py

Cases (weights / inputs regenerate from seeds via auto_avsr_b200.synthetic; only outputs and gradients are stored):
  ln    LayerNorm(768)                                  layer_norm.py:12-33
  ffn   PositionwiseFeedForward(768, 3072, 0.0)         positionwise_feed_forward.py:12-30
  conv  ConvolutionModule(768, 31) incl. BatchNorm1d in train mode (batch statistics over all B*T frames, running
        statistics updated)                             conformer_encoder.py:19-35
The loss is sum(y * r) with a seeded random r, so that dL/dy = r exercises every output element."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from espnet.nets.pytorch_backend.encoder.conformer_encoder import ConvolutionModule  # noqa: E402
from espnet.nets.pytorch_backend.transformer.layer_norm import LayerNorm  # noqa: E402
from espnet.nets.pytorch_backend.transformer.positionwise_feed_forward import PositionwiseFeedForward  # noqa: E402

from auto_avsr_b200.synthetic import encoder_state_dict  # noqa: E402

B, T, D, F, K = 3, 37, 768, 3072, 31


def tensors(seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    r = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    return x, r


def sub_state(sd, prefix):
    return {k[len(prefix):]: v.double() for k, v in sd.items() if k.startswith(prefix)}


def shrink(d):
    """Big weight-gradient matrices are stored as every 8th row (fp32) plus full-matrix checksums
    [sum, sum |.|, sum of squares] (float64): keeps the fixtures at ~4 MB instead of 30."""
    out = {}
    for k, v in d.items():
        if v.size > 500_000 and k.startswith("grad_"):
            out[k + "__rows8"] = v[::8].astype(np.float32)
            out[k + "__checksum"] = np.array([v.sum(), np.abs(v).sum(), (v.astype(np.float64) ** 2).sum()])
        else:
            out[k] = v
    return out


def run(module, x, r):
    module = module.double().train()
    x = x.clone().requires_grad_(True)
    y = module(x)
    (y * r).sum().backward()
    grads = {"grad_" + n: p.grad.numpy() for n, p in module.named_parameters()}
    bufs = {"buf_" + n: b.detach().numpy() for n, b in module.named_buffers()}
    return dict(y=y.detach().numpy(), dx=x.grad.numpy(), **grads, **bufs)


def main():
    sd = encoder_state_dict(21, D, 12, F, 1, K)
    out_dir = os.path.join(ROOT, "tests", "golden")
    x, r = tensors(31)
    ln = LayerNorm(D)
    ln.load_state_dict(sub_state(sd, "encoders.0.norm_ff."))
    np.savez_compressed(os.path.join(out_dir, "train_ln.npz"), seed=np.array(31), wseed=np.array(21), **run(ln, x, r))
    x, r = tensors(32)
    ffn = PositionwiseFeedForward(D, F, 0.0)
    ffn.load_state_dict(sub_state(sd, "encoders.0.feed_forward."))
    np.savez_compressed(os.path.join(out_dir, "train_ffn.npz"), seed=np.array(32), wseed=np.array(21), **shrink(run(ffn, x, r)))
    x, r = tensors(33)
    conv = ConvolutionModule(D, K)
    conv.load_state_dict(sub_state(sd, "encoders.0.conv_module."))
    np.savez_compressed(os.path.join(out_dir, "train_conv.npz"), seed=np.array(33), wseed=np.array(21), **shrink(run(conv, x, r)))
    for n in ("train_ln", "train_ffn", "train_conv"):
        print(n, os.path.getsize(os.path.join(out_dir, n + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
