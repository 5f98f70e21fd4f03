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

from espnet.nets.pytorch_backend.encoder.conformer_encoder import ConformerEncoder, ConvolutionModule  # noqa: E402
from espnet.nets.pytorch_backend.nets_utils import make_non_pad_mask  # noqa: E402
from espnet.nets.pytorch_backend.transformer.attention import RelPositionMultiHeadedAttention  # noqa: E402
from espnet.nets.pytorch_backend.transformer.embedding import RelPositionalEncoding  # noqa: E402
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


def shrink(d, step=8):
    """Big weight-gradient matrices are stored as every `step`-th row (fp32) plus full-matrix checksums
    [sum, sum |.|, sum of squares] (float64): keeps the fixtures at a few MB."""
    out = {}
    for k, v in d.items():
        if v.size > 500_000 and k.startswith("grad_"):
            out[k + "__rows8"] = v[::step].astype(np.float32)
            out[k + "__step"] = np.array(step)
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
    # ---- rel-pos attention (attention.py:107-193) on a ragged batch, attention dropout 0.0 as E2E builds it
    LENGTHS = [37, 29, 18]
    x, r = tensors(34)
    mask = make_non_pad_mask(LENGTHS).unsqueeze(-2)
    pos_emb = RelPositionalEncoding(D, 0.0).double().eval()(x)[1]
    att = RelPositionMultiHeadedAttention(12, D, 0.0).double().train()
    att.load_state_dict(sub_state(sd, "encoders.0.self_attn."))
    xg = x.clone().requires_grad_(True)
    y = att(xg, xg, xg, pos_emb, mask)
    (y * r).sum().backward()
    res = dict(y=y.detach().numpy(), dx=xg.grad.numpy(), lengths=np.array(LENGTHS),
               **{"grad_" + n: p.grad.numpy() for n, p in att.named_parameters()})
    np.savez_compressed(os.path.join(out_dir, "train_attn.npz"), seed=np.array(34), wseed=np.array(21), **shrink(res))
    # ---- the whole 2-layer encoder in train mode, every dropout rate 0: output, input gradient, and per-parameter
    # gradient checksums (+ every 64th row of the big matrices)
    sd2 = encoder_state_dict(22, D, 12, F, 2, K)
    enc = ConformerEncoder(attention_dim=D, attention_heads=12, linear_units=F, num_blocks=2, dropout_rate=0.0,
                           positional_dropout_rate=0.0, attention_dropout_rate=0.0, cnn_module_kernel=K)
    enc.load_state_dict(sd2, strict=True)
    enc = enc.double().train()
    x, r = tensors(35)
    xg = x.clone().requires_grad_(True)
    y, _ = enc(xg, mask)
    (y * r).sum().backward()
    res = dict(y=y.detach().numpy(), dx=xg.grad.numpy(), lengths=np.array(LENGTHS))
    for n, p_ in enc.named_parameters():
        g = p_.grad.numpy()
        res["cs_" + n] = np.array([g.sum(), np.abs(g).sum(), (g ** 2).sum()])
        if g.size > 500_000:
            res["rows64_" + n] = g[::64].astype(np.float32)
        elif g.size <= 4096:
            res["grad_" + n] = g
    for n, b_ in enc.named_buffers():
        if "running" in n:
            res["buf_" + n] = b_.detach().numpy()
    np.savez_compressed(os.path.join(out_dir, "train_enc2.npz"), seed=np.array(35), wseed=np.array(22), **res)
    for n in ("train_ln", "train_ffn", "train_conv", "train_attn", "train_enc2"):
        print(n, os.path.getsize(os.path.join(out_dir, n + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
