"""Deterministic synthetic encoder weights and inputs (no network, no checkpoints).

Keys and shapes are the reference encoder's state-dict contract
(espnet/nets/pytorch_backend/encoder/conformer_encoder.py:212-262; 40 keys per
layer + ``after_norm.{weight,bias}``).  Values are *not* the
reference's default init: LayerNorm/BatchNorm affine, running statistics and all
biases are non-trivial so that a kernel that drops one of them cannot pass parity
(default init leaves them 0/1).

Every tensor is drawn from its own ``torch.Generator`` seeded by (seed, key), so the
same (seed, config) gives the same weights on any host with this torch build; the
golden fixtures under ``tests/golden`` store only the seed.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Sequence

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def encoder_state_dict(seed: int = 0, d_model: int = 768, n_heads: int = 12, linear_units: int = 3072,
                       num_blocks: int = 12, cnn_kernel: int = 31,
                       dtype: torch.dtype = torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Synthetic ``ConformerEncoder.state_dict()`` in the reference's key order."""
    dk = d_model // n_heads
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def uni(key, shape, bound):
        sd[key] = (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)

    def nrm(key, shape, std, mean=0.0):
        sd[key] = (torch.randn(shape, generator=_gen(seed, key), dtype=torch.float64) * std + mean).to(dtype)

    def pos(key, shape, lo, hi):
        sd[key] = (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float64) * (hi - lo) + lo).to(dtype)

    def linear(pfx, out_f, in_f, bias=True):
        uni(pfx + ".weight", (out_f, in_f), 1.0 / math.sqrt(in_f))
        if bias:
            nrm(pfx + ".bias", (out_f,), 0.1)

    def lnorm(pfx):
        pos(pfx + ".weight", (d_model,), 0.5, 1.5)
        nrm(pfx + ".bias", (d_model,), 0.1)

    for l in range(num_blocks):
        p = f"encoders.{l}."
        # key order == the reference module registration order (conformer_encoder.py:61-94)
        uni(p + "self_attn.pos_bias_u", (n_heads, dk), math.sqrt(6.0 / (n_heads + dk)))
        uni(p + "self_attn.pos_bias_v", (n_heads, dk), math.sqrt(6.0 / (n_heads + dk)))
        linear(p + "self_attn.linear_q", d_model, d_model)
        linear(p + "self_attn.linear_k", d_model, d_model)
        linear(p + "self_attn.linear_v", d_model, d_model)
        linear(p + "self_attn.linear_out", d_model, d_model)
        linear(p + "self_attn.linear_pos", d_model, d_model, bias=False)
        linear(p + "feed_forward.w_1", linear_units, d_model)
        linear(p + "feed_forward.w_2", d_model, linear_units)
        uni(p + "conv_module.pointwise_cov1.weight", (2 * d_model, d_model, 1), 1.0 / math.sqrt(d_model))
        nrm(p + "conv_module.pointwise_cov1.bias", (2 * d_model,), 0.1)
        uni(p + "conv_module.depthwise_conv.weight", (d_model, 1, cnn_kernel), 1.0 / math.sqrt(cnn_kernel))
        nrm(p + "conv_module.depthwise_conv.bias", (d_model,), 0.1)
        pos(p + "conv_module.norm.weight", (d_model,), 0.5, 1.5)
        nrm(p + "conv_module.norm.bias", (d_model,), 0.1)
        nrm(p + "conv_module.norm.running_mean", (d_model,), 0.1)
        pos(p + "conv_module.norm.running_var", (d_model,), 0.5, 1.5)
        sd[p + "conv_module.norm.num_batches_tracked"] = torch.tensor(100 + l, dtype=torch.long)
        uni(p + "conv_module.pointwise_cov2.weight", (d_model, d_model, 1), 1.0 / math.sqrt(d_model))
        nrm(p + "conv_module.pointwise_cov2.bias", (d_model,), 0.1)
        lnorm(p + "norm_ff")
        lnorm(p + "norm_mha")
        linear(p + "feed_forward_macaron.w_1", linear_units, d_model)
        linear(p + "feed_forward_macaron.w_2", d_model, linear_units)
        lnorm(p + "norm_ff_macaron")
        lnorm(p + "norm_conv")
        lnorm(p + "norm_final")
    lnorm("after_norm")
    return sd


def encoder_input(lengths: Sequence[int], d_model: int = 768, seed: int = 1234,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """(B, Tmax, d) O(1)-scale frames; padded positions keep random values."""
    B, T = len(lengths), max(int(v) for v in lengths)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(B, T, d_model, generator=g, dtype=torch.float64).to(dtype)


def head_state_dict(seed: int = 0, idim: int = 512, d_model: int = 768, odim: int = 5049,
                    dtype: torch.dtype = torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Synthetic weights of the two Linear layers either side of the encoder in the reference's ``E2E`` model, under
    the reference's state-dict keys: ``proj_encoder`` (e2e_asr_conformer.py:31) and ``ctc.ctc_lo`` (ctc.py:21)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, out_f, in_f in (("proj_encoder", d_model, idim), ("ctc.ctc_lo", odim, d_model)):
        w = (torch.rand((out_f, in_f), generator=_gen(seed, key + ".weight"), dtype=torch.float64) * 2 - 1)
        sd[key + ".weight"] = w.mul_(1.0 / math.sqrt(in_f)).to(dtype)
        sd[key + ".bias"] = (torch.randn((out_f,), generator=_gen(seed, key + ".bias"), dtype=torch.float64) * 0.1).to(dtype)
    return sd


def frontend_features(lengths: Sequence[int], idim: int = 512, seed: int = 4321,
                      dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """(B, Tmax, idim) stand-in for the ResNet front-end's output (the input of ``proj_encoder``)."""
    return encoder_input(lengths, idim, seed, dtype)


def decoder_state_dict(seed: int = 0, odim: int = 5049, d_model: int = 768, n_heads: int = 12, linear_units: int = 3072,
                       num_blocks: int = 6, dtype: torch.dtype = torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Synthetic ``TransformerDecoder.state_dict()`` in the reference's key order
    (espnet/nets/pytorch_backend/decoder/transformer_decoder.py:159-229; 26 keys per ``DecoderLayer`` + ``embed.0``,
    ``after_norm``, ``output_layer``).  As for the encoder every LayerNorm affine and bias is non-trivial."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def uni(key, shape, bound):
        sd[key] = (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)

    def nrm(key, shape, std):
        sd[key] = (torch.randn(shape, generator=_gen(seed, key), dtype=torch.float64) * std).to(dtype)

    def linear(pfx, out_f, in_f):
        uni(pfx + ".weight", (out_f, in_f), 1.0 / math.sqrt(in_f))
        nrm(pfx + ".bias", (out_f,), 0.1)

    def lnorm(pfx):
        sd[pfx + ".weight"] = (torch.rand((d_model,), generator=_gen(seed, pfx + ".weight"), dtype=torch.float64) + 0.5).to(dtype)
        nrm(pfx + ".bias", (d_model,), 0.1)

    nrm("embed.0.weight", (odim, d_model), 1.0 / math.sqrt(d_model))      # x sqrt(d) in PositionalEncoding -> O(1) rows
    for l in range(num_blocks):
        p = f"decoders.{l}."
        for att in ("self_attn", "src_attn"):
            for lin in ("linear_q", "linear_k", "linear_v", "linear_out"):
                linear(p + att + "." + lin, d_model, d_model)
        linear(p + "feed_forward.w_1", linear_units, d_model)
        linear(p + "feed_forward.w_2", d_model, linear_units)
        lnorm(p + "norm1")
        lnorm(p + "norm2")
        lnorm(p + "norm3")
    lnorm("after_norm")
    linear("output_layer", odim, d_model)
    return sd


#: Canonical length sets of SURVEY.md §8 (25 Hz frames).
SHAPES: Dict[str, Sequence[int]] = {
    "S1": [100],
    "S2": [400] * 4,
    "S2r": [400, 350, 300, 250, 300],
    "S3": [100] * 16,
    "S4": [1600],
}
