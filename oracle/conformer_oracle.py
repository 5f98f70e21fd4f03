"""CPU restatement of the auto_avsr Conformer encoder forward (eval mode).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Nothing in the shipped
package imports this file.

This is a *functional* restatement written from the formulas of SURVEY.md §8a, not
a copy of the reference modules: it works on a flat ``{key: tensor}`` state dict,
gathers ``P[rel = i - j]`` explicitly instead of the reference's pad/view
``rel_shift`` trick, masks keys with ``-inf`` instead of ``finfo.min`` + zero
fill, and states BatchNorm as ``(h - mean) / sqrt(var + eps) * gamma + beta``.
Each function cites the reference lines (``/root/reference`` @ 182b628) whose
arithmetic it restates.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the reference itself,
generated in the build container by ``oracle/make_golden.py`` (which imports
``/root/reference`` read-only) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks fp64 and fp32 agreement.

dtype-generic: run it in float64 for the tight oracle, float32 for the
"reference-precision" oracle and the CPU baseline timing.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

Tensor = torch.Tensor

LN_EPS = 1e-12  # transformer/layer_norm.py:21
BN_EPS = 1e-5   # torch.nn.BatchNorm1d default, conformer_encoder.py:26


def rel_sinusoid_table(T: int, d_model: int, dtype=torch.float32) -> Tensor:
    """``pos_emb`` of shape (2T-1, d_model); row k holds the sinusoid of rel = T-1-k.

    Restates transformer/embedding.py:139-169 + the slice at :179-183.  The
    reference evaluates sin/cos in **float32** and casts afterwards
    (embedding.py:151-169), so an fp64 oracle must do the same to agree to 1e-12.
    """
    rel = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)       # (2T-1, 1)
    inv = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32)
                    * -(math.log(10000.0) / d_model))                          # (d/2,)
    # The reference computes sin(position*div) for position >= 0 and
    # sin(-1*position*div) for the negative half; sin/cos of the fp32 product is
    # what must be reproduced.  (|rel|*inv) is computed in fp32 exactly the same
    # way for both halves, then the sign is applied to the argument.
    arg = rel.abs() * inv
    arg = torch.where(rel < 0, -arg, arg)
    pe = torch.empty(2 * T - 1, d_model, dtype=torch.float32)
    pe[:, 0::2] = torch.sin(arg)
    pe[:, 1::2] = torch.cos(arg)
    return pe.to(dtype)


def layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """LayerNorm over the last dim, eps 1e-12 (transformer/layer_norm.py:12-33):
    (x - mean) / sqrt(biased_var + eps) * w + b."""
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def feed_forward(x: Tensor, sd: Dict[str, Tensor], pfx: str) -> Tensor:
    """w_2(relu(w_1 x)) -- ReLU, not Swish (positionwise_feed_forward.py:28-30)."""
    h = torch.relu(x @ sd[pfx + "w_1.weight"].T + sd[pfx + "w_1.bias"])
    return h @ sd[pfx + "w_2.weight"].T + sd[pfx + "w_2.bias"]


def rel_attention_scores(q: Tensor, k: Tensor, p: Tensor, u: Tensor, v: Tensor) -> Tensor:
    """scores[b,h,i,j] = ((q_i+u_h).k_j + (q_i+v_h).p[rel=i-j]) / sqrt(d_k).

    q,k: (B,H,T,dk); p: (H,2T-1,dk) with row m <-> rel = T-1-m; u,v: (H,dk).
    Restates attention.py:174-189 with rel_shift (attention.py:131-151) replaced by
    its closed form out[i,j] = in[i, j-i+T-1].
    """
    B, H, T, dk = q.shape
    ac = torch.einsum("bhid,bhjd->bhij", q + u[None, :, None, :], k)
    bd_raw = torch.einsum("bhid,hmd->bhim", q + v[None, :, None, :], p)       # (B,H,T,2T-1)
    i = torch.arange(T, device=q.device).unsqueeze(1)
    j = torch.arange(T, device=q.device).unsqueeze(0)
    m = (j - i + T - 1).expand(B, H, T, T)
    bd = torch.gather(bd_raw, 3, m)
    return (ac + bd) / math.sqrt(dk)


def rel_mha(x: Tensor, pos_emb: Tensor, lengths: Optional[Tensor], sd: Dict[str, Tensor],
            pfx: str, n_heads: int, return_attn: bool = False):
    """RelPositionMultiHeadedAttention.forward (attention.py:153-193, :38-57, :59-88).

    Only *keys* are masked (j >= lengths[b]); padded query rows still produce
    output.  A row whose keys are all masked yields zeros, as the
    reference's ``masked_fill(mask, 0.0)`` after the softmax does (attention.py:75-77).
    """
    B, T, D = x.shape
    dk = D // n_heads

    def heads(y):
        return y.view(B, T, n_heads, dk).transpose(1, 2)

    q = heads(x @ sd[pfx + "linear_q.weight"].T + sd[pfx + "linear_q.bias"])
    k = heads(x @ sd[pfx + "linear_k.weight"].T + sd[pfx + "linear_k.bias"])
    v = heads(x @ sd[pfx + "linear_v.weight"].T + sd[pfx + "linear_v.bias"])
    p = (pos_emb @ sd[pfx + "linear_pos.weight"].T).view(2 * T - 1, n_heads, dk).transpose(0, 1)
    scores = rel_attention_scores(q, k, p, sd[pfx + "pos_bias_u"], sd[pfx + "pos_bias_v"])
    if lengths is not None:
        key_pad = torch.arange(T, device=x.device)[None, :] >= lengths.view(B, 1)   # (B,T) True = padded
        scores = scores.masked_fill(key_pad[:, None, None, :], float("-inf"))
        smax = scores.amax(dim=-1, keepdim=True)
        smax = torch.where(torch.isinf(smax), torch.zeros_like(smax), smax)
        e = torch.exp(scores - smax)
        den = e.sum(dim=-1, keepdim=True)
        attn = torch.where(den > 0, e / den.clamp_min(1e-300 if e.dtype == torch.float64 else 1e-38),
                           torch.zeros_like(e))
    else:
        attn = torch.softmax(scores, dim=-1)
    ctx = (attn @ v).transpose(1, 2).reshape(B, T, D)
    out = ctx @ sd[pfx + "linear_out.weight"].T + sd[pfx + "linear_out.bias"]
    return (out, attn) if return_attn else out


def conv_module(x: Tensor, sd: Dict[str, Tensor], pfx: str) -> Tensor:
    """ConvolutionModule.forward, eval mode (conformer_encoder.py:30-35).

    pointwise(768->1536) -> GLU -> depthwise k (zero 'same' padding inside each
    utterance row, NO length mask: padded frames are data) ->
    BatchNorm1d with running stats -> SiLU -> pointwise(768->768).
    """
    B, T, D = x.shape
    w1 = sd[pfx + "pointwise_cov1.weight"].squeeze(-1)                          # (2D, D)
    y = x @ w1.T + sd[pfx + "pointwise_cov1.bias"]
    g = y[..., :D] * torch.sigmoid(y[..., D:])                                  # GLU over channels
    wd = sd[pfx + "depthwise_conv.weight"]                                      # (D, 1, K)
    half = (wd.shape[-1] - 1) // 2
    # depthwise cross-correlation along time, zero padding `half` frames each side of every utterance row
    acc = torch.nn.functional.conv1d(g.transpose(1, 2), wd, sd[pfx + "depthwise_conv.bias"], padding=half,
                                     groups=D).transpose(1, 2)
    scale = sd[pfx + "norm.weight"] * torch.rsqrt(sd[pfx + "norm.running_var"] + BN_EPS)
    h = (acc - sd[pfx + "norm.running_mean"]) * scale + sd[pfx + "norm.bias"]
    h = h * torch.sigmoid(h)                                                    # SiLU
    w2 = sd[pfx + "pointwise_cov2.weight"].squeeze(-1)
    return h @ w2.T + sd[pfx + "pointwise_cov2.bias"]


def encoder_layer(x: Tensor, pos_emb: Tensor, lengths: Optional[Tensor], sd: Dict[str, Tensor],
                  pfx: str, n_heads: int, stages: Optional[List[Tensor]] = None) -> Tensor:
    """EncoderLayer.forward, normalize_before + macaron + conv, eval (conformer_encoder.py:96-170)."""
    def ln(y, name):
        return layer_norm(y, sd[pfx + name + ".weight"], sd[pfx + name + ".bias"])

    x = x + 0.5 * feed_forward(ln(x, "norm_ff_macaron"), sd, pfx + "feed_forward_macaron.")   # :110-116
    if stages is not None:
        stages.append(x)
    x = x + rel_mha(ln(x, "norm_mha"), pos_emb, lengths, sd, pfx + "self_attn.", n_heads)     # :119-142
    if stages is not None:
        stages.append(x)
    x = x + conv_module(ln(x, "norm_conv"), sd, pfx + "conv_module.")                         # :145-151
    if stages is not None:
        stages.append(x)
    x = x + 0.5 * feed_forward(ln(x, "norm_ff"), sd, pfx + "feed_forward.")                   # :154-159
    if stages is not None:
        stages.append(x)
    x = ln(x, "norm_final")                                                                    # :161-162
    if stages is not None:
        stages.append(x)
    return x


def count_layers(sd: Dict[str, Tensor]) -> int:
    n = 0
    while f"encoders.{n}.norm_ff.weight" in sd:
        n += 1
    return n


def encoder_forward(sd: Dict[str, Tensor], xs: Tensor, lengths: Optional[Sequence[int]],
                    n_heads: int, stages: Optional[List[Tensor]] = None) -> Tensor:
    """ConformerEncoder.forward (conformer_encoder.py:264-282), eval mode.

    ``lengths`` is the prefix-contiguous form of the reference's ``masks`` (B,1,T)
    (``make_non_pad_mask``, nets_utils.py:183); ``None`` == mask None == all valid.
    ``stages`` (optional list) receives the 5 residual-stage outputs of layer 0.
    """
    dtype = xs.dtype
    sd = {k: (t.to(dtype) if t.is_floating_point() else t) for k, t in sd.items()}
    B, T, D = xs.shape
    x = xs * math.sqrt(D)                                                       # embedding.py:178
    pos_emb = rel_sinusoid_table(T, D, dtype).to(xs.device)                     # embedding.py:179-183
    len_t = None if lengths is None else torch.as_tensor(list(lengths), dtype=torch.long, device=xs.device)
    for l in range(count_layers(sd)):
        x = encoder_layer(x, pos_emb, len_t, sd, f"encoders.{l}.", n_heads,
                          stages if l == 0 else None)
    return layer_norm(x, sd["after_norm.weight"], sd["after_norm.bias"])        # :279-280


def non_pad_mask(lengths: Sequence[int], maxlen: Optional[int] = None) -> Tensor:
    """make_non_pad_mask(lengths) (nets_utils.py:183-269): True = valid frame."""
    lengths = [int(v) for v in lengths]
    T = max(lengths) if maxlen is None else maxlen
    return torch.arange(T)[None, :] < torch.tensor(lengths)[:, None]
