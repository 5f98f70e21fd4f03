"""Per-op entry points of the C ABI as torch-tensor functions (device memory via torch, math in the library).
Used by the per-module forwards of auto_avsr_b200.espnet_dropin and by the unit parity tests."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._cabi import check, lib
from .engine import PRECISIONS, _ptr, _stream_handle, require_cuda


def _prep(t: torch.Tensor, what: str) -> torch.Tensor:
    require_cuda(t, what)
    return t.detach().contiguous()


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    x, weight, bias = _prep(x, "x"), _prep(weight, "weight"), _prep(bias, "bias")
    y = torch.empty_like(x)
    d = x.size(-1)
    with torch.cuda.device(x.device):
        check(lib.avsr_layernorm(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(),
                                 x.numel() // max(d, 1), d, _stream_handle(x.device)))
    return y


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
           residual: Optional[torch.Tensor] = None, alpha: float = 1.0, precision: str = "f16") -> torch.Tensor:
    """y = [residual + alpha *] act(x W^T + b);  weight (n, k) or (n, k, 1)."""
    x, weight = _prep(x, "x"), _prep(weight, "weight")
    n, k = weight.size(0), weight.size(1)
    if x.size(-1) != k:
        raise ValueError(f"linear: x last dim {x.size(-1)} != weight in-features {k}")
    rows = x.numel() // max(k, 1)
    bias = None if bias is None else _prep(bias, "bias")
    residual = None if residual is None else _prep(residual, "residual")
    y = torch.empty(*x.shape[:-1], n, dtype=torch.float32, device=x.device)
    nbytes = int(lib.avsr_linear_workspace_bytes(rows, n, k, PRECISIONS[precision]))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.avsr_linear(x.data_ptr(), weight.data_ptr(), _ptr(bias), _ptr(residual), float(alpha), int(relu),
                              y.data_ptr(), rows, n, k, PRECISIONS[precision], ws.data_ptr(), ws.numel(),
                              _stream_handle(x.device)))
    return y


def relpos_attention(q, k, v, p, pos_bias_u, pos_bias_v, lengths: Optional[torch.Tensor], n_heads: int,
                     precision: str = "f16") -> torch.Tensor:
    """q,k,v (B,T,H*64) projected; p (2T-1,H*64) = linear_pos(pos_emb); lengths int32 (B) or None -> ctx (B,T,H*64)."""
    q, k, v, p = _prep(q, "q"), _prep(k, "k"), _prep(v, "v"), _prep(p, "p")
    u, vb = _prep(pos_bias_u, "pos_bias_u"), _prep(pos_bias_v, "pos_bias_v")
    B, T, D = q.shape
    if D != n_heads * 64:
        raise ValueError("relpos_attention: d_k must be 64")
    if p.numel() != (2 * T - 1) * D:
        raise ValueError(f"relpos_attention: pos table must have 2T-1={2 * T - 1} rows")
    if lengths is not None:
        lengths = lengths.to(device=q.device, dtype=torch.int32).contiguous()
    ctx = torch.empty_like(q)
    nbytes = int(lib.avsr_attention_workspace_bytes(B, T, n_heads))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        check(lib.avsr_relpos_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), p.data_ptr(), u.data_ptr(),
                                        vb.data_ptr(), _ptr(lengths), ctx.data_ptr(), B, T, n_heads, ws.data_ptr(),
                                        nbytes, PRECISIONS[precision], _stream_handle(q.device)))
    return ctx


def dwconv_bn_silu(x, weight, bias, bn_weight, bn_bias, bn_mean, bn_var) -> torch.Tensor:
    """x (B,T,C); weight (C,1,K) -> silu(bn_eval(depthwise_conv(x))) as (B,T,C)."""
    x = _prep(x, "x")
    B, T, Cc = x.shape
    K = weight.size(-1)
    args = [_prep(t, "dwconv param") for t in (weight, bias, bn_weight, bn_bias, bn_mean, bn_var)]
    y = torch.empty_like(x)
    ws = torch.empty((K + 2) * Cc, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.avsr_dwconv_bn_silu(x.data_ptr(), *[a.data_ptr() for a in args], y.data_ptr(), B, T, Cc, K,
                                      ws.data_ptr(), ws.numel() * 4, _stream_handle(x.device)))
    return y


def pointwise_glu(x, weight, bias, precision: str = "f16") -> torch.Tensor:
    """glu(x W^T + b, dim=-1) with W (2C, C[,1]) -> (…, C)   (conformer_encoder.py:32)."""
    x, weight, bias = _prep(x, "x"), _prep(weight, "weight"), _prep(bias, "bias")
    Cc = weight.size(1)
    rows = x.numel() // max(Cc, 1)
    y = torch.empty_like(x)
    ws = torch.empty(int(lib.avsr_pointwise_glu_workspace_bytes(rows, Cc)), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.avsr_pointwise_glu(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), rows, Cc,
                                     ws.data_ptr(), ws.numel(), PRECISIONS[precision], _stream_handle(x.device)))
    return y


def rel_sinusoid_table(T: int, d: int, device) -> torch.Tensor:
    pe = torch.empty(2 * T - 1, d, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib.avsr_rel_sinusoid_table(pe.data_ptr(), T, d, _stream_handle(torch.device(device))))
    return pe


def log_softmax(x: torch.Tensor, n: Optional[int] = None, want_argmax: bool = False):
    """log_softmax over the first ``n`` entries of the last dim of ``x`` (default: all of it), fp32.  ``x`` may be a
    padded GEMM output: its row stride is its last dimension.  Returns ``(…, n)`` log-probs, plus the int64 arg max
    per row when ``want_argmax`` (CTC.log_softmax / CTC.argmax, ctc.py:77-93)."""
    x = _prep(x, "x")
    if x.dtype != torch.float32:
        raise TypeError("log_softmax: fp32 logits expected")
    ld = x.size(-1)
    n = ld if n is None else int(n)
    if not 0 < n <= ld:
        raise ValueError(f"log_softmax: n={n} outside (0, {ld}]")
    rows = x.numel() // ld
    y = torch.empty(*x.shape[:-1], n, dtype=torch.float32, device=x.device)
    best = torch.empty(x.shape[:-1], dtype=torch.int32, device=x.device) if want_argmax else None
    with torch.cuda.device(x.device):
        check(lib.avsr_log_softmax(x.data_ptr(), ld, y.data_ptr(), n, _ptr(best), rows, n, _stream_handle(x.device)))
    return (y, best.long()) if want_argmax else y


def argmax_rows(x: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    """arg max over the first ``n`` entries of the last dim (first index on ties), int64."""
    x = _prep(x, "x")
    ld = x.size(-1)
    n = ld if n is None else int(n)
    rows = x.numel() // ld
    best = torch.empty(x.shape[:-1], dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.avsr_log_softmax(x.data_ptr(), ld, None, 0, best.data_ptr(), rows, n, _stream_handle(x.device)))
    return best.long()
