"""CPU restatement of the two steps either side of the encoder in the reference's ``E2E`` model:
``proj_encoder`` in front and the CTC head (``ctc_lo`` + log-softmax) behind.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Nothing in the shipped package imports this file.

Functional restatement on a flat ``{key: tensor}`` state dict with the reference's keys (``proj_encoder.*``,
``ctc.ctc_lo.*``); the log-softmax is written out as a max-shifted log-sum-exp rather than calling
``F.log_softmax``.  Pinned against the reference's own modules by ``oracle/make_golden_head.py`` ->
``tests/golden/head_*.npz`` (checked in ``tests/test_oracle_head_golden.py``).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import conformer_oracle as enc_oracle

Tensor = torch.Tensor


def proj_encoder(feats: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """``x = self.proj_encoder(x)``: Linear(512 -> 768) with bias (e2e_asr_conformer.py:31, :70)."""
    return feats @ sd["proj_encoder.weight"].to(feats.dtype).T + sd["proj_encoder.bias"].to(feats.dtype)


def ctc_logits(hs: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """``self.ctc_lo(hs_pad)``: Linear(768 -> odim) with bias (ctc.py:21, :54; dropout is identity in eval)."""
    return hs @ sd["ctc.ctc_lo.weight"].to(hs.dtype).T + sd["ctc.ctc_lo.bias"].to(hs.dtype)


def ctc_log_softmax(hs: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """``CTC.log_softmax``: log_softmax(ctc_lo(hs), dim=-1) (ctc.py:77-84; the call CTCPrefixScorer makes,
    scorers/ctc.py:35, :96, :138).  z - max - log(sum(exp(z - max)))."""
    z = ctc_logits(hs, sd)
    zs = z - z.max(dim=-1, keepdim=True).values
    return zs - zs.exp().sum(dim=-1, keepdim=True).log()


def ctc_softmax(hs: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """``CTC.softmax`` (ctc.py:67-75)."""
    return ctc_log_softmax(hs, sd).exp()


def ctc_argmax(hs: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """``CTC.argmax`` (ctc.py:86-93): greedy token per frame."""
    return ctc_logits(hs, sd).argmax(dim=-1)


def features_to_log_probs(head_sd: Dict[str, Tensor], enc_sd: Dict[str, Tensor], feats: Tensor,
                          lengths: Optional[Sequence[int]], n_heads: int) -> Tensor:
    """front-end features -> proj_encoder -> ConformerEncoder -> CTC log-probs: the inference chain of
    e2e_asr_conformer.py:70-71 followed by ``ctc.log_softmax`` (lightning.py's forward / the CTC scorer)."""
    x = proj_encoder(feats, head_sd)
    hs = enc_oracle.encoder_forward(enc_sd, x, lengths, n_heads)
    return ctc_log_softmax(hs, head_sd)
