"""Drop-in for the inference surface of espnet.nets.pytorch_backend.ctc.CTC (reference ctc.py:9-93) and for the
``proj_encoder`` Linear in front of the encoder (e2e_asr_conformer.py:31): SURVEY.md 8f #1, the steps either side of
the encoder hot path."""
from typing import Optional, Tuple

import torch

from .. import ops
from ..engine import default_precision, require_cuda

_PAD = 128   # the tensor-core GEMMs take N in multiples of 128: ctc_lo's odim (5049) is padded with zero rows


class ProjEncoder(torch.nn.Linear):
    """``torch.nn.Linear(512, 768)`` whose eval forward runs in libavsr_b200 (same parameters, same state-dict keys).
    Training / CPU tensors are refused rather than silently computed elsewhere."""

    precision: Optional[str] = None

    def forward(self, x):
        if self.training:
            raise NotImplementedError("ProjEncoder: inference forward only on the B200 path (call .eval())")
        require_cuda(x, "ProjEncoder input")
        return ops.linear(x, self.weight, self.bias, precision=self.precision or default_precision())


class CTC(torch.nn.Module):
    """Same constructor, attributes and state-dict keys (``ctc_lo.weight``, ``ctc_lo.bias``) as the reference's CTC
    module.  ``log_softmax`` / ``softmax`` / ``argmax`` (what CTCPrefixScorer and greedy decoding call) run on the
    GPU: ctc_lo as a tensor-core GEMM, then one row-wise log-sum-exp kernel.  The training loss is row 8f #2."""

    def __init__(self, odim, eprojs, dropout_rate, reduce=True):
        super().__init__()
        self.dropout_rate = dropout_rate
        self.loss = None
        self.ctc_lo = torch.nn.Linear(eprojs, odim)
        self.dropout = torch.nn.Dropout(dropout_rate)
        self.probs = None  # for visualization (reference attribute)
        self.ctc_loss = torch.nn.CTCLoss(reduction="sum" if reduce else "none", zero_infinity=True)
        self.ignore_id = -1
        self.reduce = reduce
        self.precision: Optional[str] = None
        self._padded: Optional[Tuple[tuple, torch.Tensor, torch.Tensor]] = None

    # ------------------------------------------------------------------ helpers
    def _padded_params(self):
        """ctc_lo weight / bias with the output dimension padded to a multiple of 128 (zero rows), cached until the
        parameters change (in-place update -> ``_version``; re-assignment or ``.to()`` -> data_ptr / device)."""
        w, b = self.ctc_lo.weight, self.ctc_lo.bias
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version, w.device)
        if self._padded is None or self._padded[0] != key:
            odim, k = w.shape
            npad = (odim + _PAD - 1) // _PAD * _PAD
            wp = torch.zeros(npad, k, dtype=torch.float32, device=w.device)
            bp = torch.zeros(npad, dtype=torch.float32, device=w.device)
            wp[:odim].copy_(w.detach())
            bp[:odim].copy_(b.detach())
            self._padded = (key, wp, bp)
        return self._padded[1], self._padded[2]

    def _logits(self, hs_pad):
        if self.training:
            raise NotImplementedError("CTC: inference methods only on the B200 path (call .eval())")
        require_cuda(hs_pad, "CTC input")
        wp, bp = self._padded_params()
        return ops.linear(hs_pad, wp, bp, precision=self.precision or default_precision())

    # ------------------------------------------------------------------ reference surface
    def forward(self, hs_pad, hlens, ys_pad):
        raise NotImplementedError("CTC.forward (the training loss, ctc.py:41-65) is not on the B200 inference path")

    def loss_fn(self, th_pred, th_target, th_ilen, th_olen):
        raise NotImplementedError("CTC.loss_fn (ctc.py:32-39) is not on the B200 inference path")

    def log_softmax(self, hs_pad):
        """(B, Tmax, eprojs) -> (B, Tmax, odim) log-probs (ctc.py:77-84)."""
        return ops.log_softmax(self._logits(hs_pad), self.ctc_lo.out_features)

    def softmax(self, hs_pad):
        """(B, Tmax, eprojs) -> (B, Tmax, odim) probabilities; also kept in ``self.probs`` (ctc.py:67-75)."""
        self.probs = ops.log_softmax(self._logits(hs_pad), self.ctc_lo.out_features).exp_()
        return self.probs

    def argmax(self, hs_pad):
        """(B, Tmax, eprojs) -> (B, Tmax) greedy token ids (ctc.py:86-93)."""
        return ops.argmax_rows(self._logits(hs_pad), self.ctc_lo.out_features)
