"""Drop-in for RelPositionalEncoding of espnet.nets.pytorch_backend.transformer.embedding (reference :120-184)."""
import math

import torch

from .. import ops


class RelPositionalEncoding(torch.nn.Module):
    """x * sqrt(d) and the (1, 2T-1, d) sinusoid table ordered rel = +(T-1) ... -(T-1).

    No parameters / buffers (the reference keeps ``pe`` as a plain attribute, so it is not in the state dict
    either).  The table is generated on the device for the requested T; ``max_len`` is accepted for signature
    compatibility -- there is no table to pre-extend."""

    def __init__(self, d_model, dropout_rate, max_len=5000):
        super().__init__()
        self.d_model = d_model
        self.xscale = math.sqrt(self.d_model)
        self.dropout = torch.nn.Dropout(p=dropout_rate)
        self.max_len = max_len

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("RelPositionalEncoding: CPU tensor; auto_avsr_b200 has no CPU fallback")
        pos_emb = ops.rel_sinusoid_table(x.size(1), self.d_model, x.device).unsqueeze(0)
        if self.training:      # embedding.py:183-184: both outputs pass through the positional dropout
            return self.dropout(x * self.xscale), self.dropout(pos_emb)
        return x * self.xscale, pos_emb
