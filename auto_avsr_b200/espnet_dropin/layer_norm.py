"""Drop-in for espnet.nets.pytorch_backend.transformer.layer_norm (reference layer_norm.py:12-33)."""
import torch

from .. import ops


class LayerNorm(torch.nn.LayerNorm):
    """nn.LayerNorm(nout, eps=1e-12) over the last dim; parameters ``weight`` / ``bias`` as in the reference."""

    def __init__(self, nout, dim=-1):
        super().__init__(nout, eps=1e-12)
        self.dim = dim

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("LayerNorm: CPU tensor; auto_avsr_b200 has no CPU fallback")
        if self.dim != -1:
            return self._last_dim(x.transpose(1, -1).contiguous()).transpose(1, -1)
        return self._last_dim(x)

    def _last_dim(self, x):
        if self.training and torch.is_grad_enabled():
            from ..train import LayerNormFn          # training slice: forward + backward in libavsr_b200
            return LayerNormFn.apply(x, self.weight, self.bias)
        return ops.layernorm(x, self.weight, self.bias)
