"""Drop-in for espnet.nets.pytorch_backend.transformer.repeat (reference repeat.py:8-42)."""
import torch


class MultiSequential(torch.nn.Sequential):
    """Sequential over layers that take and return an argument tuple; layer-drop only while training."""

    def __init__(self, *args, layer_drop_rate=0.0):
        super().__init__(*args)
        self.layer_drop_rate = layer_drop_rate

    def forward(self, *args):
        keep = torch.empty(len(self)).uniform_()       # same CPU RNG consumption as the reference (repeat.py:23)
        for idx, layer in enumerate(self):
            if not self.training or keep[idx] >= self.layer_drop_rate:
                args = layer(*args)
        return args


def repeat(N, fn, layer_drop_rate=0.0):
    return MultiSequential(*[fn(n) for n in range(N)], layer_drop_rate=layer_drop_rate)
