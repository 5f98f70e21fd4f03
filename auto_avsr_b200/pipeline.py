"""Host <-> device pipelining around the encoder forward (public API).

`PipelinedEncoder` runs `ConformerEncoder.forward` over a stream of pinned host batches with the host->device copy
of batch i+1 and the device->host copy of batch i-1 overlapping the forward of batch i (three CUDA streams, double-
buffered device tensors).  Every batch still pays its own H2D and D2H; only the serialisation is removed.  This is
how a serving loop drives the encoder when features are consumed on the host (bench.py's `e2e` leg uses it).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch


class PipelinedEncoder:
    def __init__(self, encoder, B: int, T: int, depth: int = 2, device: Optional[torch.device] = None):
        self.enc = encoder
        self.dev = device or next(encoder.parameters()).device
        self.B, self.T, self.depth = B, T, depth
        D = encoder._cfg[0]
        self.s_in = torch.cuda.Stream(device=self.dev)
        self.s_out = torch.cuda.Stream(device=self.dev)
        self.x_dev = [torch.empty(B, T, D, device=self.dev) for _ in range(depth)]
        self.len_dev = [torch.empty(B, dtype=torch.int32, device=self.dev) for _ in range(depth)]
        self.y_dev: List[Optional[torch.Tensor]] = [None] * depth
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]       # input slot filled
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]     # input slot consumed by the forward
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]     # forward finished (output slot ready)
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]      # output slot copied to the host
        self._col = torch.arange(T, device=self.dev)[None, :]

    @torch.no_grad()
    def run(self, host_inputs: Sequence[torch.Tensor], host_lengths: Sequence[Optional[torch.Tensor]],
            host_outputs: Sequence[torch.Tensor]) -> None:
        """host_inputs[i]: pinned (B,T,d) f32; host_lengths[i]: pinned int32 (B) or None; host_outputs[i]: pinned
        (B,T,d) f32 receiving the encoder features of batch i.  Returns after enqueueing; call synchronize()."""
        cur = torch.cuda.current_stream(self.dev)
        n = len(host_inputs)
        for i in range(n):
            k = i % self.depth
            with torch.cuda.stream(self.s_in):
                if i >= self.depth:
                    self.s_in.wait_event(self.ev_free[k])            # forward i-depth no longer reads this slot
                self.x_dev[k].copy_(host_inputs[i], non_blocking=True)
                if host_lengths[i] is not None:
                    self.len_dev[k].copy_(host_lengths[i], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            cur.wait_event(self.ev_in[k])
            if i >= self.depth:
                cur.wait_event(self.ev_out[k])                       # output slot k drained to the host
            mask = None
            if host_lengths[i] is not None:                          # make_non_pad_mask(lengths) on the device
                mask = (self._col < self.len_dev[k][:, None]).unsqueeze(1)
            self.y_dev[k], _ = self.enc(self.x_dev[k], mask)
            self.ev_free[k].record(cur)
            self.ev_done[k].record(cur)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_done[k])
                host_outputs[i].copy_(self.y_dev[k], non_blocking=True)
                self.ev_out[k].record(self.s_out)

    def synchronize(self) -> None:
        torch.cuda.synchronize(self.dev)
