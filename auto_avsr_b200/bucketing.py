"""Max-frames bucketing and on-device collate.

``max_frames_batches`` restates how the reference forms its batches (datamodule/data_module.py:44-100: bucketise the
lengths into ``num_buckets`` linspace bins, sort by length descending, stable-sort by bin, then greedily fill a batch
while the SUM of lengths stays <= max_frames) and is pinned to the reference's own functions by
tests/golden/bucketing.json.  ``assign_to_ranks`` spreads those batches over the GPUs of a node -- the reference's
layout (Lightning's distributed sampler: batch i -> rank i mod W) or a padded-work-balanced one (longest-processing-time
first over B*Tmax, the frames the encoder really computes: SURVEY.md D6) -- and ``pack_bucket`` forms the zero-padded
(B, Tmax, d) batch on the GPU from the utterances laid back to back in one flat buffer, so that only the valid frames
cross PCIe.  Host logic is plain Python / torch-free except for device memory; the collate is a CUDA kernel
(csrc/pack.cu)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from ._cabi import check, lib
from .engine import _stream_handle, require_cuda


def _linspace_bins(lo: float, hi: float, n: int) -> List[float]:
    if n == 1:
        return [float(lo)]
    t = torch.linspace(lo, hi, n)               # the reference's own bin edges (float32 arithmetic included)
    return [float(v) for v in t]


def max_frames_batches(lengths: Sequence[int], max_frames: int, num_buckets: int = 50,
                       batch_size: Optional[int] = None) -> List[List[int]]:
    """Indices of the utterances of every batch, in the reference's order (CustomBucketDataset with shuffle=False)."""
    lengths = [int(v) for v in lengths]
    if not lengths:
        return []
    if max_frames < max(lengths):
        raise ValueError(f"max_frames={max_frames} is smaller than the longest utterance ({max(lengths)})")
    edges = _linspace_bins(min(lengths), max(lengths), num_buckets)
    # torch.bucketize(v, edges) (right=False): number of edges strictly smaller than v
    def bucket(v):
        lo, hi = 0, len(edges)
        while lo < hi:
            mid = (lo + hi) // 2
            if edges[mid] < v:
                lo = mid + 1
            else:
                hi = mid
        return lo
    items = [(i, l, bucket(float(l))) for i, l in enumerate(lengths)]
    items.sort(key=lambda x: x[1], reverse=True)        # stable, like sorted(..., reverse=True)
    items.sort(key=lambda x: x[2])
    batches: List[List[int]] = []
    cur: List[int] = []
    count = 0
    for idx, length, _ in items:
        if count + length > max_frames or (batch_size and len(cur) == batch_size):
            batches.append(cur)
            cur, count = [idx], length
        else:
            cur.append(idx)
            count += length
    if cur:
        batches.append(cur)
    return batches


def padded_frames(batch: Sequence[int], lengths: Sequence[int]) -> int:
    """Frames the encoder computes for a batch: B * Tmax (padded frames are data)."""
    return len(batch) * max(int(lengths[i]) for i in batch) if batch else 0


def assign_to_ranks(batches: Sequence[Sequence[int]], lengths: Sequence[int], world_size: int,
                    policy: str = "reference") -> List[List[int]]:
    """-> per rank, the indices (into ``batches``) it processes, in order.
    ``reference``: batch i -> rank i % W (what the reference's DDP run does).  ``balanced``: longest-processing-time first
    over the padded work, so that the ranks of a step finish together (inference / evaluation order is free)."""
    if policy == "reference":
        return [list(range(r, len(batches), world_size)) for r in range(world_size)]
    if policy != "balanced":
        raise ValueError(policy)
    order = sorted(range(len(batches)), key=lambda i: -padded_frames(batches[i], lengths))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += padded_frames(batches[i], lengths)
    # the reference order is already sorted by length, so its round robin can beat the greedy bound: keep the better plan
    ref = [list(range(r, len(batches), world_size)) for r in range(world_size)]
    ref_load = max(sum(padded_frames(batches[i], lengths) for i in r) for r in ref) if batches else 0
    return out if max(load) <= ref_load else ref


def pack_bucket(flat: torch.Tensor, lengths: Sequence[int], pad_value: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """flat (sum(lengths), d) fp32 CUDA, utterances back to back -> (padded (B, Tmax, d), lengths int32 (B)) on the device.
    The device-side ``collate_pad`` (data_module.py:32-41)."""
    require_cuda(flat, "flat features")
    lengths = [int(v) for v in lengths]
    if flat.dim() != 2 or flat.size(0) != sum(lengths):
        raise ValueError(f"flat must be (sum(lengths)={sum(lengths)}, d), got {tuple(flat.shape)}")
    B, Tmax, d = len(lengths), max(lengths) if lengths else 0, flat.size(1)
    flat = flat.detach().contiguous()
    off = torch.tensor([0] + list(torch.tensor(lengths).cumsum(0).tolist()), dtype=torch.int64).to(flat.device, non_blocking=True)
    out = torch.empty(B, Tmax, d, dtype=torch.float32, device=flat.device)
    ln = torch.empty(B, dtype=torch.int32, device=flat.device)
    with torch.cuda.device(flat.device):
        check(lib.avsr_pack_padded(flat.data_ptr(), off.data_ptr(), out.data_ptr(), ln.data_ptr(), B, Tmax, d,
                                   float(pad_value), _stream_handle(flat.device)))
    return out, ln


def unpack_bucket(padded: torch.Tensor, lengths: Sequence[int]) -> torch.Tensor:
    """(B, Tmax, d) -> (sum(lengths), d): the valid frames, back to back."""
    require_cuda(padded, "padded batch")
    lengths = [int(v) for v in lengths]
    B, Tmax, d = padded.shape
    padded = padded.detach().contiguous()
    off = torch.tensor([0] + list(torch.tensor(lengths).cumsum(0).tolist()), dtype=torch.int64).to(padded.device, non_blocking=True)
    flat = torch.empty(sum(lengths), d, dtype=torch.float32, device=padded.device)
    with torch.cuda.device(padded.device):
        check(lib.avsr_unpack_padded(padded.data_ptr(), off.data_ptr(), flat.data_ptr(), B, Tmax, d, _stream_handle(padded.device)))
    return flat
