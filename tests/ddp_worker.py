"""torchrun worker of tests/test_gpu_backward.py::test_ddp_gradient_allreduce_two_ranks (also usable on CPU/gloo for
Each rank runs a different synthetic, ragged bucket through DistributedDataParallel(ConformerEncoder) in train mode;
after backward every rank must hold the SAME gradients = the mean of the per-rank gradients (checked against a manual all-reduce of a second,
un-wrapped replica)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from auto_avsr_b200 import ConformerEncoder

    # the whole 12-layer encoder (170 M parameters = 682 MB of fp32 gradients), dropout off so that the un-wrapped replica
    # sees the same forward; each rank gets its own bucket (train.py:34-37: one bucket per rank per step)
    nblocks = int(os.environ.get("AVSR_DDP_BLOCKS", "12"))
    kw = dict(num_blocks=nblocks, dropout_rate=0.0, positional_dropout_rate=0.0, attention_dropout_rate=0.0)
    os.environ.setdefault("AVSR_B200_PRECISION", "fp32" if os.environ.get("AVSR_DDP_SYNCBN", "0") == "1" else "f16")
    torch.manual_seed(0)
    model = ConformerEncoder(**kw).to(dev).train()
    replica = ConformerEncoder(**kw).to(dev).train()
    replica.load_state_dict(model.state_dict())
    sync_bn = os.environ.get("AVSR_DDP_SYNCBN", "0") == "1"
    if sync_bn:     # what Lightning's sync_batchnorm=True does (train.py:31): conv_module.norm -> torch.nn.SyncBatchNorm
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False)
    g = torch.Generator().manual_seed(100 + rank)
    lengths = [100, 80 - 10 * rank, 64, 33]
    x = torch.randn(4, 100, 768, generator=g).to(dev)
    mask = (torch.arange(100, device=dev)[None, :] < torch.tensor(lengths, device=dev)[:, None]).unsqueeze(1)
    ddp(x, mask)[0].pow(2).mean().backward()
    if sync_bn:
        # reference for SyncBatchNorm: ONE process holding both ranks' buckets in one batch sees the same statistics;
        # every rank rebuilds that: gather the inputs, run the un-wrapped BatchNorm1d replica on the concatenated batch
        # with the global-mean loss, and compare its gradients with DDP's (mean over ranks of the per-rank losses)
        xs_all = [torch.empty_like(x) for _ in range(world)]
        ln = torch.tensor(lengths, device=dev, dtype=torch.int64)
        ln_all = [torch.empty_like(ln) for _ in range(world)]
        dist.all_gather(xs_all, x)
        dist.all_gather(ln_all, ln)
        mask_all = (torch.arange(100, device=dev)[None, :] < torch.cat(ln_all)[:, None]).unsqueeze(1)
        out = replica(torch.cat(xs_all), mask_all)[0]
        (out.pow(2).mean()).backward()
        worst = 0.0
        gscale = max(q.grad.abs().max().item() for q in replica.parameters())
        for (n, p), q in zip(model.named_parameters(), replica.parameters()):
            # (biases in front of a normalisation / of the keys have mathematically zero gradients: floor the scale)
            scale = q.grad.abs().max().item() + 1e-4 * gscale
            worst = max(worst, (p.grad - q.grad).abs().max().item() / scale)
        rm = max((a - b).abs().max().item() for (na, a), (nb, b) in zip(model.named_buffers(), replica.named_buffers()) if "running" in na)
        assert worst < 2e-3 and rm < 1e-4, (worst, rm)
        if rank == 0:
            print(f"DDP-OK syncbn world={world} worst_rel={worst:.2e} running_stat_diff={rm:.2e}")
        dist.destroy_process_group()
        return
    replica(x, mask)[0].pow(2).mean().backward()
    nbytes, worst = 0, 0.0
    for (n, p), q in zip(model.named_parameters(), replica.parameters()):
        ref = q.grad.clone()
        dist.all_reduce(ref)
        ref /= world
        nbytes += p.grad.numel() * 4
        scale = ref.abs().max().item() + 1e-30
        worst = max(worst, (p.grad - ref).abs().max().item() / scale)   # same arithmetic on both sides: exact up to NCCL's sum order
    assert worst < 1e-4, worst
    if rank == 0:
        print(f"DDP-OK world={world} grad_bytes={nbytes} worst_rel={worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except Exception:          # noqa: BLE001 -- torchrun swallows the child's stderr tail: put the traceback on stdout
        import traceback
        print("DDP-WORKER-FAILED rank", os.environ.get("RANK"), traceback.format_exc(), flush=True)
        raise
