"""torchrun worker of tests/test_gpu_backward.py::test_ddp_gradient_allreduce_two_ranks (also usable on CPU/gloo for
the control flow: AVSR_DDP_BACKEND=gloo, then the modules are plain torch stand-ins).  Each rank runs a different
synthetic bucket through DDP(stack of LayerNorm -> FFN -> LayerNorm -> ConvolutionModule blocks); after backward every
rank must hold the SAME gradients = the mean of the per-rank gradients (checked against a manual all-reduce of a second,
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
    from auto_avsr_b200 import ConvolutionModule, LayerNorm, PositionwiseFeedForward

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.n1, self.ff, self.n2, self.conv = LayerNorm(768), PositionwiseFeedForward(768, 3072, 0.0), LayerNorm(768), \
                ConvolutionModule(768, 31)

        def forward(self, x):
            x = x + 0.5 * self.ff(self.n1(x))
            return x + self.conv(self.n2(x))

    nblocks = int(os.environ.get("AVSR_DDP_BLOCKS", "12"))     # 12 blocks = 113 M parameters = 454 MB of fp32 gradients
    torch.manual_seed(0)
    model = torch.nn.Sequential(*[Block() for _ in range(nblocks)]).to(dev).train()
    replica = torch.nn.Sequential(*[Block() for _ in range(nblocks)]).to(dev).train()
    replica.load_state_dict(model.state_dict())
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(4, 100, 768, generator=g).to(dev)
    ddp(x).pow(2).mean().backward()
    replica(x).pow(2).mean().backward()
    nbytes, worst = 0, 0.0
    for (n, p), q in zip(model.named_parameters(), replica.parameters()):
        ref = q.grad.clone()
        dist.all_reduce(ref)
        ref /= world
        nbytes += p.grad.numel() * 4
        scale = ref.abs().max().item() + 1e-30
        worst = max(worst, (p.grad - ref).abs().max().item() / scale)
    assert worst < 1e-4, worst
    if rank == 0:
        print(f"DDP-OK world={world} grad_bytes={nbytes} worst_rel={worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
