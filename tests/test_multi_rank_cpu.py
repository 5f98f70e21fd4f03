"""The N>1 control flow of bench.py (one bucket per rank, no data-path collective, MAX-reduced time, aggregate
frames) exercised on CPU with the gloo backend, world_size 2 -- no GPU, no CUDA library calls."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, lr, w = bench.dist_env()
    assert (r, lr, w) == (rank, rank, world)
    # each rank owns its own bucket (weak scaling): different seeds -> different inputs, same shape
    lengths = list(bench.SHAPES[bench.WORKLOAD])
    from auto_avsr_b200.synthetic import encoder_input
    x = encoder_input(lengths, 768, 1234 + rank * 100)
    # the only cross-rank exchange of the inference bench: MAX over ranks of the elapsed time
    fake_ms = torch.tensor([2.0 + rank, 3.0 + 2 * rank], dtype=torch.float64)
    dist.all_reduce(fake_ms, op=dist.ReduceOp.MAX)
    frames = sum(lengths) * 10 * world
    q.put((rank, float(x.sum()), fake_ms.tolist(), frames / (fake_ms[0].item() * 1e-3)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_weak_scaling_bookkeeping():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] != res[1][1]                       # ranks really hold different buckets
    assert res[0][2] == res[1][2] == [3.0, 5.0]         # MAX over ranks, identical on every rank
    assert res[0][3] == res[1][3] == 1600 * 10 * 2 / 3.0e-3   # aggregate frames / slowest rank's time


def test_flops_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    # S2 = 568.6 GFLOP, S3 = 533.5, S4 = 708.7, S1 = 36.0
    assert abs(bench.algorithmic_flops([400] * 4) / 1e9 - 568.6) < 0.1
    assert abs(bench.algorithmic_flops([100] * 16) / 1e9 - 533.5) < 0.1
    assert abs(bench.algorithmic_flops([1600]) / 1e9 - 708.7) < 0.1
    assert abs(bench.algorithmic_flops([100]) / 1e9 - 36.0) < 0.1
