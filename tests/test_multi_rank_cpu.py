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


def _decode_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu import build
    from helpers import load_decoder_case
    from test_decoder_dropin_cpu import CpuCTC, _dropin_decoder
    from auto_avsr_b200.beam_search import DeviceBeamSearch, decode_sharded
    c = load_decoder_case("decoder_tiny")
    cfg = c["cfg"]
    search = DeviceBeamSearch(_dropin_decoder(c, build.load()), CpuCTC(c["head_sd"]), beam_size=cfg["beam"], vocab_size=cfg["odim"])
    mem = c["memory"]
    corpus = [mem, mem[:9], mem[:15], mem[:6], mem[:12]]                  # five utterances of different lengths
    res = decode_sharded(search, corpus, rank, world)
    q.put((rank, search.stats["utterances"], [[h["yseq"] for h in nb] for nb in res], [nb[0]["score"] for nb in res]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_decoding_gathers_every_nbest():
    """SURVEY.md 8e for the decoding path: utterances shard over the ranks with no data-path collective; the host-side
    gather leaves the whole corpus' n-best lists on every rank, identical to a single-rank run (host replay, gloo)."""
    import shutil
    import pytest
    if not os.path.exists("/usr/local/cuda/bin/nvcc") and shutil.which("nvcc") is None:
        pytest.skip("nvcc is needed to build the host replay")
    sys.path.insert(0, ROOT)
    from auto_avsr_b200.beam_search import shard_utterances
    assert shard_utterances([23, 9, 15, 6, 12], 2) == [[0], [1, 2, 3, 4]]          # cost ~ T^2: 529 vs 225 + 144 + 81 + 36 = 486
    assert shard_utterances([10, 10, 10, 10], 2) == [[0, 2], [1, 3]]
    assert shard_utterances([5, 5, 5], 1) == [[0, 1, 2]] and shard_utterances([], 3) == [[], [], []]
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_decode_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[1][1]) == (1, 4)                     # each rank searched only its share
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]    # ... and holds every n-best after the gather
    assert all(len(nb) >= 1 for nb in res[0][2])
