"""World-size-2 CPU test (gloo) of the multi-rank plumbing used by bench.py --gpus N."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_waveshaping_synthesis_b200.sharding import aggregate_throughput, gather_audio, shard_bounds


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 64, 2048):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, rank, world)
        full = torch.arange(total * 4, dtype=torch.float32).view(total, 4)
        got = gather_audio(full[lo:hi].clone(), total)
        ms, n = aggregate_throughput(10.0 + rank, float((hi - lo) * 4), torch.device("cpu"))
        q.put((rank, torch.equal(got, full), ms, n))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_throughput():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world = 5, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, ms, n in res:
        assert same
        assert ms == 11.0          # max over ranks
        assert n == total * 4      # sum over ranks
