"""World-size-2 CPU test (gloo) of the multi-rank plumbing used by bench.py --gpus N."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_waveshaping_synthesis_b200.sharding import (aggregate_throughput, forward_and_gather, gather_audio, shard_bounds,
                                                       wave_bounds)


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 64, 2048):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_wave_bounds_tile_the_job():
    total, world, waves = 2048, 8, 2
    rows = sorted(i for r in range(world) for w in range(waves) for i in range(*wave_bounds(total, r, world, waves, w)))
    assert rows == list(range(total))
    assert wave_bounds(total, 3, world, waves, 1) == (1024 + 3 * 128, 1024 + 4 * 128)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, rank, world)
        full = torch.arange(total * 4, dtype=torch.float32).view(total, 4)
        got = gather_audio(full[lo:hi].clone(), total)
        ms, n = aggregate_throughput(10.0 + rank, float((hi - lo) * 4), torch.device("cpu"))
        # equal shards (one collective into the full batch) and the two-wave overlapped protocol
        tot2 = 8
        full2 = torch.arange(tot2 * 3, dtype=torch.float32).view(tot2, 3)
        lo2, hi2 = shard_bounds(tot2, rank, world)
        same2 = torch.equal(gather_audio(full2[lo2:hi2].clone(), tot2), full2)
        mine = torch.cat([torch.arange(*wave_bounds(tot2, rank, world, 2, w)) for w in range(2)])
        calls = []

        def fake_forward(f0, control):   # "audio" of utterance i = row i of full2
            calls.append(int(f0.shape[0]))
            return full2[f0.long().view(-1)].clone()

        got2 = forward_and_gather(fake_forward, mine.view(-1, 1).float(), mine.view(-1, 1).float(), tot2, waves=2)
        same2 = same2 and torch.equal(got2, full2) and calls == [2, 2]
        q.put((rank, torch.equal(got, full) and same2, ms, n))
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
