"""CPU test of the N>1 host logic: stream sharding and the counter all-reduce (gloo, world_size 2)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_ranges_partition():
    from opv_cxx_demod_b200.shard import stream_range

    for S in (1, 7, 1024, 16384):
        for W in (1, 2, 3, 4, 8):
            spans = [stream_range(r, W, S) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from opv_cxx_demod_b200.shard import reduce_counters, reduce_max_ms

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = {"samples": 1000 * (rank + 1), "frames_decoded": 10 + rank, "bit_errors": rank}
    tot = reduce_counters(local, torch.device("cpu"))
    ms = reduce_max_ms(5.0 + rank, torch.device("cpu"))
    q.put((rank, tot, ms))
    dist.destroy_process_group()


def test_counter_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tot, ms in res:
        assert tot == {"samples": 3000, "frames_decoded": 21, "bit_errors": 1}
        assert ms == 6.0
