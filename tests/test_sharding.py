"""Multi-GPU path is 'replicas only': streams are dealt to ranks, no data-path collective.
Covers the sharding + result reduction logic with a world_size-2 gloo group on CPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    streams = [(f"s{i}", 1000 + 37 * i) for i in range(23)]
    mine = bench.shard_streams(streams, rank, world)
    pixels = sum(p for _, p in mine)
    total, tmax = bench.reduce_result(float(pixels), 1.0 + rank, "cpu")
    q.put((rank, [n for n, _ in mine], total, tmax))
    dist.destroy_process_group()


def test_streams_are_dealt_without_overlap_and_reduced():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    names = sorted(n for r in res for n in r[1])
    assert names == sorted(f"s{i}" for i in range(23))
    want_total = float(sum(1000 + 37 * i for i in range(23)))
    for r in res:
        assert r[2] == want_total and r[3] == 2.0
