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


def test_every_rank_decodes_the_full_set():
    """The headline (weak-scaling) job: at world 2 / 4 / 8 every rank holds all 172 distinct
    streams exactly once.  (Round 1 dealt `world` concatenated copies round-robin, which gives a
    rank 172 / world distinct streams `world` times each whenever world divides 172.)"""
    import bench
    streams = bench.load_streams()
    assert len(streams) == 172
    for world in (1, 2, 4, 8):
        for rank in range(world):
            mine = bench.rank_job(streams, rank, world)
            names = [n for n, _, _ in mine]
            assert len(names) == 172 and len(set(names)) == 172
            assert sum(len(d) for _, d, _ in mine) == sum(len(d) for _, d, _ in streams)


def test_strong_scaling_batch_is_a_partition():
    """shard_streams(): a 64-stream batch (BASELINE configs[4]) split over 1 / 2 / 4 / 8 ranks is a
    partition: nothing dropped, nothing duplicated."""
    import bench
    batch = [f"s{i}" for i in range(64)]
    for world in (1, 2, 4, 8):
        parts = [bench.shard_streams(batch, r, world) for r in range(world)]
        assert sorted(x for p in parts for x in p) == sorted(batch)
        assert all(len(p) == 64 // world for p in parts)
