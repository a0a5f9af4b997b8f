"""The N > 1 host logic on CPU: two gloo ranks shard a track list and reduce
their timings the way bench.py does (max time, summed samples)."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("libdvd-audio_b200.shard")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    weights = [300, 120, 80, 500, 60, 60, 220]
    mine = shard.shard_tracks(weights, world)[rank]
    # pretend each rank decoded its tracks: time proportional to its bytes
    ms = sum(weights[i] for i in mine) * 0.01 * (1 + rank)
    samples = sum(weights[i] for i in mine) * 1000
    job_ms, job_samples = shard.reduce_job(ms, samples, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((job_ms, job_samples, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_reduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    job_ms, job_samples, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    weights = [300, 120, 80, 500, 60, 60, 220]
    # every track exactly once
    assert sorted(gathered[0] + gathered[1]) == list(range(len(weights)))
    loads = [sum(weights[i] for i in g) for g in gathered]
    assert abs(loads[0] - loads[1]) <= max(weights)          # balanced within one track
    assert job_samples == sum(weights) * 1000                 # SUM over ranks
    assert abs(job_ms - max(loads[0] * 0.01, loads[1] * 0.02)) < 1e-9   # MAX over ranks


def test_shard_edge_cases():
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("libdvd-audio_b200.shard")
    assert shard.shard_tracks([], 4) == [[], [], [], []]
    assert shard.shard_tracks([5], 2) == [[0], []]
    parts = shard.shard_tracks([1] * 17, 8)
    assert sorted(i for p in parts for i in p) == list(range(17))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard.reduce_job(3.5, 100) == (3.5, 100.0)
