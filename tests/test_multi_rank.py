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


def test_units_cover_every_sector_once():
    """plan_units: whole tracks, long MLP tracks cut into parts with the DVDAGPU_PART_* flags;
    assign_units: every unit on exactly one rank, loads balanced."""
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("libdvd-audio_b200.shard")
    tracks, at = [], 0
    for n in (100, 9000, 40, 30000, 700, 12000, 5, 2500):
        tracks.append((at, at + n - 1, n * 10))
        at += n
    codecs = [0, 1, 1, 1, 0, 0, 1, 1]                      # the 12000-sector track is PCM: never cut
    for world in (1, 2, 8):
        units = shard.plan_units(tracks, codecs, world, units_per_rank=4, min_part_sectors=512)
        covered = []
        for ti, (first, last, pts) in enumerate(tracks):
            mine = [u for u in units if u["track"] == ti]
            assert [u["part"] for u in mine] == list(range(len(mine))) and all(u["parts"] == len(mine) for u in mine)
            assert mine[0]["first"] == first and mine[-1]["last"] == last
            for a, b in zip(mine, mine[1:]):
                assert b["first"] == a["last"] + 1
            for i, u in enumerate(mine):
                assert u["pts"] == pts
                assert bool(u["flags"] & shard.PART_CONTINUES_PREVIOUS) == (i > 0)
                assert bool(u["flags"] & shard.PART_CONTINUED_BY_NEXT) == (i + 1 < len(mine))
            if codecs[ti] != 1 or world == 1:
                assert len(mine) == 1
            covered.append(sum(u["sectors"] for u in mine))
        assert covered == [last - first + 1 for first, last, _p in tracks]
        if world == 8:
            assert len([u for u in units if u["track"] == 3]) > 4          # the long MLP track is shared out
        per_rank = shard.assign_units(units, world)
        seen = sorted((u["track"], u["part"]) for r in per_rank for u in r)
        assert seen == sorted((u["track"], u["part"]) for u in units)
        loads = [sum(u["sectors"] for u in r) for r in per_rank]
        assert max(loads) - min(loads) <= max(u["sectors"] for u in units)
        for r in per_rank:
            assert [u["first"] for u in r] == sorted(u["first"] for u in r)
