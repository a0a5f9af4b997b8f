"""Host-side helpers for running the engine on several GPUs of one box.

Tracks (and the restart segments inside them) are independent and write
disjoint output, so multi-GPU operation is pure partitioning: no collective on
the data path.  The only cross-rank traffic is the timing reduction of bench.py.
"""


def shard_tracks(weights, world):
    """Greedy longest-first assignment of tracks to ranks, balancing `weights`
    (e.g. AOB bytes per track).  Returns a list of `world` lists of track
    indices; every track appears exactly once."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += weights[i]
    for r in range(world):
        out[r].sort()
    return out


def reduce_job(ms, samples, dist=None, device=None):
    """Whole-job numbers from per-rank ones: (max milliseconds, total samples).
    `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms), float(samples)
    import torch
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    s = torch.tensor([float(samples)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t[0]), float(s[0])
