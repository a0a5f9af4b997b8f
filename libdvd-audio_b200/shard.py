"""Host-side helpers for running the engine on several GPUs of one box.

Tracks (and the restart segments inside them) are independent and write
disjoint output, so multi-GPU operation is pure partitioning: no collective on
the data path.  The unit of work is a track, or — for a long MLP track — a
*part*: a run of consecutive sectors of the track, decoded as a "track" of its
own with the DVDAGPU_PART_* flags (include/dvdagpu.h), whose cut lands on the
first restart point behind its last sector, so that the parts' outputs
concatenate to the track (the reference's loop over tracks,
utils/dvda2wav.c:241-280, is the model: tracks are its unit of work too).
The only cross-rank traffic is the gather of the output in host memory and the
timing reduction of bench.py.
"""

PART_CONTINUES_PREVIOUS = 1
PART_CONTINUED_BY_NEXT = 2


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


def plan_units(tracks, codecs, world, units_per_rank=4, min_part_sectors=4096):
    """Cuts a title set into units of work for `world` ranks.

    tracks: [(first_sector, last_sector, pts_length), ...]; codecs: per track 1 = MLP (may be cut
    into parts), anything else = one unit.  A track longer than the target unit size (the title
    set's sectors / (world * units_per_rank)) is cut into equal parts of at least
    min_part_sectors sectors.  Returns a list of dicts
    {track, part, parts, first, last, pts, flags, sectors} in track / part order."""
    total = sum(max(0, last - first + 1) for first, last, _p in tracks)
    target = max(min_part_sectors, total // max(1, world * units_per_rank))
    units = []
    for ti, ((first, last, pts), codec) in enumerate(zip(tracks, codecs)):
        n = max(0, last - first + 1)
        parts = 1
        if codec == 1 and world > 1 and n > target + target // 2:
            parts = min((n + target - 1) // target, max(1, n // min_part_sectors))
        size = (n + parts - 1) // parts if parts else n
        for p in range(parts):
            s0 = first + p * size
            e = last if p + 1 == parts else s0 + size - 1
            flags = (PART_CONTINUES_PREVIOUS if p else 0) | (PART_CONTINUED_BY_NEXT if p + 1 < parts else 0)
            units.append(dict(track=ti, part=p, parts=parts, first=s0, last=e, pts=pts, flags=flags, sectors=e - s0 + 1))
    return units


def assign_units(units, world):
    """Greedy longest-first assignment of units to ranks by sector count.  Returns `world` lists
    of units, each in sector order; every unit appears exactly once."""
    picks = shard_tracks([u["sectors"] for u in units], world)
    return [[units[i] for i in sorted(mine, key=lambda i: units[i]["first"])] for mine in picks]


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
