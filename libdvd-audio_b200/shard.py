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


def plan_units(tracks, codecs, world, units_per_rank=4, min_part_sectors=4096, weights=None):
    """Cuts a title set into units of work for `world` ranks.

    tracks: [(first_sector, last_sector, pts_length), ...]; codecs: per track 1 = MLP (may be cut
    into parts), anything else = one unit; weights: per track, relative cost of decoding one of its
    sectors (default 1: PCM sectors are several times cheaper than MLP ones, see sector_weight).
    A track whose cost exceeds the target unit cost (the title set's cost / (world *
    units_per_rank)) is cut into equal parts of at least min_part_sectors sectors.  Returns a list
    of dicts {track, part, parts, first, last, pts, flags, sectors, cost} in track / part order."""
    if weights is None:
        weights = [1.0] * len(tracks)
    total = sum(max(0, last - first + 1) * w for (first, last, _p), w in zip(tracks, weights))
    target = max(1.0, total / max(1, world * units_per_rank))
    units = []
    for ti, ((first, last, pts), codec, w) in enumerate(zip(tracks, codecs, weights)):
        n = max(0, last - first + 1)
        parts = 1
        goal = max(min_part_sectors, int(target / w)) if w > 0 else n      # sectors of this track that cost one target
        if codec == 1 and world > 1 and n > goal + goal // 2:
            parts = min((n + goal - 1) // goal, max(1, n // min_part_sectors))
        size = (n + parts - 1) // parts if parts else n
        for p in range(parts):
            s0 = first + p * size
            e = last if p + 1 == parts else s0 + size - 1
            flags = (PART_CONTINUES_PREVIOUS if p else 0) | (PART_CONTINUED_BY_NEXT if p + 1 < parts else 0)
            units.append(dict(track=ti, part=p, parts=parts, first=s0, last=e, pts=pts, flags=flags,
                              sectors=e - s0 + 1, cost=(e - s0 + 1) * w))
    return units


def sector_weight(codec, channels):
    """Relative cost of decoding one sector on the GPU, from the single-GPU measurements of the
    bench configurations: a PCM sector is unpacked in about a fifth of the time an MLP stereo
    sector takes to decode; more channels per frame cost a little more per sector."""
    if codec != 1:
        return 0.2
    return 1.0 + 0.06 * max(0, channels - 2)


def assign_units(units, world):
    """Greedy longest-first assignment of units to ranks by cost (sector count x weight).
    Returns `world` lists of units, each in sector order; every unit appears exactly once."""
    picks = shard_tracks([u.get("cost", u["sectors"]) for u in units], world)
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
