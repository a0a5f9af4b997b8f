"""Named synthetic discs used by the CPU (oracle / golden) and GPU (parity)
tests.  Every disc is deterministic: fixed seeds, fixed generator knobs.

Sizes are small (a few thousand frames per track) so the whole CPU suite runs
in a few minutes; the shapes follow BASELINE.json's configs:
  c1_*  PCM tracks (config 1 style)
  c2_*  MLP 2ch 24/96, one substream, FIR + IIR
  c3_*  MLP 6ch 24/96, two substreams, rematrix + LSB bypass
  c4_*  MLP 2ch 24/192, max filter orders
  c5_*  mixed title set
plus edge cases the reference's behaviour defines (Appendix B of SURVEY.md).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import dvda_gen as g  # noqa: E402

RICH = (g.CHECKDATA | g.BYPASS | g.NOISE | g.QUANT | g.OUTSHIFT | g.EXTRAWORD | g.TERMINATOR |
        g.FLAGS | g.SPARSE | g.MIDAU_PARAMS)
WILD = RICH | g.MID_RESTART | g.SYNC_NO_RST | g.SS1_CHK_QUIRK | g.RANDOM_PADS | g.FIR_CARRY


def discs():
    d = {}
    # ---- PCM
    d["c1_pcm_2ch16"] = [[g.pcm(48000, bps=16, rate=48000, assignment=1, seed=1001)]]
    d["pcm_layouts"] = [
        [g.pcm(3000, bps=16, assignment=a, seed=10 + a) for a in (0, 1, 2, 3, 6, 12)],
        [g.pcm(3000, bps=24, assignment=a, seed=30 + a) for a in (0, 1, 2, 3, 6, 12, 17, 20)],
    ]
    d["pcm_rates_ragged"] = [
        [g.pcm(5000, bps=24, rate=r, assignment=1, seed=50 + i,
               features=g.RANDOM_PADS | g.TWO_PACKETS | g.PCM_RAGGED)
         for i, r in enumerate((44100, 88200, 96000, 176400, 192000))],
        [g.pcm(2, bps=16, seed=70), g.pcm(4000, bps=16, assignment=12, rate=96000, seed=71,
                                          features=g.RANDOM_PADS | g.PCM_RAGGED)],
    ]
    # ---- MLP, config shapes
    d["c2_mlp_2ch96"] = [[g.mlp(40000, rate=96000, assignment=1, seed=1002, restart_interval=16,
                                fir_max=4, iir_max=4, noise_bits=13)]]
    d["c3_mlp_6ch96"] = [[g.mlp(16000, rate=96000, assignment=12, substreams=2, seed=1003,
                                restart_interval=16, matrices=3,
                                features=g.CHECKDATA | g.BYPASS | g.NOISE | g.QUANT | g.OUTSHIFT,
                                fir_max=4, iir_max=4, noise_bits=12)]]
    d["c4_mlp_2ch192"] = [[g.mlp(64000, rate=192000, assignment=1, seed=1004, restart_interval=8,
                                 features=g.CHECKDATA | g.MAX_ORDERS, fir_max=8, iir_max=4,
                                 codebooks=0x2, min_lsbs=16, noise_bits=16)]]
    d["c5_mixed"] = [
        [g.pcm(6000, bps=16, seed=2000), g.mlp(8000, seed=2001, matrices=2, features=RICH),
         g.pcm(4000, bps=24, assignment=12, rate=96000, seed=2002),
         g.mlp(6000, rate=192000, seed=2003, restart_interval=8, features=g.CHECKDATA | g.MAX_ORDERS,
               fir_max=8, iir_max=4, min_lsbs=14)],
        [g.mlp(6000, assignment=12, substreams=2, seed=2004, matrices=4, features=RICH),
         g.mlp(5000, assignment=12, substreams=2, seed=2005, matrices=4, features=RICH, join_previous=1),
         g.pcm(3000, bps=24, seed=2006),
         g.mlp(5000, rate=48000, assignment=3, substreams=2, seed=2007, matrices=3, features=RICH,
               restart_interval=4, max_blocks=2)],
    ]
    # ---- MLP, every syntax feature at once, 1 and 2 substreams
    for i in range(4):
        six = i % 2
        d["mlp_wild_%d" % i] = [[
            g.mlp(6000, seed=100 + i, features=WILD, substreams=2 if six else 1,
                  assignment=12 if six else 1, au_frames=40 if six else 0,
                  restart_interval=4, max_blocks=3, fir_max=8, iir_max=8,
                  matrices=6 if six else 2, noise_bits=10 + i),
            g.mlp(3000, seed=200 + i, features=WILD, substreams=2 if six else 1,
                  assignment=12 if six else 1, au_frames=40 if six else 0,
                  restart_interval=3, max_blocks=4, fir_max=8, iir_max=4, matrices=4, join_previous=1)],
            [g.mlp(3000, seed=300 + i, features=WILD, substreams=2 if i == 0 else 1,
                   assignment=[20, 0, 3, 9][i] if i != 0 else 6,
                   restart_interval=5, max_blocks=2, fir_max=6, iir_max=6, matrices=3)]]
    # other channel assignments incl. the permuted RIFF WAVE orders 0x12-0x14
    d["mlp_assignments"] = [[g.mlp(2000, seed=400 + a, assignment=a, substreams=2 if a in (18, 19, 20, 6) else 1,
                                   matrices=2, features=RICH, restart_interval=6)
                             for a in (0, 2, 6, 18, 19, 20)]]
    # cross-segment FIR history (Appendix B-2): restart immediately followed by FIR order > 0
    d["mlp_fir_carry"] = [[g.mlp(12000, seed=500, features=g.CHECKDATA | g.FIR_CARRY, restart_interval=2,
                                 fir_max=8, iir_max=0),
                           g.mlp(9000, seed=501, features=g.CHECKDATA | g.FIR_CARRY | g.BYPASS, substreams=2,
                                 assignment=12, restart_interval=3, fir_max=4, iir_max=4, matrices=3)]]
    # no check data, every codebook alone, tiny and large residuals
    d["mlp_codebooks"] = [[g.mlp(3000, seed=600 + cb, features=0, codebooks=1 << cb, noise_bits=nb,
                                 restart_interval=8)
                           for cb, nb in ((0, 4), (1, 8), (2, 12), (3, 18), (0, 20))]]
    # access units larger than a packet: the reference ends the track at the first
    # packet in which no access unit ends (dvd-audio.c:766-775)
    d["mlp_zero_yield"] = [[g.mlp(8000, seed=700, assignment=12, substreams=2, au_frames=160, matrices=2,
                                  features=RICH, restart_interval=4, noise_bits=18),
                            g.mlp(6000, seed=701, assignment=12, substreams=2, features=RICH | g.TWO_PACKETS,
                                  restart_interval=4, matrices=3)]]
    # many short restart segments, one-AU segments, joined tracks
    d["mlp_short_segments"] = [[g.mlp(4000, seed=800, restart_interval=1, rate=48000),
                                g.mlp(4000, seed=801, restart_interval=1, rate=48000, join_previous=1),
                                g.mlp(4000, seed=802, restart_interval=2, max_blocks=4, features=RICH, matrices=2)]]
    # ---- behaviours of the reference that have no other fixture
    # a later major sync that states other stream parameters: the access unit is dropped (mlp.c:449-455)
    d["mlp_param_dup"] = [[g.mlp(8000, rate=48000, seed=900, features=g.CHECKDATA | g.SYNC_PARAM_DUP, restart_interval=4),
                           g.mlp(6000, seed=901, assignment=12, substreams=2, matrices=3, features=RICH | g.SYNC_PARAM_DUP,
                                 restart_interval=3),
                           g.mlp(12000, rate=192000, seed=902, features=g.CHECKDATA | g.SYNC_PARAM_DUP, restart_interval=2,
                                 noise_bits=18)]]
    # PCM packets that change the stream parameters end the track (dvd-audio.c:1049-1055)
    d["pcm_param_change"] = [[g.pcm(6000, bps=16, seed=910, features=g.PCM_PARAM_CHANGE),
                              g.pcm(5000, bps=24, assignment=12, rate=96000, seed=911, features=g.PCM_PARAM_CHANGE | g.RANDOM_PADS),
                              g.pcm(3000, bps=16, seed=912)]]
    # tracks whose tables do not point at the sector their audio starts in (reference TODO:64-80)
    d["late_start"] = [[g.mlp(8000, seed=920, restart_interval=4),
                        g.mlp(8000, seed=921, restart_interval=4, join_previous=1, start_shift=-3),
                        g.mlp(6000, seed=922, restart_interval=4, join_previous=1, start_shift=2),
                        g.pcm(6000, seed=923), g.pcm(6000, seed=924, start_shift=1)]]
    # tracks that start in, end in and span AOB file boundaries (aob.c:101-123, 181-199): see MAX_AOB_BYTES
    d["aob_split"] = [[g.mlp(20000, seed=930, restart_interval=8), g.pcm(40000, bps=24, rate=96000, seed=931),
                       g.mlp(16000, seed=932, assignment=12, substreams=2, matrices=2, features=RICH)]]
    return d


# discs whose AOB stream is cut into several files (dvda_gen's max_aob_bytes)
MAX_AOB_BYTES = {"aob_split": 60 * 2048}


GPU_LARGE = {
    # bigger cases for the GPU box only (oracle still finishes in seconds)
    "c2_large": [[g.mlp(2_000_000, rate=96000, assignment=1, seed=3002, restart_interval=16,
                        fir_max=4, iir_max=4, noise_bits=13)]],
    "c3_large": [[g.mlp(600_000, rate=96000, assignment=12, substreams=2, seed=3003, restart_interval=16,
                        matrices=3, features=g.CHECKDATA | g.BYPASS | g.NOISE | g.QUANT | g.OUTSHIFT)]],
    "c1_large": [[g.pcm(4_000_000, bps=16, seed=3001), g.pcm(1_000_000, bps=24, assignment=12, rate=96000, seed=3004)]],
}


import workloads  # noqa: E402  (gen/workloads.py: the config shapes shared with bench.py)

GPU_LARGE["c5_titleset_64"] = workloads.c5_titleset(1)


def random_disc(seed):
    """Three tracks whose shape (channel layout, substreams, access-unit size, restart interval,
    blocks per unit, filter orders, matrices, syntax features, PCM in between) is drawn at random."""
    import random
    rnd = random.Random(7700 + seed)
    layouts = [(0, 1), (1, 1), (1, 1), (2, 1), (3, 1), (9, 1), (20, 1), (6, 2), (12, 2), (12, 2), (18, 2), (19, 2), (20, 2)]
    bits = [g.CHECKDATA, g.BYPASS, g.NOISE, g.QUANT, g.OUTSHIFT, g.EXTRAWORD, g.TERMINATOR, g.FLAGS, g.SPARSE,
            g.MIDAU_PARAMS, g.MID_RESTART, g.SYNC_NO_RST, g.SS1_CHK_QUIRK, g.RANDOM_PADS, g.FIR_CARRY]

    def mlp_track(join):
        asg, nss = rnd.choice(layouts)
        feats = 0
        for b in bits:
            if rnd.random() < 0.45:
                feats |= b
        rate = rnd.choice([44100, 48000, 96000, 96000, 192000])
        fir = rnd.choice([0, 2, 4, 8])
        return g.mlp(rnd.randrange(1500, 7000), bps=rnd.choice([16, 24, 24]), rate=rate, assignment=asg, seed=rnd.randrange(1, 1 << 20),
                     features=feats, substreams=nss, au_frames=rnd.choice([0, 0, 40]) if rate <= 48000 else 0,
                     restart_interval=rnd.choice([1, 2, 3, 5, 8, 16]), max_blocks=rnd.choice([1, 1, 2, 4]),
                     fir_max=fir, iir_max=rnd.choice([0, 2, 4]) if fir <= 4 else 0,
                     matrices=rnd.choice([0, 1, 2, 3, 6]), noise_bits=rnd.randrange(4, 20), join_previous=join)

    tracks = [mlp_track(0)]
    tracks.append(g.pcm(rnd.randrange(800, 5000), bps=rnd.choice([16, 24]), rate=rnd.choice([48000, 96000]),
                        assignment=rnd.choice([0, 1, 3]), seed=rnd.randrange(1, 1 << 20)) if rnd.random() < 0.4 else mlp_track(0))
    join = 1 if tracks[1]["codec"] == 1 and rnd.random() < 0.3 else 0
    tracks.append(mlp_track(join))
    if join:                                              # a joined track continues its predecessor's stream layout
        for key in ("bps_code", "rate_code", "assignment", "substreams", "au_frames"):
            tracks[2][key] = tracks[1][key]
    return tracks
