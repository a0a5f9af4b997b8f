"""The synthetic workloads of BASELINE.json's configs, as generator specs (test / bench
infrastructure: tests/catalog.py and bench.py share them).

  c1  2ch 16-bit 48 kHz PCM track
  c2  2ch 24-bit 96 kHz MLP track, one substream, FIR + IIR prediction
  c3  6ch 24-bit 96 kHz MLP track, two substreams, rematrixing + LSB bypass
  c4  2ch 24-bit 192 kHz MLP track, max filter orders, large residuals (entropy-bound stress)
  c5  one title set of 64 mixed PCM / MLP tracks in the styles of c1 .. c4
"""
import dvda_gen as g

C3_FEATURES = g.CHECKDATA | g.BYPASS | g.NOISE | g.QUANT | g.OUTSHIFT


def c1_track(frames, seed):
    return g.pcm(frames, bps=16, rate=48000, assignment=1, seed=seed)


def c2_track(frames, seed):
    return g.mlp(frames, rate=96000, assignment=1, seed=seed, restart_interval=16, fir_max=4, iir_max=4, noise_bits=13)


def c3_track(frames, seed):
    return g.mlp(frames, rate=96000, assignment=12, substreams=2, seed=seed, restart_interval=16, matrices=3,
                 features=C3_FEATURES)


def c4_track(frames, seed):
    return g.mlp(frames, rate=192000, assignment=1, seed=seed, restart_interval=8, features=g.CHECKDATA | g.MAX_ORDERS,
                 fir_max=8, iir_max=4, codebooks=0x2, min_lsbs=16, noise_bits=16)


def c5_titleset(scale=1):
    """BASELINE.json configs[4] in shape: one title set of 64 tracks (2 titles x 32): 16 PCM tracks
    (16- and 24-bit, stereo and 6 channels) and 48 MLP tracks in the styles of c2 / c3 / c4, seeds
    2000 + track.  `scale` multiplies every track's length (scale 1: about 2 000 restart segments,
    the size the parity tests use; bench.py uses 40: about 80 000 segments, 10 000 per GPU of
    eight)."""
    titles = []
    n = 0
    for _title in range(2):
        tracks = []
        for i in range(32):
            seed = 2000 + n
            kind = i % 4
            if kind == 0:
                v = (i // 4) % 4
                tracks.append(g.pcm((30_000 + 4_000 * v) * scale, bps=(16, 24, 24, 16)[v], assignment=(1, 1, 12, 12)[v],
                                    rate=(48000, 96000, 96000, 48000)[v], seed=seed))
            elif kind == 1:
                tracks.append(c2_track((40_000 + 800 * i) * scale, seed))
            elif kind == 2:
                tracks.append(c3_track((16_000 + 400 * i) * scale, seed))
            else:
                tracks.append(c4_track((64_000 + 1_600 * i) * scale, seed))
            n += 1
        titles.append(tracks)
    return titles


NAMES = {
    "c1": "2ch 16-bit 48 kHz PCM AOB track, %d s",
    "c2": "2ch 24-bit 96 kHz MLP AOB track, 1 substream, FIR+IIR, %d s",
    "c3": "6ch 24-bit 96 kHz MLP AOB, 2 substreams, rematrix + LSB bypass, %d s",
    "c4": "2ch 24-bit 192 kHz MLP AOB, max filter orders, %d s",
    "c5": "full synthetic title set: 64 mixed PCM/MLP tracks (c1..c4 styles), lengths x%d",
}


def spec(config, seconds=600, seed=None, scale=40):
    """(titles, workload name, nominal sample rate, nominal channels) of a config.  For c1 .. c4
    one track of `seconds` seconds; for c5 the title set with its track lengths times `scale`
    (rate / channels are those of its first MLP style, for the x-realtime figure only)."""
    if config == "c1":
        return [[c1_track(int(seconds * 48000), 1001 if seed is None else seed)]], NAMES[config] % seconds, 48000, 2
    if config == "c2":
        return [[c2_track(int(seconds * 96000), 1002 if seed is None else seed)]], NAMES[config] % seconds, 96000, 2
    if config == "c3":
        return [[c3_track(int(seconds * 96000), 1003 if seed is None else seed)]], NAMES[config] % seconds, 96000, 6
    if config == "c4":
        return [[c4_track(int(seconds * 192000), 1004 if seed is None else seed)]], NAMES[config] % seconds, 192000, 2
    if config == "c5":
        return c5_titleset(scale), NAMES[config] % scale, 96000, 2
    raise ValueError("unknown config " + config)
