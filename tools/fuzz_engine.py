"""Fuzzing aid: decodes the random discs of the given seeds (the shapes of tests/test_gpu_parity.py::test_random_streams)
in one call each and prints where the engine differs from the oracle.  DBG_TRACK=n: that track alone, with the
engine's debug dump (segment and access-unit tables).

usage: python tools/fuzz_engine.py seed [seed ...]"""
import importlib, os, sys, random, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import dvda_gen as g, oracle
pkg = importlib.import_module("libdvd-audio_b200")

def disc(seed):
    rnd = random.Random(7700 + seed)
    layouts = [(0, 1), (1, 1), (1, 1), (2, 1), (3, 1), (9, 1), (20, 1), (6, 2), (12, 2), (12, 2), (18, 2), (19, 2), (20, 2)]
    bits = [g.CHECKDATA, g.BYPASS, g.NOISE, g.QUANT, g.OUTSHIFT, g.EXTRAWORD, g.TERMINATOR, g.FLAGS, g.SPARSE,
            g.MIDAU_PARAMS, g.MID_RESTART, g.SYNC_NO_RST, g.SS1_CHK_QUIRK, g.RANDOM_PADS, g.FIR_CARRY]
    names = ["CHECKDATA", "BYPASS", "NOISE", "QUANT", "OUTSHIFT", "EXTRAWORD", "TERMINATOR", "FLAGS", "SPARSE",
             "MIDAU_PARAMS", "MID_RESTART", "SYNC_NO_RST", "SS1_CHK_QUIRK", "RANDOM_PADS", "FIR_CARRY"]
    def mlp_track(join):
        asg, nss = rnd.choice(layouts)
        feats = 0
        for b in bits:
            if rnd.random() < 0.45:
                feats |= b
        rate = rnd.choice([44100, 48000, 96000, 96000, 192000])
        fir = rnd.choice([0, 2, 4, 8])
        t = g.mlp(rnd.randrange(1500, 7000), bps=rnd.choice([16, 24, 24]), rate=rate, assignment=asg, seed=rnd.randrange(1, 1 << 20),
                  features=feats, substreams=nss, au_frames=rnd.choice([0, 0, 40]) if rate <= 48000 else 0,
                  restart_interval=rnd.choice([1, 2, 3, 5, 8, 16]), max_blocks=rnd.choice([1, 1, 2, 4]),
                  fir_max=fir, iir_max=rnd.choice([0, 2, 4]) if fir <= 4 else 0,
                  matrices=rnd.choice([0, 1, 2, 3, 6]), noise_bits=rnd.randrange(4, 20), join_previous=join)
        t["_feat_names"] = [n for n, b in zip(names, bits) if feats & b]
        t["_rate"] = rate
        return t
    tracks = [mlp_track(0)]
    tracks.append(g.pcm(rnd.randrange(800, 5000), bps=rnd.choice([16, 24]), rate=rnd.choice([48000, 96000]),
                        assignment=rnd.choice([0, 1, 3]), seed=rnd.randrange(1, 1 << 20)) if rnd.random() < 0.4 else mlp_track(0))
    join = 1 if tracks[1]["codec"] == 1 and rnd.random() < 0.3 else 0
    tracks.append(mlp_track(join))
    if join:
        for key in ("bps_code", "rate_code", "assignment", "substreams", "au_frames"):
            tracks[2][key] = tracks[1][key]
    return tracks

eng = pkg.Engine(0)
ONLY = os.environ.get("DBG_TRACK")
for seed in [int(a) for a in sys.argv[1:]]:
    tracks = disc(seed)
    d = tempfile.mkdtemp()
    clean = [{k: v for k, v in t.items() if not k.startswith("_")} for t in tracks]
    info = g.make_disc(os.path.join(d, "AUDIO_TS"), [clean])
    sectors = oracle.read_aobs(os.path.join(d, "AUDIO_TS"))
    descs = [(t["first_sector"], t["last_sector"], t["pts_length"]) for t in info[0]]
    if ONLY is not None:
        i = int(ONLY)
        os.environ["DVDAGPU_DEBUG"] = "1"
        r = eng.decode_host(sectors, [descs[i]])[0]
        del os.environ["DVDAGPU_DEBUG"]
        ref = oracle.decode_track(sectors, *descs[i])
        got = eng.fetch(r)
        bad = np.argwhere(got != ref["pcm"])
        uf = np.unique(bad[:, 0]) if len(bad) else np.zeros(0, int)
        runs = np.split(uf, np.where(np.diff(uf) != 1)[0] + 1) if len(uf) else []
        print("seed", seed, "track", i, "alone:", len(bad), "samples differ; runs", [(int(x[0]), int(x[-1])) for x in runs[:12]], "channels", sorted(set(bad[:, 1].tolist())) if len(bad) else [])
        continue
    res = eng.decode_host(sectors, descs)
    for i, (r, dsc) in enumerate(zip(res, descs)):
        ref = oracle.decode_track(sectors, *dsc)
        got = eng.fetch(r)
        same = ref is not None and got.shape == ref["pcm"].shape and np.array_equal(got, ref["pcm"])
        print("seed", seed, "track", i, "OK" if same else "DIFFERS", {k: v for k, v in tracks[i].items() if k not in ("seed",)})
        if not same and ref is not None and got.shape == ref["pcm"].shape:
            bad = np.argwhere(got != ref["pcm"])
            fr = bad[:, 0]
            print("   ", len(bad), "samples differ; first frames", fr[:6], "channels", sorted(set(bad[:, 1])), "last frame", fr[-1], "of", len(got))
            f0 = fr[0]
            print("    got", got[f0 - 1:f0 + 3].tolist(), "ref", ref["pcm"][f0 - 1:f0 + 3].tolist())
            # run lengths of differing frames
            uf = np.unique(fr)
            runs = np.split(uf, np.where(np.diff(uf) != 1)[0] + 1)
            print("    runs:", [(int(x[0]), int(x[-1])) for x in runs[:12]], "n runs", len(runs))
