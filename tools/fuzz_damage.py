"""Fuzzing aid: single-bit damage on the random discs of the given seeds (five flips per track), engine
against the oracle: where the track ends, which error is flagged, every sample in front of it.

usage: python tools/fuzz_damage.py seed [seed ...]"""
import importlib, os, sys, tempfile, shutil, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import dvda_gen as g, oracle
import catalog
ns = {"_random_disc": catalog.random_disc}
pkg = importlib.import_module("libdvd-audio_b200")
eng = pkg.Engine(0)
bad = trials = 0
for seed in [int(x) for x in sys.argv[1:]]:
    tracks = ns["_random_disc"](seed)
    d = tempfile.mkdtemp()
    directory = os.path.join(d, "AUDIO_TS")
    info = g.make_disc(directory, [tracks])
    clean = oracle.read_aobs(directory)
    shutil.rmtree(d)
    rnd = random.Random(seed)
    for i, t in enumerate(info[0]):
        for _ in range(5):
            sectors = clean.copy()
            off = rnd.randrange(t["first_sector"] * 2048, (t["last_sector"] + 1) * 2048)
            bit = 1 << rnd.randrange(8)
            sectors[off] ^= bit
            dsc = (t["first_sector"], t["last_sector"], t["pts_length"])
            ref = oracle.decode_track(sectors, *dsc)
            trials += 1
            try:
                r = eng.decode_host(sectors, [dsc])[0]
            except Exception as e:                          # noqa: BLE001
                bad += 1
                print("ENGINE ERROR seed", seed, "track", i + 1, "offset", off, "bit", bit, e)
                continue
            if ref is None:
                ok = r.status != 0
                what = "oracle cannot open, engine status %d" % r.status
            else:
                got = eng.fetch(r) if r.status == 0 else None
                ok = r.status == 0 and r.frames == ref["frames"] and r.error_flags == ref["error_flags"] and np.array_equal(got, ref["pcm"])
                what = "status %d frames %d/%d flags %x/%x" % (r.status, r.frames, ref["frames"], r.error_flags, ref["error_flags"])
            if not ok:
                bad += 1
                print("MISMATCH seed", seed, "track", i + 1, "offset", off, "bit", bit, what, {k: v for k, v in tracks[i].items() if k != "seed"})
print("trials:", trials, "mismatches:", bad)
