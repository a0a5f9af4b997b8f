"""Fuzzing aid: the random discs of the given seeds read through the public API (dvda_read) in small parts,
against the oracle.

usage: python tools/fuzz_reader.py seed [seed ...]"""
import importlib, os, sys, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import dvda_gen as g, oracle
import catalog
ns = {"_random_disc": catalog.random_disc}
pkg = importlib.import_module("libdvd-audio_b200")
bad = 0
for seed in [int(x) for x in sys.argv[1:]]:
    tracks = ns["_random_disc"](seed)
    d = tempfile.mkdtemp()
    directory = os.path.join(d, "AUDIO_TS")
    info = g.make_disc(directory, [tracks])
    sectors = oracle.read_aobs(directory)
    os.environ["DVDA_B200_PART_SECTORS"] = str(3 + seed % 7)
    disc = pkg.Disc(directory)
    for i, t in enumerate(info[0]):
        ref = oracle.decode_track(sectors, t["first_sector"], t["last_sector"], t["pts_length"])
        try:
            _inf, pcm = disc.read_track(1, i + 1, chunk=997 + 31 * (seed % 5))
        except Exception as e:
            pcm = None
            err = str(e)
        if ref is None:
            ok = pcm is None
        else:
            ok = pcm is not None and pcm.shape == ref["pcm"].shape and np.array_equal(pcm, ref["pcm"])
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, "track", i + 1, "part", os.environ["DVDA_B200_PART_SECTORS"], None if pcm is None else pcm.shape, None if ref is None else ref["pcm"].shape, {k: v for k, v in tracks[i].items() if k != "seed"})
    disc.close()
    shutil.rmtree(d)
print("mismatches:", bad)
