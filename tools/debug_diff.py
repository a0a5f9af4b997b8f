"""Scratch: decode a catalog disc, print where GPU and oracle differ."""
import importlib, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import catalog, dvda_gen, oracle
pkg = importlib.import_module("libdvd-audio_b200")
name = sys.argv[1]
specs = dict(catalog.discs()); specs.update(catalog.GPU_LARGE)
with tempfile.TemporaryDirectory() as d:
    info = dvda_gen.make_disc(d, specs[name])
    sectors = oracle.read_aobs(d)
    eng = pkg.Engine(0)
    for ti, title in enumerate(info):
        for ki, t in enumerate(title):
            for rep in range(3):
                res = eng.decode_host(sectors, [(t["first_sector"], t["last_sector"], t["pts_length"])])
                ref = oracle.decode_track(sectors, t["first_sector"], t["last_sector"], t["pts_length"])
                got = eng.fetch(res[0])
                n = min(len(got), len(ref["pcm"]))
                bad = np.argwhere(got[:n] != ref["pcm"][:n])
                print(ti, ki, "rep", rep, "frames", res[0].frames, ref["frames"], "ch", res[0].channels, "diffs", len(bad), bad[:12].tolist())
                for f, c in bad[:6]:
                    print("   f=%d c=%d gpu=%d ref=%d delta=%d" % (f, c, got[f, c], ref["pcm"][f, c], got[f, c] - ref["pcm"][f, c]))
