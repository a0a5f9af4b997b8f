"""Scratch tool: decode one catalog disc on the GPU with engine debug output."""
import importlib, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import catalog, dvda_gen, oracle
os.environ["DVDAGPU_DEBUG"] = "1"
pkg = importlib.import_module("libdvd-audio_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2_mlp_2ch96"
specs = dict(catalog.discs()); specs.update(catalog.GPU_LARGE)
with tempfile.TemporaryDirectory() as d:
    info = dvda_gen.make_disc(d, specs[name])
    sectors = oracle.read_aobs(d)
    eng = pkg.Engine(0)
    for title in info:
        for t in title:
            res = eng.decode_host(sectors, [(t["first_sector"], t["last_sector"], t["pts_length"])])
            ref = oracle.decode_track(sectors, t["first_sector"], t["last_sector"], t["pts_length"])
            got = eng.fetch(res[0])
            print("frames gpu", res[0].frames, "oracle", ref["frames"], "err", res[0].error_flags, ref["error_flags"])
            n = min(len(got), len(ref["pcm"]))
            if n:
                bad = np.argwhere(got[:n] != ref["pcm"][:n])
                print("diffs:", len(bad), bad[:5].tolist())
                if len(bad):
                    f = bad[0][0]
                    print("gpu", got[f:f+3].tolist(), "ref", ref["pcm"][f:f+3].tolist())
            print(eng.stats())
