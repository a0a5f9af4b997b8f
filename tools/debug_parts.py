import importlib, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import catalog, dvda_gen, oracle
os.environ["DVDAGPU_DEBUG"] = "1"
pkg = importlib.import_module("libdvd-audio_b200")
name, part, ti = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
specs = dict(catalog.discs()); specs.update(catalog.GPU_LARGE)
with tempfile.TemporaryDirectory() as d:
    info = dvda_gen.make_disc(d, specs[name])
    sectors = oracle.read_aobs(d)
    eng = pkg.Engine(0)
    t = info[0][ti]
    ref = oracle.decode_track(sectors, t["first_sector"], t["last_sector"], t["pts_length"])
    print("ref frames", ref["frames"], "first/last", t["first_sector"], t["last_sector"])
    res = eng.decode_host(sectors, [(t["first_sector"], t["last_sector"], t["pts_length"])])
    cuts = list(range(t["first_sector"], t["last_sector"] + 1, part))
    descs = []
    for i, s0 in enumerate(cuts):
        e = t["last_sector"] if i + 1 == len(cuts) else cuts[i + 1] - 1
        descs.append((s0, e, t["pts_length"], (1 if i else 0) | (2 if i + 1 < len(cuts) else 0)))
    res = eng.decode_host(sectors, descs)
    print([(r.status, r.frames, r.stopped) for r in res])
