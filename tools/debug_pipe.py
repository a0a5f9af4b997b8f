import importlib, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import catalog, dvda_gen, oracle
os.environ["DVDAGPU_DEBUG"] = "1"
pkg = importlib.import_module("libdvd-audio_b200")
name, part = sys.argv[1], int(sys.argv[2])
specs = dict(catalog.discs()); specs.update(catalog.GPU_LARGE)
with tempfile.TemporaryDirectory() as d:
    info = dvda_gen.make_disc(d, specs[name])
    sectors = oracle.read_aobs(d)
    eng = pkg.Engine(0)
    t = info[0][0]
    ref = oracle.decode_track(sectors, t["first_sector"], t["last_sector"], t["pts_length"])
    out = np.zeros(ref["frames"] * ref["channels"] + 1024, dtype=np.int32)
    r = eng.decode_track_pipelined(sectors.ctypes.data, len(sectors)//2048, (t["first_sector"], t["last_sector"], t["pts_length"]), out.ctypes.data, len(out), part_sectors=part)
    print("frames", r.frames, ref["frames"], "stopped", r.stopped, "trunc", r.truncated)
