"""Fuzzing aid: one damaged random disc again, with the engine's debug dump.
usage: python tools/fuzz_case.py seed track(1-based) offset bit"""
import importlib, os, sys, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import dvda_gen as g, oracle
import catalog
ns = {"_random_disc": catalog.random_disc}
pkg = importlib.import_module("libdvd-audio_b200")
seed, track, off, bit = [int(x) for x in sys.argv[1:5]]
tracks = ns["_random_disc"](seed)
d = tempfile.mkdtemp()
info = g.make_disc(os.path.join(d, "AUDIO_TS"), [tracks])
clean = oracle.read_aobs(os.path.join(d, "AUDIO_TS"))
shutil.rmtree(d)
t = info[0][track - 1]
dsc = (t["first_sector"], t["last_sector"], t["pts_length"])
sectors = clean.copy()
sectors[off] ^= bit
print({k: v for k, v in tracks[track - 1].items() if k != "seed"})
print("sector", off // 2048, "byte", off % 2048, "of track sectors", t["first_sector"], "..", t["last_sector"], "clean byte %02x" % clean[off])
ref0 = oracle.decode_track(clean, *dsc)
ref = oracle.decode_track(sectors, *dsc)
print("oracle clean: frames", ref0["frames"], "aus", ref0["access_units"], "| damaged:", None if ref is None else ("frames", ref["frames"], "flags %x" % ref["error_flags"], "aus", ref["access_units"], "es", ref["es_bytes"]))
eng = pkg.Engine(0)
for mode in ("three-pass", "single-pass"):
    if mode == "single-pass":
        os.environ["DVDAGPU_SINGLE_PASS"] = "1"
    os.environ["DVDAGPU_DEBUG"] = "1"
    try:
        r = eng.decode_host(sectors, [dsc])[0]
        print(mode, ": status", r.status, "frames", r.frames, "flags %x" % r.error_flags, "stopped", r.stopped, "truncated", r.truncated)
        if ref is not None and r.status == 0:
            got = eng.fetch(r)
            n = min(len(got), len(ref["pcm"]))
            same = np.array_equal(got[:n], ref["pcm"][:n])
            print("   common prefix of", n, "frames equal:", same)
            if not same:
                bad = np.argwhere(got[:n] != ref["pcm"][:n])
                fr = np.unique(bad[:, 0])
                runs = np.split(fr, np.where(np.diff(fr) != 1)[0] + 1)
                print("   differing frames:", [(int(x[0]), int(x[-1])) for x in runs[:10]], "channels", sorted(set(bad[:, 1].tolist())))
                f0 = int(fr[0])
                print("   got", got[f0:f0 + 3].tolist(), "ref", ref["pcm"][f0:f0 + 3].tolist())
    except Exception as e:
        print(mode, ": ERROR", e)
    del os.environ["DVDAGPU_DEBUG"]
