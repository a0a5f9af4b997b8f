"""Prints the time line of one warm decode of the bench track (DVDAGPU_TRACE=1): for every
launch the host time of its enqueue and the device time at which it was done.

usage: python tools/trace_decode.py [seconds of audio] [config c1..c4]
"""
import importlib, os, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import dvda_gen as g, oracle
import workloads, bench
pkg = importlib.import_module("libdvd-audio_b200")
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 600
config = sys.argv[2] if len(sys.argv) > 2 else "c2"
titles, name, rate, ch = workloads.spec(config, secs)
d = "/dev/shm/trace_disc"; shutil.rmtree(d, ignore_errors=True)
info = g.make_disc(d, titles)
aob = oracle.read_aobs(d); shutil.rmtree(d)
n = len(aob) // 2048
dev = torch.from_numpy(np.frombuffer(aob, dtype=np.uint8).copy()).cuda()
t = info[0][0]; tr = [(t["first_sector"], t["last_sector"], t["pts_length"])]
eng = pkg.Engine(0)
for _ in range(3):
    eng.decode_device(dev.data_ptr(), n, tr)
os.environ["DVDAGPU_TRACE"] = "1"
eng.decode_device(dev.data_ptr(), n, tr)
del os.environ["DVDAGPU_TRACE"]
st = eng.stats()
eng.set_profiling(True); eng.decode_device(dev.data_ptr(), n, tr); st = eng.stats()
print("total %.3f ms, launches %d" % (st["total_ms"], st["launches"]), file=sys.stderr)
