"""Times the ways of decoding one long MLP track from host memory on a B200: plain decode +
fetch, the copies alone, and dvdagpu_decode_track_pipelined with several part sizes.

usage: python tools/time_pipe.py [seconds of audio]
"""
import importlib, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import dvda_gen as g, oracle
pkg = importlib.import_module("libdvd-audio_b200")
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 600
d = "/dev/shm/tp"; import shutil; shutil.rmtree(d, ignore_errors=True)
info = g.make_disc(d, [[g.mlp(secs * 96000, rate=96000, assignment=1, seed=1002, restart_interval=16, fir_max=4, iir_max=4, noise_bits=13)]])
aob = oracle.read_aobs(d); shutil.rmtree(d)
n = len(aob) // 2048
hin = torch.empty(len(aob), dtype=torch.uint8, pin_memory=True); hin.numpy()[:] = aob
t = info[0][0]; tr = (t["first_sector"], t["last_sector"], t["pts_length"])
eng = pkg.Engine(0)
res = eng.decode_host((hin.data_ptr(), n), [tr]); samples = int(res[0].frames) * 2
hout = torch.empty(samples, dtype=torch.int32, pin_memory=True)
def timeit(f, reps=4):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def plain():
    r = eng.decode_host((hin.data_ptr(), n), [tr]); eng.fetch_into(r[0].pcm_offset, samples, hout.data_ptr())
print("plain decode_host+fetch: %.2f ms" % timeit(plain))
def h2d_only():
    x = hin.cuda(non_blocking=True); torch.cuda.synchronize()
print("H2D only: %.2f ms" % timeit(h2d_only))
dev = torch.empty(samples, dtype=torch.int32, device="cuda")
def d2h_only():
    hout.copy_(dev, non_blocking=True); torch.cuda.synchronize()
print("D2H only: %.2f ms" % timeit(d2h_only))
for part in (0, 76000, 51000, 38000, 30500, 26000):
    def pipe():
        r = eng.decode_track_pipelined(hin.data_ptr(), n, tr, hout.data_ptr(), samples, part_sectors=part)
        assert int(r.frames) * 2 == samples
    ms = timeit(pipe); st = eng.stats()
    print("pipelined part=%d: %.2f ms, launches %d, device total %.2f ms" % (part, ms, st["launches"], st["total_ms"]))
