"""Times one long device-resident MLP track decoded as P parts at once, one engine context
(own stream, own buffers) and one host thread per part, against the plain one-context decode.

usage: python tools/time_parallel.py [seconds of audio]
"""
import importlib, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import dvda_gen as g, oracle
pkg = importlib.import_module("libdvd-audio_b200")
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 600
d = "/dev/shm/tpar"; import shutil; shutil.rmtree(d, ignore_errors=True)
info = g.make_disc(d, [[g.mlp(secs * 96000, rate=96000, assignment=1, seed=1002, restart_interval=16, fir_max=4, iir_max=4, noise_bits=13)]])
aob = oracle.read_aobs(d); shutil.rmtree(d)
n = len(aob) // 2048
dev = torch.from_numpy(np.frombuffer(aob, dtype=np.uint8).copy()).cuda()
t = info[0][0]; first, last, pts = t["first_sector"], t["last_sector"], t["pts_length"]

def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

eng = pkg.Engine(0)
r0 = eng.decode_device(dev.data_ptr(), n, [(first, last, pts)])
frames = int(r0[0].frames)
print("one context: %.3f ms (%d frames)" % (timeit(lambda: eng.decode_device(dev.data_ptr(), n, [(first, last, pts)])), frames))
whole = eng.fetch(r0[0])

for P in (2, 3, 4):
    engs = [pkg.Engine(0) for _ in range(P)]
    total = last - first + 1
    ps = (total + P - 1) // P
    jobs = []
    for i in range(P):
        s0 = first + i * ps
        e = min(s0 + ps - 1, last) if i + 1 < P else last
        stop = n if i + 1 == P else min(n, e + 1 + 64)
        flags = (1 if i else 0) | (2 if i + 1 < P else 0)
        jobs.append((s0, stop - s0, (0, e - s0, pts, flags)))
    out = [None] * P
    def work(i):
        s0, ln, desc = jobs[i]
        out[i] = engs[i].decode_device(dev.data_ptr() + s0 * 2048, ln, [desc])
    def run():
        th = [threading.Thread(target=work, args=(i,)) for i in range(P)]
        for x in th: x.start()
        for x in th: x.join()
    ms = timeit(run)
    got = np.concatenate([engs[i].fetch(out[i][0]) for i in range(P)])
    print("%d contexts at once: %.3f ms, frames %s, identical %s" % (
        P, ms, [int(o[0].frames) for o in out], got.shape == whole.shape and bool(np.array_equal(got, whole))))
    # persistent workers (no thread start in the timed region)
    import queue
    qs = [queue.Queue() for _ in range(P)]; done = queue.Queue()
    def loop(i):
        while True:
            if qs[i].get() is None: return
            work(i); done.put(i)
    th = [threading.Thread(target=loop, args=(i,), daemon=True) for i in range(P)]
    for x in th: x.start()
    def run2():
        for q in qs: q.put(1)
        for _ in range(P): done.get()
    print("%d contexts at once, standing threads: %.3f ms" % (P, timeit(run2)))
    for q in qs: q.put(None)
    for e_ in engs: e_.close()
