import importlib, os, sys, time
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch, shutil
import dvda_gen as g, oracle
pkg = importlib.import_module("libdvd-audio_b200")
secs = 600
d = "/dev/shm/tp"; shutil.rmtree(d, ignore_errors=True)
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
for env in ({}, {"DVDAGPU_DEBUG_NO_D2H": "1"}):
    for k in ("DVDAGPU_DEBUG_NO_D2H", "DVDAGPU_DEBUG_NO_H2D"): os.environ.pop(k, None)
    os.environ.update(env)
    for part in (0, 26000, 13000):
        def pipe():
            r = eng.decode_track_pipelined(hin.data_ptr(), n, tr, hout.data_ptr(), samples, part_sectors=part)
        ms = timeit(pipe); st = eng.stats()
        print(env, "part=%d: %.2f ms, launches %d, device total %.2f ms" % (part, ms, st["launches"], st["total_ms"]), {k: round(v, 2) for k, v in st.items() if k.endswith("_ms") and k != "kernel_ms"})
for k in ("DVDAGPU_DEBUG_NO_D2H", "DVDAGPU_DEBUG_NO_H2D"): os.environ.pop(k, None)
dev_in = hin.cuda()
for part in (151000, 50000, 26000, 13000):
    ns = min(part, n)
    def dd():
        eng.decode_device(dev_in.data_ptr(), ns + 64 if ns + 64 <= n else n, [(0, ns - 1, tr[2], 2 if ns < n else 0)])
    try:
        ms = timeit(dd, 6); st = eng.stats()
        print("resident part=%d: %.2f ms wall, device total %.2f ms" % (ns, ms, st["total_ms"]), {k: round(v, 2) for k, v in st.items() if k.endswith("_ms") and k != "kernel_ms"}, st["launches"])
    except Exception as e:
        print("resident", part, "failed:", e)
# decode of resident input while unrelated bulk copies run on other streams
import threading
stop = False
big_h = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
big_d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
big_h2 = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
big_d2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def copier(h2d, d2h):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    while not stop:
        if h2d:
            with torch.cuda.stream(s1): big_d.copy_(big_h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): big_h2.copy_(big_d2, non_blocking=True)
        s1.synchronize(); s2.synchronize()
for mode in ((False, False), (True, False), (False, True), (True, True)):
    stop = False
    th = threading.Thread(target=copier, args=mode)
    if any(mode): th.start()
    time.sleep(0.05)
    for part in (151000, 26000):
        ns = min(part, n)
        def dd():
            eng.decode_device(dev_in.data_ptr(), ns + 64 if ns + 64 <= n else n, [(0, ns - 1, tr[2], 2 if ns < n else 0)])
        ms = timeit(dd, 8); st = eng.stats()
        print("copies h2d=%d d2h=%d resident part=%d: %.2f ms wall, device total %.2f ms" % (mode[0], mode[1], ns, ms, st["total_ms"]), {k: round(v, 2) for k, v in st.items() if k.endswith("_ms") and k != "kernel_ms"}, {k: round(v, 3) for k, v in st["kernel_ms"].items() if v > 0.05})
    stop = True
    if any(mode): th.join()
