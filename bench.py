#!/usr/bin/env python
"""bench.py — decoded PCM samples/s of the DVD-Audio hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c1..c5]

One step = one pass of the whole hot path (AOB sector demux -> PCM unpack | MLP decode ->
interleaved int32 PCM) over one synthetic workload, generated on the box by gen/dvda_gen.c with
fixed seeds (gen/workloads.py holds the shapes of BASELINE.json's configs).

N = 1 (default): the headline workload is configs[1] — a 2-channel 24-bit 96 kHz MLP track, one
  substream, FIR + IIR prediction, 600 s (57.6 M frames, 115.2 M samples, ~310 MB of AOB).  The
  same line carries a `configs` object with the other configurations measured the same way
  (c1 PCM, c3 six channels / two substreams, c4 192 kHz, c5 the 64-track title set).
N > 1 (torchrun, one rank per GPU): configs[4] — ONE title set of 64 mixed PCM / MLP tracks is
  sharded over the ranks (whole tracks, long MLP tracks cut into parts at restart points), every
  rank decodes its shard from its own sector window, the only exchange is the host-side gather of
  the output into one shared buffer.  No collective on the data path; "scaling": "strong".

  value     samples/s with the AOB sectors already resident in HBM (dvdagpu_decode_device), CUDA
            events on the stream the kernels run on, max over ranks.
  e2e       the same metric through the C ABI with HOST buffers: pinned sectors in, every decoded
            sample back in pinned host memory, copies inside the timed region.
  parity    the samples that came back in the e2e leg, FNV-hashed track by track, against the
            hashes the unmodified reference (oracle/_ref/ref_dump) computes for the same disc.
            A mismatch is fatal (exit code 1).
  roofline  for the dominant kernel: algorithmic bytes (AOB bytes + 4 bytes per decoded sample,
            SURVEY.md 8d) / its CUDA-event duration / measured HBM copy bandwidth.
  cpu_baseline  the unmodified reference on one host core, the very disc the GPU decoded.

--impl reference: the reference's CPU decoder on all host cores (rank 0 only), same workload.
"""
import argparse
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "decoded PCM samples/sec"
UNIT = "samples/s"
E2E_CONTEXTS = int(os.environ.get("BENCH_E2E_CONTEXTS", "2"))   # engine contexts per GPU in the end-to-end leg of a job with many tracks
POOL_TRACE = bool(os.environ.get("BENCH_POOL_TRACE"))
E2E_BATCHES = int(os.environ.get("BENCH_E2E_BATCHES", "8"))     # ... which is cut into about this many batches of consecutive tracks
C5_SCALE = 40           # the title set's track lengths: ~80 000 restart segments, 10 000 per GPU of eight


# ------------------------------------------------------------------ helpers

def scratch_base():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()


def scratch_dir(tag):
    d = os.path.join(scratch_base(), "dvda_bench_%s_%d" % (tag, os.getpid()))
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    return d


def read_aobs(audio_ts, titleset=1):
    """The title set's AOB files concatenated, as one uint8 array (file bytes only)."""
    import numpy as np
    names = {n.upper(): n for n in os.listdir(audio_ts)}
    parts = []
    for i in range(1, 10):
        n = names.get("ATS_%02d_%d.AOB" % (titleset, i))
        if n is None:
            break
        a = np.fromfile(os.path.join(audio_ts, n), dtype=np.uint8)
        parts.append(a[: len(a) // 2048 * 2048])
    return np.concatenate(parts) if len(parts) != 1 else parts[0]


class ClockSampler:
    """SM clock and throttle reasons during the timed region: NVML polled from a thread every
    millisecond or so (the timed region is a few tens of milliseconds; `nvidia-smi -lms` is too
    coarse for that and serves only as the fallback)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASON_BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                   (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = False
        self.thread = None
        self.nvml = None

    def _visible_index(self):
        # NVML numbers the physical devices; CUDA_VISIBLE_DEVICES may remap them
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml

            def poll():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, name in self.REASON_BITS:
                            if bits & bit:
                                self.reasons.add(name)
                    except Exception:
                        break
                    time.sleep(0.001)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel, config):
    """dram bytes per launch of a kernel from the committed ncu capture (profiles/traffic.json), if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None
    rows = t.get("kernels") if isinstance(t, dict) and "kernels" in t else [t]
    for r in rows:
        if r.get("kernel") == kernel and r.get("config") == config:
            return r.get("dram_bytes_per_launch")
    return None


# ------------------------------------------------------------------ the reference (CPU) side

def ref_dump_path():
    p = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    return p if os.path.exists(p) else None


def parse_dump(text):
    tracks, samples, elapsed = [], 0, 0.0
    for line in text.splitlines():
        f = line.split()
        if line.startswith("track "):
            d = dict(title=int(f[1]), track=int(f[2]))
            for kv in f[3:]:
                k, v = kv.split("=")
                d[k] = v if k in ("codec", "fnv") else int(v)
            tracks.append(d)
        elif line.startswith("elapsed"):
            elapsed, samples = float(f[1]), int(f[3])
    return tracks, samples, elapsed


class ReferenceJob(threading.Thread):
    """The unmodified reference on a disc, track by track on `procs` host cores at once
    (tracks are the unit of work: api_dump's -T / -t, like dvda2wav's).  hashes=True: every
    track's frame count and FNV hash (the parity side); else decode and discard (timing)."""

    def __init__(self, disc, track_ids, procs, hashes):
        super().__init__(daemon=True)
        self.disc, self.track_ids, self.procs, self.hashes = disc, list(track_ids), max(1, procs), hashes
        self.tracks, self.samples, self.wall, self.error = {}, 0, 0.0, None

    def run(self):
        todo = list(self.track_ids)
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    if not todo or self.error:
                        return
                    t, k = todo.pop(0)
                cmd = [ref_dump_path(), self.disc, "-T", str(t), "-t", str(k)] + ([] if self.hashes else ["-n"])
                p = subprocess.run(cmd, capture_output=True, text=True)
                tr, samples, _el = parse_dump(p.stdout)
                with lock:
                    if p.returncode != 0:
                        self.error = "reference decoder failed on track %d/%d: %s" % (t, k, p.stderr[-200:])
                        return
                    self.samples += samples
                    for d in tr:
                        self.tracks[(d["title"], d["track"])] = d

        t0 = time.perf_counter()
        # longest tracks first would balance better; the list order is the caller's
        ws = [threading.Thread(target=worker, daemon=True) for _ in range(self.procs)]
        for w in ws:
            w.start()
        for w in ws:
            w.join()
        self.wall = time.perf_counter() - t0


def whole_disc_rate(disc, procs, repeat=1):
    """samples/s of `procs` concurrent reference processes, each decoding the WHOLE disc."""
    t0 = time.perf_counter()
    ps = [subprocess.Popen([ref_dump_path(), disc, "-n", "-r", str(repeat)], stdout=subprocess.PIPE, text=True)
          for _ in range(procs)]
    samples = 0
    for p in ps:
        out = p.communicate()[0]
        if p.returncode != 0:
            raise RuntimeError("reference decoder failed")
        samples += parse_dump(out)[1]
    return samples / (time.perf_counter() - t0), samples


def disc_track_ids(pkg, disc_dir):
    d = pkg.Disc(disc_dir)
    try:
        return [(t, k, info) for t, k, info in d.tracks(1)]
    finally:
        d.close()


def run_reference(args, rank):
    """--impl reference: the reference's own CPU decode of the workload on all host cores."""
    import dvda_gen as g
    import workloads
    if rank != 0:
        return
    if not ref_dump_path():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_dump was not built"}))
        return
    cores = os.cpu_count() or 1
    config = args.config or ("c2" if args.gpus == 1 else "c5")
    d = scratch_dir("ref")
    try:
        if config == "c5":
            # the title set itself, tracks dealt to one process per core
            titles, name, rate, ch = workloads.spec("c5", scale=args.scale)
            g.make_disc(d, titles)
            ids = []
            for t, title in enumerate(titles, start=1):
                ids += [(t, k) for k in range(1, len(title) + 1)]
            # longest first: the tail of the step is then one short track
            sizes = {(t, k): titles[t - 1][k - 1]["frames"] * (6 if titles[t - 1][k - 1]["assignment"] == 12 else 2) for t, k in ids}
            ids.sort(key=lambda x: -sizes[x])

            def step():
                job = ReferenceJob(d, ids, cores, hashes=False)
                job.start()
                job.join()
                if job.error:
                    raise RuntimeError(job.error)
                return job.samples
            sample = "the whole title set per step, its 64 tracks dealt to %d processes (one per core), dvda_read to memory" % cores
        else:
            # bounded sample of the workload: 60 s of the same stream shape per process and step
            sample_seconds = min(args.seconds, 60)
            titles, _n, rate, ch = workloads.spec(config, sample_seconds, args.seed)
            name = workloads.NAMES[config] % args.seconds
            g.make_disc(d, titles)

            def step():
                return whole_disc_rate(d, cores)[1]
            sample = "%d processes x %d s of the workload stream per step, dvda_read to memory" % (cores, sample_seconds)
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            total += step()
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(d, ignore_errors=True)
    value = total / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if config != "c5" else "strong", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "x_realtime": value / ch / rate,
        "config": {"workload": name, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ one workload on one GPU

def generate_async(config, seconds, seed, scale, tag):
    """Generates a workload's disc in a child process (the generator is single-threaded C; several
    discs are made side by side).  Returns (Popen, directory)."""
    d = scratch_dir(tag)
    code = ("import sys; sys.path[:0]=[%r, %r]; import dvda_gen as g, workloads; "
            "t, _n, _r, _c = workloads.spec(%r, %r, %r, %r); g.make_disc(%r, t)"
            % (os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), config, seconds, seed, scale, d))
    return subprocess.Popen([sys.executable, "-c", code]), d


class ContextPool:
    """Several engine contexts on one GPU, a host thread each (the C-ABI calls release the GIL), for
    the end-to-end leg of a job with many tracks: the job is cut into batches of consecutive tracks
    (or parts), and while one context's samples travel to the host the next batch is uploaded and
    decoded on another — what the host library's own pool does behind dvda_read()
    (DVDA_B200_CONTEXTS), here through the C ABI (dvdagpu_decode_host + dvdagpu_fetch)."""

    def __init__(self, pkg, device, first, contexts=2):
        self.engines = [first] + [pkg.Engine(device) for _ in range(max(0, contexts - 1))]

    def run(self, host_ptr, batches, out_ptr):
        """batches: [(first sector, sectors, track descriptors relative to it, [word offset of each track in the
        output])].  Returns (launches, per batch [(frames, channels, status, error_flags)])."""
        nxt, lock, errors = [0], threading.Lock(), []
        self.t_run = time.perf_counter()
        launches = [0] * len(self.engines)
        results = [None] * len(batches)

        def work(ei):
            e = self.engines[ei]
            while not errors:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(batches):
                    return
                s0, n, descs, places = batches[i]
                try:
                    t0 = time.perf_counter()
                    r2 = e.decode_host((host_ptr + s0 * 2048, n), descs)
                    t1 = time.perf_counter()
                    for r, off in zip(r2, places):
                        e.fetch_into(r.pcm_offset, int(r.frames) * int(r.channels), out_ptr + 4 * off)
                    if POOL_TRACE:
                        sys.stderr.write("[pool] ctx %d batch %d: %d sectors, %d tracks, %d launches; decode_host %.2f..%.2f ms, fetched at %.2f ms (%d samples)\n"
                                         % (ei, i, n, len(descs), e.stats()["launches"], (t0 - self.t_run) * 1e3, (t1 - self.t_run) * 1e3,
                                            (time.perf_counter() - self.t_run) * 1e3, sum(int(r.frames) * int(r.channels) for r in r2)))
                    results[i] = [(int(r.frames), int(r.channels), int(r.status), int(r.error_flags)) for r in r2]
                    launches[ei] += e.stats()["launches"]
                except Exception as ex:                       # noqa: BLE001 (reported by the caller)
                    errors.append(ex)
                    return

        ts = [threading.Thread(target=work, args=(k,)) for k in range(len(self.engines))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errors:
            raise errors[0]
        return sum(launches), results

    def close(self):
        for e in self.engines[1:]:
            e.close()


def consecutive_batches(windows, want):
    """Groups consecutive (first sector, sectors) windows into about `want` batches of similar size.
    Returns lists of indices."""
    total = sum(n for _s, n in windows)
    goal = max(1, total // max(1, want))
    out, cur, acc = [], [], 0
    for i, (_s, n) in enumerate(windows):
        cur.append(i)
        acc += n
        if acc >= goal:
            out.append(cur)
            cur, acc = [], 0
    if cur:
        out.append(cur)
    return out


def measure(pkg, torch, eng, stream, disc_dir, config, steps, warmup, dist=None, want_kernels=True):
    """Device-resident step and end-to-end step of one disc on the current GPU.  Returns a dict of
    raw numbers plus the pinned output tensor (for hashing) and the per-track results."""
    ids = disc_track_ids(pkg, disc_dir)
    tracks = [(t["first_sector"], t["last_sector"], t["pts_length"]) for _a, _b, t in ids]
    aob = read_aobs(disc_dir)
    n_sectors = len(aob) // 2048
    host_in = torch.empty(len(aob), dtype=torch.uint8, pin_memory=True)
    host_in.numpy()[:] = aob
    del aob
    dev_in = host_in.cuda(non_blocking=False)

    res = None
    for _ in range(max(warmup, 1)):
        res = eng.decode_device(dev_in.data_ptr(), n_sectors, tracks)
    frames = sum(int(r.frames) for r in res)
    samples = sum(int(r.frames) * int(r.channels) for r in res)
    if any(r.status != 0 or r.error_flags for r in res) or not frames:
        raise SystemExit("decode failed: frames=%d status=%s" % (frames, [(r.status, r.error_flags) for r in res]))
    # every track's place in the output buffer (16-byte aligned, as the engine lays them out)
    offsets, total = [], 0
    for r in res:
        total = (total + 3) & ~3
        offsets.append(total)
        total += int(r.frames) * int(r.channels)
    host_out = torch.empty(total + 64, dtype=torch.int32, pin_memory=True)

    # ---- timed: inputs resident in HBM
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches = {}, 0
    stage_ms = {"demux_ms": 0.0, "index_ms": 0.0, "decode_ms": 0.0, "output_ms": 0.0}
    e0.record(stream)
    for _ in range(steps):
        eng.decode_device(dev_in.data_ptr(), n_sectors, tracks)
        launches += eng.stats()["launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # the same steps once more with the engine's per-kernel events switched on (they cost a step a
    # few hundredths of a millisecond, so the figure above is taken without them): the kernel times
    # behind `roofline` and `kernel_ms_per_step`
    eng.set_profiling(True)
    for _ in range(steps):
        eng.decode_device(dev_in.data_ptr(), n_sectors, tracks)
        st = eng.stats()
        for k, v in st["kernel_ms"].items():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v
        for k in stage_ms:
            stage_ms[k] += st[k]
    eng.set_profiling(False)
    torch.cuda.synchronize()

    # ---- timed: end to end through the C ABI with host buffers
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    single = len(tracks) == 1
    pool, batches, groups = None, [], []
    if not single:
        # batches of consecutive tracks, dealt to two engine contexts
        pool = ContextPool(pkg, torch.cuda.current_device(), eng, E2E_CONTEXTS)
        groups = consecutive_batches([(f, l - f + 1) for f, l, _p in tracks], E2E_BATCHES)
        for idx in groups:
            s0 = min(tracks[i][0] for i in idx)
            end = max(tracks[i][1] for i in idx) + 1
            batches.append((s0, end - s0, [(tracks[i][0] - s0, tracks[i][1] - s0, tracks[i][2]) for i in idx], [offsets[i] for i in idx]))

    def e2e_step():
        if single:
            # one long track: upload / decode / download overlapped part by part
            r = eng.decode_track_pipelined(host_in.data_ptr(), n_sectors, tracks[0], host_out.data_ptr(), total)
            if int(r.frames) * int(r.channels) != samples:
                raise SystemExit("pipelined decode returned %d frames" % r.frames)
        else:
            n_l, got = pool.run(host_in.data_ptr(), batches, host_out.data_ptr())
            for idx, rs in zip(groups, got):
                for i, (fr, chn, status, flags) in zip(idx, rs):
                    if status or flags or fr != int(res[i].frames) or chn != int(res[i].channels):
                        raise SystemExit("batched decode of track %d: %d frames, status %d, flags %x" % (i, fr, status, flags))
            return n_l
        return eng.stats()["launches"]

    for _ in range(max(1, min(warmup, 2))):           # sizes the double buffers of the pipelined path
        e2e_step()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_launches = 0
    t_wall = time.perf_counter()
    e2.record(stream)
    for _ in range(steps):
        e2e_launches += e2e_step()
    e3.record(stream)
    torch.cuda.synchronize()
    # the copies run on the engine's own copy streams: take the larger of the event time on the
    # compute stream and the host wall time (every call returns only when its samples are in host memory)
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t_wall) * 1e3)
    if pool:
        pool.close()
    del dev_in
    return dict(ids=ids, res=res, offsets=offsets, frames=frames, samples=samples, n_sectors=n_sectors,
                ms=ms, ms_e2e=ms_e2e, kernel_ms=kernel_ms, stage_ms=stage_ms, launches=launches,
                e2e_launches=e2e_launches, host_out=host_out, host_in=host_in, single=single)


def check_parity(g, m, ref_tracks):
    """The samples the e2e leg left in pinned host memory against the reference's per-track
    records.  Returns (ok, list of mismatches)."""
    bad = []
    base = m["host_out"].data_ptr()
    for (t, k, _info), r, off in zip(m["ids"], m["res"], m["offsets"]):
        want = ref_tracks.get((t, k))
        n = int(r.frames) * int(r.channels)
        got = g.fnv1a_ptr(base + 4 * off, 4 * n)
        if want is None or want["frames"] != int(r.frames) or want["fnv"] != got:
            bad.append((t, k, int(r.frames), got, want and want["frames"], want and want["fnv"]))
    return not bad, bad


def summarize(m, steps, name, config, peak):
    """The per-config numbers of the JSON line from measure()'s raw ones."""
    aob_bytes = m["n_sectors"] * 2048
    alg_bytes = aob_bytes + 4 * m["samples"]                       # SURVEY.md 8d, per step
    value = m["samples"] * steps / (m["ms"] * 1e-3)
    e2e = m["samples"] * steps / (m["ms_e2e"] * 1e-3)
    return {
        "workload": name, "value": value, "unit": UNIT, "ms_per_step": m["ms"] / steps, "steps": steps,
        "frames": m["frames"], "samples": m["samples"], "aob_bytes": aob_bytes, "tracks": len(m["ids"]),
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": m["ms_e2e"] / steps, "h2d_bytes_per_step": aob_bytes,
                "d2h_bytes_per_step": 4 * m["samples"],
                "path": "dvdagpu_decode_track_pipelined (pinned host in, pinned host out)" if m["single"]
                else "batches of consecutive tracks on %d engine contexts: dvdagpu_decode_host + dvdagpu_fetch per track "
                     "(pinned host in, pinned host out)" % E2E_CONTEXTS},
        "roofline_step": {"achieved": alg_bytes * steps / (m["ms"] * 1e-3) / 1e9, "unit": "GB/s",
                          "frac": alg_bytes * steps / (m["ms"] * 1e-3) / 1e9 / peak,
                          "algorithmic_bytes_per_step": alg_bytes},
        "kernel_ms_per_step": {k: round(v / steps, 5) for k, v in m["kernel_ms"].items() if v > 0},
        "gpu_launches": m["launches"],
    }


def dominant_kernel(kernel_ms):
    # k_checkdata runs beside the chain on a low-priority stream: its events bracket the time it
    # shares the GPU, not a launch duration, so it is listed but not a candidate
    chain = {k: v for k, v in kernel_ms.items() if k != "checkdata"}
    return max(chain, key=lambda k: chain[k]) if chain else "mlp_fused"


# ------------------------------------------------------------------ N = 1

def run_single(args, pkg, torch, g, workloads, local_rank, emit):
    config = args.config or "c2"
    subs = [] if (args.config or args.no_sub_configs) else ["c1", "c3", "c4", "c5"]
    cores = os.cpu_count() or 1
    peak, peak_src = measured_hbm_peak()

    # all discs are generated side by side, before anything is timed
    t_gen = time.perf_counter()
    jobs = {c: generate_async(c, args.seconds, args.seed if c == config else None, args.scale, c) for c in [config] + subs}
    for c, (p, _d) in jobs.items():
        if p.wait() != 0:
            raise SystemExit("generator failed for " + c)
    t_gen = time.perf_counter() - t_gen
    dirs = {c: d for c, (_p, d) in jobs.items()}
    try:
        # the reference's hashes of every disc, in the background on the host cores
        refs = {}
        if ref_dump_path():
            for c in [config] + subs:
                ids = [(t, k) for t, k, _i in disc_track_ids(pkg, dirs[c])]
                refs[c] = ReferenceJob(dirs[c], ids, max(1, min(len(ids), cores // 2)), hashes=True)
            # one at a time, the headline's first (they share the cores with each other, not with the timed GPU work:
            # the GPU legs need one core)
            def chain():
                for c in [config] + subs:
                    refs[c].run()
            ref_thread = threading.Thread(target=chain, daemon=True)
            ref_thread.start()

        eng = pkg.Engine(local_rank)
        stream = torch.cuda.current_stream()
        eng.set_stream(stream.cuda_stream)
        sampler = ClockSampler(local_rank)
        sampler.start()
        m = measure(pkg, torch, eng, stream, dirs[config], config, args.steps, args.warmup)
        clocks = sampler.stop()
        name = workloads.spec(config, args.seconds, args.seed, args.scale)[1]
        _t, _n, rate, ch = workloads.spec(config, 1, args.seed, 1)
        head = summarize(m, args.steps, name, config, peak)

        # roofline of the dominant kernel
        top = dominant_kernel(m["kernel_ms"])
        top_ms = m["kernel_ms"].get(top, 0.0) / args.steps
        alg_bytes = head["roofline_step"]["algorithmic_bytes_per_step"]
        achieved = alg_bytes / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0

        parity, parity_detail = None, {}
        sub_out = {}
        results = {config: m}
        for c in subs:
            steps_c = max(3, args.steps // 4)
            mc = measure(pkg, torch, eng, stream, dirs[c], c, steps_c, max(3, min(args.warmup, 3)))
            results[c] = mc
            sub_out[c] = summarize(mc, steps_c, workloads.spec(c, args.seconds, None, args.scale)[1], c, peak)
            mc["host_in"] = None
        eng.close()

        if refs:
            ref_thread.join()
            parity = True
            for c in [config] + subs:
                if refs[c].error:
                    raise SystemExit(refs[c].error)
                ok, bad = check_parity(g, results[c], refs[c].tracks)
                parity_detail[c] = ok
                if c != config:
                    sub_out[c]["parity"] = ok
                if not ok:
                    parity = False
                    sys.stderr.write("PARITY MISMATCH in %s: %s\n" % (c, bad[:4]))

        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "x_realtime": head["value"] / ch / rate,
            "config": {"workload": name, "frames_per_track": m["frames"] // max(1, len(m["ids"])),
                       "aob_bytes": head["aob_bytes"], "tracks_per_gpu": len(m["ids"]),
                       "l2": "inputs larger than L2 (no flush needed)", "parallelism": "one GPU, no collective"},
            "e2e": dict(head["e2e"], gpu_launches=m["e2e_launches"]),
            "gpu_launches": m["launches"],
            "parity": parity,
            "parity_against": "oracle/_ref/ref_dump (unmodified reference), FNV-1a per track of the samples the e2e leg returned" if refs else "unavailable: oracle/_ref not built",
            "roofline": {"bound": "hbm", "kernel": "k_" + top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": recorded_traffic("k_" + top, config), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": top_ms},
            "roofline_step": dict(head["roofline_step"], note="same algorithmic bytes over the whole device-resident step (all kernels)"),
            "kernel_ms_per_step": head["kernel_ms_per_step"],
            "stage_ms_per_step": {k: v / args.steps for k, v in m["stage_ms"].items()},
            "clocks": clocks, "generator_s": t_gen,
        }
        if sub_out:
            line["configs"] = sub_out
        if not args.no_cpu_baseline and ref_dump_path():
            # the reference on one host core: the very disc the GPU decoded when that is a bounded
            # amount of CPU work (the default 600 s track: about 6 s), else a shorter disc of the same shape
            if config != "c5" and args.seconds <= 1200:
                v, _s = whole_disc_rate(dirs[config], 1)
                sample = "the whole workload (%d s track), one process, dvda_read to memory" % args.seconds
            else:
                ds = scratch_dir("cpu")
                try:
                    if config == "c5":
                        st, _n, _r, _c = workloads.spec("c5", scale=max(1, args.scale // 4))
                        sample = "the title set at a quarter of the lengths, one process, dvda_read to memory"
                    else:
                        st, _n, _r, _c = workloads.spec(config, 1200, args.seed)
                        sample = "1200 s of the same stream shape, one process, dvda_read to memory"
                    g.make_disc(ds, st)
                    v, _s = whole_disc_rate(ds, 1)
                finally:
                    shutil.rmtree(ds, ignore_errors=True)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample}
        emit(line)
        return 0 if parity is not False else 1
    finally:
        for d in dirs.values():
            shutil.rmtree(d, ignore_errors=True)


# ------------------------------------------------------------------ N > 1: one title set over the ranks

def run_sharded(args, pkg, torch, g, workloads, dist, rank, world, local_rank, emit):
    import numpy as np
    shard = importlib.import_module("libdvd-audio_b200.shard")
    config = args.config or "c5"
    job = os.environ.get("MASTER_PORT", "0") + "_" + os.environ.get("TORCHELASTIC_RUN_ID", "x")
    disc = os.path.join(scratch_base(), "dvda_bench_job_" + job)
    out_path = disc + ".pcm"
    peak, peak_src = measured_hbm_peak()
    cores = os.cpu_count() or 1
    t_gen = 0.0
    ref = None
    try:
        if rank == 0:
            shutil.rmtree(disc, ignore_errors=True)
            os.makedirs(disc)
            t_gen = time.perf_counter()
            titles, name, rate, ch = workloads.spec(config, args.seconds, args.seed, args.scale)
            g.make_disc(disc, titles)
            t_gen = time.perf_counter() - t_gen
        else:
            _t, name, rate, ch = workloads.spec(config, 1, args.seed, 1)[0], workloads.spec(config, args.seconds, args.seed, args.scale)[1], 0, 0
            rate, ch = workloads.spec(config, 1, args.seed, 1)[2:]
        dist.barrier()
        ids = disc_track_ids(pkg, disc)
        if rank == 0 and ref_dump_path():
            ref = ReferenceJob(disc, [(t, k) for t, k, _i in ids], max(1, cores - 2 * world), hashes=True)
            ref.start()
        aob = read_aobs(disc)
        n_total = len(aob) // 2048

        eng = pkg.Engine(local_rank)
        stream = torch.cuda.current_stream()
        eng.set_stream(stream.cuda_stream)

        # ---- the plan: which tracks are MLP (those may be cut), units, their ranks
        tracks = [(i["first_sector"], i["last_sector"], i["pts_length"]) for _t, _k, i in ids]
        codecs, weights = [], []
        for first, last, pts in tracks:
            n = min(8, n_total - first)
            r = eng.decode_host(aob[first * 2048:(first + n) * 2048], [(0, min(n - 1, last - first), pts)])
            codecs.append(int(r[0].codec) if r[0].status == 0 else -1)
            weights.append(shard.sector_weight(codecs[-1], int(r[0].channels)))
        units = shard.plan_units(tracks, codecs, world, weights=weights)
        mine = shard.assign_units(units, world)[rank]
        # ---- this rank's sector window: its units' sector ranges (+ margin for the run to the next sync), back to back
        margin = 64
        pieces, descs, at = [], [], 0
        for u in mine:
            first, last = u["first"], u["last"]
            stop = min(n_total, last + 1 + margin)
            pieces.append(aob[first * 2048: stop * 2048])
            descs.append((at, at + (last - first), u["pts"], u["flags"]))
            at += stop - first
        n_sectors = at
        host_in = torch.empty(max(1, n_sectors) * 2048, dtype=torch.uint8, pin_memory=True)
        if pieces:
            host_in.numpy()[: n_sectors * 2048] = np.concatenate(pieces)
        del aob, pieces
        dev_in = host_in.cuda(non_blocking=False)

        # ---- warm-up: also tells every unit's length, from which follow the places in the gathered output
        res = None
        for _ in range(max(args.warmup, 1)):
            res = eng.decode_device(dev_in.data_ptr(), n_sectors, descs) if descs else []
        for u, r in zip(mine, res):
            if r.status != 0 or r.error_flags or r.stopped == 2:
                raise SystemExit("rank %d: unit %r failed: status %d flags %x stopped %d" % (rank, u, r.status, r.error_flags, r.stopped))
            u["frames"], u["channels"] = int(r.frames), int(r.channels)
        gathered = [None] * world
        dist.all_gather_object(gathered, [(u["track"], u["part"], u["frames"], u["channels"]) for u in mine])
        size = {}
        for lst in gathered:
            for tr, part, fr, chn in lst:
                size[(tr, part)] = fr * chn
        track_off, unit_off, total = [], {}, 0
        for ti in range(len(tracks)):
            total = (total + 3) & ~3
            track_off.append(total)
            part = 0
            while (ti, part) in size:
                unit_off[(ti, part)] = total
                total += size[(ti, part)]
                part += 1
        track_len = [(track_off[i + 1] if i + 1 < len(tracks) else total) - track_off[i] for i in range(len(tracks))]
        # (the alignment gap of the next track is not part of this one)
        for ti in range(len(tracks)):
            track_len[ti] = sum(size[(ti, p)] for p in range(64) if (ti, p) in size)
        my_samples = sum(u["frames"] * u["channels"] for u in mine)

        # ---- the gathered output: one buffer in shared host memory, page-locked by every rank
        if rank == 0:
            with open(out_path, "wb") as f:
                f.truncate((total + 64) * 4)
        dist.barrier()
        out = np.memmap(out_path, dtype=np.int32, mode="r+")
        rt = torch.cuda.cudart()
        registered = int(rt.cudaHostRegister(out.ctypes.data, out.nbytes, 0)) == 0
        staging = None if registered else torch.empty(max(1, my_samples) + 64, dtype=torch.int32, pin_memory=True)

        # ---- timed: inputs resident in HBM
        dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, kernel_ms = 0, {}
        e0.record(stream)
        for _ in range(args.steps):
            if descs:
                eng.decode_device(dev_in.data_ptr(), n_sectors, descs)
                launches += eng.stats()["launches"]
        e1.record(stream)
        torch.cuda.synchronize()
        # (once more with the engine's per-kernel events on: the kernel times of the line)
        eng.set_profiling(True)
        for _ in range(args.steps):
            if descs:
                eng.decode_device(dev_in.data_ptr(), n_sectors, descs)
                for k, v in eng.stats()["kernel_ms"].items():
                    kernel_ms[k] = kernel_ms.get(k, 0.0) + v
        eng.set_profiling(False)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)

        # ---- timed: end to end — pinned sectors in, every sample into its place of the gathered buffer
        # batches of consecutive units, dealt to two engine contexts: upload, decode and download overlap
        pool = ContextPool(pkg, local_rank, eng, E2E_CONTEXTS)
        windows = [(d[0], (descs[i + 1][0] if i + 1 < len(descs) else n_sectors) - d[0]) for i, d in enumerate(descs)]
        groups = consecutive_batches(windows, E2E_BATCHES)
        stage_off, at_s = [], 0
        for u in mine:
            stage_off.append(at_s)
            at_s += u["frames"] * u["channels"]
        batches = []
        for idx in groups:
            s0 = windows[idx[0]][0]
            n = windows[idx[-1]][0] + windows[idx[-1]][1] - s0
            places = [unit_off[(mine[i]["track"], mine[i]["part"])] if registered else stage_off[i] for i in idx]
            batches.append((s0, n, [(descs[i][0] - s0, descs[i][1] - s0, descs[i][2], descs[i][3]) for i in idx], places))
        e2e_out_ptr = out.ctypes.data if registered else staging.data_ptr()

        def e2e_step():
            if not descs:
                return 0
            n_l, got = pool.run(host_in.data_ptr(), batches, e2e_out_ptr)
            for idx, rs in zip(groups, got):
                for i, (fr, chn, status, flags) in zip(idx, rs):
                    if status or flags or fr != mine[i]["frames"] or chn != mine[i]["channels"]:
                        raise SystemExit("rank %d: batched decode of unit %r: %d frames, status %d, flags %x" % (rank, mine[i], fr, status, flags))
            return n_l

        e2e_step()
        dist.barrier()
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        e2e_launches = 0
        for _ in range(args.steps):
            e2e_launches += e2e_step()
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t_wall) * 1e3
        pool.close()
        if not registered:
            at_s = 0
            for u in mine:
                n = u["frames"] * u["channels"]
                o = unit_off[(u["track"], u["part"])]
                out[o:o + n] = staging.numpy()[at_s:at_s + n]
                at_s += n
        out.flush()

        # ---- the floor of the end-to-end leg on this box: the same bytes as plain pinned copies, all ranks at once
        dist.barrier()
        torch.cuda.synchronize()
        dev_out = torch.empty(max(1, my_samples), dtype=torch.int32, device="cuda")
        pin_out = staging if staging is not None else torch.empty(max(1, my_samples), dtype=torch.int32, pin_memory=True)
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        t_wall = time.perf_counter()
        for _ in range(args.steps):
            with torch.cuda.stream(s_up):
                dev_in.copy_(host_in, non_blocking=True)
            with torch.cuda.stream(s_down):
                pin_out[: my_samples].copy_(dev_out[: my_samples], non_blocking=True)
        torch.cuda.synchronize()
        ms_floor = (time.perf_counter() - t_wall) * 1e3
        clocks = sampler.stop()

        rank_ms = [None] * world
        dist.all_gather_object(rank_ms, (round(ms / args.steps, 4), round(ms_e2e / args.steps, 3)))
        ms, total_samples = shard.reduce_job(ms, my_samples, dist, "cuda")            # MAX time, SUM samples
        ms_e2e, _ = shard.reduce_job(ms_e2e, my_samples, dist, "cuda")
        ms_floor, _ = shard.reduce_job(ms_floor, my_samples, dist, "cuda")
        in_bytes = torch.tensor([float(n_sectors * 2048)], dtype=torch.float64, device="cuda")
        dist.all_reduce(in_bytes)
        loads = [None] * world
        dist.all_gather_object(loads, (n_sectors, my_samples, len(mine), launches))
        dist.barrier()

        rc = 0
        if rank == 0:
            # ---- the same title set on ONE GPU, measured in the same job (the ranks above stay idle meanwhile)
            n1 = None
            if not args.no_n1:
                m1 = measure(pkg, torch, eng, stream, disc, config, max(3, args.steps // 2), 3)
                n1 = {"value": m1["samples"] * max(3, args.steps // 2) / (m1["ms"] * 1e-3), "ms_per_step": m1["ms"] / max(3, args.steps // 2),
                      "e2e_value": m1["samples"] * max(3, args.steps // 2) / (m1["ms_e2e"] * 1e-3),
                      "e2e_ms_per_step": m1["ms_e2e"] / max(3, args.steps // 2),
                      "note": "the whole title set on rank 0's GPU alone, same job, same build"}
                del m1
            parity = None
            if ref is not None:
                ref.join()
                if ref.error:
                    raise SystemExit(ref.error)
                parity, bad = True, []
                for (t, k, _i), off, n in zip(ids, track_off, track_len):
                    want = ref.tracks.get((t, k))
                    got = g.fnv1a_ptr(out.ctypes.data + 4 * off, 4 * n)
                    chn = want["ch"] if want else 1
                    if want is None or want["frames"] * chn != n or want["fnv"] != got:
                        parity = False
                        bad.append((t, k, n, got, want and want["frames"], want and want["fnv"]))
                if not parity:
                    sys.stderr.write("PARITY MISMATCH: %s\n" % (bad[:4],))
                    rc = 1
            aob_bytes = float(in_bytes[0])
            alg_bytes = n_total * 2048 + 4 * total_samples
            value = total_samples * args.steps / (ms * 1e-3)
            e2e = total_samples * args.steps / (ms_e2e * 1e-3)
            top = dominant_kernel(kernel_ms)
            top_ms = kernel_ms.get(top, 0.0) / args.steps
            my_alg = n_sectors * 2048 + 4 * my_samples
            achieved = my_alg / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "x_realtime": value / ch / rate,
                "config": {"workload": name, "tracks": len(tracks), "units": len(units),
                           "units_per_rank": [l[2] for l in loads], "sectors_per_rank": [l[0] for l in loads],
                           "samples_per_rank": [l[1] for l in loads], "step_ms_per_rank": [r[0] for r in rank_ms],
                           "e2e_ms_per_rank": [r[1] for r in rank_ms], "aob_bytes": n_total * 2048,
                           "l2": "inputs larger than L2 (no flush needed)",
                           "parallelism": "one title set sharded over %d ranks by track and by parts of long MLP tracks "
                                          "(cut at restart points); host-side gather of the output, no collective" % world},
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(aob_bytes),
                        "d2h_bytes_per_step": int(4 * total_samples), "gpu_launches": e2e_launches,
                        "path": "per rank: batches of consecutive units on %d engine contexts, dvdagpu_decode_host from pinned sectors + "
                                "dvdagpu_fetch of every unit into its place of the gathered buffer (%s)" % (E2E_CONTEXTS, "shared host memory page-locked by every rank" if registered else "through a pinned staging buffer"),
                        "copy_floor_ms_per_step": ms_floor / args.steps,
                        "copy_floor_note": "the same bytes as plain pinned copies (H2D and D2H side by side) on all ranks at once"},
                "gpu_launches": sum(l[3] for l in loads),
                "parity": parity,
                "parity_against": "oracle/_ref/ref_dump (unmodified reference), FNV-1a per track of the gathered output",
                "roofline": {"bound": "hbm", "kernel": "k_" + top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": my_alg, "kernel_ms": top_ms, "note": "rank 0's shard"},
                "roofline_step": {"achieved": alg_bytes * args.steps / (ms * 1e-3) / 1e9 / world, "unit": "GB/s",
                                  "frac": alg_bytes * args.steps / (ms * 1e-3) / 1e9 / world / peak,
                                  "note": "algorithmic bytes of the title set over the device-resident step, per GPU"},
                "kernel_ms_per_step": {k: round(v / args.steps, 5) for k, v in kernel_ms.items() if v > 0},
                "clocks": clocks, "generator_s": t_gen,
            }
            if n1:
                line["single_gpu"] = n1
            emit(line)
        dist.barrier()
        if registered:
            rt.cudaHostUnregister(out.ctypes.data)
        eng.close()
        return rc
    finally:
        try:
            dist.barrier()
        except Exception:
            pass
        if rank == 0:
            shutil.rmtree(disc, ignore_errors=True)
            try:
                os.remove(out_path)
            except OSError:
                pass


# ------------------------------------------------------------------ N > 1, one track per rank (configs c1 .. c4 on request)

def run_replicas(args, pkg, torch, g, workloads, dist, rank, world, local_rank, emit):
    shard = importlib.import_module("libdvd-audio_b200.shard")
    config = args.config
    peak, peak_src = measured_hbm_peak()
    titles, name, rate, ch = workloads.spec(config, args.seconds, args.seed + rank, args.scale)
    d = scratch_dir("r%d" % rank)
    try:
        g.make_disc(d, titles)
        eng = pkg.Engine(local_rank)
        stream = torch.cuda.current_stream()
        eng.set_stream(stream.cuda_stream)
        sampler = ClockSampler(local_rank)
        sampler.start()
        m = measure(pkg, torch, eng, stream, d, config, args.steps, args.warmup, dist)
        clocks = sampler.stop()
        ms, total_samples = shard.reduce_job(m["ms"], m["samples"], dist, "cuda")
        ms_e2e, _ = shard.reduce_job(m["ms_e2e"], m["samples"], dist, "cuda")
        if rank == 0:
            head = summarize(m, args.steps, name, config, peak)
            value = total_samples * args.steps / (ms * 1e-3)
            top = dominant_kernel(m["kernel_ms"])
            top_ms = m["kernel_ms"].get(top, 0.0) / args.steps
            alg = head["roofline_step"]["algorithmic_bytes_per_step"]
            achieved = alg / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
            emit({
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic", "x_realtime": value / ch / rate,
                "config": {"workload": name, "tracks_per_gpu": len(m["ids"]), "l2": "inputs larger than L2 (no flush needed)",
                           "parallelism": "one track per GPU, no collective"},
                "e2e": {"value": total_samples * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": head["aob_bytes"] * world, "d2h_bytes_per_step": 4 * m["samples"] * world,
                        "path": head["e2e"]["path"], "gpu_launches": m["e2e_launches"]},
                "gpu_launches": m["launches"] * world,
                "roofline": {"bound": "hbm", "kernel": "k_" + top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": recorded_traffic("k_" + top, config), "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg, "kernel_ms": top_ms},
                "kernel_ms_per_step": head["kernel_ms_per_step"], "clocks": clocks,
            })
        eng.close()
        return 0
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=["c1", "c2", "c3", "c4", "c5"],
                    help="default: c2 (+ the others as the line's `configs` object) on one GPU, c5 sharded on several")
    ap.add_argument("--seconds", type=int, default=600, help="length of the synthetic track (c1 .. c4)")
    ap.add_argument("--scale", type=int, default=C5_SCALE, help="track length multiplier of the c5 title set")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-configs", action="store_true", help="N = 1: the headline workload only")
    ap.add_argument("--no-n1", action="store_true", help="N > 1: skip the single-GPU measurement of the same title set")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.seed is None:
            args.seed = None
        run_reference(args, rank)
        return 0

    import torch
    import dvda_gen as g
    import workloads
    pkg = importlib.import_module("libdvd-audio_b200")
    g.build()

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one line (rank 0's JSON): whatever libraries print while the job
    # runs (NCCL's version banner, for one) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if world == 1:
        return run_single(args, pkg, torch, g, workloads, local_rank, emit)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if (args.config or "c5") == "c5":
            return run_sharded(args, pkg, torch, g, workloads, dist, rank, world, local_rank, emit)
        if args.seed is None:
            args.seed = {"c1": 1001, "c2": 1002, "c3": 1003, "c4": 1004}[args.config]
        return run_replicas(args, pkg, torch, g, workloads, dist, rank, world, local_rank, emit)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
