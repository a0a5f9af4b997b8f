#!/usr/bin/env python
"""bench.py — decoded PCM samples/s of the DVD-Audio hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one pass of the whole hot path (AOB sector demux -> MLP decode ->
interleaved int32 PCM) over one synthetic track.  The workload is
BASELINE.json's configs[1]: a 2-channel 24-bit 96 kHz MLP track, one substream,
FIR + IIR prediction, 600 s (57.6 M frames, 115.2 M samples, ~310 MB of AOB),
generated on the box by gen/dvda_gen.c with a fixed seed.

  value   samples/s with the AOB sectors already resident in HBM
          (dvdagpu_decode_device), timed with CUDA events on the stream the
          kernels run on, max over ranks.
  e2e     the same metric through the C ABI with HOST buffers:
          dvdagpu_decode_host from pinned memory + dvdagpu_fetch of every
          sample into pinned memory, copies inside the timed region.
  roofline  for the dominant kernel (k_mlp_decode): algorithmic bytes
          (AOB bytes of the track + 4 bytes per decoded sample)
          / its CUDA-event duration / measured HBM copy bandwidth.
  cpu_baseline  the unmodified reference (oracle/_ref/ref_dump = reference
          library behind our raw dumper) on one host core, the very disc the GPU decoded.

N > 1 (torchrun): every rank decodes its own track of the same shape on its own
GPU — tracks shard with no data-path collective (weak scaling).
--impl reference: the reference's CPU decoder on all host cores (rank 0 only).
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


def workload_spec(g, config, seconds, seed):
    if config == "c2":
        rate, ch = 96000, 2
        tracks = [g.mlp(int(seconds * rate), rate=rate, assignment=1, seed=seed, restart_interval=16,
                        fir_max=4, iir_max=4, noise_bits=13)]
        name = "2ch 24-bit 96 kHz MLP AOB track, 1 substream, FIR+IIR, %d s" % seconds
    elif config == "c3":
        rate, ch = 96000, 6
        tracks = [g.mlp(int(seconds * rate), rate=rate, assignment=12, substreams=2, seed=seed, restart_interval=16,
                        matrices=3, features=g.CHECKDATA | g.BYPASS | g.NOISE | g.QUANT | g.OUTSHIFT)]
        name = "6ch 24-bit 96 kHz MLP AOB, 2 substreams, rematrix + LSB bypass, %d s" % seconds
    elif config == "c4":
        rate, ch = 192000, 2
        tracks = [g.mlp(int(seconds * rate), rate=rate, assignment=1, seed=seed, restart_interval=8,
                        features=g.CHECKDATA | g.MAX_ORDERS, fir_max=8, iir_max=4, codebooks=0x2, min_lsbs=16,
                        noise_bits=16)]
        name = "2ch 24-bit 192 kHz MLP AOB, max filter orders, %d s" % seconds
    elif config == "c1":
        rate, ch = 48000, 2
        tracks = [g.pcm(int(seconds * rate), bps=16, rate=rate, assignment=1, seed=seed)]
        name = "2ch 16-bit 48 kHz PCM AOB track, %d s" % seconds
    else:
        raise SystemExit("unknown config " + config)
    return [tracks], name, rate, ch


def scratch_dir(tag):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    d = os.path.join(base, "dvda_bench_%s_%d" % (tag, os.getpid()))
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    return d


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


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def reference_rate(ref_dump, disc, procs, repeat=1):
    """samples/s of the unmodified reference: `procs` concurrent processes, each decoding
    the whole disc `repeat` times through dvda_read (no file output)."""
    t0 = time.perf_counter()
    ps = [subprocess.Popen([ref_dump, disc, "-n", "-r", str(repeat)], stdout=subprocess.PIPE, text=True)
          for _ in range(procs)]
    samples = 0
    for p in ps:
        out = p.communicate()[0]
        if p.returncode != 0:
            raise RuntimeError("reference decoder failed")
        for line in out.splitlines():
            if line.startswith("elapsed"):
                samples += int(line.split()[3])
    dt = time.perf_counter() - t0
    return samples / dt, samples, dt


def run_reference(args, rank, world):
    import dvda_gen as g
    import oracle
    if rank != 0:
        return
    if not oracle.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_dump was not built"}))
        return
    cores = os.cpu_count() or 1
    # bounded sample of the workload: 60 s of the same stream shape per process and step
    sample_seconds = min(args.seconds, 60)
    titles, name, rate, ch = workload_spec(g, args.config, sample_seconds, args.seed)
    d = scratch_dir("ref")
    try:
        g.make_disc(d, titles)
        for _ in range(args.warmup):
            reference_rate(oracle.REF_DUMP, d, cores)
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            _r, samples, _dt = reference_rate(oracle.REF_DUMP, d, cores)
            total += samples
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(d, ignore_errors=True)
    value = total / dt
    sample = "%d processes x %d s of the workload stream per step, dvda_read to memory" % (cores, sample_seconds)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "x_realtime": value / ch / rate,
        "config": {"workload": name, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--seconds", type=int, default=600, help="length of the synthetic track")
    ap.add_argument("--seed", type=int, default=1002)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import dvda_gen as g
    import oracle
    pkg = importlib.import_module("libdvd-audio_b200")

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    # stdout carries exactly one line (rank 0's JSON): whatever libraries print while the job
    # runs (NCCL's version banner, for one) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload: generated on the box, one track per rank
    titles, name, rate, ch = workload_spec(g, args.config, args.seconds, args.seed + rank)
    d = scratch_dir("r%d" % rank)
    try:
        t_gen = time.perf_counter()
        g.make_disc(d, titles)
        t_gen = time.perf_counter() - t_gen
        disc = pkg.Disc(d)
        tracks = [(t["first_sector"], t["last_sector"], t["pts_length"]) for _a, _b, t in disc.tracks(1)]
        disc.close()
        aob = oracle.read_aobs(d)            # file bytes only; nothing of oracle/ decodes here
        n_sectors = len(aob) // 2048
        host_in = torch.empty(len(aob), dtype=torch.uint8, pin_memory=True)
        host_in.numpy()[:] = aob
        del aob
        dev_in = host_in.cuda(non_blocking=False)

        eng = pkg.Engine(local_rank)
        stream = torch.cuda.current_stream()
        eng.set_stream(stream.cuda_stream)

        # ---- warm-up (also sizes every device buffer)
        res = None
        for _ in range(max(args.warmup, 1)):
            res = eng.decode_device(dev_in.data_ptr(), n_sectors, tracks)
        frames = sum(int(r.frames) for r in res)
        samples = sum(int(r.frames) * int(r.channels) for r in res)
        if any(r.status != 0 or r.error_flags for r in res) or frames < args.seconds * rate:
            raise SystemExit("decode failed: frames=%d status=%s" % (frames, [(r.status, r.error_flags) for r in res]))
        host_out = torch.empty(samples, dtype=torch.int32, pin_memory=True)

        # ---- timed: inputs resident in HBM
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ms = {}
        launches = 0
        stage_ms = {"demux_ms": 0.0, "index_ms": 0.0, "decode_ms": 0.0, "output_ms": 0.0}
        e0.record(stream)
        for _ in range(args.steps):
            eng.decode_device(dev_in.data_ptr(), n_sectors, tracks)
            st = eng.stats()
            launches += st["launches"]
            for k, v in st["kernel_ms"].items():
                kernel_ms[k] = kernel_ms.get(k, 0.0) + v
            for k in stage_ms:
                stage_ms[k] += st[k]
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)

        # ---- timed: end to end through the C ABI with host buffers
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if len(tracks) == 1:
            for _ in range(max(1, min(args.warmup, 2))):      # sizes the double buffers of the pipelined path
                eng.decode_track_pipelined(host_in.data_ptr(), n_sectors, tracks[0], host_out.data_ptr(), samples)
        e2e_launches = 0
        t_wall = time.perf_counter()
        e2.record(stream)
        for _ in range(args.steps):
            if len(tracks) == 1:
                # one long track: upload / decode / download overlapped part by part
                r = eng.decode_track_pipelined(host_in.data_ptr(), n_sectors, tracks[0],
                                               host_out.data_ptr(), samples)
                if int(r.frames) * int(r.channels) != samples:
                    raise SystemExit("pipelined decode returned %d frames" % r.frames)
                st = eng.stats()
                e2e_launches += st["launches"]
            else:
                r2 = eng.decode_host((host_in.data_ptr(), n_sectors), tracks)
                for r in r2:
                    eng.fetch_into(r.pcm_offset, int(r.frames) * int(r.channels),
                                   host_out.data_ptr() + 4 * int(r.pcm_offset))
        e3.record(stream)
        torch.cuda.synchronize()
        # the copies run on the engine's own copy streams: take the larger of the event time on the
        # compute stream and the host wall time (every call returns only when its samples are in host memory)
        ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t_wall) * 1e3)
        clocks = sampler.stop()                       # sampled across both timed regions (a query takes ~10 ms)
        # the two paths must agree with each other
        check = int(host_out[:: max(1, samples // 65536)].to(torch.int64).sum())

        shard = importlib.import_module("libdvd-audio_b200.shard")
        ms, total_samples = shard.reduce_job(ms, samples, dist, "cuda")            # MAX time, SUM samples
        ms_e2e, _ = shard.reduce_job(ms_e2e, samples, dist, "cuda")

        if rank == 0:
            value = total_samples * args.steps / (ms * 1e-3)
            e2e = total_samples * args.steps / (ms_e2e * 1e-3)
            aob_bytes = n_sectors * 2048
            alg_bytes = aob_bytes + 4 * samples                       # SURVEY.md §8d, per launch
            # the dominant kernel of the step: largest event-timed share among the kernels of the decode
            # chain.  (k_checkdata runs beside the chain on a low-priority stream: its events bracket the
            # time it shares the GPU, not a launch duration, so it is listed but not a candidate.)
            chain = {k: v for k, v in kernel_ms.items() if k != "checkdata"}
            top = max(chain, key=lambda k: chain[k]) if chain else "mlp_decode"
            top_ms = kernel_ms.get(top, 0.0) / args.steps
            peak, peak_src = measured_hbm_peak()
            achieved = alg_bytes / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "x_realtime": value / ch / rate,
                "config": {"workload": name, "frames_per_track": frames, "aob_bytes_per_track": aob_bytes,
                           "tracks_per_gpu": len(tracks), "l2": "inputs larger than L2 (no flush needed)",
                           "parallelism": "one track per GPU, no collective"},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": aob_bytes * world,
                        "d2h_bytes_per_step": 4 * samples * world, "ms_per_step": ms_e2e / args.steps,
                        "path": "dvdagpu_decode_track_pipelined (pinned host in, pinned host out)",
                        "gpu_launches": e2e_launches},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "k_" + top, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": top_ms},
                "roofline_step": {"achieved": alg_bytes * args.steps / (ms * 1e-3) / 1e9, "unit": "GB/s",
                                  "frac": alg_bytes * args.steps / (ms * 1e-3) / 1e9 / peak,
                                  "note": "same algorithmic bytes over the whole device-resident step (all kernels), per GPU"},
                "kernel_ms_per_step": {k: v / args.steps for k, v in kernel_ms.items()},
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "clocks": clocks,
                "generator_s": t_gen, "checksum": check,
            }
            tr = recorded_traffic()
            if tr and tr.get("kernel") == "k_" + top and tr.get("config") == args.config:
                line["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
            if world == 1 and not args.no_cpu_baseline and oracle.have_ref():
                # the reference on one host core: the very disc the GPU decoded when that is a bounded
                # amount of CPU work (the default 600 s track: about 6 s), else its first 1200 s of stream shape
                if args.seconds <= 1200:
                    v, _s, _dt = reference_rate(oracle.REF_DUMP, d, 1)
                    sample = "the whole workload (%d s track), one process, dvda_read to memory" % args.seconds
                else:
                    ds = scratch_dir("cpu")
                    try:
                        st_titles, _n, _r, _c = workload_spec(g, args.config, 1200, args.seed)
                        g.make_disc(ds, st_titles)
                        v, _s, _dt = reference_rate(oracle.REF_DUMP, ds, 1)
                    finally:
                        shutil.rmtree(ds, ignore_errors=True)
                    sample = "1200 s of the same stream shape, one process, dvda_read to memory"
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample}
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(line), flush=True)
            os.dup2(2, 1)
        eng.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)
        if dist:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
