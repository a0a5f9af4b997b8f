"""libdvd-audio_b200 — B200-native DVD-Audio decode engine.

The product is two shared libraries built by build.py:

  lib/libdvdagpu.so    the sm_100a kernels behind the C ABI of include/dvdagpu.h
  lib/libdvd-audio.so  the C host library with the public API of include/dvd-audio.h

This module is only a thin ctypes binding of both, for tests and bench.py; it
adds no decode logic of its own and refuses to load when the CUDA library has
not been built (there is no CPU path to fall back to).

The directory name carries a hyphen (it mirrors the reference's name), so import
it with importlib.import_module("libdvd-audio_b200") or through the alias module
libdvd_audio_b200.py at the repository root.
"""
import ctypes
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_LIB = _build.ENGINE_LIB
HOST_LIB = _build.HOST_LIB
DUMP_BIN = _build.DUMP_BIN
WAV_BIN = _build.WAV_BIN

ERR_PARITY, ERR_CRC, ERR_SYNTAX = 1 << 4, 1 << 5, 1 << 6

# every symbol include/dvdagpu.h declares
ENGINE_SYMBOLS = [
    "dvdagpu_device_count", "dvdagpu_create", "dvdagpu_destroy", "dvdagpu_set_stream",
    "dvdagpu_decode_host", "dvdagpu_decode_device", "dvdagpu_decode_track_pipelined",
    "dvdagpu_fetch", "dvdagpu_pcm_device",
    "dvdagpu_host_alloc", "dvdagpu_host_free", "dvdagpu_host_usage", "dvdagpu_get_stats", "dvdagpu_set_profiling", "dvdagpu_last_error",
]
# every function include/dvd-audio.h declares
API_SYMBOLS = [
    "dvda_open", "dvda_close", "dvda_titleset_count",
    "dvda_open_titleset", "dvda_close_titleset", "dvda_titleset_number", "dvda_title_count",
    "dvda_open_title", "dvda_close_title", "dvda_title_number", "dvda_track_count", "dvda_title_pts_length",
    "dvda_open_track", "dvda_close_track", "dvda_track_number", "dvda_track_pts_index",
    "dvda_track_pts_length", "dvda_track_first_sector", "dvda_track_last_sector",
    "dvda_open_track_reader", "dvda_close_track_reader", "dvda_codec", "dvda_bits_per_sample",
    "dvda_sample_rate", "dvda_channel_count", "dvda_riff_wave_channel_mask", "dvda_read",
]


class TrackDesc(ctypes.Structure):
    _fields_ = [("first_sector", ctypes.c_uint32), ("last_sector", ctypes.c_uint32),
                ("pts_length", ctypes.c_uint32), ("flags", ctypes.c_uint32)]


class TrackResult(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("error_flags", ctypes.c_int32), ("codec", ctypes.c_int32),
                ("group_0_bps", ctypes.c_uint32), ("group_1_bps", ctypes.c_uint32),
                ("group_0_rate", ctypes.c_uint32), ("group_1_rate", ctypes.c_uint32),
                ("channel_assignment", ctypes.c_uint32),
                ("channels", ctypes.c_uint32), ("bits_per_sample", ctypes.c_uint32),
                ("sample_rate", ctypes.c_uint32), ("truncated", ctypes.c_uint32),
                ("stopped", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("frames", ctypes.c_uint64), ("pcm_offset", ctypes.c_uint64)]


class Stats(ctypes.Structure):
    _fields_ = [("demux_ms", ctypes.c_float), ("index_ms", ctypes.c_float), ("decode_ms", ctypes.c_float),
                ("output_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("launches", ctypes.c_uint32), ("segments", ctypes.c_uint32),
                ("access_units", ctypes.c_uint64), ("es_bytes", ctypes.c_uint64), ("samples", ctypes.c_uint64),
                ("kernel_ms", ctypes.c_float * 16)]

KERNEL_NAMES = ["es_gather", "sync_scan", "au_chase", "checkdata", "mlp_decode", "carry_fix", "rematrix", "pcm_unpack",
                "mlp_segctx", "mlp_entropy", "mlp_filter", "mlp_filter_out", "mlp_au_parse", "mlp_resolve", "mlp_fused"]


_engine = None
_api = None


def engine_lib():
    """The CUDA engine's C ABI.  Raises if the library has not been built."""
    global _engine
    if _engine is None:
        if not os.path.exists(ENGINE_LIB):
            raise RuntimeError(
                "libdvdagpu.so is missing: run __graft_entry__.build(); this engine has no CPU fallback")
        L = ctypes.CDLL(ENGINE_LIB)
        L.dvdagpu_device_count.restype = ctypes.c_int
        L.dvdagpu_create.restype = ctypes.c_void_p
        L.dvdagpu_create.argtypes = [ctypes.c_int]
        L.dvdagpu_destroy.argtypes = [ctypes.c_void_p]
        L.dvdagpu_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        for name in ("dvdagpu_decode_host", "dvdagpu_decode_device"):
            f = getattr(L, name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32,
                          ctypes.POINTER(TrackDesc), ctypes.POINTER(TrackResult)]
        L.dvdagpu_decode_track_pipelined.restype = ctypes.c_int
        L.dvdagpu_decode_track_pipelined.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(TrackDesc), ctypes.c_uint32,
            ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(TrackResult)]
        L.dvdagpu_fetch.restype = ctypes.c_int
        L.dvdagpu_fetch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
        L.dvdagpu_pcm_device.restype = ctypes.c_void_p
        L.dvdagpu_pcm_device.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.dvdagpu_host_alloc.restype = ctypes.c_void_p
        L.dvdagpu_host_alloc.argtypes = [ctypes.c_size_t]
        L.dvdagpu_host_free.argtypes = [ctypes.c_void_p]
        L.dvdagpu_host_usage.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
        L.dvdagpu_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
        L.dvdagpu_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.dvdagpu_last_error.restype = ctypes.c_char_p
        _engine = L
    return _engine


def api_lib():
    """The public dvd-audio.h API (GPU backed)."""
    global _api
    if _api is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError("libdvd-audio.so is missing: run __graft_entry__.build()")
        L = ctypes.CDLL(HOST_LIB)
        P = ctypes.c_void_p
        for name in ("dvda_open",):
            getattr(L, name).restype = P
            getattr(L, name).argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        for name in ("dvda_open_titleset", "dvda_open_title", "dvda_open_track"):
            getattr(L, name).restype = P
            getattr(L, name).argtypes = [P, ctypes.c_uint]
        L.dvda_open_track_reader.restype = P
        L.dvda_open_track_reader.argtypes = [P]
        for name in ("dvda_close", "dvda_close_titleset", "dvda_close_title", "dvda_close_track",
                     "dvda_close_track_reader"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [P]
        for name in ("dvda_titleset_count", "dvda_titleset_number", "dvda_title_count", "dvda_title_number",
                     "dvda_track_count", "dvda_title_pts_length", "dvda_track_number", "dvda_track_pts_index",
                     "dvda_track_pts_length", "dvda_track_first_sector", "dvda_track_last_sector",
                     "dvda_codec", "dvda_bits_per_sample", "dvda_sample_rate", "dvda_channel_count",
                     "dvda_riff_wave_channel_mask"):
            getattr(L, name).restype = ctypes.c_uint
            getattr(L, name).argtypes = [P]
        L.dvda_read.restype = ctypes.c_uint
        L.dvda_read.argtypes = [P, ctypes.c_uint, ctypes.c_void_p]
        _api = L
    return _api


class EngineError(RuntimeError):
    pass


def host_usage(reset_peak=False):
    """(live, peak) bytes of pinned host memory handed out by dvdagpu_host_alloc()."""
    live, peak = ctypes.c_uint64(), ctypes.c_uint64()
    engine_lib().dvdagpu_host_usage(ctypes.byref(live), ctypes.byref(peak), 1 if reset_peak else 0)
    return live.value, peak.value


class Engine:
    """One engine context on one CUDA device (dvdagpu_create)."""

    def __init__(self, device=0):
        self.lib = engine_lib()
        self.ctx = self.lib.dvdagpu_create(device)
        if not self.ctx:
            raise EngineError(self.lib.dvdagpu_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.dvdagpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self.lib.dvdagpu_set_stream(self.ctx, ctypes.c_void_p(cuda_stream))

    def _descs(self, tracks):
        arr = (TrackDesc * len(tracks))()
        for i, t in enumerate(tracks):
            arr[i].first_sector, arr[i].last_sector, arr[i].pts_length = t[0], t[1], t[2]
            arr[i].flags = t[3] if len(t) > 3 else 0
        return arr

    def decode_host(self, sectors, tracks):
        """sectors: uint8 numpy array (n * 2048) or (pointer, n_sectors); tracks:
        [(first_sector, last_sector, pts_length), ...].  Returns the TrackResult array."""
        if isinstance(sectors, tuple):
            ptr, n = sectors
        else:
            sectors = np.ascontiguousarray(sectors, dtype=np.uint8)
            ptr, n = sectors.ctypes.data, len(sectors) // 2048
        descs = self._descs(tracks)
        res = (TrackResult * len(tracks))()
        rc = self.lib.dvdagpu_decode_host(self.ctx, ctypes.c_void_p(ptr), n, len(tracks), descs, res)
        if rc:
            raise EngineError(self.lib.dvdagpu_last_error().decode())
        return res

    def decode_track_pipelined(self, sectors_ptr, n_sectors, track, pcm_ptr, pcm_capacity, part_sectors=0):
        """One track from (pinned) host sectors into (pinned) host PCM with overlapped copies.
        Returns the merged TrackResult."""
        descs = self._descs([track])
        res = TrackResult()
        rc = self.lib.dvdagpu_decode_track_pipelined(self.ctx, ctypes.c_void_p(sectors_ptr), n_sectors, descs,
                                                     part_sectors, ctypes.c_void_p(pcm_ptr), pcm_capacity,
                                                     ctypes.byref(res))
        if rc:
            raise EngineError(self.lib.dvdagpu_last_error().decode())
        return res

    def decode_device(self, device_ptr, n_sectors, tracks):
        descs = self._descs(tracks)
        res = (TrackResult * len(tracks))()
        rc = self.lib.dvdagpu_decode_device(self.ctx, ctypes.c_void_p(device_ptr), n_sectors, len(tracks), descs, res)
        if rc:
            raise EngineError(self.lib.dvdagpu_last_error().decode())
        return res

    def fetch(self, result, out=None):
        """Interleaved samples of one decoded track as an int32 array [frames, channels]."""
        n = int(result.frames) * int(result.channels)
        if out is None:
            out = np.empty(n, dtype=np.int32)
        if n and self.lib.dvdagpu_fetch(self.ctx, result.pcm_offset, n, ctypes.c_void_p(out.ctypes.data)):
            raise EngineError(self.lib.dvdagpu_last_error().decode())
        return out[:n].reshape(-1, max(1, int(result.channels)))

    def fetch_into(self, offset, count, host_ptr):
        if count and self.lib.dvdagpu_fetch(self.ctx, offset, count, ctypes.c_void_p(host_ptr)):
            raise EngineError(self.lib.dvdagpu_last_error().decode())

    def pcm_device(self):
        n = ctypes.c_uint64()
        p = self.lib.dvdagpu_pcm_device(self.ctx, ctypes.byref(n))
        return p, n.value

    def set_profiling(self, on):
        """Per-stage / per-kernel CUDA-event times in stats() for the following decodes."""
        self.lib.dvdagpu_set_profiling(self.ctx, 1 if on else 0)

    def stats(self):
        s = Stats()
        self.lib.dvdagpu_get_stats(self.ctx, ctypes.byref(s))
        d = {f: getattr(s, f) for f, _ in Stats._fields_ if f != "kernel_ms"}
        d["kernel_ms"] = {n: float(s.kernel_ms[i]) for i, n in enumerate(KERNEL_NAMES)}
        return d


class Disc:
    """Python mirror of the dvd-audio.h object model, for tests (same names,
    same 1-based numbering, None where the C API returns NULL)."""

    def __init__(self, audio_ts):
        self.L = api_lib()
        self.h = self.L.dvda_open(os.fsencode(audio_ts), None)
        if not self.h:
            raise FileNotFoundError(audio_ts)

    def close(self):
        if self.h:
            self.L.dvda_close(self.h)
            self.h = None

    def titleset_count(self):
        return self.L.dvda_titleset_count(self.h)

    def tracks(self, titleset=1):
        """[(title, track, dict(first_sector, last_sector, pts_index, pts_length)), ...]"""
        L = self.L
        ts = L.dvda_open_titleset(self.h, titleset)
        if not ts:
            return None
        out = []
        for t in range(1, L.dvda_title_count(ts) + 1):
            title = L.dvda_open_title(ts, t)
            for k in range(1, L.dvda_track_count(title) + 1):
                tr = L.dvda_open_track(title, k)
                out.append((t, k, dict(first_sector=L.dvda_track_first_sector(tr),
                                       last_sector=L.dvda_track_last_sector(tr),
                                       pts_index=L.dvda_track_pts_index(tr),
                                       pts_length=L.dvda_track_pts_length(tr))))
                L.dvda_close_track(tr)
            L.dvda_close_title(title)
        L.dvda_close_titleset(ts)
        return out

    def read_track(self, title, track, titleset=1, chunk=4096):
        """Decodes one track through dvda_open_track_reader / dvda_read.  Returns
        (info dict, int32 array [frames, ch]) or None."""
        L = self.L
        ts = L.dvda_open_titleset(self.h, titleset)
        if not ts:
            return None
        ti = L.dvda_open_title(ts, title)
        tr = L.dvda_open_track(ti, track) if ti else None
        rd = L.dvda_open_track_reader(tr) if tr else None
        result = None
        if rd:
            ch = L.dvda_channel_count(rd)
            info = dict(codec="MLP" if L.dvda_codec(rd) else "PCM", bits_per_sample=L.dvda_bits_per_sample(rd),
                        sample_rate=L.dvda_sample_rate(rd), channels=ch, mask=L.dvda_riff_wave_channel_mask(rd))
            buf = np.empty(chunk * max(ch, 1), dtype=np.int32)
            parts = []
            while True:
                got = L.dvda_read(rd, chunk, ctypes.c_void_p(buf.ctypes.data))
                if not got:
                    break
                parts.append(buf[:got * ch].copy())
            pcm = np.concatenate(parts) if parts else np.zeros(0, np.int32)
            result = (info, pcm.reshape(-1, max(ch, 1)))
            L.dvda_close_track_reader(rd)
        if tr:
            L.dvda_close_track(tr)
        if ti:
            L.dvda_close_title(ti)
        L.dvda_close_titleset(ts)
        return result
