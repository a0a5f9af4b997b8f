"""ctypes front end of the CPU oracle (oracle/dvda_oracle.c) and of the
reference build in oracle/_ref.  TEST INFRASTRUCTURE ONLY — see dvda_oracle.h.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdvda_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_DUMP = os.path.join(REF_DIR, "ref_dump")
REFERENCE_SRC = "/root/reference"

ERR_PARITY, ERR_CRC, ERR_SYNTAX = 1 << 4, 1 << 5, 1 << 6


class Result(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int), ("error_flags", ctypes.c_int), ("codec", ctypes.c_int),
        ("group_0_bps", ctypes.c_uint), ("group_1_bps", ctypes.c_uint),
        ("group_0_rate", ctypes.c_uint), ("group_1_rate", ctypes.c_uint),
        ("channel_assignment", ctypes.c_uint),
        ("channels", ctypes.c_uint), ("bits_per_sample", ctypes.c_uint), ("sample_rate", ctypes.c_uint),
        ("frames", ctypes.c_uint64), ("pcm", ctypes.POINTER(ctypes.c_int32)),
        ("access_units", ctypes.c_uint64), ("es_bytes", ctypes.c_uint64),
    ]


def build(force=False):
    src = [os.path.join(HERE, "dvda_oracle.c"), os.path.join(HERE, "dvda_oracle.h")]
    if (force or not os.path.exists(LIB)
            or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in src)):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB


def build_ref():
    """(Re)build oracle/_ref from /root/reference when the sources are present;
    on the GPU box only the prebuilt files exist.  Returns True if usable."""
    if os.path.isdir(os.path.join(REFERENCE_SRC, "src")):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(REF_DUMP)


def have_ref():
    return os.path.exists(REF_DUMP)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.dvda_oracle_decode_track.restype = ctypes.c_int
        L.dvda_oracle_decode_track.argtypes = [
            ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
            ctypes.POINTER(Result)]
        L.dvda_oracle_free.argtypes = [ctypes.POINTER(Result)]
        L.dvda_oracle_read_bits.restype = ctypes.c_uint32
        L.dvda_oracle_read_bits.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint]
        L.dvda_oracle_read_signed.restype = ctypes.c_int32
        L.dvda_oracle_read_signed.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint]
        L.dvda_oracle_huffman.restype = ctypes.c_int
        L.dvda_oracle_huffman.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_uint)]
        L.dvda_oracle_crc8_table.restype = ctypes.c_uint8
        L.dvda_oracle_crc8_table.argtypes = [ctypes.c_uint]
        L.dvda_oracle_pcm_permutation.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        _lib = L
    return _lib


def read_aobs(audio_ts, titleset=1):
    """The title set's AOB files concatenated, as a uint8 array (aob.c:90-123)."""
    parts = []
    names = {n.upper(): n for n in os.listdir(audio_ts)}
    for i in range(1, 10):
        n = names.get("ATS_%02d_%d.AOB" % (titleset, i))
        if n is None:
            break
        a = np.fromfile(os.path.join(audio_ts, n), dtype=np.uint8)
        parts.append(a[: len(a) // 2048 * 2048])
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def decode_track(sectors, first_sector, last_sector, pts_length):
    """Returns dict(codec, channels, bits_per_sample, sample_rate, frames, pcm[frames, ch],
    error_flags, access_units, es_bytes) or None if the track cannot be opened."""
    sectors = np.ascontiguousarray(sectors, dtype=np.uint8)
    r = Result()
    rc = lib().dvda_oracle_decode_track(sectors.ctypes.data, len(sectors) // 2048,
                                        first_sector, last_sector, pts_length, ctypes.byref(r))
    if rc != 0:
        return None
    n = int(r.frames) * int(r.channels)
    pcm = np.ctypeslib.as_array(r.pcm, shape=(n,)).copy() if n else np.zeros(0, np.int32)
    out = dict(codec="MLP" if r.codec else "PCM", channels=int(r.channels),
               bits_per_sample=int(r.bits_per_sample), sample_rate=int(r.sample_rate),
               assignment=int(r.channel_assignment),
               frames=int(r.frames), pcm=pcm.reshape(-1, max(1, int(r.channels))),
               error_flags=int(r.error_flags), access_units=int(r.access_units),
               es_bytes=int(r.es_bytes))
    lib().dvda_oracle_free(ctypes.byref(r))
    return out


def parse_dump_lines(text):
    """Parses api_dump's 'track ...' lines into dicts."""
    tracks = []
    for line in text.splitlines():
        if not line.startswith("track "):
            continue
        f = line.split()
        d = dict(title=int(f[1]), track=int(f[2]))
        for kv in f[3:]:
            k, v = kv.split("=")
            d[k] = v if k in ("codec", "fnv") else int(v)
        tracks.append(d)
    return tracks


def run_dump(binary, audio_ts, out_path=None, chunk=4096, extra=()):
    """Runs an api_dump binary (reference- or engine-linked) on a disc.
    Returns (returncode, tracks, samples[int32] or None, stderr)."""
    cmd = [binary, audio_ts, "-c", str(chunk)] + list(extra)
    if out_path:
        cmd += ["-o", out_path]
    p = subprocess.run(cmd, capture_output=True, text=True)
    tracks = parse_dump_lines(p.stdout)
    samples = None
    if out_path and os.path.exists(out_path):
        samples = np.fromfile(out_path, dtype=np.int32)
    return p.returncode, tracks, samples, p.stderr


def reference_decode(audio_ts, tmp_path, chunk=4096):
    """All tracks of title set 1 through the UNMODIFIED reference.  Returns
    (tracks, list of [frames, ch] arrays)."""
    out = os.path.join(str(tmp_path), "ref.raw")
    if os.path.exists(out):
        os.remove(out)
    rc, tracks, samples, err = run_dump(REF_DUMP, audio_ts, out, chunk)
    if rc != 0:
        raise RuntimeError("reference failed rc=%d: %s" % (rc, err))
    pcm, k = [], 0
    for t in tracks:
        n = t["frames"] * t["ch"]
        pcm.append(samples[k:k + n].reshape(-1, max(1, t["ch"])))
        k += n
    return tracks, pcm


def fnv1a(samples):
    """64-bit FNV-1a over int32 little-endian bytes, as api_dump prints it."""
    h = 0xCBF29CE484222325
    for b in np.ascontiguousarray(samples, dtype="<i4").tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
