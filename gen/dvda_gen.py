"""ctypes front end of the synthetic disc generator (gen/dvda_gen.c).

Test / bench infrastructure.  `make_disc(dir, titles)` writes AUDIO_TS.IFO,
ATS_01_0.IFO and the AOB files for a list of titles, each a list of track
dicts (keys = dvda_gen_track_t fields, see dvda_gen.h; helpers pcm()/mlp()
fill the defaults).
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdvda_gen.so")

# feature bits (dvda_gen.h)
CHECKDATA = 1 << 0
BYPASS = 1 << 1
NOISE = 1 << 2
QUANT = 1 << 3
OUTSHIFT = 1 << 4
EXTRAWORD = 1 << 5
TERMINATOR = 1 << 6
FLAGS = 1 << 7
FIR_CARRY = 1 << 8
SPARSE = 1 << 9
MID_RESTART = 1 << 10
SYNC_NO_RST = 1 << 11
SS1_CHK_QUIRK = 1 << 12
MIDAU_PARAMS = 1 << 13
RANDOM_PADS = 1 << 14
PCM_RAGGED = 1 << 15
TWO_PACKETS = 1 << 16
MAX_ORDERS = 1 << 17
SYNC_PARAM_DUP = 1 << 18
PCM_PARAM_CHANGE = 1 << 19

RATE_CODE = {48000: 0, 96000: 1, 192000: 2, 44100: 8, 88200: 9, 176400: 10}
BPS_CODE = {16: 0, 24: 2}
CHANNELS = [1, 2, 3, 4, 3, 4, 5, 3, 4, 5, 4, 5, 6, 4, 5, 4, 5, 6, 5, 5, 6]


class Track(ctypes.Structure):
    _fields_ = [
        ("codec", ctypes.c_int32),
        ("bps_code", ctypes.c_int32),
        ("rate_code", ctypes.c_int32),
        ("assignment", ctypes.c_int32),
        ("frames", ctypes.c_int64),
        ("seed", ctypes.c_uint64),
        ("join_previous", ctypes.c_int32),
        ("features", ctypes.c_int32),
        ("substreams", ctypes.c_int32),
        ("au_frames", ctypes.c_int32),
        ("restart_interval", ctypes.c_int32),
        ("max_blocks", ctypes.c_int32),
        ("fir_max", ctypes.c_int32),
        ("iir_max", ctypes.c_int32),
        ("codebooks", ctypes.c_int32),
        ("matrices", ctypes.c_int32),
        ("noise_bits", ctypes.c_int32),
        ("min_lsbs", ctypes.c_int32),
        ("start_shift", ctypes.c_int32),
    ]


class Info(ctypes.Structure):
    _fields_ = [
        ("first_sector", ctypes.c_uint32),
        ("last_sector", ctypes.c_uint32),
        ("pts_length", ctypes.c_uint32),
        ("channels", ctypes.c_uint32),
        ("frames", ctypes.c_int64),
        ("payload_bytes", ctypes.c_int64),
    ]


def build(force=False):
    src = [os.path.join(HERE, "dvda_gen.c"), os.path.join(HERE, "dvda_gen.h")]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in src)):
        return LIB
    subprocess.check_call(
        ["gcc", "-O2", "-g", "-Wall", "-std=c11", "-fPIC", "-shared", "-o", LIB, src[0], "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.dvda_gen_disc.restype = ctypes.c_int
        _lib.dvda_gen_disc.argtypes = [
            ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int32),
            ctypes.POINTER(Track), ctypes.POINTER(Info), ctypes.c_uint64]
        _lib.dvda_gen_error.restype = ctypes.c_char_p
        _lib.dvda_gen_fnv1a.restype = ctypes.c_uint64
        _lib.dvda_gen_fnv1a.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
    return _lib


def pcm(frames, bps=16, rate=48000, assignment=1, seed=1, features=0, start_shift=0):
    return dict(codec=0, bps_code=BPS_CODE[bps], rate_code=RATE_CODE[rate],
                assignment=assignment, frames=frames, seed=seed, features=features, start_shift=start_shift)


def mlp(frames, bps=24, rate=96000, assignment=1, seed=1, features=CHECKDATA,
        substreams=1, au_frames=0, restart_interval=16, max_blocks=1,
        fir_max=4, iir_max=4, codebooks=0xF, matrices=0, noise_bits=12,
        min_lsbs=0, join_previous=0, start_shift=0):
    return dict(codec=1, bps_code=BPS_CODE[bps], rate_code=RATE_CODE[rate],
                assignment=assignment, frames=frames, seed=seed, features=features,
                substreams=substreams, au_frames=au_frames,
                restart_interval=restart_interval, max_blocks=max_blocks,
                fir_max=fir_max, iir_max=iir_max, codebooks=codebooks,
                matrices=matrices, noise_bits=noise_bits, min_lsbs=min_lsbs,
                join_previous=join_previous, start_shift=start_shift)


def make_disc(directory, titles, max_aob_bytes=0):
    """titles: list of lists of track dicts.  Returns a list (per title) of
    lists of info dicts (first_sector, last_sector, pts_length, channels,
    frames, payload_bytes)."""
    os.makedirs(directory, exist_ok=True)
    flat = [t for title in titles for t in title]
    n = len(flat)
    arr = (Track * n)()
    for i, t in enumerate(flat):
        for k, v in t.items():
            setattr(arr[i], k, v)
    tpt = (ctypes.c_int32 * len(titles))(*[len(t) for t in titles])
    info = (Info * n)()
    rc = lib().dvda_gen_disc(os.fsencode(directory), len(titles), tpt, arr, info, max_aob_bytes)
    if rc != 0:
        raise RuntimeError("dvda_gen: " + lib().dvda_gen_error().decode())
    out, k = [], 0
    for title in titles:
        row = []
        for _ in title:
            row.append({f: getattr(info[k], f) for f, _t in Info._fields_})
            k += 1
        out.append(row)
    return out


def make_disc_multi(directory, titlesets, max_aob_bytes=0):
    """A disc with several title sets (ATS_01 .. ATS_nn): titlesets is a list of
    `titles` arguments of make_disc().  Every title set is generated on its own
    and renamed; AUDIO_TS.IFO gets the count.  Returns the list of make_disc()
    results."""
    import shutil
    import tempfile
    os.makedirs(directory, exist_ok=True)
    out = []
    for n, titles in enumerate(titlesets, start=1):
        tmp = tempfile.mkdtemp(prefix="dvda_ts%02d_" % n, dir=directory)
        try:
            out.append(make_disc(tmp, titles, max_aob_bytes))
            for name in sorted(os.listdir(tmp)):
                if name.startswith("ATS_01_"):
                    shutil.move(os.path.join(tmp, name), os.path.join(directory, "ATS_%02d_%s" % (n, name[7:])))
                elif n == 1 and name == "AUDIO_TS.IFO":
                    shutil.move(os.path.join(tmp, name), os.path.join(directory, name))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    # title set count: one byte of the manager file (reference dvd-audio.c:824-858)
    path = os.path.join(directory, "AUDIO_TS.IFO")
    data = bytearray(open(path, "rb").read())
    data[63] = len(titlesets)
    open(path, "wb").write(bytes(data))
    return out


def fnv1a(samples):
    """64-bit FNV-1a over the little-endian bytes of an int32 array (numpy, or anything with
    the buffer protocol), as oracle/api_dump.c prints it: 16 hex digits."""
    import numpy as np
    a = np.ascontiguousarray(samples)
    if a.dtype != np.dtype("<i4"):
        a = a.astype("<i4")
    h = lib().dvda_gen_fnv1a(ctypes.c_void_p(a.ctypes.data), a.nbytes, 0xCBF29CE484222325)
    return "%016x" % h


def fnv1a_ptr(ptr, nbytes):
    """The same over raw memory (e.g. a pinned torch tensor's data_ptr())."""
    return "%016x" % lib().dvda_gen_fnv1a(ctypes.c_void_p(ptr), nbytes, 0xCBF29CE484222325)
