"""CPU tests of the host side: the C ABI libraries load and export every
declared symbol, and the disc model (IFO parsing, sector ranges) agrees with the
reference.  No compute call is made here — that needs a GPU."""
import ctypes
import json
import os
import re
import subprocess

import pytest

import catalog

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = json.load(open(os.path.join(HERE, "golden", "catalog_golden.json")))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dvda(?:gpu)?_[a-z_0-9]+)\s*\(", text)))


def test_engine_exports_every_declared_symbol(pkg):
    names = declared_functions("dvdagpu.h")
    assert names == sorted(pkg.ENGINE_SYMBOLS)
    lib = ctypes.CDLL(pkg.ENGINE_LIB)
    for n in names:
        assert hasattr(lib, n), n


def test_api_exports_every_declared_symbol(pkg):
    names = declared_functions("dvd-audio.h")
    assert names == sorted(pkg.API_SYMBOLS)
    assert len(names) == 27                         # the reference's public API
    lib = ctypes.CDLL(pkg.HOST_LIB)
    for n in names:
        assert hasattr(lib, n), n


def test_api_matches_reference_header(pkg):
    """Same function names as the reference's public header, when it is available."""
    ref = "/root/reference/include/dvd-audio.h"
    if not os.path.exists(ref):
        pytest.skip("reference header not present")
    text = re.sub(r"/\*.*?\*/", "", open(ref).read(), flags=re.S)
    ref_names = sorted(set(re.findall(r"\b(dvda_[a-z_0-9]+)\s*\(", text)))
    assert ref_names == declared_functions("dvd-audio.h")


def test_engine_is_a_cuda_library(pkg):
    """The engine carries sm_100a code and no CPU decode path."""
    out = subprocess.run(["cuobjdump", "-lelf", pkg.ENGINE_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("name", ["c5_mixed", "pcm_layouts", "mlp_wild_0", "mlp_short_segments"])
def test_disc_model_matches_reference(pkg, disc_cache, name):
    directory, _ = disc_cache(name)
    d = pkg.Disc(directory)
    assert d.titleset_count() == 1
    mine = d.tracks(1)
    d.close()
    golden = GOLDEN[name]["tracks"]
    assert len(mine) == len(golden)
    for (title, track, info), g in zip(mine, golden):
        assert (title, track) == (g["title"], g["track"])
        assert info["first_sector"] == g["first"]
        assert info["last_sector"] == g["last"]
        assert info["pts_length"] == g["pts"]


def test_open_errors(pkg, tmp_path, disc_cache):
    L = pkg.api_lib()
    assert not L.dvda_open(os.fsencode(str(tmp_path)), None)          # no AUDIO_TS.IFO
    (tmp_path / "AUDIO_TS.IFO").write_bytes(b"NOTADISC" + bytes(200))
    assert not L.dvda_open(os.fsencode(str(tmp_path)), None)          # wrong identifier
    directory, _ = disc_cache("c1_pcm_2ch16")
    h = L.dvda_open(os.fsencode(directory), None)
    assert h
    assert not L.dvda_open_titleset(h, 2)                             # no ATS_02_0.IFO
    ts = L.dvda_open_titleset(h, 1)
    assert ts
    assert not L.dvda_open_title(ts, 0)
    assert not L.dvda_open_title(ts, 99)
    t = L.dvda_open_title(ts, 1)
    assert not L.dvda_open_track(t, 0)
    assert not L.dvda_open_track(t, 50)
    # children outlive parents (reference dvda2wav.c:278-280 closes the track before reading)
    k = L.dvda_open_track(t, 1)
    L.dvda_close_title(t)
    L.dvda_close_titleset(ts)
    L.dvda_close(h)
    assert L.dvda_track_number(k) == 1
    L.dvda_close_track(k)


def test_case_insensitive_lookup(pkg, tmp_path, gen):
    d = str(tmp_path / "disc")
    gen.make_disc(d, [[gen.pcm(500)]])
    for n in os.listdir(d):
        os.rename(os.path.join(d, n), os.path.join(d, n.lower()))
    disc = pkg.Disc(d)
    assert len(disc.tracks(1)) == 1
    disc.close()


def test_no_cpu_fallback(pkg, disc_cache):
    """Without a CUDA device the reader must fail (NULL), never decode on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert pkg.engine_lib().dvdagpu_device_count() == 0
    with pytest.raises(pkg.EngineError):
        pkg.Engine(0)
    directory, _ = disc_cache("c1_pcm_2ch16")
    d = pkg.Disc(directory)
    assert d.read_track(1, 1) is None
    d.close()
