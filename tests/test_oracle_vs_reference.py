"""Pins the CPU oracle: against the golden records produced by the unmodified
reference (always) and against a live run of the reference build in oracle/_ref
(when it is present).  CPU only."""
import json
import os

import numpy as np
import pytest

import catalog

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "catalog_golden.json")))
# the small catalog plus the 64-track title set of BASELINE.json configs[4] (scaled in length)
CPU_DISCS = sorted(catalog.discs().keys()) + ["c5_titleset_64"]


def oracle_tracks(oracle, directory, golden_tracks):
    sectors = oracle.read_aobs(directory)
    out = []
    for t in golden_tracks:
        out.append(oracle.decode_track(sectors, t["first"], t["last"], t["pts"]))
    return out


@pytest.mark.parametrize("name", CPU_DISCS)
def test_oracle_matches_golden(oracle, disc_cache, name):
    directory, _info = disc_cache(name)
    golden = GOLDEN[name]["tracks"]
    decoded = oracle_tracks(oracle, directory, golden)
    assert len(decoded) == len(golden)
    for g, r in zip(golden, decoded):
        assert r is not None, g
        assert (r["codec"], r["channels"], r["bits_per_sample"], r["sample_rate"], r["frames"]) == \
               (g["codec"], g["ch"], g["bps"], g["rate"], g["frames"]), g
        assert oracle.fnv1a(r["pcm"]) == g["fnv"], g
        assert r["error_flags"] == 0


@pytest.mark.parametrize("name", CPU_DISCS)
def test_oracle_matches_live_reference(oracle, disc_cache, tmp_path, name):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    directory, _info = disc_cache(name)
    tracks, pcm = oracle.reference_decode(directory, tmp_path)
    # the live run must also reproduce the committed golden records
    assert [(t["frames"], t["fnv"]) for t in tracks] == [(t["frames"], t["fnv"]) for t in GOLDEN[name]["tracks"]]
    decoded = oracle_tracks(oracle, directory, tracks)
    for t, p, r in zip(tracks, pcm, decoded):
        assert r is not None
        assert r["frames"] == t["frames"]
        assert np.array_equal(r["pcm"], p), t


@pytest.mark.parametrize("seed", range(64))
def test_oracle_matches_live_reference_on_random_discs(oracle, tmp_path, seed):
    """The discs of the GPU suite's random-stream tests (tests/catalog.py: random_disc): the oracle is
    pinned on them too, by the unmodified reference run on the very disc.  (Three of the 64 shapes —
    joined tracks behind certain predecessors — make the reference segfault: nothing to compare with.)"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import dvda_gen as g
    directory = str(tmp_path / "AUDIO_TS")
    info = g.make_disc(directory, [catalog.random_disc(seed)])
    rc, ref_tracks, _samples, err = oracle.run_dump(oracle.REF_DUMP, directory)
    if rc < 0:
        pytest.skip("the unmodified reference does not survive this stream (signal %d)" % -rc)
    assert rc == 0, err
    sectors = oracle.read_aobs(directory)
    assert len(ref_tracks) == len(info[0])
    for rt, t in zip(ref_tracks, info[0]):
        r = oracle.decode_track(sectors, t["first_sector"], t["last_sector"], t["pts_length"])
        assert r is not None, rt
        assert (r["channels"], r["bits_per_sample"], r["sample_rate"], r["frames"]) == (rt["ch"], rt["bps"], rt["rate"], rt["frames"]), rt
        assert oracle.fnv1a(r["pcm"]) == rt["fnv"], rt
        assert r["error_flags"] == 0


def test_reference_read_size_does_not_matter(oracle, disc_cache, tmp_path):
    """dvda_read(…, n) with odd n returns the same samples (the zero-yield rule does not
    depend on the read pattern)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    directory, _ = disc_cache("mlp_zero_yield")
    a = oracle.reference_decode(directory, tmp_path, chunk=4096)
    b = oracle.reference_decode(directory, tmp_path, chunk=37)
    assert [t["fnv"] for t in a[0]] == [t["fnv"] for t in b[0]]


def test_oracle_flags_damage(oracle, disc_cache):
    """A flipped payload bit is caught by parity/CRC; the track ends in front of the
    damaged access unit (the stock reference prints the message and aborts)."""
    directory, _ = disc_cache("c2_mlp_2ch96")
    g = GOLDEN["c2_mlp_2ch96"]["tracks"][0]
    sectors = oracle.read_aobs(directory).copy()
    good = oracle.decode_track(sectors, g["first"], g["last"], g["pts"])
    # damage a byte in the middle of sector 20's payload
    sectors[20 * 2048 + 1000] ^= 0x10
    bad = oracle.decode_track(sectors, g["first"], g["last"], g["pts"])
    assert bad["error_flags"] & (oracle.ERR_PARITY | oracle.ERR_CRC)
    assert 0 < bad["frames"] < good["frames"]
    assert np.array_equal(bad["pcm"], good["pcm"][:bad["frames"]])


def test_oracle_open_failures(oracle, disc_cache):
    directory, _ = disc_cache("c1_pcm_2ch16")
    sectors = oracle.read_aobs(directory)
    assert oracle.decode_track(sectors, len(sectors) // 2048 + 5, 10 ** 6, 1000) is None   # seek past the end
    junk = np.zeros(4 * 2048, np.uint8)
    assert oracle.decode_track(junk, 0, 3, 1000) is None                                    # no pack header
