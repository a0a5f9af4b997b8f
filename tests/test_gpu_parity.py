"""GPU parity tests: the CUDA engine against the CPU oracle and the golden
records of the reference, bit for bit.  Everything goes through the C ABI
(include/dvdagpu.h) or the public API (include/dvd-audio.h)."""
import json
import os

import numpy as np
import pytest

import catalog

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "catalog_golden.json")))
DISCS = sorted(catalog.discs().keys())


@pytest.fixture(scope="module", params=["three-pass", "single-pass"])
def engine(pkg, request):
    """Every test runs against both MLP decode paths: the access-unit-parallel three-pass
    path with the complete decoder as its fall-back (default), and the complete single-pass
    decoder alone (DVDAGPU_SINGLE_PASS=1)."""
    os.environ.pop("DVDAGPU_SINGLE_PASS", None)
    if request.param == "single-pass":
        os.environ["DVDAGPU_SINGLE_PASS"] = "1"
    e = pkg.Engine(0)
    yield e
    e.close()
    os.environ.pop("DVDAGPU_SINGLE_PASS", None)


def check_track(oracle, eng, res, sectors, g, label):
    """One decoded track (TrackResult res) against the oracle run on the same sectors."""
    ref = oracle.decode_track(sectors, g["first"], g["last"], g["pts"])
    if ref is None:
        assert res.status != 0, label
        return
    assert res.status == 0, label
    got = eng.fetch(res)
    assert ("MLP" if res.codec else "PCM") == ref["codec"], label
    assert (res.channels, res.bits_per_sample, res.sample_rate, res.channel_assignment) == \
           (ref["channels"], ref["bits_per_sample"], ref["sample_rate"], ref["assignment"]), label
    assert res.frames == ref["frames"], (label, res.frames, ref["frames"])
    assert res.error_flags == ref["error_flags"], label
    if not np.array_equal(got, ref["pcm"]):
        bad = np.argwhere(got != ref["pcm"])
        raise AssertionError("%s: %d samples differ, first at frame %d channel %d: got %d want %d" % (
            label, len(bad), bad[0][0], bad[0][1], got[tuple(bad[0])], ref["pcm"][tuple(bad[0])]))


@pytest.mark.parametrize("name", DISCS)
def test_tracks_one_by_one(pkg, oracle, engine, disc_cache, name):
    directory, _ = disc_cache(name)
    sectors = oracle.read_aobs(directory)
    for g in GOLDEN[name]["tracks"]:
        res = engine.decode_host(sectors, [(g["first"], g["last"], g["pts"])])
        check_track(oracle, engine, res[0], sectors, g, "%s %d/%d" % (name, g["title"], g["track"]))
        # and against the reference's own hash
        assert res[0].frames == g["frames"]
        assert oracle.fnv1a(engine.fetch(res[0])) == g["fnv"]


@pytest.mark.parametrize("name", DISCS)
def test_whole_titleset_in_one_call(pkg, oracle, engine, disc_cache, name):
    """All tracks of the disc as one batch over one sector buffer (the multi-track path)."""
    directory, _ = disc_cache(name)
    sectors = oracle.read_aobs(directory)
    golden = GOLDEN[name]["tracks"]
    res = engine.decode_host(sectors, [(g["first"], g["last"], g["pts"]) for g in golden])
    for r, g in zip(res, golden):
        assert r.status == 0
        assert r.frames == g["frames"], (name, g["title"], g["track"])
        assert oracle.fnv1a(engine.fetch(r)) == g["fnv"], (name, g["title"], g["track"])


@pytest.mark.parametrize("name", ["mlp_wild_0", "mlp_short_segments", "c5_mixed"])
def test_sync_search_overflow_path(pkg, oracle, disc_cache, name, monkeypatch):
    """With no slots per chunk every chunk that holds a sync pattern is searched again by one
    thread (the path taken when a 2 KiB chunk holds more patterns than it has slots)."""
    monkeypatch.setenv("DVDAGPU_SYNC_SLOTS", "0")
    directory, _ = disc_cache(name)
    sectors = oracle.read_aobs(directory)
    golden = GOLDEN[name]["tracks"]
    eng = pkg.Engine(0)
    try:
        res = eng.decode_host(sectors, [(g["first"], g["last"], g["pts"]) for g in golden])
        for r, g in zip(res, golden):
            assert r.status == 0
            assert r.frames == g["frames"], (name, g["title"], g["track"])
            assert oracle.fnv1a(eng.fetch(r)) == g["fnv"], (name, g["title"], g["track"])
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["c5_mixed", "mlp_wild_0", "pcm_rates_ragged", "mlp_zero_yield", "mlp_short_segments"])
def test_tables_sized_in_advance_grow(pkg, oracle, disc_cache, name, monkeypatch):
    """The packet table and the sync lists are sized before their sizes are known; too small, they grow
    and the step is repeated.  DVDAGPU_SMALL_TABLES=1 starts them with room for one entry."""
    monkeypatch.setenv("DVDAGPU_SMALL_TABLES", "1")
    eng = pkg.Engine(0)
    try:
        directory, _ = disc_cache(name)
        sectors = oracle.read_aobs(directory)
        golden = GOLDEN[name]["tracks"]
        for _ in range(2):                      # (both decodes grow their tables: the hook applies to every decode)
            res = eng.decode_host(sectors, [(g["first"], g["last"], g["pts"]) for g in golden])
            for r, g in zip(res, golden):
                assert r.status == 0 and r.frames == g["frames"], (name, g["title"], g["track"])
                assert oracle.fnv1a(eng.fetch(r)) == g["fnv"], (name, g["title"], g["track"])
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["c5_mixed", "mlp_wild_1", "pcm_rates_ragged", "mlp_zero_yield", "aob_split", "late_start",
                                  "pcm_param_change", "mlp_param_dup"])
def test_public_api(pkg, oracle, disc_cache, name):
    """dvda_open .. dvda_open_track_reader .. dvda_read, as a program written for the
    reference would call it (odd read sizes included)."""
    directory, _ = disc_cache(name)
    d = pkg.Disc(directory)
    for g in GOLDEN[name]["tracks"]:
        info, pcm = d.read_track(g["title"], g["track"], chunk=4096 if g["track"] % 2 else 1013)
        assert (info["codec"], info["channels"], info["bits_per_sample"], info["sample_rate"], info["mask"]) == \
               (g["codec"], g["ch"], g["bps"], g["rate"], g["mask"])
        assert len(pcm) == g["frames"]
        assert oracle.fnv1a(pcm) == g["fnv"]
    d.close()


@pytest.mark.parametrize("name,part", [("c5_mixed", 16), ("mlp_wild_0", 16), ("mlp_wild_1", 24), ("mlp_fir_carry", 16),
                                       ("mlp_zero_yield", 16), ("pcm_rates_ragged", 16), ("pcm_layouts", 16),
                                       ("c1_large", 300), ("c2_large", 256), ("c3_large", 500), ("mlp_short_segments", 16),
                                       ("aob_split", 16), ("late_start", 16), ("pcm_param_change", 16), ("mlp_param_dup", 16)])
def test_public_api_reads_long_tracks_in_parts(pkg, oracle, disc_cache, name, part, monkeypatch):
    """The track reader cuts a track into parts, decoded ahead of dvda_read() by the pool of engine
    contexts and gathered in order (small parts here, so that every catalog track is "long"): PCM
    windows with the frame budget carried over, MLP parts cut at restart points, a part that needs its
    predecessor's filter history decoded together with it."""
    monkeypatch.setenv("DVDA_B200_PART_SECTORS", str(part))
    directory, _ = disc_cache(name)
    d = pkg.Disc(directory)
    try:
        for g in GOLDEN[name]["tracks"]:
            info, pcm = d.read_track(g["title"], g["track"], chunk=4096 if g["track"] % 2 else 1013)
            assert (info["codec"], info["channels"], info["bits_per_sample"], info["sample_rate"], info["mask"]) == \
                   (g["codec"], g["ch"], g["bps"], g["rate"], g["mask"])
            assert len(pcm) == g["frames"], (name, g["title"], g["track"], len(pcm), g["frames"])
            assert oracle.fnv1a(pcm) == g["fnv"], (name, g["title"], g["track"])
    finally:
        d.close()


def test_reader_memory_is_bounded(pkg, oracle, tmp_path):
    """A track of more than a gigabyte read through the public API: the pinned host memory of the
    reader stays under 256 MB (the reference streams with O(1) state, src/dvd-audio.c:751-795), and
    the samples are the reference's."""
    import subprocess
    import dvda_gen as g
    import workloads
    directory = str(tmp_path / "AUDIO_TS")
    g.make_disc(directory, [[workloads.c2_track(96000 * 2200, 4242)]])          # 37 minutes of 24/96 stereo: ~1.12 GB of AOB
    assert sum(os.path.getsize(os.path.join(directory, f)) for f in os.listdir(directory)) > 1 << 30
    ref = subprocess.Popen([oracle.REF_DUMP, directory], stdout=subprocess.PIPE, text=True)
    pkg.host_usage(reset_peak=True)
    live0, _ = pkg.host_usage()
    d = pkg.Disc(directory)
    L = d.L
    ts = L.dvda_open_titleset(d.h, 1)
    ti = L.dvda_open_title(ts, 1)
    tr = L.dvda_open_track(ti, 1)
    rd = L.dvda_open_track_reader(tr)
    assert rd
    import ctypes
    chunk = 1 << 16
    buf = np.empty(chunk * 2, dtype=np.int32)
    frames, h = 0, 0xCBF29CE484222325
    while True:
        got = L.dvda_read(rd, chunk, ctypes.c_void_p(buf.ctypes.data))
        if not got:
            break
        h = g.lib().dvda_gen_fnv1a(ctypes.c_void_p(buf.ctypes.data), got * 8, h)
        frames += got
    _live, peak = pkg.host_usage()
    L.dvda_close_track_reader(rd)
    L.dvda_close_track(tr); L.dvda_close_title(ti); L.dvda_close_titleset(ts)
    d.close()
    assert peak - live0 < 256 << 20, "reader held %d MB of pinned memory" % ((peak - live0) >> 20)
    out = ref.communicate()[0]
    want = oracle.parse_dump_lines(out)[0]
    assert frames == want["frames"] == 96000 * 2200
    assert "%016x" % h == want["fnv"]


def test_readers_on_several_threads(pkg, oracle, disc_cache):
    """Distinct track readers may be used from distinct threads (the reference keeps no
    global state on the read path; here they share one engine behind a lock)."""
    import threading
    name = "c5_mixed"
    directory, _ = disc_cache(name)
    golden = GOLDEN[name]["tracks"]
    failures = []

    def work(tid):
        try:
            d = pkg.Disc(directory)
            for rep in range(2):
                for g in golden[tid::4]:
                    info, pcm = d.read_track(g["title"], g["track"], chunk=2048 + 17 * tid)
                    if len(pcm) != g["frames"] or oracle.fnv1a(pcm) != g["fnv"]:
                        failures.append((tid, g["title"], g["track"]))
            d.close()
        except Exception as e:                      # noqa: BLE001 - reported below
            failures.append((tid, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures, failures


def test_drop_in_dumper(pkg, oracle, disc_cache, tmp_path):
    """The api_dump program, linked against OUR library, prints what it prints when
    linked against the reference (golden records)."""
    directory, _ = disc_cache("c5_mixed")
    rc, tracks, _samples, err = oracle.run_dump(pkg.DUMP_BIN, directory, str(tmp_path / "out.raw"))
    assert rc == 0, err
    assert tracks == GOLDEN["c5_mixed"]["tracks"]


@pytest.mark.parametrize("devices,contexts", [("all", "3"), ("0", "1")])
def test_drop_in_dumper_on_every_device(pkg, oracle, disc_cache, tmp_path, monkeypatch, devices, contexts):
    """The library's own pool of engine contexts over every GPU of the box (DVDA_B200_DEVICES=all;
    one GPU: the same code with one device), parts small enough that every track is dealt out
    across the workers and gathered in order: same records as the reference's."""
    monkeypatch.setenv("DVDA_B200_DEVICES", devices)
    monkeypatch.setenv("DVDA_B200_CONTEXTS", contexts)
    monkeypatch.setenv("DVDA_B200_PART_SECTORS", "32")
    for name in ("c5_mixed", "c3_mlp_6ch96", "aob_split"):
        directory, _ = disc_cache(name)
        rc, tracks, _samples, err = oracle.run_dump(pkg.DUMP_BIN, directory, str(tmp_path / (name + ".raw")))
        assert rc == 0, err
        assert tracks == GOLDEN[name]["tracks"], name


def test_several_title_sets(pkg, oracle, tmp_path):
    """dvda_titleset_count / dvda_open_titleset(n) on a disc with three title sets, every
    track of every set against the reference library through the same dumper."""
    import dvda_gen as g
    directory = str(tmp_path / "AUDIO_TS")
    g.make_disc_multi(directory, [
        [[g.pcm(3000, bps=16, seed=9001), g.mlp(4000, seed=9002, restart_interval=4)]],
        [[g.mlp(5000, seed=9003, assignment=12, substreams=2, matrices=3, features=catalog.RICH, restart_interval=4)],
         [g.pcm(2000, bps=24, assignment=3, seed=9004)]],
        [[g.mlp(3000, rate=192000, seed=9005, restart_interval=8, features=g.CHECKDATA | g.MAX_ORDERS, fir_max=8, iir_max=4)]],
    ])
    disc = pkg.Disc(directory)
    try:
        assert disc.titleset_count() == 3
    finally:
        disc.close()
    for ts in (1, 2, 3):
        rc_r, tracks_r, pcm_r, err_r = oracle.run_dump(oracle.REF_DUMP, directory, str(tmp_path / ("ref%d.raw" % ts)), extra=("-s", str(ts)))
        rc_b, tracks_b, pcm_b, err_b = oracle.run_dump(pkg.DUMP_BIN, directory, str(tmp_path / ("b200_%d.raw" % ts)), extra=("-s", str(ts)))
        assert rc_r == 0, err_r
        assert rc_b == 0, err_b
        assert tracks_b == tracks_r and len(tracks_r) > 0
        assert np.array_equal(pcm_b, pcm_r)
    # a title set that does not exist
    rc_b, _t, _p, _e = oracle.run_dump(pkg.DUMP_BIN, directory, extra=("-s", "4"))
    assert rc_b != 0


@pytest.mark.parametrize("name,offset", [("c2_mlp_2ch96", 20 * 2048 + 1000), ("c3_mlp_6ch96", 31 * 2048 + 700),
                                         ("c3_mlp_6ch96", 40 * 2048 + 1500)])
def test_damage_is_caught_like_the_oracle(pkg, oracle, engine, disc_cache, name, offset):
    directory, _ = disc_cache(name)
    g = GOLDEN[name]["tracks"][0]
    sectors = oracle.read_aobs(directory).copy()
    sectors[offset] ^= 0x04
    res = engine.decode_host(sectors, [(g["first"], g["last"], g["pts"])])
    ref = oracle.decode_track(sectors, g["first"], g["last"], g["pts"])
    assert ref["error_flags"] & (oracle.ERR_PARITY | oracle.ERR_CRC)
    check_track(oracle, engine, res[0], sectors, g, name + " damaged")


@pytest.mark.parametrize("name", ["c2_mlp_2ch96", "c3_mlp_6ch96", "mlp_wild_0", "mlp_wild_1", "c5_mixed", "pcm_rates_ragged"])
def test_random_damage(pkg, oracle, engine, disc_cache, name):
    """Sixty single-bit flips per disc, anywhere in a track's sectors (pack and packet headers,
    access-unit lengths, major syncs, parameters, residuals, check bytes): where the track ends,
    which error is flagged and every sample delivered before it are the oracle's."""
    import random
    directory, _ = disc_cache(name)
    clean = oracle.read_aobs(directory)
    rnd = random.Random(4242 + len(name))
    for _ in range(60):
        g = rnd.choice(GOLDEN[name]["tracks"])
        sectors = clean.copy()
        off = rnd.randrange(g["first"] * 2048, (g["last"] + 1) * 2048)
        bit = 1 << rnd.randrange(8)
        sectors[off] ^= bit
        ref = oracle.decode_track(sectors, g["first"], g["last"], g["pts"])
        r = engine.decode_host(sectors, [(g["first"], g["last"], g["pts"])])[0]
        where = (name, g["track"], off, bit)
        if ref is None:
            assert r.status != 0, where
            continue
        assert r.status == 0, where
        assert (r.frames, r.error_flags, r.channels) == (ref["frames"], ref["error_flags"], ref["channels"]), where
        assert np.array_equal(engine.fetch(r), ref["pcm"]), where


@pytest.mark.parametrize("seed,track,offset,bit", [(581, 3, 40608, 1), (548, 1, 64447, 16)])
def test_damage_found_by_fuzzing(pkg, oracle, engine, tmp_path, seed, track, offset, bit):
    """Two single-bit flips on random discs that tools/fuzz_damage.py found the engine wrong on: a major
    sync whose parameters no longer match (the access unit is dropped) in front of a segment that starts
    with FIR taps — the history has to come from the segment before the dropped one (the unmodified
    reference agrees with the oracle there); and a block size that overruns the tile in the very access
    unit that ends the track with a syntax error (the engine kept asking for a larger tile)."""
    import dvda_gen as g
    tracks = _random_disc(seed)
    directory = str(tmp_path / "AUDIO_TS")
    info = g.make_disc(directory, [tracks])
    sectors = oracle.read_aobs(directory).copy()
    sectors[offset] ^= bit
    t = info[0][track - 1]
    desc = (t["first_sector"], t["last_sector"], t["pts_length"])
    ref = oracle.decode_track(sectors, *desc)
    assert ref is not None
    r = engine.decode_host(sectors, [desc])[0]
    assert r.status == 0
    assert (r.frames, r.error_flags, r.channels) == (ref["frames"], ref["error_flags"], ref["channels"])
    assert np.array_equal(engine.fetch(r), ref["pcm"])


def test_truncated_window_is_reported(pkg, oracle, engine, disc_cache):
    directory, _ = disc_cache("c2_mlp_2ch96")
    g = GOLDEN["c2_mlp_2ch96"]["tracks"][0]
    sectors = oracle.read_aobs(directory)
    part = sectors[: 30 * 2048]
    res = engine.decode_host(part, [(0, g["last"], g["pts"])])
    assert res[0].status == 0 and res[0].truncated == 1
    ref = oracle.decode_track(part, 0, g["last"], g["pts"])
    assert res[0].frames == ref["frames"]
    assert np.array_equal(engine.fetch(res[0]), ref["pcm"])


def test_open_failures(pkg, oracle, engine):
    junk = np.zeros(8 * 2048, np.uint8)
    res = engine.decode_host(junk, [(0, 7, 1000), (100, 200, 1000)])
    assert res[0].status != 0 and res[1].status != 0


@pytest.mark.parametrize("name", sorted(catalog.GPU_LARGE.keys()))
def test_large(pkg, oracle, engine, disc_cache, name):
    directory, _ = disc_cache(name)
    sectors = oracle.read_aobs(directory)
    golden = GOLDEN[name]["tracks"]
    res = engine.decode_host(sectors, [(g["first"], g["last"], g["pts"]) for g in golden])
    for r, g in zip(res, golden):
        check_track(oracle, engine, r, sectors, g, name)
        assert r.frames == g["frames"]


def test_title_set_sharded_like_eight_ranks(pkg, oracle, engine, disc_cache):
    """BASELINE.json configs[4]: the 64-track title set cut into 8 rank shards (shard.py, weights =
    sectors per track).  Every shard is decoded by its own call, as a rank would, from the sector
    window its tracks span; the union must be the reference's output for every track."""
    import importlib
    shard = importlib.import_module("libdvd-audio_b200.shard")
    directory, _ = disc_cache("c5_titleset_64")
    sectors = oracle.read_aobs(directory)
    golden = GOLDEN["c5_titleset_64"]["tracks"]
    assert len(golden) == 64
    shards = shard.shard_tracks([g["last"] - g["first"] + 1 for g in golden], 8)
    seen = []
    for mine in shards:
        assert mine
        res = engine.decode_host(sectors, [(golden[i]["first"], golden[i]["last"], golden[i]["pts"]) for i in mine])
        for r, i in zip(res, mine):
            g = golden[i]
            assert r.status == 0 and r.frames == g["frames"], (i, r.status, r.frames, g["frames"])
            got = engine.fetch(r)
            assert oracle.fnv1a(got) == g["fnv"], "track %d differs from the reference" % i
            seen.append(i)
    assert sorted(seen) == list(range(64))


def test_title_set_sharded_by_units(pkg, oracle, engine, disc_cache):
    """The way bench.py shards configs[4] over ranks: units = tracks and parts of long MLP tracks
    (shard.plan_units), every rank decodes its units from its own compact sector window (the units'
    sector ranges plus a margin, back to back); parts concatenate to their tracks."""
    import importlib
    shard = importlib.import_module("libdvd-audio_b200.shard")
    directory, _ = disc_cache("c5_titleset_64")
    sectors = oracle.read_aobs(directory)
    n_total = len(sectors) // 2048
    golden = GOLDEN["c5_titleset_64"]["tracks"]
    tracks = [(g["first"], g["last"], g["pts"]) for g in golden]
    codecs = [1 if g["codec"] == "MLP" else 0 for g in golden]
    units = shard.plan_units(tracks, codecs, 8, units_per_rank=12, min_part_sectors=16)
    assert any(u["parts"] > 1 for u in units)
    got = {}
    for mine in shard.assign_units(units, 8):
        pieces, descs, at = [], [], 0
        for u in mine:
            stop = min(n_total, u["last"] + 1 + 64)
            pieces.append(sectors[u["first"] * 2048: stop * 2048])
            descs.append((at, at + (u["last"] - u["first"]), u["pts"], u["flags"]))
            at += stop - u["first"]
        res = engine.decode_host(np.concatenate(pieces), descs)
        for u, r in zip(mine, res):
            assert r.status == 0 and r.stopped != 2, (u, r.status, r.stopped)
            got[(u["track"], u["part"])] = engine.fetch(r).reshape(-1) if r.frames else np.zeros(0, np.int32)
    for ti, g in enumerate(golden):
        pcm = np.concatenate([got[(ti, p)] for p in range(64) if (ti, p) in got])
        assert len(pcm) == g["frames"] * g["ch"], (ti, len(pcm), g["frames"])
        assert oracle.fnv1a(pcm) == g["fnv"], "track %d differs from the reference" % ti


def test_two_devices_in_one_process(pkg, oracle, disc_cache):
    """Engines on two GPUs of the box in one process (kernel attributes are per device)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    directory, _ = disc_cache("c5_mixed")
    sectors = oracle.read_aobs(directory)
    golden = GOLDEN["c5_mixed"]["tracks"]
    engines = [pkg.Engine(0), pkg.Engine(1)]
    try:
        for eng in engines + engines[:1]:
            res = eng.decode_host(sectors, [(g["first"], g["last"], g["pts"]) for g in golden])
            for r, g in zip(res, golden):
                assert r.status == 0 and r.frames == g["frames"]
                assert oracle.fnv1a(eng.fetch(r)) == g["fnv"], (g["title"], g["track"])
    finally:
        for eng in engines:
            eng.close()


def test_device_resident_input(pkg, oracle, engine, disc_cache):
    """Sectors already in HBM (a torch tensor), engine on torch's stream."""
    import torch
    directory, _ = disc_cache("c2_large")
    g = GOLDEN["c2_large"]["tracks"][0]
    sectors = oracle.read_aobs(directory)
    dev = torch.from_numpy(sectors).cuda()
    engine.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        res = engine.decode_device(dev.data_ptr(), len(sectors) // 2048, [(g["first"], g["last"], g["pts"])])
        assert res[0].frames == g["frames"]
        ptr, n = engine.pcm_device()
        assert n == g["frames"] * g["ch"]
        assert oracle.fnv1a(engine.fetch(res[0])) == g["fnv"]
        st = engine.stats()
        assert st["launches"] > 10 and st["samples"] == n
    finally:
        engine.set_stream(0)


# ---- long tracks decoded in parts (multi-GPU sharding / pipelined host path) -----------

PART_CASES = [("c2_large", 200), ("c3_large", 300), ("mlp_wild_0", 7), ("mlp_wild_1", 5), ("mlp_wild_2", 9),
              ("mlp_fir_carry", 6), ("mlp_zero_yield", 4), ("mlp_short_segments", 3), ("c4_mlp_2ch192", 11),
              ("mlp_codebooks", 2)]


@pytest.mark.parametrize("name,part", PART_CASES)
def test_parts_concatenate_to_the_track(pkg, oracle, engine, disc_cache, name, part):
    """A track cut into consecutive sector ranges, every range decoded as its own
    "track" with the DVDAGPU_PART_* flags, gives the same samples as the track in
    one piece — unless a part reports that it needs its predecessor's filter
    history (stopped == 2), which is the documented signal to decode in one piece."""
    directory, _ = disc_cache(name)
    sectors = oracle.read_aobs(directory)
    for g in GOLDEN[name]["tracks"]:
        if g["codec"] != "MLP" or g["last"] - g["first"] + 1 < 2 * part:
            continue
        ref = oracle.decode_track(sectors, g["first"], g["last"], g["pts"])
        cuts = list(range(g["first"], g["last"] + 1, part))
        descs = []
        for i, s0 in enumerate(cuts):
            e = g["last"] if i + 1 == len(cuts) else cuts[i + 1] - 1
            flags = (1 if i else 0) | (2 if i + 1 < len(cuts) else 0)
            descs.append((s0, e, g["pts"], flags))
        res = engine.decode_host(sectors, descs)
        if any(r.stopped == 2 for r in res):
            assert name == "mlp_fir_carry" or "wild" in name      # only streams with FIR carried over restarts
            continue
        out = []
        for r in res:
            assert r.status == 0
            if r.frames:                         # parts without a major sync are empty
                out.append(engine.fetch(r))
            if r.stopped == 1:
                break
        got = np.concatenate(out)
        assert len(got) == ref["frames"], (name, g["track"], len(got), ref["frames"])
        assert np.array_equal(got, ref["pcm"]), (name, g["track"])


@pytest.mark.parametrize("name,part", [("c2_large", 256), ("c3_large", 500), ("c1_large", 300), ("mlp_fir_carry", 6),
                                       ("mlp_wild_0", 7), ("mlp_zero_yield", 4), ("c5_mixed", 5),
                                       ("pcm_rates_ragged", 3), ("pcm_layouts", 2), ("pcm_param_change", 2), ("c1_pcm_2ch16", 4)])
def test_pipelined_track_decode(pkg, oracle, engine, disc_cache, name, part):
    """dvdagpu_decode_track_pipelined (overlapped upload / decode / download, with its
    own fallbacks) against the oracle: MLP tracks in parts cut at restart points, PCM tracks in
    windows that carry the frame budget along."""
    directory, _ = disc_cache(name)
    sectors = oracle.read_aobs(directory)
    n_sectors = len(sectors) // 2048
    for g in GOLDEN[name]["tracks"]:
        ref = oracle.decode_track(sectors, g["first"], g["last"], g["pts"])
        out = np.zeros(ref["frames"] * ref["channels"] + 1024, dtype=np.int32)
        r = engine.decode_track_pipelined(sectors.ctypes.data, n_sectors, (g["first"], g["last"], g["pts"]),
                                          out.ctypes.data, len(out), part_sectors=part)
        assert r.status == 0
        assert r.frames == ref["frames"], (name, g["track"], r.frames, ref["frames"])
        assert (r.channels, r.bits_per_sample, r.sample_rate) == (ref["channels"], ref["bits_per_sample"], ref["sample_rate"])
        got = out[: r.frames * r.channels].reshape(-1, r.channels)
        assert np.array_equal(got, ref["pcm"]), (name, g["track"])


_random_disc = catalog.random_disc


@pytest.mark.parametrize("seed", range(40))
def test_random_streams(pkg, oracle, engine, tmp_path, seed):
    """Streams nobody wrote down: per seed a disc of three tracks of random shape, every track
    against the oracle (itself checked against the unmodified reference on the same disc) — in
    one call for the title set and through the pipelined path track by track."""
    import random
    import dvda_gen as g
    rnd = random.Random(99 + seed)
    tracks = _random_disc(seed)
    directory = str(tmp_path / "AUDIO_TS")
    info = g.make_disc(directory, [tracks])
    sectors = oracle.read_aobs(directory)
    n_sectors = len(sectors) // 2048
    descs = [(t["first_sector"], t["last_sector"], t["pts_length"]) for t in info[0]]
    refs = [oracle.decode_track(sectors, *d) for d in descs]
    # (the oracle is pinned on the catalog; on a stream drawn here the unmodified reference has the word)
    rc, ref_tracks, _s, err = oracle.run_dump(oracle.REF_DUMP, directory)
    if rc < 0:
        pytest.skip("the unmodified reference does not survive this stream (signal %d): nothing to compare with" % -rc)
    assert rc == 0, err
    assert len(ref_tracks) == len(refs)
    for rt, ref in zip(ref_tracks, refs):
        assert ref is not None and rt["frames"] == ref["frames"] and rt["fnv"] == oracle.fnv1a(ref["pcm"]), (seed, rt)
    res = engine.decode_host(sectors, descs)
    for i, (r, ref) in enumerate(zip(res, refs)):
        assert (r.status == 0) == (ref is not None), (seed, i, tracks[i])
        if ref is None:
            continue
        assert (r.frames, r.channels, r.bits_per_sample, r.sample_rate, r.error_flags) == \
               (ref["frames"], ref["channels"], ref["bits_per_sample"], ref["sample_rate"], ref["error_flags"]), (seed, i, tracks[i])
        assert np.array_equal(engine.fetch(r), ref["pcm"]), (seed, i, tracks[i])
    for i, (d, ref) in enumerate(zip(descs, refs)):
        if ref is None or not ref["frames"]:
            continue
        out = np.zeros(ref["frames"] * ref["channels"] + 1024, dtype=np.int32)
        r = engine.decode_track_pipelined(sectors.ctypes.data, n_sectors, d, out.ctypes.data, len(out), part_sectors=rnd.choice([3, 5, 9]))
        assert r.status == 0 and r.frames == ref["frames"], (seed, i, r.frames, ref["frames"], tracks[i])
        assert np.array_equal(out[: r.frames * r.channels].reshape(-1, r.channels), ref["pcm"]), (seed, i, tracks[i])


@pytest.mark.parametrize("seed", range(40, 64))
def test_random_streams_through_the_reader(pkg, oracle, tmp_path, monkeypatch, seed):
    """The same kind of disc read through the public API, the tracks cut into small parts that the
    pool of engine contexts decodes ahead of dvda_read(): frame counts and hashes of the unmodified
    reference."""
    import dvda_gen as g
    tracks = _random_disc(seed)
    directory = str(tmp_path / "AUDIO_TS")
    g.make_disc(directory, [tracks])
    rc, ref_tracks, _s, err = oracle.run_dump(oracle.REF_DUMP, directory)
    if rc < 0:
        pytest.skip("the unmodified reference does not survive this stream (signal %d)" % -rc)
    assert rc == 0, err
    monkeypatch.setenv("DVDA_B200_PART_SECTORS", str(3 + seed % 7))
    d = pkg.Disc(directory)
    try:
        for rt in ref_tracks:
            info, pcm = d.read_track(rt["title"], rt["track"], chunk=997 + 31 * (seed % 5))
            assert (info["channels"], info["bits_per_sample"], info["sample_rate"]) == (rt["ch"], rt["bps"], rt["rate"]), (seed, rt)
            assert len(pcm) == rt["frames"], (seed, rt, len(pcm), tracks[rt["track"] - 1])
            assert oracle.fnv1a(pcm) == rt["fnv"], (seed, rt, tracks[rt["track"] - 1])
    finally:
        d.close()


@pytest.mark.parametrize("name", ["c1_pcm_2ch16", "c2_mlp_2ch96", "pcm_layouts"])
def test_dvda2wav_matches_the_reference_tool(pkg, oracle, disc_cache, tmp_path, name):
    """tools/dvda2wav.c on the GPU library writes the same .wav files, byte for byte, as the
    reference's own dvda2wav (unmodified, built into oracle/_ref) — for discs whose
    samples fit their bits per sample (the reference mangles the others, App. B-12)."""
    import subprocess
    ref_tool = os.path.join(oracle.REF_DIR, "dvda2wav")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref/dvda2wav not built")
    directory, _ = disc_cache(name)
    ours, theirs = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir()
    theirs.mkdir()
    a = subprocess.run([pkg.WAV_BIN, "-A", directory, "-d", str(ours)], capture_output=True, text=True)
    b = subprocess.run([ref_tool, "-A", directory, "-d", str(theirs)], capture_output=True, text=True)
    assert a.returncode == 0, a.stderr
    assert b.returncode == 0, b.stderr
    files = sorted(os.listdir(theirs))
    assert files and files == sorted(os.listdir(ours))
    for f in files:
        assert (ours / f).read_bytes() == (theirs / f).read_bytes(), f
    # same progress lines too
    assert [l for l in a.stdout.splitlines() if l.startswith("* Extracting")] == \
           [l for l in b.stdout.splitlines() if l.startswith("* Extracting")]
