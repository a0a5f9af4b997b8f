"""Fuzzing aid (CPU only): single-bit damage on the random discs of the given seeds (five flips per track);
the unmodified reference (oracle/_ref/ref_dump) against the oracle.  Counts: same samples / reference
aborts (assert, signal) while the oracle flags an error / reference aborts while the oracle sees none /
both deliver but differ.

usage: python tools/fuzz_oracle_vs_reference.py seed [seed ...]"""
import os, random, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import dvda_gen as g, oracle, catalog

stats = {"same": 0, "reference aborts, oracle flags an error": 0, "reference aborts, oracle sees no error": 0,
         "reference cannot open": 0, "differ": 0}
for seed in [int(x) for x in sys.argv[1:]]:
    d = tempfile.mkdtemp()
    directory = os.path.join(d, "AUDIO_TS")
    info = g.make_disc(directory, [catalog.random_disc(seed)])
    aob = [f for f in sorted(os.listdir(directory)) if f.endswith(".AOB")]
    path = os.path.join(directory, aob[0])
    clean = open(path, "rb").read()
    rnd = random.Random(seed)
    for i, t in enumerate(info[0]):
        for _ in range(5):
            off = rnd.randrange(t["first_sector"] * 2048, (t["last_sector"] + 1) * 2048)
            bit = 1 << rnd.randrange(8)
            data = bytearray(clean)
            data[off] ^= bit
            open(path, "wb").write(data)
            ref = oracle.decode_track(np.frombuffer(bytes(data), dtype=np.uint8), t["first_sector"], t["last_sector"], t["pts_length"])
            out = os.path.join(d, "ref.raw")
            if os.path.exists(out):
                os.remove(out)
            p = subprocess.run([oracle.REF_DUMP, directory, "-T", "1", "-t", str(i + 1), "-o", out], capture_output=True, text=True)
            if p.returncode < 0:
                stats["reference aborts, oracle flags an error" if ref is not None and ref["error_flags"] else "reference aborts, oracle sees no error"] += 1
                continue
            if p.returncode != 0:
                stats["reference cannot open"] += 1
                continue
            got = np.fromfile(out, dtype=np.int32) if os.path.exists(out) else np.zeros(0, np.int32)
            if ref is not None and len(got) == ref["pcm"].size and np.array_equal(got, ref["pcm"].reshape(-1)):
                stats["same"] += 1
            else:
                stats["differ"] += 1
                print("DIFFER seed", seed, "track", i + 1, "offset", off, "bit", bit, "reference samples", len(got),
                      "oracle", None if ref is None else (ref["frames"], "flags %x" % ref["error_flags"]))
    shutil.rmtree(d)
print(stats)
