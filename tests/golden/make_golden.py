"""Regenerates tests/golden/catalog_golden.json from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref, i.e. /root/reference):

    python tests/golden/make_golden.py

For every disc of tests/catalog.py the generator writes the disc, the reference
build (oracle/_ref/ref_dump: the reference library behind our raw-int dumper)
decodes every track, and the per-track facts are recorded: codec, sample
format, frame count, sector range and a 64-bit FNV-1a hash of the int32 samples
dvda_read() returned.  The tests compare the oracle (CPU) and the CUDA engine
(GPU) against these records, so parity stays pinned where the reference itself
is absent.
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import catalog  # noqa: E402
import dvda_gen  # noqa: E402
import oracle  # noqa: E402


def main():
    if not oracle.build_ref():
        raise SystemExit("oracle/_ref is missing and /root/reference is not available")
    out = {}
    specs = dict(catalog.discs())
    specs.update(catalog.GPU_LARGE)
    for name, titles in specs.items():
        with tempfile.TemporaryDirectory() as d:
            dvda_gen.make_disc(d, titles, catalog.MAX_AOB_BYTES.get(name, 0))
            rc, tracks, _samples, err = oracle.run_dump(oracle.REF_DUMP, d)
            if rc != 0:
                raise SystemExit("reference failed on %s: %s" % (name, err))
            out[name] = dict(tracks=tracks, stderr=err.strip())
            print(name, [(t["codec"], t["frames"]) for t in tracks])
    with open(os.path.join(HERE, "catalog_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
