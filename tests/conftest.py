"""pytest configuration: the `gpu` marker, import paths, build fixtures.

CPU tests (-m "not gpu") cover the generator, the oracle against the reference
build / golden hashes, the host-side disc model and the exported C ABI.  GPU
tests (-m gpu) are the parity tests proper: the CUDA engine, called through its
C ABI and through the public dvd-audio.h API, against the oracle.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gen"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def gen():
    import dvda_gen
    dvda_gen.build()
    return dvda_gen


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def pkg():
    """The engine package (built libraries must exist: __graft_entry__.build())."""
    mod = importlib.import_module("libdvd-audio_b200")
    if not os.path.exists(mod.ENGINE_LIB) or not os.path.exists(mod.HOST_LIB):
        mod._build.build_all()
    return mod


@pytest.fixture(scope="session")
def disc_cache(tmp_path_factory, gen):
    """name -> (directory, info) for the catalog discs, generated once per session."""
    import catalog
    specs = dict(catalog.discs())
    specs.update(catalog.GPU_LARGE)
    made = {}

    def get(name):
        if name not in made:
            d = str(tmp_path_factory.mktemp(name))
            made[name] = (d, gen.make_disc(d, specs[name], catalog.MAX_AOB_BYTES.get(name, 0)))
        return made[name]

    return get
