"""Importable alias of the package directory `libdvd-audio_b200/` (whose name,
mirroring the reference's, is not a valid Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("libdvd-audio_b200")
sys.modules[__name__] = _pkg
