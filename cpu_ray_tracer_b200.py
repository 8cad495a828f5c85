"""Importable alias for the package directory `cpu-ray-tracer_b200/` (hyphenated like the reference's name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("cpu-ray-tracer_b200")
sys.modules[__name__] = _pkg
