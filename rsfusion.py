"""Importable alias of the ``ra-slam_b200`` package directory (a hyphen cannot be imported by name).

    import rsfusion
    grid = rsfusion.TSDFGrid(0.01, 0.06)
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("ra-slam_b200")
sys.modules[__name__] = _pkg
