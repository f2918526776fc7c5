"""Import alias: ``import sfmnerf_b200`` -> the package in ``sfm-nerf_b200/`` (hyphenated dir)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("sfm-nerf_b200")
sys.modules[__name__] = _pkg
