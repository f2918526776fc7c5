"""sfm-nerf_b200 -- B200-native (sm_100a) NeRF render / train-step hot path, drop-in for the
Phase2 call surface of ColinBalfour/SfM-NeRF.  The directory name carries a hyphen (the project
name); import it as ``sfmnerf_b200`` (alias module at the repo root) or with
``importlib.import_module("sfm-nerf_b200")``.
"""
from . import _lib, engine, legacy, parallel, synthetic  # noqa: F401
from . import Wrapper  # noqa: F401  (name-compatible mirror of the reference's older Phase2/Wrapper.py driver)
from .NeRFModel import NeRFmodel  # noqa: F401
from .legacy import convert_legacy_state_dict, load_legacy_checkpoint  # noqa: F401
from .train import (PixelToRay, configParser, generateBatch, loadDataset, loss, main, render,  # noqa: F401
                    render_full, test, test_image, train)

__all__ = ["NeRFmodel", "PixelToRay", "generateBatch", "render", "render_full", "loss", "test_image", "train",
           "test", "main", "configParser", "loadDataset", "engine", "synthetic", "legacy", "load_legacy_checkpoint",
           "convert_legacy_state_dict"]
