"""Importer for the reference's SHIPPED checkpoints: bare ``OrderedDict`` state_dicts with legacy key
names that the reference's own loaders (`train.py:421,603`) cannot read.

  * LEGACY_PE   (`Phase2/checkpoints/chair`, `checkpoints_tmp/...`): the commented-out class `NeRFModel.py:7-99` --
    ``fc1.0`` 63->256, ``block1.{0,2,4}``, ``skip_layer.0`` 319->256 on cat([h, pos]), ``block2.{0,2}``, ``block2.4``
    256->257 (ReLU on all 257 outputs; density = channel 0, feature = channels 1..256), ``rgb_fc.0`` 283->128,
    ``rgb_fc.2`` 128->3 + sigmoid; ``density_fc`` is present but unused.
  * LEGACY_NOPE (`Phase2/checkpoints_no_pos_encoding`): ``fc1.0`` 3->256, ``block1.{0,..,10}`` six 256->256 layers,
    ``block1.12`` 256->257, no skip layer, ``rgb_fc.0`` 259->128, ``rgb_fc.2``; raw xyz / direction inputs.

Both are evaluated by the SAME kernels as the live model: their weights are remapped onto the live layer table
(NeRFModel.py:102-140), which is exact --
  * a missing 256->256 layer becomes the identity: its input is a ReLU output (>= 0), so relu(I h + 0) = h, also after
    bf16 rounding (1.0 * h accumulates exactly and rounds back to the same bf16 value);
  * LEGACY_NOPE has no skip input: the live ``skip_layer`` gets zero weights on its 3 position columns;
  * row 0 of the 257-wide layer is the density head, rows 1..256 the feature layer;
and the density activation is ReLU instead of softplus (``NeRFmodel.density_activation = "relu"`` -> layouts 2 / 3 of
libnerf_b200).  No arithmetic happens here: tensors are sliced and copied.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .NeRFModel import NeRFmodel


def legacy_kind(sd) -> str:
    """"pe" / "nope" for a legacy state_dict, ValueError otherwise (e.g. a live-format dict)."""
    keys = set(sd.keys())
    if "fc1.0.weight" not in keys or "rgb_fc.2.weight" not in keys:
        raise ValueError("not a legacy SfM-NeRF checkpoint (expected keys like 'fc1.0.weight', 'rgb_fc.2.weight')")
    if "skip_layer.0.weight" in keys and "block2.4.weight" in keys:
        if tuple(sd["fc1.0.weight"].shape) != (256, 63) or tuple(sd["block2.4.weight"].shape) != (257, 256):
            raise ValueError("legacy PE checkpoint with unsupported shapes")
        return "pe"
    if "block1.12.weight" in keys:
        if tuple(sd["fc1.0.weight"].shape) != (256, 3) or tuple(sd["block1.12.weight"].shape) != (257, 256):
            raise ValueError("legacy no-PE checkpoint with unsupported shapes")
        return "nope"
    raise ValueError("unrecognised legacy checkpoint layout")


def convert_legacy_state_dict(sd):
    """legacy state_dict -> (live-format state_dict, pos_encoding flag).  Exact (see module docstring)."""
    kind = legacy_kind(sd)
    f = lambda k: sd[k].detach().to(torch.float32).clone()
    eye, zero = torch.eye(256, dtype=torch.float32), torch.zeros(256, dtype=torch.float32)
    out = OrderedDict()

    def put(name, w, b):
        out[name + ".weight"], out[name + ".bias"] = w, b

    put("fc1", f("fc1.0.weight"), f("fc1.0.bias"))
    if kind == "pe":
        for i, k in enumerate(("block1.0", "block1.2", "block1.4")):
            put(f"fc_layers.{i}", f(k + ".weight"), f(k + ".bias"))
        put("fc_layers.3", eye.clone(), zero.clone())
        put("skip_layer", f("skip_layer.0.weight"), f("skip_layer.0.bias"))
        put("fc_layers2.0", f("block2.0.weight"), f("block2.0.bias"))
        put("fc_layers2.1", f("block2.2.weight"), f("block2.2.bias"))
        put("fc_layers2.2", eye.clone(), zero.clone())
        last = "block2.4"
    else:
        for i, k in enumerate(("block1.0", "block1.2", "block1.4", "block1.6")):
            put(f"fc_layers.{i}", f(k + ".weight"), f(k + ".bias"))
        put("skip_layer", torch.cat([f("block1.8.weight"), torch.zeros(256, 3)], dim=1), f("block1.8.bias"))   # [h, pos]: pos unused
        put("fc_layers2.0", f("block1.10.weight"), f("block1.10.bias"))
        put("fc_layers2.1", eye.clone(), zero.clone())
        put("fc_layers2.2", eye.clone(), zero.clone())
        last = "block1.12"
    w, b = f(last + ".weight"), f(last + ".bias")
    put("density_layer", w[0:1].clone(), b[0:1].clone())
    put("feature_layer", w[1:].clone(), b[1:].clone())
    put("rgb_layer1", f("rgb_fc.0.weight"), f("rgb_fc.0.bias"))
    put("rgb_layer2", f("rgb_fc.2.weight"), f("rgb_fc.2.bias"))
    return out, kind == "pe"


def load_legacy_checkpoint(path_or_state_dict, device="cuda") -> NeRFmodel:
    """A ``NeRFmodel`` (live class, ``density_activation="relu"``) that evaluates a shipped legacy checkpoint."""
    sd = path_or_state_dict
    if not isinstance(sd, dict):
        sd = torch.load(sd, map_location="cpu", weights_only=True)
    if "model_state_dict" in sd:                       # the dict format of train.py:510-528 wrapping a legacy state_dict
        sd = sd["model_state_dict"]
    live, pos_encoding = convert_legacy_state_dict(sd)
    model = NeRFmodel(pos_encoding=pos_encoding)
    model.load_state_dict(live)
    model.density_activation = "relu"
    return model.to(device)
