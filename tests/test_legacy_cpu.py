"""Legacy checkpoint layouts on CPU: the oracle's restatement against vectors recorded from the reference's own
(commented-out) legacy class, and the exactness of the weight remapping onto the live layer table (sfm-nerf_b200/legacy.py)."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "legacy_pe.npz")
REF = "/root/reference/Phase2"


def _inputs(n=300, seed=1):
    g = torch.Generator().manual_seed(seed)
    pos = (torch.rand(n, 3, generator=g) - 0.5) * 4.0
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return pos, dirs


def test_oracle_legacy_forward_matches_reference_commented_class():
    g = np.load(GOLD)
    sd = orc.legacy_init_state_dict("pe", 0)
    assert abs(sum(float(v.double().abs().sum()) for v in sd.values()) - float(g["weight_checksum"])) < 1e-6
    with torch.no_grad():
        density, rgb = orc.legacy_forward(sd, torch.tensor(g["pos"]), torch.tensor(g["dirs"]))
    np.testing.assert_allclose(density.numpy(), g["density"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(rgb.numpy(), g["rgb"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("kind", ["pe", "nope"])
def test_legacy_remap_onto_live_layout_is_exact(kind):
    legacy = importlib.import_module("sfm-nerf_b200.legacy")
    sd = orc.legacy_init_state_dict(kind, 3)
    assert legacy.legacy_kind(sd) == kind
    live, pos_encoding = legacy.convert_legacy_state_dict(sd)
    assert pos_encoding == (kind == "pe")
    assert sum(v.numel() for v in live.values()) == (661636 if kind == "pe" else 627844)
    pos, dirs = _inputs()
    with torch.no_grad():
        d0, c0 = orc.legacy_forward(sd, pos, dirs)
        d1, c1 = orc.model_forward(live, pos, dirs, pos_encoding=pos_encoding, density_activation="relu")
    # identity layers and zero skip columns are exact; only BLAS blocking of the 257-row layer vs its 1 + 256 split differs
    np.testing.assert_allclose(d1.numpy(), d0.numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(c1.numpy(), c0.numpy(), rtol=0, atol=1e-6)
    assert float(d0.max()) > 0.0                      # the ReLU density head is active for some samples
    with pytest.raises(ValueError):
        legacy.legacy_kind(orc.init_state_dict(0))    # a live-format dict is not a legacy checkpoint


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference's shipped checkpoints are only present in the build container")
def test_shipped_checkpoints_load_and_remap():
    legacy = importlib.import_module("sfm-nerf_b200.legacy")
    g = np.load(GOLD)
    for rel, kind in (("checkpoints_tmp/final_model.pth", "pe"), ("checkpoints/chair/final_model.pth", "pe"),
                      ("checkpoints_no_pos_encoding/final_model.pth", "nope")):
        path = os.path.join(REF, rel)
        if not os.path.exists(path):
            continue
        sd = torch.load(path, map_location="cpu", weights_only=True)
        assert legacy.legacy_kind(sd) == kind
        live, pos_encoding = legacy.convert_legacy_state_dict(sd)
        if kind == "pe":
            pos, dirs = torch.tensor(g["pos"]).reshape(-1, 3), torch.tensor(g["dirs"]).reshape(-1, 3)
        else:
            pos, dirs = _inputs(120, 7)
        with torch.no_grad():
            d0, c0 = orc.legacy_forward(sd, pos, dirs)
            d1, c1 = orc.model_forward(live, pos, dirs, pos_encoding=pos_encoding, density_activation="relu")
        scale = max(1.0, float(d0.abs().max()))
        np.testing.assert_allclose(d1.numpy(), d0.numpy(), rtol=0, atol=2e-5 * scale)
        np.testing.assert_allclose(c1.numpy(), c0.numpy(), rtol=0, atol=2e-5)
        if rel == "checkpoints_tmp/final_model.pth" and "ckpt_rgb" in g:      # recorded from the reference's own class
            np.testing.assert_allclose(d0.numpy().reshape(g["ckpt_density"].shape), g["ckpt_density"], rtol=0, atol=2e-5 * scale)
            np.testing.assert_allclose(c0.numpy().reshape(g["ckpt_rgb"].shape), g["ckpt_rgb"], rtol=0, atol=2e-5)
