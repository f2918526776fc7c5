"""Drop-in for the reference ``Phase2/NeRFModel.py`` (live class, NeRFModel.py:101-202).

Same constructor, attribute names, ``state_dict`` keys/shapes, default ``nn.Linear`` init order and
``forward(pos, direction) -> (density, rgb)`` contract -- but ``forward`` runs the fused sm_100a
tcgen05 kernel (``nerf_mlp_forward`` / ``nerf_mlp_backward`` of libnerf_b200.so) instead of 13
eager ``addmm`` launches.  All 26 parameter tensors are views into ONE flat fp32 device buffer in
``state_dict`` order, which is what the C ABI consumes; ``torch.optim`` and ``load_state_dict``
keep working on the views.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import engine

# (attribute path, out_features, in_features as a function of (C, pos_in, dir_in)); NeRFModel.py:116-140
_LAYER_TABLE = (
    ("fc1", lambda C, p, d: (C, p)),
    ("fc_layers.0", lambda C, p, d: (C, C)),
    ("fc_layers.1", lambda C, p, d: (C, C)),
    ("fc_layers.2", lambda C, p, d: (C, C)),
    ("fc_layers.3", lambda C, p, d: (C, C)),
    ("skip_layer", lambda C, p, d: (C, C + p)),
    ("fc_layers2.0", lambda C, p, d: (C, C)),
    ("fc_layers2.1", lambda C, p, d: (C, C)),
    ("fc_layers2.2", lambda C, p, d: (C, C)),
    ("density_layer", lambda C, p, d: (1, C)),
    ("feature_layer", lambda C, p, d: (C, C)),
    ("rgb_layer1", lambda C, p, d: (128, C + d)),
    ("rgb_layer2", lambda C, p, d: (3, 128)),
)


class NeRFmodel(nn.Module):
    def __init__(self, embed_pos_L=10, embed_direction_L=4, num_channels=256, pos_encoding=True):
        super().__init__()
        self.embed_pos_L = embed_pos_L
        self.embed_direction_L = embed_direction_L
        self.pos_encoding = pos_encoding
        self.num_channels = num_channels
        enc = lambda L: 3 * (2 * L + 1) if pos_encoding else 3
        pos_in, dir_in = enc(embed_pos_L), enc(embed_direction_L)

        # Build the Linear modules in the reference's construction order so that
        # torch.manual_seed(s); NeRFmodel() draws identical initial weights.
        made = {}
        for name, shape in _LAYER_TABLE:
            out_f, in_f = shape(num_channels, pos_in, dir_in)
            made[name] = nn.Linear(in_f, out_f)
        self.fc1 = made["fc1"]
        self.fc_layers = nn.ModuleList([made[f"fc_layers.{i}"] for i in range(4)])
        self.skip_layer = made["skip_layer"]
        self.fc_layers2 = nn.ModuleList([made[f"fc_layers2.{i}"] for i in range(3)])
        self.density_layer = made["density_layer"]
        self.feature_layer = made["feature_layer"]
        self.rgb_layer1 = made["rgb_layer1"]
        self.rgb_layer2 = made["rgb_layer2"]
        # "softplus" (NeRFModel.py:192) or "relu": the density activation of the reference's legacy checkpoints (legacy.py)
        self.density_activation = "softplus"
        self._kernel_state = None        # engine.ModelState, created lazily on the CUDA device

    # ------------------------------------------------------------------ kernel plumbing
    @property
    def layout_id(self) -> int:
        """Which fixed-function layout of libnerf_b200 this instance maps to (raises otherwise)."""
        if self.num_channels not in (256, 128):
            raise NotImplementedError("libnerf_b200 kernels are built for num_channels 256 (the reference default) and 128 (--hidden_size)")
        if self.density_activation not in ("softplus", "relu"):
            raise NotImplementedError("density_activation must be 'softplus' or 'relu'")
        lid = int(_lib.load().nerf_layout_id(int(self.embed_pos_L), int(self.embed_direction_L), 1 if self.pos_encoding else 0,
                                             1 if self.density_activation == "relu" else 0, int(self.num_channels)))
        if lid < 0:
            raise NotImplementedError("libnerf_b200 supports embed_pos_L / embed_direction_L (--n_pos_freq / --n_dirc_freq) from 0 to 10: "
                                      "the encoding must fit the 64-channel operand tile")
        return lid

    def ordered_parameters(self):
        """The 26 parameter tensors in state_dict order (weight, bias per layer)."""
        out = []
        for name, _ in _LAYER_TABLE:
            mod = self.get_submodule(name)
            out += [mod.weight, mod.bias]
        return out

    def kernel_state(self) -> "engine.ModelState":
        """Flat parameter buffer + packed bf16 weight images for the current parameter values."""
        params = self.ordered_parameters()
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.NerfB200Error("NeRFmodel parameters are on %s: the sm_100a kernels need a CUDA device "
                                     "(there is no CPU fallback)" % dev)
        st = self._kernel_state
        if st is None or st.device != dev or st.layout != self.layout_id or not st.owns(params):
            st = engine.ModelState.adopt(self.layout_id, params)
            self._kernel_state = st
        st.refresh_packed(params)
        return st

    # ------------------------------------------------------------------ reference API
    def position_encoding(self, x, L):
        """[x, sin(2^0 pi x), cos(2^0 pi x), ..., sin(2^(L-1) pi x), cos(2^(L-1) pi x)] on the last
        axis (NeRFModel.py:142-156), computed by the K3 kernel (``nerf_encode``)."""
        return engine.encode(x, L)

    def forward(self, pos, direction):
        """density [...,1] (softplus) and rgb [...,3] (sigmoid) for positions / view directions
        [...,3] (NeRFModel.py:158-202); differentiable w.r.t. the parameters."""
        st = self.kernel_state()
        lead = pos.shape[:-1]
        p = pos.reshape(-1, 3).to(torch.float32).contiguous()
        d = direction.reshape(-1, 3).to(torch.float32).contiguous()
        params = self.ordered_parameters()
        out = engine.MLPFunction.apply(st, p, d, *params) if engine.wants_grad(params) else engine.mlp_inference(st, p, d)
        return out[:, 0:1].reshape(*lead, 1), out[:, 1:4].reshape(*lead, 3)


__all__ = ["NeRFmodel", "np"]
