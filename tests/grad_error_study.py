"""Where does the gradient error of a bf16-operand evaluation come from?  (CPU study with the oracle)

    python tests/grad_error_study.py [n_rays=48]

Per parameter tensor: rel-L2 error against the fp32 reference gradients of
  bf16      : the operand rounding the CUDA kernels use (weights, biases, layer inputs -> bf16; fp32 accumulate)
  +pe_split : the same, but the positional / direction encodings kept at ~16 mantissa bits (what a hi + lo pair of bf16
              K-slabs for the PE inputs of fc1 / skip_layer / rgb_layer1 would give)
  w_only    : only the weights rounded (activations fp32)
  act_only  : only the activations rounded (weights fp32)
The CUDA path agrees with `bf16` to <= 0.5 % on every tensor (tests/test_gpu_parity.py), so this table is a proxy
for what each change would buy on the GPU.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from oracle import nerf_oracle as orc


def forward_variant(sd, pos, direction, round_w, round_act, round_enc):
    bf = torch.bfloat16
    qa = (lambda x: orc._round_st(x, bf)) if round_act else (lambda x: x)
    qw = (lambda x: orc._round_st(x, bf)) if round_w else (lambda x: x)

    def qe(x):
        if round_enc == "bf16":
            return orc._round_st(x, bf)
        if round_enc == "split":                       # hi + lo bf16 pair
            hi = x.to(bf).float()
            lo = (x - hi).to(bf).float()
            return x + ((hi + lo) - x).detach()
        return x
    lin = lambda name, x: F.linear(x, qw(sd[name + ".weight"]), qw(sd[name + ".bias"]))
    head = lambda name, x: F.linear(x, sd[name + ".weight"], sd[name + ".bias"])
    pe, de = qe(orc.position_encoding(pos, 10)), qe(orc.position_encoding(direction, 4))
    h = qa(F.relu(lin("fc1", pe)))
    for i in range(4):
        h = qa(F.relu(lin(f"fc_layers.{i}", h)))
    h = qa(F.relu(lin("skip_layer", torch.cat([h, pe], -1))))
    for i in range(3):
        h = qa(F.relu(lin(f"fc_layers2.{i}", h)))
    density = F.softplus(head("density_layer", h))
    feature = qa(F.relu(lin("feature_layer", h)))
    r = qa(F.relu(lin("rgb_layer1", torch.cat([feature, de], -1))))
    return density, torch.sigmoid(head("rgb_layer2", r))


def grads_variant(sd, o, d, target, S, **kw):
    from collections import OrderedDict
    p = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in sd.items())
    N = o.shape[0]
    z = orc.sample_z(N, S, 2.0, 6.0)
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).reshape(-1, 3)
    dirs = d[:, None, :].expand(-1, S, -1).reshape(-1, 3)
    dens, rgb = forward_variant(p, pts, dirs, **kw)
    out = orc.composite(dens.reshape(N, S), rgb.reshape(N, S, 3), z)["rgb"]
    loss = torch.mean((out - target) ** 2)
    g = torch.autograd.grad(loss, list(p.values()))
    return dict(zip(p.keys(), g))


def main():
    n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    torch.set_num_threads(os.cpu_count())
    sd = orc.init_state_dict(0)
    gen = torch.Generator().manual_seed(5)
    cam = orc.make_camera_info(100, 100)
    pose = orc.look_at_pose(30.0, -30.0)
    px = torch.randint(0, 100, (n_rays, 2), generator=gen).numpy()
    rays = orc.rays_from_indices(np.zeros((1, 100, 100, 3)), pose[None], cam, np.zeros(n_rays, int), px[:, 0], px[:, 1])
    o, d = torch.tensor(rays[:, 0:3]), torch.tensor(rays[:, 3:6])
    target = torch.rand(n_rays, 3, generator=gen)
    variants = {
        "fp32": dict(round_w=False, round_act=False, round_enc="none"),
        "bf16": dict(round_w=True, round_act=True, round_enc="bf16"),
        "+pe_split": dict(round_w=True, round_act=True, round_enc="split"),
        "w_only": dict(round_w=True, round_act=False, round_enc="none"),
        "act_only": dict(round_w=False, round_act=True, round_enc="bf16"),
    }
    res = {k: grads_variant(sd, o, d, target, 64, **v) for k, v in variants.items()}
    names = [k for k in variants if k != "fp32"]
    print(f"{n_rays} rays x 64 samples; rel-L2 of the gradient vs fp32 autograd")
    print(f"{'tensor':24s}" + "".join(f"{n:>11s}" for n in names))
    for k in sd:
        ref = res["fp32"][k]
        print(f"{k:24s}" + "".join(f"{float((res[n][k] - ref).norm() / ref.norm()):11.2e}" for n in names))
    allv = lambda r: torch.cat([v.flatten() for v in r.values()])
    ref = allv(res["fp32"])
    print(f"{'ALL (flat vector)':24s}" + "".join(f"{float((allv(res[n]) - ref).norm() / ref.norm()):11.2e}" for n in names))


if __name__ == "__main__":
    main()
