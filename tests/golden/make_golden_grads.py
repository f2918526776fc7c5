"""Full fp32 gradient tensors of the UNMODIFIED reference train step (train.py:456-463), for the per-tensor gradient
parity test (north_star: "per-layer gradients match within 2 % relative error").  Run in the build container:

    python tests/golden/make_golden_grads.py

  grads_pe256.npz        the 48-ray batch of render_pe256.npz (same rays / target / seed-0 weights): all 26 tensors
  grads_pe256_b1024.npz  1024 rays from 4 poses x 64 samples (a quarter of the C2 batch): rays, target and all 26 tensors
Both are outputs of ref.render + nn.MSELoss + backward on ref.NeRFmodel under torch.manual_seed(0), fp32, one thread.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    from make_golden import import_reference
    from oracle import nerf_oracle as orc
    ref = import_reference()
    torch.set_num_threads(1)
    args = argparse.Namespace(n_sample=64, perturb=False, chunk_size=32768, n_rays_batch=48, near=2.0, far=6.0)

    def grads_of(o, d, target):
        torch.manual_seed(0)
        model = ref.NeRFmodel()
        model.zero_grad()
        loss = torch.nn.MSELoss()(ref.render(model, o, d, args), target)
        loss.backward()
        return loss.item(), {k.replace(".", "__"): p.grad.numpy().copy() for k, p in model.named_parameters()}

    g = np.load(os.path.join(HERE, "render_pe256.npz"))
    loss, grads = grads_of(torch.tensor(g["o"]), torch.tensor(g["d"]), torch.tensor(g["target"]))
    assert abs(loss - float(g["loss"])) < 1e-7, (loss, float(g["loss"]))
    for k, v in grads.items():                                  # consistent with the norms / samples recorded by make_golden.py
        dev = abs(np.linalg.norm(v) - float(g[k + "__gnorm"])) / (float(g[k + "__gnorm"]) + 1e-30)
        samp = v.reshape(-1)[:: max(1, v.size // 64)][:64]
        dev_s = np.abs(samp - g[k + "__g"]).max() / (np.abs(g[k + "__g"]).max() + 1e-30)
        assert dev < 1e-4 and dev_s < 1e-3, (k, dev, dev_s)      # two fp32 CPU evaluations differ by summation order only
    np.savez(os.path.join(HERE, "grads_pe256.npz"), loss=loss, **grads)

    cam = orc.make_camera_info(100, 100)
    poses = orc.sphere_poses(4, seed=7)
    rr = np.random.default_rng(8)
    o, d = [], []
    for i in range(1024):
        a, b = ref.PixelToRay(cam, poses[i % 4], (int(rr.integers(0, 100)), int(rr.integers(0, 100))), None)
        o.append(a); d.append(b)
    o = np.array(o, dtype=np.float32); d = np.array(d, dtype=np.float32)
    target = rr.random((1024, 3)).astype(np.float32)
    loss, grads = grads_of(torch.tensor(o), torch.tensor(d), torch.tensor(target))
    np.savez(os.path.join(HERE, "grads_pe256_b1024.npz"), o=o, d=d, target=target, loss=loss, **grads)
    print("gradient goldens written")


if __name__ == "__main__":
    main()
