"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py

Imports /root/reference/Phase2/train.py with `imageio` / `matplotlib` stubbed (they are
top-level imports the image lacks; no path logic lives in them) and records the reference's
own outputs on seeded synthetic inputs.  /root/reference does not travel to the GPU box, so
these vectors are what pins oracle/nerf_oracle.py (and through it the CUDA path) there.
"""
import argparse
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("NERF_REFERENCE", "/root/reference")


def import_reference():
    for n in ("imageio", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, os.path.join(REF, "Phase2"))
    import train as ref  # noqa
    return ref


def main():
    ref = import_reference()
    from oracle import nerf_oracle as orc
    torch.set_num_threads(1)          # fixed reduction order for the recorded matmuls
    ns = lambda **kw: argparse.Namespace(**kw)

    # ---- G1: rays (PixelToRay, test_image ray loop, generateBatch) ----------------------
    cam = orc.make_camera_info(12, 10)
    pose = orc.look_at_pose(30.0, -30.0)
    o = np.array([ref.PixelToRay(cam, pose, (x, y), None)[0] for y in range(10) for x in range(12)])
    d = np.array([ref.PixelToRay(cam, pose, (x, y), None)[1] for y in range(10) for x in range(12)])
    poses = orc.sphere_poses(5, seed=1)
    rng = np.random.default_rng(2)
    images = rng.random((5, 10, 12, 3))
    random.seed(1234)
    batch = ref.generateBatch(images, poses, cam, ns(n_rays_batch=64))
    np.savez(os.path.join(HERE, "rays.npz"), K=cam["camera_matrix"], pose=pose, o64=o, d64=d,
             poses=poses, images=images, batch=batch, py_seed=1234)

    # ---- G2: positional encoding ---------------------------------------------------------
    m = ref.NeRFmodel()
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(257, 3, generator=g) - 0.5) * 12.0        # |x| up to 6 (far bound)
    np.savez(os.path.join(HERE, "encoding.npz"), x=x.numpy(),
             pe10=m.position_encoding(x, 10).numpy(), pe4=m.position_encoding(x, 4).numpy())

    # ---- G3: live model forward, render, gradients (hidden 256, PE on) -----------------
    torch.manual_seed(0)
    model = ref.NeRFmodel()
    sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    np.savez(os.path.join(HERE, "weights_pe256_seed0.npz"), **sd)
    N, S = 48, 64
    cam2 = orc.make_camera_info(100, 100)
    o2, d2 = [], []
    rr = np.random.default_rng(4)
    for _ in range(N):
        a, b = ref.PixelToRay(cam2, pose, (int(rr.integers(0, 100)), int(rr.integers(0, 100))), None)
        o2.append(a); d2.append(b)
    o2 = torch.tensor(np.array(o2, dtype=np.float32)); d2 = torch.tensor(np.array(d2, dtype=np.float32))
    args = ns(n_sample=S, perturb=False, chunk_size=32768, n_rays_batch=N, near=2.0, far=6.0)
    with torch.no_grad():
        rgb_det = ref.render(model, o2, d2, args)
        pts = (o2[:, None, :] + d2[:, None, :] * torch.linspace(2.0, 6.0, S)[None, :, None]).reshape(-1, 3)
        dens, cols = model(pts[:512], d2.repeat_interleave(S, 0)[:512])
    args_p = ns(n_sample=S, perturb=True, chunk_size=1000, n_rays_batch=N, near=2.0, far=6.0)
    torch.manual_seed(77)
    with torch.no_grad():
        rgb_jit = ref.render(model, o2, d2, args_p)
    torch.manual_seed(77)
    t_rand = torch.rand(N, S)
    # gradients of the reference train step (train.py:456-463) at deterministic depths
    target = torch.tensor(rr.random((N, 3)).astype(np.float32))
    model.zero_grad()
    loss = torch.nn.MSELoss()(ref.render(model, o2, d2, args), target)
    loss.backward()
    gnorm = {k.replace(".", "__") + "__gnorm": p.grad.norm().item() for k, p in model.named_parameters()}
    gsamp = {k.replace(".", "__") + "__g": p.grad.flatten()[:: max(1, p.numel() // 64)][:64].numpy().copy()
             for k, p in model.named_parameters()}
    # one Adam step (train.py:401,464)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    opt.zero_grad()
    loss2 = torch.nn.MSELoss()(ref.render(model, o2, d2, args), target)
    loss2.backward(); opt.step()
    w_after = {k.replace(".", "__") + "__w1": p.detach().flatten()[:: max(1, p.numel() // 64)][:64].numpy().copy()
               for k, p in model.named_parameters()}
    np.savez(os.path.join(HERE, "render_pe256.npz"), o=o2.numpy(), d=d2.numpy(), S=S,
             rgb_det=rgb_det.numpy(), rgb_jit=rgb_jit.numpy(), t_rand=t_rand.numpy(),
             pts512=pts[:512].numpy(), dens512=dens.numpy(), cols512=cols.numpy(),
             target=target.numpy(), loss=loss.item(), **gnorm, **gsamp, **w_after)

    # ---- G4: no-PE live variant (pos_encoding=False) -----------------------------------
    torch.manual_seed(0)
    model_np = ref.NeRFmodel(pos_encoding=False)
    with torch.no_grad():
        rgb_np = ref.render(model_np, o2, d2, args)
    w0 = model_np.fc1.weight.detach().numpy().copy()
    np.savez(os.path.join(HERE, "render_nope256.npz"), rgb_det=rgb_np.numpy(), fc1_w=w0,
             n_params=sum(p.numel() for p in model_np.parameters()),
             checksum=float(sum(p.double().sum().item() for p in model_np.parameters())))

    # ---- G5: test_image on a small frame -------------------------------------------------
    torch.manual_seed(0)
    model = ref.NeRFmodel()
    cam3 = orc.make_camera_info(16, 12)
    img = np.random.default_rng(5).random((12, 16, 3))
    args_t = ns(n_sample=32, perturb=False, chunk_size=32768, n_rays_batch=50, near=2.0, far=6.0)
    pred, mse = ref.test_image(model, img, pose, cam3, args_t)
    np.savez(os.path.join(HERE, "test_image.npz"), image=img, pred=pred.numpy(), mse=mse,
             K=cam3["camera_matrix"], pose=pose)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
