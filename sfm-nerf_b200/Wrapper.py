"""Name / signature compatibility with the reference's older driver `Phase2/Wrapper.py`.

Only the SURFACE is mirrored.  `Wrapper.render` upstream differs numerically from `train.render` (near 1 / far 10 defaults,
always-jittered t = (i + U)/S, last delta 1.0) and contains a bug -- its transmittance cumprod runs over a size-1 axis and
is therefore the identity (`Wrapper.py:176`) -- so its numerics are deliberately NOT reproduced: every function here
delegates to the `train.py` mirror (the kernels of libnerf_b200), keeping `Wrapper.py`'s names, argument lists and
defaults.  `test_single_image` (the "gif" mode, `Wrapper.py:381-507`) renders the reference's 10-view orbit
`pose_spherical(th, -30 deg, 4.0)` (converted to `PixelToRay`'s camera convention by default, see `orbit_poses`) and writes
PNGs with cv2 when it is available (imageio / matplotlib are not needed).
"""
from __future__ import annotations

import argparse
import glob
import os

import numpy as np
import torch

from . import train as _t
from .NeRFModel import NeRFmodel  # noqa: F401  (Wrapper.py does `from NeRFModel import *`)

device = _t.device
loadDataset = _t.loadDataset          # Wrapper.py:18-65 == train.py:19-65
PixelToRay = _t.PixelToRay            # Wrapper.py:67-101
generateBatch = _t.generateBatch      # Wrapper.py:103-127
loss = _t.loss                        # Wrapper.py:188-189
test_image = _t.test_image            # Wrapper.py:349-379
train = _t.train                      # Wrapper.py:191-295 (loop body identical in structure to train.py:380-584)
test = _t.test                        # Wrapper.py:297-347


def render(model, rays_origin, rays_direction, args, near=1.0, far=10.0):
    """`Wrapper.render` signature (near 1, far 10 defaults); numerics of `train.render` (see module docstring)."""
    return _t.render(model, rays_origin, rays_direction, args, near, far)


def pose_spherical(theta, phi, radius):
    """Camera-to-world matrix of the reference's orbit (`Wrapper.py:443-470`, restated): translate by `radius` along z, rotate by
    `phi` about x and `theta` about y (degrees), then the fixed axis permutation -- float32 factors like the reference's."""
    trans_t = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=np.float32)
    ph, th = phi / 180.0 * np.pi, theta / 180.0 * np.pi
    rot_phi = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]], dtype=np.float32)
    rot_theta = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=np.float32)
    c2w = rot_theta @ (rot_phi @ trans_t)
    return np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w


def orbit_poses(n_views=10, phi=-30.0, radius=4.0, convention="pixel_to_ray"):
    """The `n_views` poses of the gif orbit: `pose_spherical(th, -30, 4.0)` for th in linspace(0, 360, 10, endpoint=False)
    (`Wrapper.py:472-475`).  `pose_spherical` builds Blender / OpenGL camera frames (the camera looks along its -z, +y up), while
    `PixelToRay` shoots rays along +z with +y down (`train.py:87-94`): the reference feeds the one into the other unconverted, so
    its orbit looks AWAY from the object (SURVEY A.5 -- the degenerate image/*.png it ships).  convention="pixel_to_ray"
    (default) flips the camera's y and z axes so that the orbit looks at the origin in PixelToRay's convention;
    convention="reference" returns the matrices exactly as the reference uses them."""
    out = []
    for th in np.linspace(0.0, 360.0, n_views, endpoint=False):
        c2w = np.asarray(pose_spherical(th, phi, radius), dtype=np.float64)
        if convention == "pixel_to_ray":
            c2w = c2w @ np.diag([1.0, -1.0, -1.0, 1.0])
        elif convention != "reference":
            raise ValueError("convention must be 'pixel_to_ray' or 'reference'")
        out.append(c2w)
    return np.stack(out)


def test_single_image(images, poses, camera_info, args):
    """"gif" mode (`Wrapper.py:381-507`): the newest checkpoint under `args.checkpoint_path` renders the 10 novel views of the
    orbit `pose_spherical(th, -30 deg, 4.0)` (see `orbit_poses`; `args.orbit_convention`, default "pixel_to_ray").  Every frame is
    one pass of ray generation + depths + fused MLP + compositing on the GPU per batch of `args.n_rays_batch` rays.  Returns the
    list of [H,W,3] frames scaled to 0..255 like the reference's and writes image_<k>.png (k from 1) when cv2 is present; the
    reference's matplotlib ray plots are not reproduced."""
    final = os.path.join(args.checkpoint_path, "final_model.pth")
    found = [final] if os.path.exists(final) else sorted(glob.glob(os.path.join(args.checkpoint_path, "model_*.pth")))[-1:]
    if not found:
        print("No checkpoint found")
        return None
    ckpt = torch.load(found[0], map_location=device, weights_only=False)
    sd = ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt
    if "fc1.0.weight" in sd:
        from .legacy import load_legacy_checkpoint
        model = load_legacy_checkpoint(sd, device)
    else:
        model = NeRFmodel().to(device)
        model.load_state_dict(sd)
    print(f"Checkpoint {found[0]} loaded")
    os.makedirs(args.images_path, exist_ok=True)
    H, W = int(camera_info["height"]), int(camera_info["width"])
    blank = np.zeros((H, W, 3), dtype=np.float32)             # test_image also reports an MSE; there is no ground truth for novel views
    frames = []
    for count, c2w in enumerate(orbit_poses(convention=getattr(args, "orbit_convention", "pixel_to_ray")), start=1):
        pred, _ = test_image(model, blank, c2w, camera_info, args)
        frame = pred.detach().cpu().numpy().astype(np.float32) * 255.0
        frames.append(frame)
        try:
            import cv2
            cv2.imwrite(os.path.join(args.images_path, f"image_{count}.png"), frame)      # Wrapper.py:505-507 (RGB order as the reference writes it)
        except ImportError:
            pass
    return frames


def main(args):
    """`Wrapper.main` (`Wrapper.py:509-536`): always loads the 'train' split; modes train / test / gif."""
    print("Loading data...")
    images, poses, camera_info = loadDataset(args.data_path, "train")
    for k in ("n_rays_batch", "n_sample", "max_iters", "n_pos_freq", "n_dirc_freq"):
        setattr(args, k, int(getattr(args, k)))
    for k, v in (("perturb", True), ("near", 1.0), ("far", 10.0), ("chunk_size", 32768), ("hidden_size", 256), ("lrate_decay", 0.1),
                 ("lrate_decay_steps", 1000), ("exp_name", "wrapper")):
        if not hasattr(args, k):
            setattr(args, k, v)                       # flags train.py has and Wrapper.py hard-codes (StepLR(1000, 0.1): Wrapper.py:220)
    args.checkpoint_path = os.path.join(args.checkpoint_path, args.data_path.split("/")[-2])
    if args.mode == "train":
        print("Start training")
        train(images, poses, camera_info, args)
    elif args.mode == "test":
        print("Start testing")
        args.load_checkpoint = True
        test(images, poses, camera_info, args)
    elif args.mode == "gif":
        print("Start gif")
        args.load_checkpoint = True
        test_single_image(images, poses, camera_info, args)


def configParser():
    """Flag names and defaults of `Wrapper.py:538-555`."""
    p = argparse.ArgumentParser()
    for flag, default in (("--data_path", "./Phase2/nerf_synthetic/lego/"), ("--mode", "train"), ("--lrate", 5e-4), ("--n_pos_freq", 10),
                          ("--n_dirc_freq", 4), ("--n_rays_batch", 32 * 32 * 8), ("--n_sample", 256), ("--max_iters", 100001),
                          ("--logs_path", "./logs/"), ("--log_id", ""), ("--checkpoint_path", "./Phase2/checkpoints/"),
                          ("--load_checkpoint", True), ("--save_ckpt_iter", 1000), ("--images_path", "./image/")):
        p.add_argument(flag, default=default)
    return p


if __name__ == "__main__":
    main(configParser().parse_args())
