"""Synthetic inputs: the reference ships no dataset and there is no network,
so tests and the benchmark use poses on a sphere in PixelToRay's convention (+z forward, +y down,
train.py:87-94), the reference's focal formula (train.py:53) and an analytic "lego-like" scene
(boxes + studs, black background as in train.py:47-48/269).  Data generation only -- not hot path.
"""
from __future__ import annotations

import numpy as np
import torch

CAMERA_ANGLE_X = 0.6911112070083618


def make_camera_info(width, height, camera_angle_x=CAMERA_ANGLE_X):
    focal = 0.5 * camera_angle_x * width / np.tan(0.5 * camera_angle_x)
    return {"width": int(width), "height": int(height),
            "camera_matrix": np.array([[focal, 0, width / 2], [0, focal, height / 2], [0, 0, 1]])}


def look_at_pose(theta_deg, phi_deg, radius=4.0):
    """Camera-to-world matrix of a camera on a sphere looking at the origin (+z forward, +y down)."""
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    c = radius * np.array([np.cos(ph) * np.sin(th), np.sin(ph), np.cos(ph) * np.cos(th)])
    fwd = -c / np.linalg.norm(c)
    right = np.cross(np.array([0.0, -1.0, 0.0]), fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, down, fwd, c
    return m


def sphere_poses(n, seed=0, radius=4.0):
    rng = np.random.default_rng(seed)
    return np.stack([look_at_pose(rng.uniform(0, 360), rng.uniform(-60, -10), radius) for _ in range(n)])


def lego_field(pts):
    """Analytic density / albedo of a brick-like object inside [-1.2,1.2]^3: pts [...,3] -> (sigma [...], rgb [...,3])."""
    x, y, z = pts[..., 0], pts[..., 1], pts[..., 2]
    base = (x.abs() < 1.0) & (z.abs() < 0.6) & (y > -0.1) & (y < 0.5)
    top = (x.abs() < 0.5) & (z.abs() < 0.6) & (y > -0.6) & (y <= -0.1)
    stud = torch.zeros_like(base)
    for cx in (-0.75, -0.25, 0.25, 0.75):
        for cz in (-0.3, 0.3):
            on_top = (cx > -0.5) & (cx < 0.5)
            y0, y1 = (-0.75, -0.6) if on_top else (-0.25, -0.1)
            stud |= (((x - cx) ** 2 + (z - cz) ** 2) < 0.15 ** 2) & (y > y0) & (y <= y1)
    inside = base | top | stud
    sigma = torch.where(inside, torch.full_like(x, 40.0), torch.zeros_like(x))
    rgb = torch.zeros(pts.shape, dtype=pts.dtype, device=pts.device)
    rgb[base] = torch.tensor([0.85, 0.15, 0.10], dtype=pts.dtype, device=pts.device)
    rgb[top] = torch.tensor([0.95, 0.75, 0.10], dtype=pts.dtype, device=pts.device)
    rgb[stud] = torch.tensor([0.15, 0.35, 0.85], dtype=pts.dtype, device=pts.device)
    return sigma, rgb


def dense_field(pts):
    """A frame-filling test scene (box of half-size 1.2 with smoothly varying colour): unlike the mostly
    black lego views it has signal in every ray, so short trainings do not collapse to the black solution."""
    inside = (pts.abs() < 1.2).all(dim=-1)
    sigma = torch.where(inside, torch.full_like(pts[..., 0], 30.0), torch.zeros_like(pts[..., 0]))
    phase = torch.tensor([0.0, 1.0, 2.0], dtype=pts.dtype, device=pts.device)
    rgb = (0.5 + 0.4 * torch.sin(3.0 * pts + phase)) * inside.unsqueeze(-1)
    return sigma, rgb


def textured_field(pts):
    """Like dense_field but with finer colour detail (spatial frequency 9 instead of 3) and two spherical cut-outs, so a
    2k-step fit ends near 25-30 dB instead of > 40 dB: differences between two trainings then reflect the optimiser's
    path, not the bf16 rounding floor of the rendered colours."""
    inside = (pts.abs() < 1.2).all(dim=-1)
    c1 = torch.tensor([0.5, 0.4, 0.3], dtype=pts.dtype, device=pts.device)
    c2 = torch.tensor([-0.6, -0.2, 0.5], dtype=pts.dtype, device=pts.device)
    inside = inside & ((pts - c1).norm(dim=-1) > 0.55) & ((pts - c2).norm(dim=-1) > 0.45)
    sigma = torch.where(inside, torch.full_like(pts[..., 0], 30.0), torch.zeros_like(pts[..., 0]))
    phase = torch.tensor([0.0, 1.0, 2.0], dtype=pts.dtype, device=pts.device)
    rgb = (0.5 + 0.4 * torch.sin(9.0 * pts + phase) * torch.cos(5.0 * pts.roll(1, -1))) * inside.unsqueeze(-1)
    return sigma, rgb


def render_lego_images(poses, camera_info, n_sample=256, near=2.0, far=6.0, device="cpu", rays_per_chunk=65536, field=None):
    """Ground-truth images [n,H,W,3] float32 of the analytic scene, composited with the reference's
    formula (train.py:248-269) in fp64 on ``device`` (plain torch: dataset generation, not hot path)."""
    W, H = camera_info["width"], camera_info["height"]
    K = camera_info["camera_matrix"]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    p = torch.stack([(xs - K[0][2]) / K[0][0], (ys - K[1][2]) / K[1][1], torch.ones_like(xs)], -1).reshape(-1, 3).to(device)
    t = torch.linspace(0, 1, n_sample, dtype=torch.float64, device=device)
    zv = near + (far - near) * t
    dists = torch.cat([zv[1:] - zv[:-1], torch.full((1,), 1e-3, dtype=torch.float64, device=device)])
    out = []
    for pose in poses:
        pose_t = torch.as_tensor(pose, dtype=torch.float64, device=device)
        d = p @ pose_t[:3, :3].T
        d = d / d.norm(dim=-1, keepdim=True)
        o = pose_t[:3, 3]
        img = torch.empty((H * W, 3), dtype=torch.float64, device=device)
        for i in range(0, H * W, rays_per_chunk):
            dd = d[i:i + rays_per_chunk]
            pts = o + dd[:, None, :] * zv[None, :, None]
            sigma, rgb = (field or lego_field)(pts)
            alpha = 1.0 - torch.exp(-sigma * dists)
            T = torch.cumprod(1.0 - alpha + 1e-10, dim=1)
            T = torch.cat([torch.ones_like(T[:, :1]), T[:, :-1]], dim=1)
            img[i:i + rays_per_chunk] = ((alpha * T)[..., None] * rgb).sum(1)
        out.append(img.reshape(H, W, 3).to(torch.float32))
    return torch.stack(out)
