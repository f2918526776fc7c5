"""CPU oracle for the Phase2 NeRF render / train-step hot path.

TEST INFRASTRUCTURE ONLY.  This file restates, op for op, the reference's PyTorch
algorithm (ColinBalfour/SfM-NeRF, ``Phase2/train.py`` and ``Phase2/NeRFModel.py``) on the
CPU in fp32/fp64.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or as the
CPU arm being timed -- never as the shipped path.  The product (``sfm-nerf_b200``) does not
import anything from ``oracle/`` and raises when its CUDA library is missing.

Pinning: the reference publishes no golden vectors or tests, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container
by ``tests/golden/make_golden.py`` (which imports ``/root/reference/Phase2/train.py``
unmodified) and committed under ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this
file against those vectors, and -- when ``/root/reference`` is present -- against the live
reference as well.  Depth/acc outputs do not exist upstream (``render`` returns RGB only,
train.py:269-271); they are derived here from the reference's own ``weights`` (train.py:266)
and are therefore pinned by restatement only.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
import random
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# Parameter layout (NeRFModel.py:102-140) -- names, order and shapes of the live class.
# ----------------------------------------------------------------------------------------

PARAM_ORDER = (
    "fc1",
    "fc_layers.0", "fc_layers.1", "fc_layers.2", "fc_layers.3",
    "skip_layer",
    "fc_layers2.0", "fc_layers2.1", "fc_layers2.2",
    "density_layer", "feature_layer", "rgb_layer1", "rgb_layer2",
)


def layer_shapes(embed_pos_L=10, embed_direction_L=4, num_channels=256, pos_encoding=True):
    """(name, out_features, in_features) in state_dict order (NeRFModel.py:112-140)."""
    pos_in = 3 * (2 * embed_pos_L + 1) if pos_encoding else 3
    dir_in = 3 * (2 * embed_direction_L + 1) if pos_encoding else 3
    C = num_channels
    return [
        ("fc1", C, pos_in),
        ("fc_layers.0", C, C), ("fc_layers.1", C, C), ("fc_layers.2", C, C), ("fc_layers.3", C, C),
        ("skip_layer", C, C + pos_in),
        ("fc_layers2.0", C, C), ("fc_layers2.1", C, C), ("fc_layers2.2", C, C),
        ("density_layer", 1, C),
        ("feature_layer", C, C),
        ("rgb_layer1", 128, C + dir_in),
        ("rgb_layer2", 3, 128),
    ]


def init_state_dict(seed=0, **kw):
    """Random-init weights with torch's default ``nn.Linear`` init, drawn in the same order
    as ``NeRFmodel.__init__`` constructs its layers (NeRFModel.py:116-140), so that
    ``torch.manual_seed(seed); NeRFmodel(**kw)`` and this function agree bit for bit."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    for name, out_f, in_f in layer_shapes(**kw):
        lin = torch.nn.Linear(in_f, out_f)
        sd[name + ".weight"] = lin.weight.detach().clone()
        sd[name + ".bias"] = lin.bias.detach().clone()
    return sd


# ----------------------------------------------------------------------------------------
# (a1)-(a3) Ray generation -- train.py:68-102, 104-127, 735-742.  fp64 numpy, cast once.
# ----------------------------------------------------------------------------------------

def pixel_to_ray(camera_info, pose, pixel):
    """train.py:68-102: d = normalize(R @ [(px-cx)/fx, (py-cy)/fy, 1]), o = pose[:3, 3]."""
    px, py = pixel
    K = camera_info["camera_matrix"]
    fx, fy, cx, cy = K[0][0], K[1][1], K[0][2], K[1][2]
    p = np.array([(px - cx) / fx, (py - cy) / fy, 1.0])
    d = np.dot(pose[:3, :3], p)
    d = d / np.linalg.norm(d)
    return pose[:3, 3], d


def image_rays(camera_info, pose):
    """train.py:735-742: every pixel of the image, ray index = y*W + x, float32 [H*W,3] x2."""
    W, H = int(camera_info["width"]), int(camera_info["height"])
    o = np.empty((H * W, 3), np.float64)
    d = np.empty((H * W, 3), np.float64)
    i = 0
    for y in range(H):
        for x in range(W):
            o[i], d[i] = pixel_to_ray(camera_info, pose, (x, y))
            i += 1
    return o.astype(np.float32), d.astype(np.float32)


def rays_from_indices(images, poses, camera_info, img_idx, px, py):
    """train.py:119-127 for pre-drawn (image, x, y) triples -> float32 [N,9] = o|d|rgb."""
    out = np.empty((len(img_idx), 9), np.float64)
    for i, (k, x, y) in enumerate(zip(img_idx, px, py)):
        o, d = pixel_to_ray(camera_info, poses[k], (int(x), int(y)))
        out[i, 0:3], out[i, 3:6], out[i, 6:9] = o, d, images[k][int(y), int(x)]
    return out.astype(np.float32)


def draw_batch_indices(n_rays, n_images, width, height, rng=random):
    """The three ``random.randint`` draws per ray, in the order of train.py:120-121."""
    img_idx, px, py = [], [], []
    for _ in range(n_rays):
        img_idx.append(rng.randint(0, n_images - 1))
        px.append(rng.randint(0, width - 1))
        py.append(rng.randint(0, height - 1))
    return np.array(img_idx), np.array(px), np.array(py)


def generate_batch(images, poses, camera_info, n_rays_batch, rng=random):
    """train.py:104-127."""
    W, H = int(camera_info["width"]), int(camera_info["height"])
    k, x, y = draw_batch_indices(n_rays_batch, len(images), W, H, rng)
    return rays_from_indices(images, poses, camera_info, k, x, y)


# ----------------------------------------------------------------------------------------
# (a4) Stratified depths -- train.py:205-217.
# ----------------------------------------------------------------------------------------

def sample_z(n_rays, n_sample, near, far, t_rand=None):
    """t = linspace(0,1,S); optional jitter inside mid-point bins with the supplied uniform
    draws ``t_rand`` [N,S] (train.py:209-214 draws them with torch.rand); z = near+(far-near)t."""
    t_vals = torch.linspace(0, 1, n_sample).expand(n_rays, n_sample)
    if t_rand is not None:
        mids = 0.5 * (t_vals[..., 1:] + t_vals[..., :-1])
        upper = torch.cat([mids, t_vals[..., -1:]], -1)
        lower = torch.cat([t_vals[..., :1], mids], -1)
        t_vals = lower + (upper - lower) * t_rand
    return near + (far - near) * t_vals


# ----------------------------------------------------------------------------------------
# (a6) Positional encoding -- NeRFModel.py:142-156.
# ----------------------------------------------------------------------------------------

def position_encoding(x, L):
    """[x, sin(2^0 pi x), cos(2^0 pi x), ..., sin(2^(L-1) pi x), cos(2^(L-1) pi x)]."""
    enc = [x]
    for i in range(L):
        freq = 2.0 ** i
        for fn in (torch.sin, torch.cos):
            enc.append(fn(freq * np.pi * x))
    return torch.cat(enc, dim=-1)


# ----------------------------------------------------------------------------------------
# (a7) MLP -- NeRFModel.py:158-202, functional over a state_dict.
# ----------------------------------------------------------------------------------------

def _round_st(x, dtype):
    """Round to ``dtype`` and back with a straight-through gradient (identity when dtype is None)."""
    if dtype is None:
        return x
    return x + (x.to(dtype).float() - x).detach()


def model_forward(sd, pos, direction, embed_pos_L=10, embed_direction_L=4, pos_encoding=True,
                  return_hidden=False, operand_dtype=None, density_activation="softplus"):
    """density [...,1] (softplus), rgb [...,3] (sigmoid).  ``return_hidden`` also returns the
    post-activation tensors in layer order (used by per-layer parity tests).

    ``operand_dtype=torch.bfloat16`` restates the SAME network with the operand rounding of a
    bf16-tensor-core evaluation (weights, biases and inputs of the eleven 128/256-wide layers rounded
    to bf16, fp32 accumulation; the fp32 density / rgb heads read the bf16-rounded activations).  It is NOT reference behaviour; tests use it
    to separate "kernel is wrong" from "bf16 operands differ from fp32" when comparing gradients,
    whose element-wise error is dominated by ReLU masks flipping for near-zero pre-activations."""
    q = lambda x: _round_st(x, operand_dtype)
    lin = lambda name, x: F.linear(x, q(sd[name + ".weight"]), q(sd[name + ".bias"]))   # bias rides on the MMA too
    head = lambda name, x: F.linear(x, sd[name + ".weight"], sd[name + ".bias"])
    if pos_encoding:
        pe = position_encoding(pos, embed_pos_L)
        de = position_encoding(direction, embed_direction_L)
    else:
        pe, de = pos, direction
    pe, de = q(pe), q(de)
    hidden = []
    h = q(F.relu(lin("fc1", pe))); hidden.append(h)
    for i in range(4):
        h = q(F.relu(lin(f"fc_layers.{i}", h))); hidden.append(h)
        if i == 3:
            h = torch.cat([h, pe], dim=-1)
            h = q(F.relu(lin("skip_layer", h))); hidden.append(h)
    for i in range(3):
        h_full = F.relu(lin(f"fc_layers2.{i}", h))
        h = q(h_full); hidden.append(h)
    dens_act = F.softplus if density_activation == "softplus" else F.relu    # "relu": remapped legacy checkpoints
    density = dens_act(head("density_layer", h))               # head reads the (rounded) activation tile
    feature = q(F.relu(lin("feature_layer", h))); hidden.append(feature)
    r = q(F.relu(lin("rgb_layer1", torch.cat([feature, de], dim=-1)))); hidden.append(r)
    rgb = torch.sigmoid(head("rgb_layer2", r))
    if return_hidden:
        return density, rgb, hidden
    return density, rgb


# ----------------------------------------------------------------------------------------
# Legacy checkpoint layouts: the commented-out class NeRFModel.py:7-99 (LEGACY_PE) and the layout of
# Phase2/checkpoints_no_pos_encoding/final_model.pth (LEGACY_NOPE: no source, by analogy).  ReLU after every Linear of the
# trunk including the 257-wide one; density = channel 0, feature = channels 1..256; density_fc unused.
# ----------------------------------------------------------------------------------------

def legacy_init_state_dict(kind, seed=0):
    """Random legacy-format state_dict: nn.Linear default init, modules created in the order of NeRFModel.py:24-59
    (fc1, block1, skip_layer, block2, density_fc, rgb_fc), so torch.manual_seed(seed) reproduces the commented class."""
    torch.manual_seed(seed)
    sd = OrderedDict()

    def lin(name, i, o):
        m = torch.nn.Linear(i, o)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach().clone(), m.bias.detach().clone()

    if kind == "pe":
        lin("fc1.0", 63, 256)
        for k in (0, 2): lin(f"block1.{k}", 256, 256)
        lin("block1.4", 256, 256)
        lin("skip_layer.0", 319, 256)
        for k in (0, 2): lin(f"block2.{k}", 256, 256)
        lin("block2.4", 256, 257)
        lin("density_fc", 256, 1)
        lin("rgb_fc.0", 283, 128); lin("rgb_fc.2", 128, 3)
    else:
        lin("fc1.0", 3, 256)
        for k in (0, 2, 4, 6, 8, 10): lin(f"block1.{k}", 256, 256)
        lin("block1.12", 256, 257)
        lin("density_fc", 256, 1)
        lin("rgb_fc.0", 259, 128); lin("rgb_fc.2", 128, 3)
    return sd


def legacy_forward(sd, pos, direction):
    """(density [...,1], rgb [...,3]) of a legacy state_dict -- NeRFModel.py:74-99 (commented) for the PE layout."""
    pe_kind = "skip_layer.0.weight" in sd
    lin = lambda name, x: F.linear(x, sd[name + ".weight"], sd[name + ".bias"])
    if pe_kind:
        p, d = position_encoding(pos, 10), position_encoding(direction, 4)
        out = F.relu(lin("fc1.0", p))
        for k in (0, 2, 4):
            out = F.relu(lin(f"block1.{k}", out))
        out = F.relu(lin("skip_layer.0", torch.cat([out, p], dim=-1)))
        for k in (0, 2, 4):
            out = F.relu(lin(f"block2.{k}", out))
    else:
        p, d = pos, direction
        out = F.relu(lin("fc1.0", p))
        for k in (0, 2, 4, 6, 8, 10, 12):
            out = F.relu(lin(f"block1.{k}", out))
    density = out[..., 0:1]
    feature = out[..., 1:]
    h = F.relu(lin("rgb_fc.0", torch.cat([feature, d], dim=-1)))
    rgb = torch.sigmoid(lin("rgb_fc.2", h))
    return density, rgb


# ----------------------------------------------------------------------------------------
# (a9) Compositing -- train.py:248-269 (+ depth/acc, which the reference does not output).
# ----------------------------------------------------------------------------------------

def composite(sigma, rgb, z_vals):
    """sigma [N,S], rgb [N,S,3], z [N,S] -> dict(rgb [N,3], depth [N], acc [N], weights [N,S])."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.ones_like(dists[..., :1]) * 1e-3], -1)
    alpha = 1.0 - torch.exp(-sigma * dists)
    T = torch.ones_like(alpha)
    cumprod_term = torch.cumprod(1.0 - alpha + 1e-10, dim=1)
    T = torch.cat([T[:, :1], cumprod_term[:, :-1]], dim=1)   # == T[:,1:] = cumprod[:,:-1] (261-263)
    weights = alpha * T
    out_rgb = torch.sum(weights.unsqueeze(-1) * rgb, dim=1)
    depth = torch.sum(weights * z_vals, dim=1)               # extension, see header
    acc = torch.sum(weights, dim=1)                          # extension, see header
    return {"rgb": out_rgb, "depth": depth, "acc": acc, "weights": weights}


def composite_backward(sigma, rgb, z_vals, g_rgb):
    """Closed-form backward of ``composite`` w.r.t. sigma and rgb for a cotangent g_rgb [N,3]
    (SURVEY.md A.1 step 7); used to check the CUDA backward without autograd."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.ones_like(dists[..., :1]) * 1e-3], -1)
    alpha = 1.0 - torch.exp(-sigma * dists)
    om = 1.0 - alpha + 1e-10
    T = torch.cat([torch.ones_like(alpha[:, :1]), torch.cumprod(om, dim=1)[:, :-1]], dim=1)
    w = alpha * T
    gc = (g_rgb.unsqueeze(1) * rgb).sum(-1)                    # g . c_i
    wgc = w * gc
    suffix = torch.flip(torch.cumsum(torch.flip(wgc, [1]), 1), [1]) - wgc   # sum_{j>i}
    d_alpha = T * gc - suffix / om
    d_sigma = dists * (1.0 - alpha) * d_alpha
    d_rgb = w.unsqueeze(-1) * g_rgb.unsqueeze(1)
    return d_sigma, d_rgb


# ----------------------------------------------------------------------------------------
# (a4)-(a9) render -- train.py:187-271.
# ----------------------------------------------------------------------------------------

def render_full(sd, rays_o, rays_d, n_sample, near=2.0, far=6.0, t_rand=None, chunk_size=32768,
                **model_kw):
    """Restated ``render``; returns dict with rgb/depth/acc/weights/sigma/rgb_s/z."""
    N = rays_o.shape[0]
    z_vals = sample_z(N, n_sample, near, far, t_rand)
    ro = rays_o.unsqueeze(1).expand(-1, n_sample, -1)
    rd = rays_d.unsqueeze(1).expand(-1, n_sample, -1)
    pts = ro + rd * z_vals.unsqueeze(-1)
    pts_flat, rd_flat = pts.reshape(-1, 3), rd.reshape(-1, 3)
    chunk = min(pts_flat.shape[0], chunk_size)
    dens, cols = [], []
    for i in range(0, pts_flat.shape[0], chunk):
        dd, cc = model_forward(sd, pts_flat[i:i + chunk], rd_flat[i:i + chunk], **model_kw)
        dens.append(dd); cols.append(cc)
    sigma = torch.cat(dens, 0).reshape(N, n_sample)
    rgb_s = torch.cat(cols, 0).reshape(N, n_sample, 3)
    out = composite(sigma, rgb_s, z_vals)
    out.update(sigma=sigma, rgb_s=rgb_s, z=z_vals)
    return out


def render(sd, rays_o, rays_d, n_sample, near=2.0, far=6.0, t_rand=None, **kw):
    return render_full(sd, rays_o, rays_d, n_sample, near, far, t_rand, **kw)["rgb"]


# ----------------------------------------------------------------------------------------
# (a10) loss / backward / Adam / StepLR -- train.py:401-404, 456-465.
# ----------------------------------------------------------------------------------------

def loss_and_grads(sd, rays_o, rays_d, target, n_sample, near=2.0, far=6.0, t_rand=None, **kw):
    """MSE over N*3 (train.py:460) and autograd gradients of the 26 parameter tensors."""
    p = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in sd.items())
    rgb = render(p, rays_o, rays_d, n_sample, near, far, t_rand, **kw)
    loss = torch.mean((rgb - target) ** 2)
    grads = torch.autograd.grad(loss, list(p.values()), allow_unused=True)
    return loss.detach(), OrderedDict(
        (k, (g if g is not None else torch.zeros_like(v))) for (k, v), g in zip(p.items(), grads))


class Adam:
    """torch.optim.Adam defaults as used at train.py:401 (betas .9/.999, eps 1e-8, no weight
    decay), written out; StepLR(step_size, gamma) as at train.py:404."""

    def __init__(self, sd, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, step_size=50000, gamma=0.1):
        self.lr0, self.betas, self.eps = lr, betas, eps
        self.step_size, self.gamma = step_size, gamma
        self.t = 0
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in sd.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in sd.items())

    def lr(self):
        return self.lr0 * self.gamma ** (self.t // self.step_size)

    def step(self, sd, grads):
        lr = self.lr()
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1 ** self.t
        bc2 = 1.0 - b2 ** self.t
        for k in sd:
            g = grads[k]
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            sd[k] = sd[k] - (lr / bc1) * (self.m[k] / denom)
        return sd


def test_image(sd, image, pose, camera_info, n_sample, n_rays_batch, near=2.0, far=6.0, **kw):
    """train.py:714-763 with deterministic depths: (pred [H,W,3] float32, mse float)."""
    W, H = int(camera_info["width"]), int(camera_info["height"])
    o, d = image_rays(camera_info, pose)
    o, d = torch.tensor(o), torch.tensor(d)
    pred = torch.zeros((H * W, 3))
    with torch.no_grad():
        for i in range(0, H * W, n_rays_batch):
            pred[i:i + n_rays_batch] = render(sd, o[i:i + n_rays_batch], d[i:i + n_rays_batch],
                                              n_sample, near, far, **kw)
    gt = torch.tensor(image.reshape(-1, 3))
    mse = torch.nn.MSELoss()(gt, pred)
    return pred.reshape(H, W, 3), mse.item()


# ----------------------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench.  Not reference code:
# the reference has no dataset in-tree, so poses/scenes are generated in PixelToRay's
# convention (+z forward, +y down, no axis flip; train.py:87-94).
# ----------------------------------------------------------------------------------------

CAMERA_ANGLE_X = 0.6911112070083618


def make_camera_info(width, height, camera_angle_x=CAMERA_ANGLE_X):
    """Reference focal formula, train.py:53 (note the extra camera_angle_x factor)."""
    focal = 0.5 * camera_angle_x * width / np.tan(0.5 * camera_angle_x)
    K = np.array([[focal, 0, width / 2], [0, focal, height / 2], [0, 0, 1]])
    return {"width": width, "height": height, "camera_matrix": K}


def look_at_pose(theta_deg, phi_deg, radius=4.0):
    """OpenCV-convention camera-to-world matrix looking at the origin from a sphere."""
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    C = radius * np.array([np.cos(ph) * np.sin(th), np.sin(ph), np.cos(ph) * np.cos(th)])
    z = -C / np.linalg.norm(C)
    up = np.array([0.0, -1.0, 0.0])
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = x, y, z, C
    return pose


def sphere_poses(n, seed=0, radius=4.0):
    rng = np.random.default_rng(seed)
    return np.stack([look_at_pose(rng.uniform(0, 360), rng.uniform(-60, -10), radius)
                     for _ in range(n)])
