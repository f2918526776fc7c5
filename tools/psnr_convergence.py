"""north_star check: PSNR after N training steps on a synthetic scene, CUDA path vs the fp32 oracle
(CPU autograd + Adam) on IDENTICAL ray batches, initial weights and deterministic depths.
    python tools/psnr_convergence.py [steps=2000] [rays=256] [samples=32]"""
import os, sys, time, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfmnerf_b200 as s
from oracle import nerf_oracle as orc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S = int(sys.argv[3]) if len(sys.argv) > 3 else 32
torch.set_num_threads(os.cpu_count())
torch.manual_seed(0)
m = s.NeRFmodel()
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.cuda()
W = 48
cam = s.synthetic.make_camera_info(W, W)
poses = s.synthetic.sphere_poses(13, seed=0)
images = s.synthetic.render_lego_images(poses, cam, n_sample=128, device="cuda", field=s.synthetic.dense_field).cpu().numpy()
train_idx, test_idx = np.arange(12), 12
rng = np.random.default_rng(0)
tr = s.engine.Trainer(m, None, None, cam, n_rays, S, perturb=False, lr=5e-4)
opt = orc.Adam(sd, lr=5e-4)
t0 = time.time()
for it in range(steps):
    k, x, y = rng.integers(0, 12, n_rays), rng.integers(0, W, n_rays), rng.integers(0, W, n_rays)
    rays = orc.rays_from_indices(images, poses, cam, k, x, y)
    o, d, t = (torch.tensor(rays[:, i:i + 3]) for i in (0, 3, 6))
    tr.step(rays=(o.cuda(), d.cuda(), t.cuda()))
    loss, grads = orc.loss_and_grads(sd, o, d, t, S)
    sd = opt.step(sd, grads)
    if it % 200 == 0:
        print(it, "cuda loss", float(tr.loss.item()), "oracle loss", float(loss), f"{time.time()-t0:.0f}s", flush=True)
tr.close()
args = argparse.Namespace(n_sample=S, perturb=False, chunk_size=32768, n_rays_batch=W * W, near=2.0, far=6.0)
psnr = lambda mse: -10 * np.log10(mse)
out = {"steps": steps, "rays": n_rays, "samples": S}
for name, idx in (("train_view", 0), ("held_out_view", test_idx)):
    _, mse_c = s.test_image(m, images[idx], poses[idx], cam, args)
    _, mse_o = orc.test_image(sd, images[idx], poses[idx], cam, S, W * W)
    out[name] = {"psnr_cuda": psnr(mse_c), "psnr_oracle": psnr(mse_o), "delta_db": psnr(mse_c) - psnr(mse_o)}
print(json.dumps(out, indent=1))
