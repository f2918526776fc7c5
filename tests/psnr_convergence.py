"""north_star check: PSNR after N training steps on a synthetic scene, CUDA path vs the fp32 oracle (CPU autograd + Adam)
on IDENTICAL ray batches, initial weights and deterministic depths.  Training is chaotic, so a second fp32 oracle
trajectory started from weights perturbed by 1e-7 (relative) measures how far two fp32 runs drift apart by themselves;
PSNR is averaged over several training views and several held-out views.
    python tests/psnr_convergence.py [steps=2000] [rays=512] [samples=48] [scene=textured|dense]"""
import os, sys, time, json, argparse, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfmnerf_b200 as s
from oracle import nerf_oracle as orc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 512
S = int(sys.argv[3]) if len(sys.argv) > 3 else 48
scene = sys.argv[4] if len(sys.argv) > 4 else "textured"
field = {"textured": s.synthetic.textured_field, "dense": s.synthetic.dense_field}[scene]
torch.set_num_threads(os.cpu_count())
torch.manual_seed(0)
m = s.NeRFmodel()
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
g = torch.Generator().manual_seed(123)
n_pert = int(os.environ.get("N_PERTURBED", 1))
sds = [{k: v * (1.0 + 1e-7 * torch.randn(v.shape, generator=g)) for k, v in sd.items()} for _ in range(n_pert)]   # "chaos floor" trajectories
sd2 = sds[0]
m = m.cuda()
W, n_train, n_test = 48, 12, 4
cam = s.synthetic.make_camera_info(W, W)
poses = s.synthetic.sphere_poses(n_train + n_test, seed=0)
images = s.synthetic.render_lego_images(poses, cam, n_sample=128, device="cuda", field=field).cpu().numpy()
rng = np.random.default_rng(0)
tr = s.engine.Trainer(m, None, None, cam, n_rays, S, perturb=False, lr=5e-4)
opt, opts = orc.Adam(sd, lr=5e-4), [orc.Adam(x, lr=5e-4) for x in sds]
m2 = s.NeRFmodel().cuda(); m2.load_state_dict(m.state_dict())                     # second CUDA run: same start, different atomics order
tr2 = s.engine.Trainer(m2, None, None, cam, n_rays, S, perturb=False, lr=5e-4)
t0 = time.time()
for it in range(steps):
    k, x, y = rng.integers(0, n_train, n_rays), rng.integers(0, W, n_rays), rng.integers(0, W, n_rays)
    rays = orc.rays_from_indices(images, poses, cam, k, x, y)
    o, d, t = (torch.tensor(rays[:, i:i + 3]) for i in (0, 3, 6))
    tr.step(rays=torch.tensor(np.asarray(rays, dtype=np.float32)).cuda())   # [N,9] array: nerf_train_step, CUDA-graph replay
    loss, grads = orc.loss_and_grads(sd, o, d, t, S)
    sd = opt.step(sd, grads)
    tr2.step(rays=(o.cuda(), d.cuda(), t.cuda()))                            # three tensors: the per-kernel sequence
    for j in range(n_pert):
        loss2, grads2 = orc.loss_and_grads(sds[j], o, d, t, S)
        sds[j] = opts[j].step(sds[j], grads2)
    if it % 250 == 0:
        print(it, "loss cuda", float(tr.loss.item()), "oracle", float(loss), "oracle'", float(loss2), f"{time.time()-t0:.0f}s", flush=True)
print("first CUDA trainer: cuda graphs", len(tr._graphs), tr.graph_error, "| second: per-kernel path, graphs", len(tr2._graphs), flush=True)
tr.close(); tr2.close()
args = argparse.Namespace(n_sample=S, perturb=False, chunk_size=32768, n_rays_batch=W * W, near=2.0, far=6.0)
psnr = lambda mse: -10 * np.log10(mse)
out = {"steps": steps, "rays": n_rays, "samples": S, "scene": scene, "views": {"train": 4, "held_out": n_test}}
for name, idxs in (("train_views", [0, 3, 6, 9]), ("held_out_views", list(range(n_train, n_train + n_test)))):
    mc = np.mean([s.test_image(m, images[i], poses[i], cam, args)[1] for i in idxs])
    mo = np.mean([orc.test_image(sd, images[i], poses[i], cam, S, W * W)[1] for i in idxs])
    mc2 = np.mean([s.test_image(m2, images[i], poses[i], cam, args)[1] for i in idxs])
    po = [psnr(mo)] + [psnr(np.mean([orc.test_image(x, images[i], poses[i], cam, S, W * W)[1] for i in idxs])) for x in sds]
    out[name] = {"psnr_cuda": [psnr(mc), psnr(mc2)], "psnr_oracle_fp32_trajectories": po,
                 "oracle_mean": float(np.mean(po)), "oracle_std": float(np.std(po, ddof=1)) if len(po) > 1 else None,
                 "delta_cuda_mean_vs_oracle_mean_db": float(np.mean([psnr(mc), psnr(mc2)]) - np.mean(po)),
                 "oracle_spread_db": float(np.max(po) - np.min(po))}
print(json.dumps(out, indent=1))
