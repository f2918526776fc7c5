"""Small end-to-end workload that launches every kernel of the library once, on ragged sizes -- written for
`compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py`.  (compute-sanitizer is closed on the
GPU pool this round was built on; the script runs clean on its own, and the bounds are covered by the ragged / unaligned /
empty-input parity tests in tests/test_gpu_parity.py.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfmnerf_b200 as s
E = s.engine
torch.manual_seed(0)
cam = s.synthetic.make_camera_info(20, 15)
poses_np = s.synthetic.sphere_poses(3, seed=1)
poses = E.poses_to_device(poses_np, "cuda")
images = torch.rand(3, 15, 20, 3, device="cuda")
m = s.NeRFmodel().cuda()
st = m.kernel_state()
with torch.no_grad():
    for S in (64, 48, 7):                                         # exact-length and generic compositing / sampling kernels
        rgb, depth, acc = E.render_image(st, poses, 1, cam, S, ray_begin=3, ray_count=37, perturb=True, seed=3, offset=5)
        o, d = E.raygen_image(poses, 0, cam, 0, 33)
        z = E.sample_z(33, S, 2.0, 6.0, "cuda", perturb=True, seed=1)
        sr = E.mlp_forward_rays(st, o, d, z, None)
        E.composite_backward(sr, z, torch.rand(33, 3, device="cuda"))
    enc = m.position_encoding(torch.rand(11, 3, device="cuda"), 10)
tr = E.Trainer(m, images, poses, cam, 96, 64, perturb=True, seed=2, graph=False)
for _ in range(2):
    tr.step()                                                     # nerf_train_step: batch, fwd, composite_train, dgrad, wgrad, heads, adam, pack
tr.step(timers={})                                                # the per-kernel sequence
tr.step(rays=torch.rand(96, 9, device="cuda"))
loss = float(tr.loss.item())
tr.close()
out = m(torch.rand(5, 3, device="cuda"), torch.rand(5, 3, device="cuda"))
(out[0].sum() + out[1].sum()).backward()                          # autograd path: nerf_mlp_backward
torch.cuda.synchronize()
print("sanitize_smoke ok", loss, float(rgb.sum()))
