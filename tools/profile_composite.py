"""Launches the compositing kernels alone at 640,000 rays x 64 samples for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sfmnerf_b200 as s
E = s.engine
n, S = int(os.environ.get("RAYS", 640000)), int(os.environ.get("SAMPLES", 64))
sr = torch.rand(n, S, 4, device="cuda"); sr[..., 0] *= 4.0
z = E.sample_z(n, S, 2.0, 6.0, "cuda", perturb=True, seed=9)
g = torch.rand(n, 3, device="cuda")
for _ in range(3):
    E.composite_forward(sr, z, extras=True)
    E.composite_backward(sr, z, g)
torch.cuda.synchronize()
print("ok")
