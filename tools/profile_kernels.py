"""Launches each MLP kernel a few times at the C2 shape (4096 rays x 64) for ncu captures:
order per iteration: fwd(infer), fwd(train), dgrad, wgrad."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sfmnerf_b200 as s
E = s.engine
n_rays, S = int(os.environ.get("RAYS", 4096)), 64
torch.manual_seed(0)
m = s.NeRFmodel().cuda()
st = m.kernel_state()
o = torch.zeros(n_rays, 3, device="cuda"); o[:, 2] = -4
d = torch.nn.functional.normalize(torch.randn(n_rays, 3, device="cuda") * 0.2 + torch.tensor([0, 0, 1.0], device="cuda"), dim=-1)
z = E.sample_z(n_rays, S, 2.0, 6.0, "cuda", perturb=True, seed=1)
rows = n_rays * S
stash, dstash = E.alloc_stash(st, rows), E.alloc_dstash(st, rows)
grads = torch.zeros(st.n_params, device="cuda")
g = torch.randn(n_rays, 3, device="cuda")
call, ptr, sp = s._lib.call, s._lib.ptr, s._lib.stream_ptr
for it in range(int(os.environ.get("ITERS", 3))):
    sr = E.mlp_forward_rays(st, o, d, z, None)
    sr = E.mlp_forward_rays(st, o, d, z, stash)
    dsr = E.composite_backward(sr, z, g)
    call("nerf_mlp_dgrad", ptr(st.packed), st.layout, ptr(stash), ptr(sr), ptr(dsr), rows, ptr(dstash), sp())
    call("nerf_mlp_wgrad", st.layout, ptr(stash), ptr(dstash), rows, ptr(grads), 0, sp())
torch.cuda.synchronize()
print("ok", float(sr.sum()), float(grads.abs().sum()))
