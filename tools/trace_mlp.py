"""Per-layer timeline of CTA 0 of the fused MLP kernel (clock64 stamps via nerf_debug_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
os.environ.setdefault("NERF_B200_LIB", importlib.import_module("sfm-nerf_b200._lib").build(diagnostics=True))   # -DNERF_DIAGNOSTICS build
import torch
import sfmnerf_b200 as s
E = s.engine
n_rays, S = 4096, 64
torch.manual_seed(0)
m = s.NeRFmodel().cuda(); st = m.kernel_state()
o = torch.zeros(n_rays, 3, device="cuda"); o[:, 2] = -4
d = torch.nn.functional.normalize(torch.randn(n_rays, 3, device="cuda") * 0.2 + torch.tensor([0, 0, 1.0], device="cuda"), dim=-1)
z = E.sample_z(n_rays, S, 2.0, 6.0, "cuda", perturb=True, seed=1)
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
stash = E.alloc_stash(st, n_rays * S) if mode != "infer" else None
for _ in range(3): E.mlp_forward_rays(st, o, d, z, stash)
buf = torch.zeros(8 * 12 * 2 * 4 + 576 + 288, dtype=torch.int64, device="cuda")
s._lib.call("nerf_debug_trace", s._lib.ptr(buf))
sr = E.mlp_forward_rays(st, o, d, z, stash)
if mode == "dgrad":
    g = torch.randn(n_rays, 3, device="cuda"); dsr = E.composite_backward(sr, z, g)
    dst = E.alloc_dstash(st, n_rays * S); buf.zero_()
    s._lib.call("nerf_mlp_dgrad", s._lib.ptr(st.packed), st.layout, s._lib.ptr(stash), s._lib.ptr(sr), s._lib.ptr(dsr), n_rays * S, s._lib.ptr(dst), s._lib.stream_ptr())
torch.cuda.synchronize()
s._lib.call("nerf_debug_trace", None)
full = buf.cpu()
t = full[:768].reshape(8, 12, 2, 4)
t0 = int(t[t > 0].min())
print("mode", mode, "(cycles relative to first stamp; per step: slot: issue_start issue_end acc_seen released | issue_len epi_len)")
for pair in (0, 1, 3):
    for si in range(11 if mode != "dgrad" else 10):
        row = []
        for slot in range(2):
            a, b, c, e = (int(x) - t0 for x in t[pair, si, slot])
            row.append(f"s{slot}: {a:7d} {b:7d} {c:7d} {e:7d} | {b-a:5d} {e-c:5d}")
        print(f"p{pair} L{si:2d}  " + "   ".join(row))

g = t[1]  # steady-state iteration
import numpy as np
starts = np.array([int(g[si, 0, 0]) for si in range(1, 9)])
print("steady-state period per layer (cycles, slot 0 issue_start deltas):", np.diff(starts).tolist())
