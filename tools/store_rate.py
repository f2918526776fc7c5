"""Store-rate microbenchmark (nerf_debug_store_rate, -DNERF_DIAGNOSTICS build): bytes per clock per SM that W warps of one CTA
per SM can stream to HBM with the stash store pattern (512 contiguous bytes per warp instruction)."""
import os, sys, ctypes, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NERF_B200_LIB", importlib.import_module("sfm-nerf_b200._lib").build(diagnostics=True))
import torch
import sfmnerf_b200 as s
lib = s._lib.load()
buf = torch.empty(6 << 30, dtype=torch.uint8, device="cuda")
sms = torch.cuda.get_device_properties(0).multi_processor_count
for mode, name in ((0, "st.global.cs from smem"), (1, "st.global from smem"), (2, "st.global.cs from registers"),
                   (3, "64 KB blocks interleaved over CTAs"), (4, "2 KB pieces interleaved over CTAs")):
    for warps in ((1, 2, 4, 8, 12, 16, 32) if mode < 3 else (4, 8, 16)):
        iters = (32 << 20) // (warps * 512)            # 32 MB per CTA
        out = (ctypes.c_longlong * 3)()
        rc = lib.nerf_debug_store_rate(warps, iters, mode, ctypes.c_void_p(buf.data_ptr()), buf.numel(), out)
        tot = sms * out[1]
        print(f"{name:30s} warps {warps:2d}: rc={rc} {out[1] / max(out[0], 1):6.1f} B/clk/SM  {tot / max(out[2], 1):7.1f} GB/s", flush=True)
