"""UMMA issue-rate microbenchmark (nerf_debug_umma_rate): cycles per 128x256x16 MMA for different issue patterns."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
os.environ.setdefault("NERF_B200_LIB", importlib.import_module("sfm-nerf_b200._lib").build(diagnostics=True))   # -DNERF_DIAGNOSTICS build
import torch
import sfmnerf_b200 as s
lib = s._lib.load()
names = ["1 thread, 1 accumulator", "1 thread, 2 accumulators", "2 threads, 1 accumulator each", "1 thread, N=128 halves", "2 threads, same accumulator"]
names += ["PAIR 1 thread, 1 accumulator", "PAIR 1 thread, 2 accumulators", "PAIR 2 threads, 1 acc each", "PAIR 1 thread, commit per MMA", "PAIR 2 threads, commit per MMA"]
names += ["PAIR flags=%d (1 mbar wait, 4 pollers, 8 fence, 16 smem-word poll, 32 MMA-wait-commit order)" % f for f in range(64)]
for v in  [10 + f for f in (0, 1, 16, 32)]:
    for n in (256,):
        out = (ctypes.c_longlong * 2)()
        rc = lib.nerf_debug_umma_rate(v, n, out)
        per = out[0] / (n if v != 3 else n / 2)
        print(f"variant {v} ({names[v]:56s}) n={n:4d}: rc={rc} {out[0]:7d} cycles  -> {per:6.1f} cycles per 128x256x16")
