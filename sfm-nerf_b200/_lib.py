"""ctypes binding of libnerf_b200.so (include/nerf_b200.h).

The shared library is the product: hand-written sm_100a kernels behind a plain C ABI.  This module
only locates/loads it, declares the prototypes and turns non-zero return codes into exceptions.
There is no CPU or eager-PyTorch fallback: if the library is missing or the device is not sm_100,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libnerf_b200.so")
DIAG_LIB_PATH = os.path.join(HERE, "libnerf_b200_diag.so")    # the same sources with -DNERF_DIAGNOSTICS (tools/ only)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

LAYOUT_LIVE_PE = 0
LAYOUT_LIVE_NOPE = 1
LAYOUT_LIVE_PE_RELU = 2       # same parameters, density = relu(.): remapped legacy checkpoints (legacy.py)
LAYOUT_LIVE_NOPE_RELU = 3


class NerfB200Error(RuntimeError):
    pass


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
        [os.path.join(os.path.dirname(HERE), "include", "nerf_b200.h")]


def build(force: bool = False, verbose: bool = False, diagnostics: bool = False) -> str:
    """Compile csrc/api.cu -> libnerf_b200.so in-tree with nvcc for sm_100a (cross-compiles
    without a GPU).  Rebuilds only when a source is newer than the library.  ``diagnostics``: the same sources with
    -DNERF_DIAGNOSTICS -> libnerf_b200_diag.so (clock-stamp trace of the fused MLP kernels, UMMA issue-rate microbenchmarks;
    used by tools/ through NERF_B200_LIB, never by the product path)."""
    out = DIAG_LIB_PATH if diagnostics else LIB_PATH
    if not force and os.path.exists(out):
        newest = max(os.path.getmtime(p) for p in _sources())
        if os.path.getmtime(out) >= newest:
            return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-DNERF_DIAGNOSTICS"] if diagnostics else []) + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", out, os.path.join(CSRC, "api.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise NerfB200Error("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_U = C.c_uint64
_F = C.c_float
_D = C.c_double



class NerfTrainStep(C.Structure):
    """Mirror of `struct NerfTrainStep` (include/nerf_b200.h); struct_bytes guards against drift."""
    _fields_ = [("struct_bytes", _L)] + [(n, C.c_int32) for n in ("layout", "n_sample", "perturb", "world", "rank", "n_images", "width", "height")] + \
        [("n_rays", _L), ("fx", _D), ("fy", _D), ("cx", _D), ("cy", _D), ("near", _F), ("far", _F), ("seed", _U), ("z_seed", _U),
         ("lr0", _D), ("lr_decay", _D), ("lr_decay_steps", _L), ("beta1", _D), ("beta2", _D), ("eps", _D), ("grad_scale", _F), ("pad_", _F)] + \
        [(n, _P) for n in ("poses", "images", "batch9", "rays_o", "rays_d", "target", "z", "sigma_rgb", "d_sigma_rgb", "loss",
                           "composite_ws", "stash", "dstash", "packed", "params", "grads", "exp_avg", "exp_avg_sq", "step_state")] + \
        [("peer_grads", _P * 8), ("multicast_grads", _P), ("peer_flags", _P * 8), ("exchange_epoch", _P)]


# name -> (restype, argtypes); must list every symbol include/nerf_b200.h declares
PROTOTYPES = {
    "nerf_abi_version": (_I, []),
    "nerf_error_string": (C.c_char_p, [_I]),
    "nerf_device_check": (_I, []),
    "nerf_layout_id": (_I, [_I, _I, _I, _I, _I]),
    "nerf_param_count": (_L, [_I]),
    "nerf_param_offset": (_L, [_I, _I]),
    "nerf_raygen_image": (_I, [_P, _I, _D, _D, _D, _D, _I, _I, _L, _L, _P, _P, _P]),
    "nerf_raygen_batch": (_I, [_P, _I, _D, _D, _D, _D, _I, _I, _P, _P, _P, _P, _L, _P, _P, _P, _P]),
    "nerf_draw_pixels": (_I, [_U, _U, _I, _I, _I, _L, _P, _P, _P, _P]),
    "nerf_train_batch": (_I, [_P, _I, _D, _D, _D, _D, _I, _I, _P, _U, _U, _L, _P, _P, _P, _I, _F, _F, _I, _U, _U, _P, _P]),
    "nerf_unpack_batch": (_I, [_P, _L, _P, _P, _P, _I, _F, _F, _I, _U, _U, _P, _P]),
    "nerf_sample_z": (_I, [_L, _I, _F, _F, _I, _P, _U, _U, _P, _P]),
    "nerf_encode": (_I, [_P, _L, _I, _P, _P]),
    "nerf_mlp_packed_bytes": (_L, [_I]),
    "nerf_mlp_pack_weights": (_I, [_P, _I, _P, _P]),
    "nerf_mlp_stash_bytes": (_L, [_I, _L]),
    "nerf_mlp_dstash_bytes": (_L, [_I, _L]),
    "nerf_mlp_forward": (_I, [_P, _I, _P, _P, _L, _P, _P, _P]),
    "nerf_mlp_forward_rays": (_I, [_P, _I, _P, _P, _P, _L, _I, _P, _P, _P]),
    "nerf_mlp_render_rays": (_I, [_P, _I, _P, _P, _P, _L, _I, _P, _P, _P, _P]),
    "nerf_mlp_backward": (_I, [_P, _I, _P, _P, _P, _L, _P, _P, _I, _P]),
    "nerf_mlp_dgrad": (_I, [_P, _I, _P, _P, _P, _L, _P, _P]),
    "nerf_mlp_wgrad": (_I, [_I, _P, _P, _L, _P, _I, _P]),
    "nerf_composite_forward": (_I, [_P, _P, _L, _I, _P, _P, _P, _P]),
    "nerf_composite_backward": (_I, [_P, _P, _P, _L, _I, _P, _P]),
    "nerf_adam_step_reduce": (_I, [_P, _P, _I, _P, _P, _P, _L, _D, _D, _D, _D, _L, _F, _P, _I, _P, _P]),
    "nerf_composite_train_workspace_bytes": (_L, [_L]),
    "nerf_composite_train": (_I, [_P, _P, _P, _L, _I, _F, _P, _P, _P, _P]),
    "nerf_mse_loss": (_I, [_P, _P, _L, _F, _P, _P, _P]),
    "nerf_adam_step": (_I, [_P, _P, _P, _P, _L, _D, _D, _D, _D, _L, _F, _P]),
    "nerf_render_workspace_bytes": (_L, [_L, _I]),
    "nerf_render_image": (_I, [_P, _I, _P, _I, _D, _D, _D, _D, _I, _I, _L, _L, _I, _F, _F, _I, _U, _U, _P, _P, _P, _P, _P]),
    "nerf_train_step": (_I, [C.POINTER(NerfTrainStep), _I, _P]),
    "nerf_selftest_umma": (_I, [_I, _P]),
}
# only exported by libnerf_b200_diag.so (-DNERF_DIAGNOSTICS)
DIAG_PROTOTYPES = {
    "nerf_debug_umma_rate": (_I, [_I, _I, _P]),
    "nerf_debug_trace": (_I, [_P]),
    "nerf_debug_store_rate": (_I, [_I, _I, _I, _P, C.c_longlong, _P]),
}

_lib = None


def load():
    """Load the library (raises if it has not been built) and declare prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfB200Error(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / eager fallback for the sm_100a kernels)")
    lib = C.CDLL(os.environ.get("NERF_B200_LIB", LIB_PATH))     # override: A/B runs of two builds of the same ABI
    for name, (res, args) in list(PROTOTYPES.items()) + list(DIAG_PROTOTYPES.items()):
        if name in DIAG_PROTOTYPES and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def error_string(code: int) -> str:
    return load().nerf_error_string(int(code)).decode()


def check(code: int, what: str = ""):
    if code != 0:
        raise NerfB200Error(f"{what or 'libnerf_b200'} failed with code {code}: {error_string(code)}")


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero code."""
    check(getattr(load(), name)(*args), name)


def ptr(t):
    """Device (or host) address of a tensor / None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
