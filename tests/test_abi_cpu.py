"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/nerf_b200.h
declares (no compute calls -- there is no GPU here), the parameter layout matches the reference
class, and the drop-in Python surface has the reference's names and defaults."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libmod():
    mod = importlib.import_module("sfm-nerf_b200._lib")
    mod.build()
    return mod


def test_header_symbols_are_exported(libmod):
    header = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
    product, diag = re.split(r"#ifdef NERF_DIAGNOSTICS", header, maxsplit=1)
    diag = diag.split("#endif", 1)[0]
    declared = set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", product))
    declared_diag = set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", diag))
    assert len(declared) >= 20 and declared_diag == set(libmod.DIAG_PROTOTYPES)
    lib = ctypes.CDLL(libmod.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in nerf_b200.h but not exported"
    assert declared == set(libmod.PROTOTYPES), declared ^ set(libmod.PROTOTYPES)
    # the product library exports exactly the declared C ABI: no diagnostics, nothing undeclared
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", libmod.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("nerf_")}
    assert exported == declared, exported ^ declared
    # the diagnostics build (-DNERF_DIAGNOSTICS) adds the nerf_debug_* entry points
    dlib = ctypes.CDLL(libmod.build(diagnostics=True))
    for name in declared | declared_diag:
        assert hasattr(dlib, name), name


def test_two_threads_first_calls_do_not_race(libmod):
    """Header promise: re-entrant and thread-safe.  Two host threads make their FIRST calls into a freshly loaded copy of the
    library at the same time (layout tables are built on first use); every query returns the right value in both."""
    import shutil, tempfile, threading
    with tempfile.TemporaryDirectory() as d:
        for trial in range(4):
            path = os.path.join(d, f"libnerf_copy{trial}.so")          # a new path = a new image with uninitialised statics
            shutil.copyfile(libmod.LIB_PATH, path)
            lib = ctypes.CDLL(path)
            lib.nerf_param_count.restype = ctypes.c_int64
            lib.nerf_param_offset.restype = ctypes.c_int64
            lib.nerf_mlp_stash_bytes.restype = ctypes.c_int64
            lib.nerf_mlp_stash_bytes.argtypes = [ctypes.c_int, ctypes.c_int64]
            go = threading.Barrier(2)
            res = [None, None]

            def worker(i):
                go.wait()
                res[i] = (lib.nerf_param_count(i), lib.nerf_param_count(1 - i), lib.nerf_param_offset(0, 25), lib.nerf_mlp_stash_bytes(i, 128))
            ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
            for t in ts: t.start()
            for t in ts: t.join()
            assert res[0][0] == 661636 and res[0][1] == 627844 and res[1][0] == 627844 and res[1][1] == 661636
            assert res[0][2] == res[1][2] == 661636 - 3 and res[0][3] == res[1][3] > 0


def test_train_step_struct_matches_header(libmod, tmp_path):
    """The ctypes mirror of NerfTrainStep has the size and field offsets gcc gives the header's struct."""
    import subprocess
    fields = [f[0] for f in libmod.NerfTrainStep._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nerf_b200.h"\nint main(void){printf("%zu\\n", sizeof(NerfTrainStep));\n' +
                   "".join(f'printf("%zu\\n", offsetof(NerfTrainStep, {f}));\n' for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(libmod.NerfTrainStep)
    assert out[1:] == [getattr(libmod.NerfTrainStep, f).offset for f in fields]


def test_abi_queries_without_gpu(libmod):
    lib = libmod.load()
    assert lib.nerf_abi_version() == 4
    assert lib.nerf_param_count(0) == 661636 and lib.nerf_param_count(1) == 627844
    assert lib.nerf_param_count(2) == 661636 and lib.nerf_param_count(3) == 627844      # ReLU-density (legacy checkpoint) layouts
    assert lib.nerf_param_count(4) == -4
    assert b"sm_100" in lib.nerf_error_string(-3)
    # tile rounding of the stash queries
    assert lib.nerf_mlp_stash_bytes(0, 1) == lib.nerf_mlp_stash_bytes(0, 128)
    assert lib.nerf_mlp_stash_bytes(0, 129) == 2 * lib.nerf_mlp_stash_bytes(0, 128)
    assert lib.nerf_mlp_stash_bytes(0, 0) == 0
    assert lib.nerf_mlp_packed_bytes(0) % 16 == 0


def test_param_offsets_follow_state_dict_order(libmod):
    import sfmnerf_b200 as s
    lib = libmod.load()
    cases = [(0, {}), (1, {"pos_encoding": False})]
    for pl, dl, ch in ((6, 2, 256), (10, 10, 256), (0, 3, 256), (4, 4, 256), (10, 4, 128), (2, 9, 128)):   # --n_pos_freq / --n_dirc_freq / --hidden_size
        lid = lib.nerf_layout_id(pl, dl, 1, 0, ch)
        assert lid > 3
        cases.append((lid, {"embed_pos_L": pl, "embed_direction_L": dl, "num_channels": ch}))
    cases.append((lib.nerf_layout_id(10, 4, 0, 0, 128), {"pos_encoding": False, "num_channels": 128}))
    assert lib.nerf_layout_id(10, 4, 1, 0, 256) == 0 and lib.nerf_layout_id(10, 4, 1, 1, 256) == 2 and lib.nerf_layout_id(7, 7, 0, 0, 256) == 1
    assert lib.nerf_layout_id(11, 4, 1, 0, 256) == -4 and lib.nerf_layout_id(10, -1, 1, 0, 256) == -4 and lib.nerf_layout_id(10, 4, 1, 0, 64) == -4
    for layout, kw in cases:
        m = s.NeRFmodel(**kw)
        assert m.layout_id == layout
        off = 0
        for i, p in enumerate(m.ordered_parameters()):
            assert lib.nerf_param_offset(layout, i) == off
            off += p.numel()
        assert off == lib.nerf_param_count(layout)
        assert [id(p) for p in m.ordered_parameters()] == [id(p) for p in m.state_dict(keep_vars=True).values()]


def test_no_gpu_means_loud_failure(libmod):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = libmod.load()
    assert lib.nerf_device_check() != 0                 # cudaError / NERF_ERR_ARCH, never a silent fallback
    import sfmnerf_b200 as s
    with pytest.raises(s._lib.NerfB200Error):
        s.NeRFmodel()(torch.zeros(2, 3), torch.zeros(2, 3))
    with pytest.raises(s._lib.NerfB200Error):
        s.engine.encode(torch.zeros(2, 3), 4)


def test_cli_flags_match_reference_defaults():
    import sfmnerf_b200 as s
    a = s.configParser().parse_args([])
    assert (a.lrate, a.n_pos_freq, a.n_dirc_freq, a.n_rays_batch, a.n_sample, a.max_iters) == (5e-4, 10, 4, 8192, 256, 200001)
    assert (a.near, a.far, a.chunk_size, a.hidden_size, a.lrate_decay, a.lrate_decay_steps) == (2.0, 6.0, 32768, 256, 0.1, 50000)
    assert a.perturb is True and a.mode == "train" and a.save_ckpt_iter == 1000
    for name in ("PixelToRay", "generateBatch", "render", "loss", "test_image", "train", "test", "main", "configParser", "NeRFmodel"):
        assert hasattr(s, name)


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "sfm-nerf_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "oracle" not in src.replace("oracle/", ""), fn


def test_synthetic_scene_and_poses():
    import sfmnerf_b200 as s
    from oracle import nerf_oracle as orc
    np.testing.assert_allclose(s.synthetic.look_at_pose(30, -30), orc.look_at_pose(30, -30), atol=1e-12)
    cam = s.synthetic.make_camera_info(100, 100)
    assert abs(cam["camera_matrix"][0][0] - 95.987661) < 1e-5
    imgs = s.synthetic.render_lego_images(s.synthetic.sphere_poses(1), s.synthetic.make_camera_info(24, 24), n_sample=96)
    assert imgs.shape == (1, 24, 24, 3) and 0.01 < float(imgs.mean()) < 0.5


def test_reference_python_surface_names_and_signatures():
    """The drop-in boundary is the reference's Python call surface: every public function of Phase2/train.py and
    Phase2/Wrapper.py exists here with the same parameter names and defaults (checked against the reference source when it
    is present, else against the recorded list)."""
    import ast, importlib, inspect, os
    pkg = importlib.import_module("sfm-nerf_b200")
    recorded = {
        "train": {"loadDataset": ["data_path", "mode"], "PixelToRay": ["camera_info", "pose", "pixelPosition", "args"],
                  "generateBatch": ["images", "poses", "camera_info", "args"],
                  "render": ["model", "rays_origin", "rays_direction", "args", "near", "far"], "loss": ["groundtruth", "prediction"],
                  "train": ["images", "poses", "camera_info", "args"], "test": ["images", "poses", "camera_info", "args"],
                  "test_image": ["model", "image", "pose", "camera_info", "args"], "main": ["args"], "configParser": []},
        "Wrapper": {"loadDataset": ["data_path", "mode"], "PixelToRay": ["camera_info", "pose", "pixelPosition", "args"],
                    "generateBatch": ["images", "poses", "camera_info", "args"],
                    "render": ["model", "rays_origin", "rays_direction", "args", "near", "far"], "loss": ["groundtruth", "prediction"],
                    "train": ["images", "poses", "camera_info", "args"], "test": ["images", "poses", "camera_info", "args"],
                    "test_image": ["model", "image", "pose", "camera_info", "args"],
                    "test_single_image": ["images", "poses", "camera_info", "args"], "main": ["args"], "configParser": []},
    }
    for modname, funcs in recorded.items():
        ref_path = f"/root/reference/Phase2/{modname}.py"
        if os.path.exists(ref_path):                       # the live reference wins over the recorded list
            tree = ast.parse(open(ref_path).read())
            funcs = {n.name: [a.arg for a in n.args.args] for n in tree.body if isinstance(n, ast.FunctionDef)}
            assert set(funcs) == set(recorded[modname]), (modname, sorted(funcs))
        mod = getattr(pkg, modname) if modname == "Wrapper" else importlib.import_module("sfm-nerf_b200.train")
        for name, params in funcs.items():
            fn = getattr(mod, name)
            assert list(inspect.signature(fn).parameters) == params, (modname, name)
    assert inspect.signature(pkg.Wrapper.render).parameters["near"].default == 1.0       # Wrapper.py:129
    assert inspect.signature(importlib.import_module("sfm-nerf_b200.train").render).parameters["far"].default == 6.0   # train.py:187
    assert pkg.Wrapper.configParser().parse_args([]).max_iters == 100001


def test_vendored_reference_is_unmodified():
    """oracle/_ref (git-ignored, built by oracle/make_ref.py / build()) holds byte-identical copies of the reference sources the
    CPU arm of bench.py times; when the live reference is present the copies are compared with it."""
    import hashlib, json
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not vendored (run python oracle/make_ref.py in the build container)")
    man = json.load(open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")))
    for rel, digest in man["sha256"].items():
        data = open(os.path.join(ROOT, "oracle", "_ref", rel), "rb").read()
        assert hashlib.sha256(data).hexdigest() == digest, rel
        live = os.path.join("/root/reference", rel)
        if os.path.exists(live):
            assert open(live, "rb").read() == data, rel
    ref = ref_loader.import_reference()
    assert callable(ref.render) and callable(ref.generateBatch) and callable(ref.test_image)
