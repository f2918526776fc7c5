"""Generate tests/golden/legacy_pe.npz from the reference's OWN legacy model code (run in the build container).

    python tests/golden/make_golden_legacy.py

The legacy positional-encoding model is the commented-out class at /root/reference/Phase2/NeRFModel.py:7-99 (it is the
layout of the shipped checkpoints).  This script un-comments those lines in memory, executes the class with
torch.manual_seed(0) and records its outputs on seeded inputs; only inputs / outputs and a weight checksum are stored --
oracle.legacy_init_state_dict("pe", 0) re-creates the identical weights (same nn.Linear construction order).
When the shipped checkpoints are present their outputs under the same class (PE) are recorded too."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("NERF_REFERENCE", "/root/reference")


def legacy_class():
    lines = open(os.path.join(REF, "Phase2", "NeRFModel.py")).read().splitlines()[6:99]      # lines 7..99
    src = "\n".join(l[2:] if l.startswith("# ") else ("" if l.strip() == "#" else l) for l in lines)
    ns = {}
    exec("import torch\nimport torch.nn as nn\nimport torch.nn.functional as F\nimport numpy as np\n" + src, ns)
    return ns["NeRFmodel"]


def main():
    from oracle import nerf_oracle as orc
    torch.set_num_threads(1)
    Legacy = legacy_class()
    torch.manual_seed(0)
    m = Legacy()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    mine = orc.legacy_init_state_dict("pe", 0)
    assert list(sd.keys()) == list(mine.keys()), (list(sd.keys()), list(mine.keys()))
    assert all(torch.equal(sd[k], mine[k]) for k in sd), "oracle.legacy_init_state_dict does not reproduce the commented class"
    g = torch.Generator().manual_seed(5)
    pos = (torch.rand(3, 40, 3, generator=g) - 0.5) * 4.0          # the class indexes out[:, :, 0]: 3-D inputs
    dirs = torch.nn.functional.normalize(torch.randn(3, 40, 3, generator=g), dim=-1)
    with torch.no_grad():
        density, rgb = m(pos, dirs)
    out = dict(pos=pos.numpy(), dirs=dirs.numpy(), density=density.numpy(), rgb=rgb.numpy(),
               weight_checksum=np.float64(sum(float(v.double().abs().sum()) for v in sd.values())))
    ck = os.path.join(REF, "Phase2", "checkpoints_tmp", "final_model.pth")
    if os.path.exists(ck):
        csd = torch.load(ck, map_location="cpu", weights_only=True)
        m.load_state_dict(csd)
        with torch.no_grad():
            cd, cr = m(pos, dirs)
        out.update(ckpt_density=cd.numpy(), ckpt_rgb=cr.numpy(),
                   ckpt_checksum=np.float64(sum(float(v.double().abs().sum()) for v in csd.values())))
    np.savez(os.path.join(HERE, "legacy_pe.npz"), **out)
    print("wrote legacy_pe.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
