"""Vendors the UNMODIFIED reference sources of the hot path into oracle/_ref/ (git-ignored; travels to the GPU box with the
snapshot like a built .so) so that `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own code
on the GPU box's host cores.  TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under sfm-nerf_b200/ imports it.

    python oracle/make_ref.py            # needs /root/reference (build container); byte-for-byte copies + sha256 manifest

Only Phase2/train.py and Phase2/NeRFModel.py are needed: train.py holds PixelToRay / generateBatch / render / test_image
(train.py:68-127, 187-271, 714-763) and imports NeRFModel.  No reference source enters the git history.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("NERF_REFERENCE", "/root/reference")
FILES = ("Phase2/train.py", "Phase2/NeRFModel.py")


def vendor(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "Phase2")):
        if verbose:
            print(f"oracle/make_ref.py: {SRC}/Phase2 not found; nothing vendored")
        return False
    dst_root = os.path.join(HERE, "_ref")
    manifest = {}
    for rel in FILES:
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(dst_root, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print("vendored", ", ".join(FILES), "->", dst_root)
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
