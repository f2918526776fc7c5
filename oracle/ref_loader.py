"""Imports the vendored, unmodified reference (oracle/_ref/Phase2/train.py; see oracle/make_ref.py) for the CPU arm of
bench.py.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.  `imageio` / `matplotlib` are top-level imports of the reference that this
image lacks and that hold no logic of the path: they are stubbed."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "Phase2")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "train.py")) and os.path.isfile(os.path.join(REF_DIR, "NeRFModel.py"))


def import_reference():
    """The reference's `train` module (prints its device and seeds numpy on import, train.py:16-17)."""
    if not available():
        raise FileNotFoundError(f"{REF_DIR} is missing: run `python oracle/make_ref.py` in the build container")
    for n in ("imageio", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import train as ref  # noqa: E402
    return ref
