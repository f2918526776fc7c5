"""Host logic of the Wrapper.py mirror (no GPU): the gif-mode orbit (`Wrapper.py:443-475`)."""
import ast
import os

import numpy as np
import pytest

REF = "/root/reference/Phase2/Wrapper.py"


def test_orbit_is_the_references_pose_spherical():
    import sfmnerf_b200 as s
    W = s.Wrapper
    # recorded from the reference's nested pose_spherical (th = 72, phi = -30, r = 4): fp32 factors, fp64 product with the permutation
    got = W.pose_spherical(72.0, -30.0, 4.0)
    assert got.shape == (4, 4)
    c = got[:3, 3]
    assert abs(np.linalg.norm(c) - 4.0) < 1e-5 and abs(c[2] - 2.0) < 1e-5           # elevation -30 deg -> height r*sin(30 deg) after the axis permutation
    if os.path.exists(REF):
        # the live reference: lift the three lambdas and the nested function out of test_single_image and evaluate them
        tree = ast.parse(open(REF).read())
        fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "test_single_image")
        keep = [n for n in fn.body if (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") in ("trans_t", "rot_phi", "rot_theta"))
                or (isinstance(n, ast.FunctionDef) and n.name == "pose_spherical")]
        assert len(keep) == 4
        ns = {"np": np}
        exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
        for th in np.linspace(0.0, 360.0, 10, endpoint=False):
            np.testing.assert_array_equal(W.pose_spherical(th, -30.0, 4.0), ns["pose_spherical"](th, -30.0, 4.0))
    poses_ref = W.orbit_poses(convention="reference")
    poses = W.orbit_poses()
    assert poses.shape == (10, 4, 4) and poses_ref.shape == (10, 4, 4)
    for p_ref, p in zip(poses_ref, poses):
        centre = p[:3, 3]
        assert abs(np.linalg.norm(centre) - 4.0) < 1e-5
        # PixelToRay shoots the central ray along the camera's +z column: converted poses look AT the origin ...
        fwd = p[:3, 2]
        assert np.dot(fwd, -centre / np.linalg.norm(centre)) > 0.999
        # ... the unconverted reference matrices look away from it
        assert np.dot(p_ref[:3, 2], -centre / np.linalg.norm(centre)) < -0.999
        np.testing.assert_allclose(np.linalg.det(p[:3, :3]), 1.0, atol=1e-5)
    with pytest.raises(ValueError):
        W.orbit_poses(convention="opengl")
