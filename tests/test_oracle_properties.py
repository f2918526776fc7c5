"""Size-independent properties of the compositing / sampling restatement (hypothesis, CPU): the invariants the GPU tests rely
on at BASELINE sizes, checked here on random ray lengths, densities and depths.  (train.py:205-217, 248-269)"""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import nerf_oracle as orc


def _inputs(seed, n, S, sigma_scale):
    g = torch.Generator().manual_seed(seed)
    t_rand = torch.rand(n, S, generator=g)
    z = orc.sample_z(n, S, 2.0, 6.0, t_rand)
    sigma = torch.rand(n, S, generator=g) * sigma_scale
    rgb = torch.rand(n, S, 3, generator=g)
    return sigma, rgb, z


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 9), S=st.integers(2, 200), sigma_scale=st.sampled_from([0.0, 0.5, 5.0, 200.0]))
def test_compositing_invariants(seed, n, S, sigma_scale):
    sigma, rgb, z = _inputs(seed, n, S, sigma_scale)
    out = orc.composite(sigma, rgb, z)
    w = out["weights"]
    assert bool((w >= 0).all())
    assert bool((out["acc"] <= 1.0 + 1e-4).all())                         # transmittance never creates light
    # rgb is a sub-convex combination of the sample colours; depth of the depths
    assert bool((out["rgb"] <= rgb.max(dim=1).values * out["acc"][:, None] + 1e-5).all())
    assert bool((out["depth"] <= z.max(dim=1).values * out["acc"] + 1e-4).all())
    # linear in the colours for fixed densities
    rgb2 = torch.rand_like(rgb)
    both = orc.composite(sigma, rgb + rgb2, z)["rgb"]
    torch.testing.assert_close(both, out["rgb"] + orc.composite(sigma, rgb2, z)["rgb"], rtol=1e-5, atol=1e-6)
    # rays are independent: any split of the batch gives the same rows
    k = n // 2
    parts = [orc.composite(sigma[a:b], rgb[a:b], z[a:b])["rgb"] for a, b in ((0, k), (k, n)) if b > a]
    assert torch.equal(torch.cat(parts), out["rgb"])
    # zero density renders nothing
    zero = orc.composite(torch.zeros_like(sigma), rgb, z)
    assert float(zero["acc"].abs().max()) < 1e-6


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 5), S=st.integers(2, 96))
def test_closed_form_backward_matches_autograd(seed, n, S):
    sigma, rgb, z = _inputs(seed, n, S, 3.0)
    sigma = sigma.double().requires_grad_(True)
    rgb = rgb.double().requires_grad_(True)
    g = torch.rand(n, 3, dtype=torch.float64)
    (orc.composite(sigma, rgb, z.double())["rgb"] * g).sum().backward()
    d_sigma, d_rgb = orc.composite_backward(sigma.detach(), rgb.detach(), z.double(), g)
    torch.testing.assert_close(d_sigma, sigma.grad, rtol=1e-7, atol=1e-9)
    torch.testing.assert_close(d_rgb, rgb.grad, rtol=1e-7, atol=1e-9)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 6), S=st.integers(1, 300))
def test_stratified_depths_are_sorted_and_bounded(seed, n, S):
    g = torch.Generator().manual_seed(seed)
    z = orc.sample_z(n, S, 2.0, 6.0, torch.rand(n, S, generator=g))
    assert float(z.min()) >= 2.0 and float(z.max()) <= 6.0
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    z0 = orc.sample_z(n, S, 2.0, 6.0)                                       # deterministic: the same row for every ray
    assert torch.equal(z0, z0[:1].expand(n, S))
    if S > 1:
        assert float(z0[0, 0]) == 2.0 and float(z0[0, -1]) == 6.0


def test_single_sample_ray_renders_zero():
    """Reference quirk (train.py:248-250): with S == 1 `dists` is empty, so is the appended 1e-3 column, and the sums are 0."""
    out = orc.composite(torch.rand(3, 1) * 10, torch.rand(3, 1, 3), torch.full((3, 1), 2.0))
    assert float(out["rgb"].abs().max()) == 0.0 and float(out["acc"].abs().max()) == 0.0
