"""The CPU oracle's ray generation against golden vectors minted from the reference's own ``get_rays``
(nerf/provider_utils.py:238-302, imported unmodified by tests/golden/make_golden_rays.py).

Tolerance: fp32 abs 1e-6 on unit-norm directions (the reference's bmm and the restatement's einsum may order the three
products of the rotation differently: 1-2 ulp); origins bit-exact (a copy of the pose's translation)."""
import os

import numpy as np

from oracle import cpu_ops

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ("full", "subset", "offset0")


def load():
    return np.load(os.path.join(HERE, "golden", "ref_get_rays.npz"))


def test_oracle_get_rays_matches_reference_vectors():
    G = load()
    for name in CASES:
        H, W, N = [int(v) for v in G[name + "_HWN"]]
        inds = G[name + "_inds"] if N > 0 else None
        o, d = cpu_ops.get_rays(G[name + "_poses"], G[name + "_intrinsics"], H, W, inds, tuple(G[name + "_offset"]))
        assert o.shape == G[name + "_rays_o"].shape and d.shape == G[name + "_rays_d"].shape
        assert np.array_equal(o, G[name + "_rays_o"]), name
        assert np.abs(d - G[name + "_rays_d"]).max() <= 1e-6, (name, np.abs(d - G[name + "_rays_d"]).max())
        assert np.abs(np.linalg.norm(d, axis=-1) - 1).max() < 1e-6


def test_pixel_sampling_draws_the_reference_indices():
    """the product's training-batch pixel sampler (nerf/provider_utils.py: _sample_pixels, plain torch, run here on the CPU)
    makes the reference's draws in the reference's order: same generator state -> same indices (:263-284), exact"""
    import torch
    from customnerf_b200.nerf import provider_utils as pu
    G = load()
    H, W, B, N = [int(v) for v in G["sample_HWBN"]]
    torch.manual_seed(4)
    inds, cells = pu._sample_pixels(B, H, W, N, torch.from_numpy(G["sample_error_map"]), torch.device("cpu"))
    assert np.array_equal(inds.numpy(), G["sample_em_inds"]) and np.array_equal(cells.numpy(), G["sample_em_coarse"])
    torch.manual_seed(5)
    inds, cells = pu._sample_pixels(B, H, W, N, None, torch.device("cpu"))
    assert cells is None and np.array_equal(inds.numpy(), G["sample_uniform_inds"])
