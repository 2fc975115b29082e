import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def scene():
    """bear scene: density grid [2,128^3], bitfield, one 142x105 camera (configs[1] of BASELINE.json)"""
    import torch
    from customnerf_b200 import synthetic as syn
    from oracle import cpu_ops
    grid = syn.density_grid(2, 128)
    thr = min(float(grid.mean()), 10.0)
    bf = cpu_ops.packbits(grid.numpy(), thr)
    o, d = syn.camera_rays(105, 142)
    aabb = np.array([-2, -2, -2, 2, 2, 2], np.float32)
    nears, fars = cpu_ops.near_far_from_aabb(o.numpy(), d.numpy(), aabb)
    return dict(grid=grid.numpy(), thresh=thr, bitfield=bf, rays_o=o.numpy(), rays_d=d.numpy(), aabb=aabb,
                nears=nears, fars=fars, bound=2.0, cascade=2, H=128)


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d mismatches, max err %.3e (tol %.3e) at %s" % (
        what, bad.sum(), bad.size, err.max(), tol.flat[err.argmax()], np.unravel_index(err.argmax(), err.shape))
