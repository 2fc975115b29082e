"""The CPU oracle against golden vectors minted from the reference's own CUDA kernels (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_cases as gc
from oracle import cpu_ops

G = gc.load()
pytestmark = pytest.mark.skipif(G is None, reason="tests/golden/ref_ext_vectors.npz not minted yet")


class OracleOps:
    morton3D = staticmethod(cpu_ops.morton3D)
    morton3D_invert = staticmethod(cpu_ops.morton3D_invert)
    packbits = staticmethod(cpu_ops.packbits)
    near_far_from_aabb = staticmethod(cpu_ops.near_far_from_aabb)

    @staticmethod
    def march_rays_train(o, d, bound, bf, C, H, nears, fars, counter, noises):
        return cpu_ops.march_rays_train(o, d, bound, bf, C, H, nears, fars, counter, -1, noises, 128, True, 0, 1024)

    composite_forward = staticmethod(cpu_ops.composite_rays_train_forward)
    composite_backward = staticmethod(cpu_ops.composite_rays_train_backward)

    @staticmethod
    def march_rays(n_alive, n_step, alive, rays_t, o, d, bound, bf, C, H, nears, fars):
        return cpu_ops.march_rays(n_alive, n_step, alive, rays_t, o, d, bound, bf, C, H, nears, fars, -1, None, 0, 1024)

    @staticmethod
    def composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, dl, ws, depth, image, T):
        cpu_ops.composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, dl, ws, depth, image, T)
        return alive, rays_t, ws, depth, image

    @staticmethod
    def grid(x, emb, offs, pls, gridtype, grad, scales, half):
        e = emb.astype(np.float16).astype(np.float32) if half else emb
        out, _ = cpu_ops.grid_encode_forward(x, e, offs, pls, 16, gridtype=gridtype, half=half, scales=scales)
        gemb, _ = cpu_ops.grid_encode_backward(grad, x, emb.shape, offs, pls, 16, gridtype=gridtype, scales=scales)
        return out, gemb


def test_oracle_integer_ops_match_reference_kernels():
    gc.check_integer_ops(G, OracleOps)


def test_oracle_march_matches_reference_kernels():
    gc.check_march(G, OracleOps)


def test_oracle_composite_matches_reference_kernels():
    gc.check_composite(G, OracleOps)


def test_oracle_inference_ops_match_reference_kernels():
    gc.check_inference(G, OracleOps)


def test_oracle_grid_encoder_matches_reference_kernels():
    gc.check_grid(G, OracleOps)
