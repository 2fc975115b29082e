"""-m gpu: the CUDA path against golden vectors minted from the reference's own CUDA kernels."""
import numpy as np
import pytest
import torch

import golden_cases as gc

G = gc.load()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(G is None, reason="tests/golden/ref_ext_vectors.npz not minted yet")]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(*ts):
    return tuple(t.detach().cpu().numpy() for t in ts)


class CudaOps:
    @staticmethod
    def _rm():
        from customnerf_b200 import raymarching
        return raymarching

    @classmethod
    def morton3D(cls, c):
        return cls._rm().morton3D(cu(c)).cpu().numpy()

    @classmethod
    def morton3D_invert(cls, i):
        return cls._rm().morton3D_invert(cu(i)).cpu().numpy()

    @classmethod
    def packbits(cls, g, t):
        return cls._rm().packbits(cu(g), t).cpu().numpy()

    @classmethod
    def near_far_from_aabb(cls, o, d, aabb, mn):
        return host(*cls._rm().near_far_from_aabb(cu(o), cu(d), cu(aabb), mn))

    @classmethod
    def march_rays_train(cls, o, d, bound, bf, C, H, nears, fars, counter, noises):
        cnt = cu(counter)
        out = cls._rm().march_rays_train(cu(o), cu(d), bound, cu(bf), C, H, cu(nears), cu(fars), cnt, -1,
                                         noises is not None, 128, True, 0, 1024,
                                         noises=None if noises is None else cu(noises))
        counter[:] = cnt.cpu().numpy()
        return host(*out)

    @classmethod
    def composite_forward(cls, sig, rgb, dl, rays, T):
        return host(*cls._rm().composite_rays_train(cu(sig), cu(rgb), cu(dl), cu(rays), T))

    @classmethod
    def composite_backward(cls, g_ws, g_img, sig, rgb, dl, rays, ws, img, T):
        s = cu(sig).requires_grad_()
        c = cu(rgb).requires_grad_()
        w, _, im = cls._rm().composite_rays_train(s, c, cu(dl), cu(rays), T)
        torch.autograd.backward([w, im], [cu(g_ws), cu(g_img)])
        return host(s.grad, c.grad)

    @classmethod
    def march_rays(cls, n_alive, n_step, alive, rays_t, o, d, bound, bf, C, H, nears, fars):
        return host(*cls._rm().march_rays(n_alive, n_step, cu(alive), cu(rays_t), cu(o), cu(d), bound, cu(bf), C, H,
                                          cu(nears), cu(fars), -1, False, 0, 1024))

    @classmethod
    def composite_rays(cls, n_alive, n_step, alive, rays_t, sig, rgb, dl, ws, depth, image, T):
        a, t, w, dp, im = cu(alive), cu(rays_t), cu(ws), cu(depth), cu(image)
        cls._rm().composite_rays(n_alive, n_step, a, t, cu(sig), cu(rgb), cu(dl), w, dp, im, T)
        return host(a, t, w, dp, im)

    @staticmethod
    def grid(x, emb, offs, pls, gridtype, grad, scales, half):
        from customnerf_b200.gridencoder import grid_encode
        e = cu(emb).requires_grad_()
        with torch.autocast("cuda", dtype=torch.float16, enabled=half):
            out = grid_encode(cu(x), e, cu(offs), pls, 16, False, gridtype, False, 0, None)
        assert out.dtype == (torch.float16 if half else torch.float32)
        out.backward(cu(grad).to(out.dtype))
        return out.float().detach().cpu().numpy(), e.grad.cpu().numpy()


def test_cuda_integer_ops_match_reference_kernels():
    gc.check_integer_ops(G, CudaOps)


def test_cuda_march_matches_reference_kernels():
    gc.check_march(G, CudaOps)


def test_cuda_composite_matches_reference_kernels():
    gc.check_composite(G, CudaOps)


def test_cuda_inference_ops_match_reference_kernels():
    gc.check_inference(G, CudaOps)


def test_cuda_grid_encoder_matches_reference_kernels():
    gc.check_grid(G, CudaOps)
