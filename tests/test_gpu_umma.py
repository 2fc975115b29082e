"""-m gpu: the tcgen05 (UMMA) building block against torch.matmul: K-major and MN-major 128B-swizzled operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(N, K, a_mn, b_mn, seed=0):
    from customnerf_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(128, K, device="cuda", generator=g)).half()
    B = (torch.randn(N, K, device="cuda", generator=g)).half()
    D = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    L.check(L.lib().nb200_umma_selftest(L.ptr(Ain), L.ptr(Bin), L.ptr(D), L.u32(N), L.u32(K), L.u32(a_mn), L.u32(b_mn),
                                        L.ptr(status), L.stream()), "umma_selftest")
    torch.cuda.synchronize()
    assert int(status[0]) == 0, "MMA completion barrier timed out"
    ref = A.float() @ B.float().t()
    return (D - ref).abs().max().item(), ref.abs().max().item()


@pytest.mark.parametrize("N,K", [(64, 64), (64, 32), (16, 64), (64, 96), (128, 128), (32, 16)])
def test_umma_k_major(N, K):
    err, mag = _run(N, K, 0, 0)
    assert err < 1e-3 * max(mag, 1.0), (err, mag)


@pytest.mark.parametrize("N,K,a_mn,b_mn", [(64, 128, 1, 1), (16, 128, 1, 1), (64, 64, 1, 0), (64, 64, 0, 1),
                                           (128, 128, 1, 1), (32, 32, 1, 1)])
def test_umma_mn_major(N, K, a_mn, b_mn):
    err, mag = _run(N, K, a_mn, b_mn)
    assert err < 1e-3 * max(mag, 1.0), (err, mag)
