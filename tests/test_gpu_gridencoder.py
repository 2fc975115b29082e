"""-m gpu: CUDA grid encoder (through the C ABI / GridEncoder module) vs the CPU oracle.

Tolerances: fp32 forward rel 1e-4 (abs floor 1e-6); fp16 forward rel 2e-3;
fp32 grad rel 1e-4 with abs floor 1e-6*max|g|; fp16-path grads rel 1e-2 with abs floor 1e-3*max|g|.
The oracle is given the device-evaluated per-level scales (exp2f is the one libm-dependent value, see
oracle/nerf_oracle.c orc_locate).
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import cpu_ops

pytestmark = pytest.mark.gpu

CONFIGS = {
    "hash19": dict(log2_hashmap_size=19, desired_resolution=2048, gridtype="hash"),      # configs[0..3]
    "tiled21": dict(log2_hashmap_size=21, desired_resolution=8192, gridtype="tiled"),    # network_grid.py:89-96
    "hash14": dict(log2_hashmap_size=14, desired_resolution=512, gridtype="hash"),
}


def _points(rng, B):
    x = rng.uniform(0, 1, (B, 3)).astype(np.float32)
    x[:8] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1.0001, 0.5, 0.5], [-1e-4, 0.2, 0.3], [0.25, 0.75, 1.0],
             [1.0, 0.0, 0.5], [0.999999, 0.999999, 0.999999]]
    # a "ray": consecutive samples 0.00085 apart, the pattern the warp-aggregated backward is built for
    t = np.arange(256, dtype=np.float32)[:, None]
    x[8:264] = np.array([0.3, 0.4, 0.2], np.float32) + t * np.array([0.0006, 0.0005, 0.0003], np.float32)
    return x


def _make(cfg, device="cuda"):
    from customnerf_b200.gridencoder import GridEncoder, level_scales
    torch.manual_seed(0)
    enc = GridEncoder(**cfg).to(device)
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    sc = level_scales(enc.num_levels, enc.per_level_scale, enc.base_resolution).cpu().numpy()
    return enc, sc


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_backward_fp32(name):
    enc, sc = _make(CONFIGS[name])
    rng = np.random.RandomState(1)
    x = _points(rng, 5000)
    xb = torch.from_numpy(x * 4 - 2).cuda()                  # GridEncoder maps [-bound, bound] -> [0, 1]
    out = enc(xb, bound=2)
    x01 = ((xb + 2) / 4).cpu().numpy()
    emb = enc.embeddings.detach().cpu().numpy()
    offs = enc.offsets.cpu().numpy()
    out0, _ = cpu_ops.grid_encode_forward(x01, emb, offs, enc.per_level_scale, 16, gridtype=enc.gridtype_id, scales=sc)
    assert out.dtype == torch.float32 and out.shape == (5000, 32)
    assert_close(out.detach().cpu().numpy(), out0, 1e-4, 1e-6, "forward fp32")
    g = rng.normal(size=out0.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    ge0, _ = cpu_ops.grid_encode_backward(g, x01, emb.shape, offs, enc.per_level_scale, 16, gridtype=enc.gridtype_id,
                                          scales=sc)
    assert_close(enc.embeddings.grad.cpu().numpy(), ge0, 1e-4, 1e-6 * np.abs(ge0).max(), "grad_embeddings fp32")


@pytest.mark.parametrize("name", ["hash19", "tiled21"])
def test_forward_backward_autocast_fp16(name):
    enc, sc = _make(CONFIGS[name])
    rng = np.random.RandomState(2)
    x = _points(rng, 4096)
    xb = torch.from_numpy(x * 4 - 2).cuda()
    with torch.autocast("cuda", dtype=torch.float16):
        out = enc(xb, bound=2)
    assert out.dtype == torch.float16                        # grid.py:45-49
    x01 = ((xb + 2) / 4).cpu().numpy()
    emb16 = enc.embeddings.detach().half().float().cpu().numpy()
    offs = enc.offsets.cpu().numpy()
    out0, _ = cpu_ops.grid_encode_forward(x01, emb16, offs, enc.per_level_scale, 16, gridtype=enc.gridtype_id, scales=sc)
    assert_close(out.float().detach().cpu().numpy(), out0, 2e-3, 1e-3, "forward fp16")
    g = rng.normal(size=out0.shape).astype(np.float16)
    out.backward(torch.from_numpy(g).cuda())
    assert enc.embeddings.grad.dtype == torch.float32
    ge0, _ = cpu_ops.grid_encode_backward(g.astype(np.float32), x01, emb16.shape, offs, enc.per_level_scale, 16,
                                          gridtype=enc.gridtype_id, scales=sc)
    assert_close(enc.embeddings.grad.cpu().numpy(), ge0, 1e-2, 1e-3 * np.abs(ge0).max(), "grad_embeddings fp16 path")
    # no stale copies: an in-place update that does NOT bump the tensor version (what torch's fused Adam does) is seen
    enc.embeddings.data.view(-1).view(torch.int32)  # (touch: keep the parameter object identical)
    torch._foreach_mul_([enc.embeddings.data], 0.5)
    with torch.autocast("cuda", dtype=torch.float16):
        out2 = enc(xb, bound=2)
    assert_close(out2.float().detach().cpu().numpy(), out0 * 0.5, 4e-3, 1e-3, "after in-place update")


def test_backward_aggregated_equals_plain_atomics():
    """agg=1 (warp-aggregated scatter) and agg=0 give the same gradient up to fp32 summation order"""
    from customnerf_b200 import _lib as L
    enc, sc = _make(CONFIGS["hash19"])
    rng = np.random.RandomState(3)
    x = torch.from_numpy(_points(rng, 20000)).cuda()
    g = torch.from_numpy(rng.normal(size=(20000, 32)).astype(np.float32)).cuda()
    outs = []
    for agg in (0, 1):
        ge = torch.zeros_like(enc.embeddings)
        L.check(L.lib().nb200_grid_encode_backward(
            L.ptr(g), L.ptr(x), L.ptr(enc.offsets), L.ptr(ge), L.u32(20000), L.u32(3), L.u32(2), L.u32(16), L.u32(16),
            L.f32(float(np.log2(enc.per_level_scale))), L.u32(16), L.ptr(None), L.ptr(None), L.u32(0), L.i32(0),
            L.u32(0), L.i32(L.F32), L.i32(L.LAYOUT_BLC), L.i32(agg), L.stream()), "bwd")
        outs.append(ge.cpu().numpy())
    assert_close(outs[1], outs[0], 1e-4, 1e-6 * np.abs(outs[0]).max())


@pytest.mark.parametrize("D,C,interp,align", [(2, 1, 0, False), (2, 4, 1, False), (3, 8, 0, True), (3, 2, 1, False),
                                              (4, 2, 0, False), (5, 1, 0, False), (3, 4, 0, False)])
def test_generic_shapes_with_dy_dx(D, C, interp, align):
    from customnerf_b200.gridencoder import GridEncoder, level_scales
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=D, num_levels=6, level_dim=C, per_level_scale=1.6, base_resolution=4,
                      log2_hashmap_size=12, gridtype="hash", align_corners=align,
                      interpolation="smoothstep" if interp else "linear").cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    sc = level_scales(6, 1.6, 4).cpu().numpy()
    rng = np.random.RandomState(4)
    x = rng.uniform(0, 1, (777, D)).astype(np.float32)
    x[0] = 1.5
    xt = torch.from_numpy(x * 2 - 1).cuda().requires_grad_()
    out = enc(xt, bound=1)
    x01 = ((xt.detach() + 1) / 2).cpu().numpy()
    emb = enc.embeddings.detach().cpu().numpy()
    offs = enc.offsets.cpu().numpy()
    out0, dy0 = cpu_ops.grid_encode_forward(x01, emb, offs, 1.6, 4, calc_grad_inputs=True, align_corners=align,
                                            interpolation=interp, scales=sc)
    assert_close(out.detach().cpu().numpy(), out0, 1e-4, 1e-6, "forward")
    g = rng.normal(size=out0.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    ge0, gi0 = cpu_ops.grid_encode_backward(g, x01, emb.shape, offs, 1.6, 4, dy_dx=dy0, align_corners=align,
                                            interpolation=interp, scales=sc)
    assert_close(enc.embeddings.grad.cpu().numpy(), ge0, 1e-4, 1e-6 * np.abs(ge0).max(), "grad_embeddings")
    # d(x01)/d(x) = 1/(2*bound)
    assert_close(xt.grad.cpu().numpy(), gi0 * 0.5, 1e-4, 1e-5 * np.abs(gi0).max(), "grad_inputs")


def test_max_level_and_lbc_layout():
    from customnerf_b200 import _lib as L
    enc, sc = _make(CONFIGS["hash14"])
    rng = np.random.RandomState(5)
    x = rng.uniform(0, 1, (1000, 3)).astype(np.float32)
    xb = torch.from_numpy(x * 2 - 1).cuda()
    out = enc(xb, bound=1, max_level=5)
    x01 = ((xb + 1) / 2).cpu().numpy()
    emb = enc.embeddings.detach().cpu().numpy()
    offs = enc.offsets.cpu().numpy()
    out0, _ = cpu_ops.grid_encode_forward(x01, emb, offs, enc.per_level_scale, 16, max_level=5, scales=sc)
    assert not out0[:, 10:].any()
    assert_close(out.detach().cpu().numpy(), out0, 1e-4, 1e-6, "max_level")
    # the reference's native [L, B, C] layout through the C ABI
    lbc = torch.empty(16, 1000, 2, device="cuda")
    xd = torch.from_numpy(x01).cuda()
    L.check(L.lib().nb200_grid_encode_forward(
        L.ptr(xd), L.ptr(enc.embeddings), L.ptr(enc.offsets), L.ptr(lbc), L.u32(1000), L.u32(3), L.u32(2), L.u32(16),
        L.u32(16), L.f32(float(np.log2(enc.per_level_scale))), L.u32(16), L.ptr(None), L.u32(0), L.i32(0), L.u32(0),
        L.i32(L.F32), L.i32(L.LAYOUT_LBC), L.stream()), "fwd")
    full0, _ = cpu_ops.grid_encode_forward(x01, emb, offs, enc.per_level_scale, 16, scales=sc)
    assert_close(lbc.permute(1, 0, 2).reshape(1000, 32).cpu().numpy(), full0, 1e-4, 1e-6, "LBC layout")


def test_error_behaviour():
    from customnerf_b200.gridencoder import GridEncoder, grid_encode
    enc = GridEncoder(log2_hashmap_size=12, desired_resolution=64).cuda()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        grid_encode(torch.rand(4, 3), enc.embeddings, enc.offsets, enc.per_level_scale, 16)
    with pytest.raises(RuntimeError, match="C must be 1, 2, 4, or 8"):
        grid_encode(torch.rand(4, 3).cuda(), torch.rand(100, 3).cuda(), enc.offsets, 2.0, 16)
    assert enc(torch.zeros(0, 3).cuda()).shape == (0, 32)
    with pytest.raises(ValueError):
        enc.grad_total_variation()


def test_grad_total_variation():
    from customnerf_b200.gridencoder import GridEncoder, level_scales
    torch.manual_seed(0)
    enc = GridEncoder(num_levels=4, per_level_scale=2, base_resolution=8, log2_hashmap_size=14).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    enc.embeddings.grad = torch.zeros_like(enc.embeddings)
    rng = np.random.RandomState(6)
    x = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    enc.grad_total_variation(1e-2, torch.from_numpy(x).cuda(), bound=1)
    sc = level_scales(4, 2, 8).cpu().numpy()
    x01 = ((torch.from_numpy(x) + 1) / 2).numpy()
    g0 = cpu_ops.grad_total_variation(x01, enc.embeddings.detach().cpu().numpy(), np.zeros((enc.embeddings.shape[0], 2), np.float32),
                                      enc.offsets.cpu().numpy(), 1e-2, 2, 8, scales=sc)
    assert_close(enc.embeddings.grad.cpu().numpy(), g0, 1e-4, 1e-6 * np.abs(g0).max(), "tv grad")
