"""Mint golden vectors for ray generation from the UNMODIFIED reference function ``get_rays``
(/root/reference/nerf/provider_utils.py:238-302), imported here on the CPU (it is plain PyTorch).  The module's
top-level import of ``torchtyping`` (absent from this image, used only in annotations) is stubbed.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_rays.py
    -> tests/golden/ref_get_rays.npz   (inputs stored next to the outputs)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/nerf/provider_utils.py"


def load_reference():
    stub = types.ModuleType("torchtyping")
    stub.TensorType = type("TensorType", (), {"__class_getitem__": classmethod(lambda cls, item: cls)})
    sys.modules.setdefault("torchtyping", stub)
    spec = importlib.util.spec_from_file_location("ref_provider_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pose(theta, phi, radius):
    """camera-to-world looking at the origin (any valid rotation will do: the vectors pin arithmetic, not cameras)"""
    c = np.array([radius * np.sin(theta) * np.sin(phi), radius * np.cos(theta), radius * np.sin(theta) * np.cos(phi)])
    f = -c / np.linalg.norm(c)
    r = np.cross(f, np.array([0.0, 1.0, 0.0])); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    m = np.eye(4, dtype=np.float32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r, -u, f, c
    return m


def main():
    ref = load_reference()
    torch.manual_seed(0)
    G = {}
    cases = [("full", 12, 16, 2, -1, (0.5, 0.5)), ("subset", 105, 142, 1, 64, (0.5, 0.5)), ("offset0", 7, 5, 3, -1, (0.0, 0.0))]
    for name, H, W, B, N, off in cases:
        poses = np.stack([pose(1.1 + 0.2 * b, 0.3 + 1.7 * b, 1.5 + 0.1 * b) for b in range(B)]).astype(np.float32)
        intr = np.array([0.9 * W, 0.95 * W, W / 2 + 0.25, H / 2 - 0.5], np.float32)
        out = ref.get_rays(torch.from_numpy(poses), intr, H, W, N, offset=off)
        G[name + "_poses"], G[name + "_intrinsics"] = poses, intr
        G[name + "_HWN"] = np.array([H, W, N], np.int64)
        G[name + "_offset"] = np.array(off, np.float32)
        G[name + "_rays_o"] = out["rays_o"].contiguous().numpy()
        G[name + "_rays_d"] = out["rays_d"].contiguous().numpy()
        if "inds" in out:
            G[name + "_inds"] = out["inds"].contiguous().numpy()
    # the pixel sampling of a training batch (:263-284): uniform, and error-map importance sampling; seeded CPU generator
    H, W, B, N = 105, 142, 2, 300
    poses = torch.from_numpy(np.stack([pose(1.0, 0.4, 1.5), pose(1.3, 2.0, 1.6)]))
    intr = np.array([120.0, 120.0, W / 2, H / 2], np.float32)
    em = torch.rand(B, 128 * 128, generator=torch.Generator().manual_seed(9))
    torch.manual_seed(4)
    out = ref.get_rays(poses, intr, H, W, N, error_map=em)
    G["sample_error_map"], G["sample_em_inds"], G["sample_em_coarse"] = em.numpy(), out["inds"].numpy(), out["inds_coarse"].numpy()
    torch.manual_seed(5)
    G["sample_uniform_inds"] = ref.get_rays(poses, intr, H, W, N)["inds"].contiguous().numpy()
    G["sample_HWBN"] = np.array([H, W, B, N], np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_get_rays.npz"), **G)
    print("wrote", os.path.join(HERE, "ref_get_rays.npz"), {k: v.shape for k, v in G.items()})


if __name__ == "__main__":
    main()
