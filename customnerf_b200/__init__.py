"""customnerf_b200 -- B200 (sm_100a) implementation of CustomNeRF's NeRF render/train hot path.

Drop-in surface (same names, arguments and error behaviour as the reference):
    customnerf_b200.gridencoder   <->  /root/reference/gridencoder   (GridEncoder, grid_encode)
    customnerf_b200.raymarching   <->  /root/reference/raymarching   (near_far_from_aabb, morton3D, packbits,
                                                                      march_rays_train, composite_rays_train, ...)
    customnerf_b200.nerf          <->  nerf/network_grid.py + the hot-path half of nerf/renderer.py
                                       + get_rays of nerf/provider_utils.py (ray generation on the device)

Beyond the drop-in surface (B200-first execution of the same computations):
    customnerf_b200.fused_trainer.FusedTrainStep   the reconstruction train step as one replayable CUDA graph
    customnerf_b200.fused_infer.FusedInference     full-image rendering with device-driven rounds (no per-round host sync)
    customnerf_b200.fused_edit.FusedEditStep       the LGIE editing step (all / fg / bg renders, caller-supplied loss) likewise
    customnerf_b200.parallel                       ray sharding; PeerMemory: the multi-GPU gradient reduce + Adam + parameter
                                                   broadcast as ONE kernel over NVLink peer memory (NCCL all-reduce optional)

``install_aliases()`` registers the two op packages under their reference names so that
``from gridencoder import GridEncoder`` (nerf/encoding.py:62) and ``import raymarching``
(nerf/renderer.py:16) resolve to this implementation.  See INTEGRATION.md.
"""
import sys

__version__ = "0.1.0"


def install_aliases():
    from . import gridencoder as _g, raymarching as _r
    sys.modules.setdefault("gridencoder", _g)
    sys.modules.setdefault("raymarching", _r)
    return _g, _r
