"""``Network``: stand-in for ``tinycudann.Network(n_in, n_out, {"otype": "FullyFusedMLP", ...})`` as the reference
uses it (nerf/network_grid.py:98-139).

tiny-cuda-nn is not vendored by the reference and not pinned (README.md:50), so its arithmetic is restated from
its published behaviour [upstream, unverified -- parity unpinned, see DESIGN.md]:
  * bias-free layers, hidden activation ReLU, output activation None | Sigmoid, width 64;
  * ONE flat fp32 ``params`` Parameter: per layer a row-major [out_padded, in_padded] matrix, layers concatenated,
    input / output widths padded up to a multiple of 16;
  * input lanes [n_in, in_padded) are fed the constant 1.0 (their weight columns act as a learnable bias);
  * fp16 operands, fp32 accumulation (tcnn accumulates in fp16; this contract is stricter), fp16 output.

Two execution paths share this module: the per-layer library path below (cuBLAS through torch, used for
arbitrary shapes), and the fused tcgen05 field kernel (``fused_field.py``) that NeRFNetwork uses for the
trunk + density head + colour head in one launch.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def pad16(n):
    return (n + 15) // 16 * 16


def layer_shapes(n_in, n_out, n_neurons=64, n_hidden_layers=1):
    dims = [pad16(n_in)] + [n_neurons] * n_hidden_layers + [pad16(n_out)]
    return [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


def xavier_uniform_flat(n_in, n_out, n_neurons=64, n_hidden_layers=1, seed=1337):
    g = torch.Generator().manual_seed(seed)
    parts = []
    for (o, i) in layer_shapes(n_in, n_out, n_neurons, n_hidden_layers):
        s = math.sqrt(6.0 / (i + o))
        parts.append(((torch.rand(o, i, generator=g) * 2 - 1) * s).reshape(-1))
    return torch.cat(parts)


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        if network_config.get("otype", "FullyFusedMLP") not in ("FullyFusedMLP", "CutlassMLP"):
            raise ValueError("only MLP networks are supported")
        if network_config.get("activation", "ReLU") != "ReLU":
            raise ValueError("hidden activation must be ReLU")
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.n_neurons = int(network_config.get("n_neurons", 64))
        self.n_hidden_layers = int(network_config.get("n_hidden_layers", 1))
        self.output_activation = network_config.get("output_activation", "None")
        if self.output_activation not in ("None", "Sigmoid"):
            raise ValueError("output activation must be None or Sigmoid")
        self.shapes = layer_shapes(n_input_dims, n_output_dims, self.n_neurons, self.n_hidden_layers)
        self.params = nn.Parameter(xavier_uniform_flat(n_input_dims, n_output_dims, self.n_neurons,
                                                       self.n_hidden_layers, seed))

    def weights(self):
        """list of [out_padded, in_padded] views into the flat parameter vector"""
        out, off = [], 0
        for (o, i) in self.shapes:
            out.append(self.params[off:off + o * i].view(o, i))
            off += o * i
        return out

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("customnerf_b200.nerf.Network runs on CUDA only (no CPU fallback)")
        with torch.amp.autocast('cuda', enabled=False):
            h = x.half()
            pad = self.shapes[0][1] - self.n_input_dims
            if pad:
                h = torch.cat([h, torch.ones(h.shape[0], pad, dtype=h.dtype, device=h.device)], -1)
            ws = self.weights()
            for li, W in enumerate(ws):
                h = F.linear(h, W.half())
                if li < len(ws) - 1:
                    h = torch.relu(h)
            h = h[:, :self.n_output_dims]
            if self.output_activation == "Sigmoid":
                h = torch.sigmoid(h.float()).half()
            return h
