"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import anything from here, and only as the checker / the timed CPU baseline.  The
product package ``customnerf_b200`` never imports this package; it fails loudly when its CUDA
library is missing instead of falling back to any of this.

Contents
  nerf_oracle.c   plain-C restatement of the two reference CUDA extensions (cites file:line)
  cpu_ops.py      numpy/ctypes front-end mirroring the reference's Python wrappers
  torch_ref.py    fp32 PyTorch-on-CPU restatement of the field network and both renderers
  build.py        gcc recipe for nerf_oracle.c
  build_ref.py    nvcc recipe compiling the UNMODIFIED reference extensions into oracle/_ref/
  ref_ext.py      loader for oracle/_ref (GPU box only)

Parity pin: the reference has no golden vectors or tests.  The restatement is
pinned against the reference's own CUDA extensions (oracle/_ref, run on the GPU box) both live
(tests/test_ref_ext_parity.py) and through vectors minted from them and committed under
tests/golden/.  The MLP (tinycudann, un-vendored, un-pinned) is the one component whose parity is
UNPINNED: see DESIGN.md.
"""
