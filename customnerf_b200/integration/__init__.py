"""Reference-side bindings: the `_backend` objects the reference's own wrappers import (gridencoder/grid.py:9-12,
raymarching/raymarching.py:10-13), served by libnerf_b200.so.  Copy `gridencoder_backend.py` / `raymarching_backend.py`
next to the reference's wrappers (as `backend_b200.py`) or put them on sys.modules under the pybind module names
(`_gridencoder`, `_raymarching`); INTEGRATION.md section 2.  tests/test_gpu_reference_wrappers.py runs the reference's
unmodified grid.py / raymarching.py over them."""
