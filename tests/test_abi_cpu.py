"""The C-ABI library loads without a GPU and exports every symbol include/nerf_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from customnerf_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_version_and_error_strings():
    from customnerf_b200 import _lib
    lib = _lib.lib()
    assert lib.nb200_version() >= 1
    assert lib.nb200_error_string(ctypes.c_int(0)) == b"ok"
    assert b"C must be 1, 2, 4, or 8" in lib.nb200_error_string(ctypes.c_int(-1))
    assert lib.nb200_march_scratch_ints(ctypes.c_uint32(1000)) >= 2 * 4


def test_no_oracle_import_in_product():
    """the product package must never import the oracle (no CPU fallback)"""
    pkg = os.path.join(ROOT, "customnerf_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_missing_library_fails_loudly(monkeypatch):
    from customnerf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnerf_b200.so")
    try:
        _lib.lib()
    except ImportError as e:
        assert "no CPU" in str(e)
    else:
        raise AssertionError("expected ImportError")


def test_plan_structures_match_the_library_layout():
    """the ctypes mirrors of the three plan structs have the size the library compiled (no GPU needed)"""
    import ctypes as C
    from customnerf_b200 import _lib, fused_trainer, fused_edit, parallel
    lib = _lib.lib()
    for fn in (lib.nb200_train_plan_bytes, lib.nb200_peer_plan_bytes, lib.nb200_lgie_plan_bytes, lib.nb200_peer_handle_bytes):
        fn.restype = C.c_uint32
    assert C.sizeof(fused_trainer.TrainPlan) == lib.nb200_train_plan_bytes()
    assert C.sizeof(parallel.PeerPlan) == lib.nb200_peer_plan_bytes()
    assert C.sizeof(fused_edit.LgiePlan) == lib.nb200_lgie_plan_bytes()
    assert lib.nb200_peer_handle_bytes() == 64
    lib.nb200_peer_signal_bytes.restype = C.c_uint64
    assert lib.nb200_peer_signal_bytes(C.c_uint32(296)) == 3 * 296 * 8 * 4
