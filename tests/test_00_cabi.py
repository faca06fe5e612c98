"""The C-ABI library loads and exports every symbol include/toued.h declares (no compute, no GPU)."""
import ctypes, os, re


def test_library_exports_every_declared_symbol(built_lib):
    from to_ued_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "toued.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(toued_\w+)\s*\(", hdr))
    assert declared, "no declarations found"
    l = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(l, name), f"{name} declared in toued.h but not exported"
    assert declared == set(_lib.SIGNATURES) | {"toued_last_error", "toued_version"}
    assert l.toued_version() >= 1


def test_ops_refuse_cpu_tensors(built_lib):
    import torch, pytest
    from to_ued_b200 import _lib
    with pytest.raises(_lib.TouedError):
        _lib.ptr(torch.zeros(4))
