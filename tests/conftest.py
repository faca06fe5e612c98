import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree libtoued.so; built on demand when nvcc is available."""
    from to_ued_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from to_ued_b200.csrc.build import build
        build()
    return _lib.LIB_PATH


@pytest.fixture
def fp32_gru(monkeypatch):
    """Run the exact-fp32 SIMT GRU kernels (the tight-tolerance parity tests use these)."""
    import to_ued_b200
    monkeypatch.setattr(to_ued_b200, "GRU_PRECISION", "fp32")
    monkeypatch.setattr(to_ued_b200, "ES_PRECISION", "fp32")
