import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _cuda_context():
    import zwz_b200
    return zwz_b200.Context(0)  # raises loudly when the extension or the GPU is missing — no fallback


@pytest.fixture(scope="session")
def cuda_ctx():
    ctx = _cuda_context()
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def emu_ctx():
    """The same kernel sources run by the CPU SIMT emulator (tests/simt). Test infrastructure, never the product path."""
    import emu_lib
    ctx = emu_lib.emu_context()
    yield ctx
    ctx.close()


# Backend parametrisation shared by the kernel tests: "emu" runs here (CPU, small sizes), "cuda" on the GPU box.
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def ctx(request):
    if request.param == "cuda":
        return request.getfixturevalue("cuda_ctx")
    return request.getfixturevalue("emu_ctx")


@pytest.fixture
def is_gpu(request):
    return "cuda" in request.node.name
