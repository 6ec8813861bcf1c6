import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The emulator fills shared memory with garbage at CTA start, as the hardware leaves it. 0xA5A5... is out of range for most
    # indices and once hid a read of unset state that the device turned into a corrupt stream (profiles/round2_notes.md item 5):
    # the whole CPU suite runs with small look-alike values instead.
    os.environ.setdefault("ZWZ_EMU_POISON", "5")


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


def _with_mode(make, mode):
    old = os.environ.get("ZWZ_INFLATE_MODE")
    os.environ["ZWZ_INFLATE_MODE"] = mode  # read once, by zwz_init
    try:
        return make()
    finally:
        if old is None:
            del os.environ["ZWZ_INFLATE_MODE"]
        else:
            os.environ["ZWZ_INFLATE_MODE"] = old


@pytest.fixture(scope="session")
def cuda_ctx_lanes():
    ctx = _with_mode(_cuda_context, "lanes")
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def cuda_ctx_warp():
    ctx = _with_mode(_cuda_context, "warp")
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def cuda_ctx_careful():
    ctx = _with_mode(_cuda_context, "careful")
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def emu_ctx_careful():
    import emu_lib
    ctx = _with_mode(emu_lib.emu_context, "careful")
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def emu_ctx_lanes():
    import emu_lib
    ctx = _with_mode(emu_lib.emu_context, "lanes")
    yield ctx
    ctx.close()


# The inflate decoders: one warp per stream with the lane-parallel block decoder (the product default: inflate_fast.cuh inside
# inflate.cuh), the same without it ("careful": every symbol decoded warp-redundantly — also what the default hands
# truncated/invalid streams to), and one lane per stream (inflate_lanes.cuh). The inflate tests run all three on the same inputs.
INFLATE_BACKENDS = [pytest.param("emu_ctx", id="emu-warp"), pytest.param("emu_ctx_careful", id="emu-careful"),
                    pytest.param("emu_ctx_lanes", id="emu-lanes"),
                    pytest.param("cuda_ctx_warp", id="cuda-warp", marks=pytest.mark.gpu),
                    pytest.param("cuda_ctx_careful", id="cuda-careful", marks=pytest.mark.gpu),
                    pytest.param("cuda_ctx_lanes", id="cuda-lanes", marks=pytest.mark.gpu)]


@pytest.fixture(params=INFLATE_BACKENDS)
def ctx_inf(request):
    return request.getfixturevalue(request.param)


@pytest.fixture
def is_gpu(request):
    return "cuda" in request.node.name
