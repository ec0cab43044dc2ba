import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement oracle (always buildable: gcc only)."""
    from oracle import bind
    bind.build("port")
    return bind.PortOracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled in place; skipped when neither /root/reference nor a prebuilt _ref exists."""
    from oracle import bind
    if os.path.isdir("/root/reference/inMyRoom_vulkan"):
        bind.build("ref")
    if not os.path.exists(bind.REF_SO):
        pytest.skip("oracle/_ref/libimr_ref.so not available")
    return bind.RefOracle()


@pytest.fixture(scope="session")
def oracle(port):
    """Strongest checker available: the real reference if its library is present, else the port."""
    from oracle import bind
    if os.path.exists(bind.REF_SO):
        return bind.RefOracle()
    return port


@pytest.fixture(scope="session")
def gpu_ctx():
    from inmyroom_vulkan_b200.collision import Context
    return Context(0)
