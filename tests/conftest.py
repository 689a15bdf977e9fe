import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """the CPU restatement (oracle/liboracle.so), built on demand"""
    import orc
    if not os.path.exists(orc.PORT_SO):
        orc.build()
    return orc.Oracle(orc.PORT_SO)


@pytest.fixture(scope="session")
def reference():
    """the unmodified reference (oracle/_ref/libquids_ref.so); skipped where it cannot exist"""
    import orc
    if not orc.have_reference():
        if os.path.isdir("/root/reference/src"):
            orc.build()
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return orc.Oracle(orc.REF_SO)
