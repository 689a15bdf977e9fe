import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _usable_gpus():
    """CUDA devices the product library can use (0 when the library is not built or there is no driver)"""
    try:
        import quids_b200
        return int(quids_b200.lib().qb_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """tests marked `gpu` need a device: on a box without one they are skipped (not failed) unless they were asked for
    explicitly with -m gpu -- there a missing device must fail loudly, the CUDA path has no fallback"""
    if "gpu" in (config.getoption("-m") or "") or _usable_gpus() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
