import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the reference compiled in place)")


def pytest_collection_modifyitems(config, items):
    import orc
    if not orc.have_ref():
        # built from /root/reference when present (this container); prebuilt on the GPU box
        try:
            orc.build_ref()
        except Exception:
            pass
    if not orc.have_ref():
        skip = pytest.mark.skip(reason="oracle/_ref not built and /root/reference absent")
        for it in items:
            if "ref" in it.keywords:
                it.add_marker(skip)
