import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the reference checkout (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import refshim
    have_ref = refshim.available()
    skip_ref = pytest.mark.skip(reason="reference checkout not present")
    for item in items:
        if "ref" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
