import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/code/smplx")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def syn_model():
    from mvsmplfitting_b200 import synthetic as S
    return S.make_model(0)


@pytest.fixture(scope="session")
def syn_gmm():
    from mvsmplfitting_b200 import synthetic as S
    return S.make_gmm(7)
