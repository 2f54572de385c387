import os
import sys

import pytest

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # see imagecaptionlearn_py_b200/__init__.py
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT
