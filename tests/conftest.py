"""pytest configuration: the `gpu` marker and the import of the (dash-named) package."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rtb():
    """The product's Python driver package (r-star-tree_b200/)."""
    return importlib.import_module("r-star-tree_b200")
