import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # checkers (gcc only; seconds)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "_mathtest")], check=True, capture_output=True)
    # the product library: built in-tree by __graft_entry__.build(); build it here if a fresh checkout lacks it
    if not os.path.exists(os.path.join(ROOT, "epimethex_b200", "libepx.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "epimethex_b200", "csrc")], check=True, capture_output=True)


@pytest.fixture(scope="session")
def readme_frames():
    import __graft_entry__ as ge
    return ge.analysis_fixture()
