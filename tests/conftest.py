import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Build the checkers (oracle restatement; reference build when its sources are present).
    subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "liboracle.so"], check=True)
    if os.path.isdir("/root/reference") and not os.path.exists(os.path.join(REPO, "oracle", "_ref", "libedlib_ref.so")):
        subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "ref"], check=True)
