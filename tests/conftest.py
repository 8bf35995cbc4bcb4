import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


BITS = os.path.join(ROOT, "tests", "golden", "bits")


def load_md5_table():
    table = {}
    for line in open(os.path.join(BITS, "bits.md5")):
        parts = line.split()
        if len(parts) == 2:
            table[parts[1]] = parts[0]
    return table


def all_streams():
    return sorted(f for f in os.listdir(BITS) if f.endswith(".ivf"))


@pytest.fixture(scope="session")
def md5_table():
    return load_md5_table()
