import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "coeb-slam_b200", "python"),
          os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def load_pattern():
    import numpy as np
    txt = open(os.path.join(ROOT, "include", "coeb_orb_pattern.inc")).read()
    nums = [int(t) for line in txt.splitlines() if not line.startswith("//") for t in line.split(",") if t.strip()]
    assert len(nums) == 1024
    return np.array(nums, np.int32).reshape(512, 2)


@pytest.fixture(scope="session")
def orb_pattern():
    return load_pattern()
