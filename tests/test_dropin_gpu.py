"""GPU: the C++ drop-in classes (include/ORBextractor.h, include/ORBmatcher.h) driven like the reference's
Frame/Tracking code by tests/cpp/dropin_test.cpp (mock Frame/MapPoint types), every result compared with the oracle."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "build", "dropin_test")


def test_cpp_dropin_classes_against_oracle():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")])
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "dropin_test: PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
