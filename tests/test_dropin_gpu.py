"""GPU: the C++ drop-in classes (include/ORBextractor.h, include/ORBmatcher.h) driven like the reference's
Frame/Tracking code by tests/cpp/dropin_test.cpp (mock Frame/MapPoint types), every result compared with the oracle."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "build")


@pytest.mark.parametrize("binary", ["dropin_test", "dropin_test_cv"])
def test_cpp_dropin_classes_against_oracle(binary):
    """dropin_test: the adapters over their own stand-in types; dropin_test_cv: the same source with COEB_WITH_OPENCV, i.e.
    the cv::Mat / cv::KeyPoint / cv::InputArray branch a maintainer uses (OpenCV headers stood in by oracle/ref_shim)."""
    path = os.path.join(BUILD, binary)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "all"])
    out = subprocess.run([path], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "dropin_test: PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
