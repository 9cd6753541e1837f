"""CPU: the C-ABI library loads, exports every symbol include/coeb_frontend.h declares, keeps the POD layouts the
reference's containers need, and fails loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import coeb_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "coeb_frontend.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(coeb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = cb.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in coeb_frontend.h but not exported: %s" % missing
    assert b"sm_100a" in lib.coeb_version()


def test_pod_layouts():
    assert cb.KP_DTYPE.itemsize == 28                       # cv::KeyPoint
    assert C.sizeof(cb.OrbParams) == 20 and C.sizeof(cb.Camera) == 40
    assert C.sizeof(cb.DynInfo) == 4 + 4 + 32 * 16 + 4


def test_product_does_not_import_or_link_the_oracle():
    pkg = os.path.join(ROOT, "coeb-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import orc" not in src and "coeb_oracle" not in src and "oracle/" not in src, os.path.join(dirpath, f)
    for h in ("coeb_frontend.h", "ORBextractor.h", "ORBmatcher.h"):
        assert "oracle" not in open(os.path.join(ROOT, "include", h)).read().lower()


@pytest.mark.skipif(cb.device_count() > 0, reason="a GPU is present")
def test_no_device_is_a_loud_failure_not_a_fallback():
    with pytest.raises(cb.CoebError) as e:
        cb.Extractor()
    assert e.value.status == cb.ERR_NO_DEVICE
    with pytest.raises(cb.CoebError) as e:
        cb.Matcher()
    assert e.value.status == cb.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_argument_validation_without_device():
    h = C.c_void_p()
    bad = cb.OrbParams(1000, 1.2, 99, 20, 7)
    assert cb.lib().coeb_extractor_create(C.byref(bad), 0, C.byref(h)) == cb.ERR_INVALID_ARG
    assert cb.lib().coeb_extractor_create(None, 0, C.byref(h)) == cb.ERR_INVALID_ARG
    assert b"nlevels" in cb.lib().coeb_last_error() or b"null" in cb.lib().coeb_last_error()
