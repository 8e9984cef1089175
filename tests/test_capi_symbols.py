"""CPU: the C-ABI library loads and exports every symbol include/ictrack.h declares; no compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ictrack.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ict_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(ict):
    from invcompcamtrack_b200 import api
    names = header_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(ict.lib_path())
    for n in names:
        assert hasattr(L, n), "libictrack.so does not export " + n
    assert sorted(api.SIGNATURES) == names, "api.SIGNATURES and include/ictrack.h disagree"


def test_optparam_layout_and_derivations(ict):
    from oracle import oracle as O
    assert ctypes.sizeof(ict.OptParam) == 44 == ctypes.sizeof(O.OptParam)
    op = ict.make_optparam(lv_f=3, lv_l=0, psz=8, maxiter=5, normdp_ratio=0.1, maxpttrack=99)
    ref = O.make_optparam(lv_f=3, lv_l=0, psz=8, maxiter=5, normdp_ratio=0.1, maxpttrack=99)
    assert bytes(op) == bytes(ref)
    assert (op.pszd2, op.pszd2m3, op.novals, op.maxpttrack) == (4, 11, 64, 100)   # run_io_reprojection_test.cpp:115-126
    op = ict.make_optparam(psz=5, maxpttrack=8)
    assert (op.pszd2, op.pszd2m3, op.novals) == (2, 6, 25)


def test_host_only_entry_points(ict, orc):
    assert ict.lib().ict_version() == 100
    tot, off, sw, sh = ict.pyramid_layout(640, 480, 3, 8)
    assert (tot, off, sw, sh) == (442624, [0, 325376, 411392, 435328], [656, 336, 176, 96], [496, 256, 136, 76])
    with pytest.raises(ict.IctError):
        ict.pyramid_layout(642, 480, 3, 8)        # not divisible by 2^lv_f (camera.h:12-13)
    a = ict.camera_levels(4, (600.5, 610.25), (321.5, 239.25), (640, 480), 8)
    assert np.array_equal(a, orc.camera_levels(4, (600.5, 610.25), (321.5, 239.25), (640, 480), 8))


def test_no_cpu_fallback(ict):
    """Without a device every compute entry point must fail loudly (the product never routes through the oracle)."""
    if ict.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ict.IctError, match="no CUDA device"):
        ict.Frames(1, 64, 48, 1, 4)
    with pytest.raises(ict.IctError, match="no CUDA device"):
        ict.pyramid_build(np.zeros((48, 64), np.float32), 1, 4)
    with pytest.raises(ict.IctError):
        ict.Tracker(ict.make_optparam(), (100, 100), (32, 24), (64, 48))


def test_product_does_not_reference_the_oracle():
    """Nothing under invcompcamtrack_b200/ may include, import, link or load oracle/ (comments may cite it)."""
    pkg = os.path.join(ROOT, "invcompcamtrack_b200")
    bad = re.compile(r'#include\s*[<"][^>"]*oracle|from\s+oracle|import\s+oracle|libictrack_oracle|libictrack_ref|dlopen')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), f
