"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mvstereovision3_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mvsv.h")).read()
    declared = set(re.findall(r"\b(mvsv_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared == set(api.ABI_SYMBOLS), declared ^ set(api.ABI_SYMBOLS)
    lib = api.load_library()
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.MvsvError) as e:
        api.Engine(752, 480)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    lib = api.load_library()
    assert lib.mvsv_set_sgbm_params(None, None) == -1
    assert lib.mvsv_compute(None, None, 0, None, 0, 0, 1, 2) == -1


def test_struct_layout_matches_reference_order():
    # Disparity::sgbmParameters field order (reference inc/disparity.h:17-27)
    names = [f[0] for f in api.SgbmParams._fields_]
    assert names[:9] == ["minDisp", "numDisp", "blockSize", "disp12MaxDiff", "preFilterCap", "uniquenessRatio",
                         "speckleWindowSize", "speckleRange", "disparityMode"]
    assert C.sizeof(api.SgbmParams) == 11 * 4


class _FakeEngine:
    def __init__(self):
        self.calls = []

    def set_sgbm_params(self, **kw):
        self.calls.append(("sgbm", kw))

    def set_bm_params(self, **kw):
        self.calls.append(("bm", kw))


def test_load_sgbm_parameters(tmp_path):
    f = tmp_path / "sgbm.yml"
    # verbatim content of the reference's configs/sgbm.yml
    f.write_text("%YAML:1.0\nminDisp: 1\nnumDisp: 128\nblockSize: 13\ndisp12MaxDiff: 0\npreFilterCap: 0\n"
                 "uniquenessRatio: 0\nspeckleWindowSize: 150\nspeckleWindowRange: 2\nmode: 0\n")
    e, para = _FakeEngine(), {}
    assert api.loadSGBMParameters(str(f), e, para) is True
    assert para == dict(minDisp=1, numDisp=128, blockSize=13, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
                        speckleWindowSize=150, speckleRange=2, disparityMode=0)
    assert "P1" not in para and "P2" not in para        # src/disparity.cpp:83-90 never sets them
    assert e.calls == [("sgbm", para)]
    # missing required node -> False, engine untouched (src/disparity.cpp:67-71)
    g = tmp_path / "bad.yml"
    g.write_text("%YAML:1.0\nminDisp: 1\nnumDisp: 128\nblockSize: 13\n")
    e2 = _FakeEngine()
    assert api.loadSGBMParameters(str(g), e2, {}) is False and e2.calls == []
    assert api.loadSGBMParameters(str(tmp_path / "nope.yml"), e2, {}) is False


def test_load_bm_parameters(tmp_path):
    f = tmp_path / "bm.yml"
    f.write_text("%YAML:1.0\nnumDisp: 80\nblockSize: 21\npreFilterCap: 2\npreFilterSize: 51\nuniquenessRatio: 0\n"
                 "textureThreshold: 30\n")
    e, para = _FakeEngine(), {}
    assert api.loadBMParameters(str(f), e, para)
    assert para == dict(numDisp=80, blockSize=21, preFilterCap=2, textureThreshold=30, uniquenessRatio=0)


def test_roi_helpers_follow_reference():
    # MeanDisparityDetection::init: 9x9 grid of cols/9 x rows/9, remainder ignored
    rois = api.subimage_rois(688, 480, x_offset=64)
    assert len(rois) == 81 and rois[0] == (64, 0, 76, 53) and rois[80] == (64 + 8 * 76, 8 * 53, 76, 53)
    # SamplepointDetection::init: c in [1, cols/8), r in [1, rows/8); centre (c*(cols/(cols/8)), r*(rows/(rows/8)))
    sp = api.samplepoint_rois(688, 480)
    assert len(sp) == (688 // 8 - 1) * (480 // 8 - 1)
    assert sp[0] == (8 - 2, 8 - 2, 5, 5)
    assert api.dmap_roi_offset(128, 752) == 64
    assert api.dmap_roi_offset(64, 752) == 32


def _build_shim_demo(tmp_path):
    import subprocess
    exe = str(tmp_path / "shim_demo")
    libdir = os.path.join(ROOT, "mvstereovision3_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_demo.cpp"), "-o", exe, "-L" + libdir, "-lmvsv",
                           "-Wl,-rpath," + libdir])
    return exe


def test_cpp_mirror_compiles_as_cxx11(tmp_path):
    """include/mvsv_disparity.hpp (the C++ mirror of reference inc/disparity.h) builds with -std=c++11 like the
    reference (Makefile:7) and links against libmvsv.so only."""
    assert os.path.exists(_build_shim_demo(tmp_path))


def test_cpp_dmap_roi_offset_matches_reference_arithmetic(tmp_path):
    """mvsv::dMapRoiOffset (include/mvsv_detection.hpp) == createDMapROIS' pixelShift (reference trgt/demo.cpp:87-96),
    including the odd-half branch, for every numDisp the contract allows and a few widths."""
    import subprocess
    src = tmp_path / "off.cpp"
    src.write_text('#include "mvsv_detection.hpp"\n#include <cstdio>\n#include <cstdlib>\n'
                   'int main(int c, char** v) { for (int i = 1; i + 1 < c; i += 2) '
                   'std::printf("%d\\n", mvsv::dMapRoiOffset(std::atoi(v[i]), std::atoi(v[i + 1]))); return 0; }\n')
    exe = str(tmp_path / "off")
    subprocess.check_call(["g++", "-std=c++11", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    pairs = [(d, w) for d in range(8, 257, 8) for w in (752, 376, 1920, 3840, 201)]
    args = [str(x) for p in pairs for x in p]
    got = [int(x) for x in subprocess.check_output([exe] + args).decode().split()]
    assert got == [api.dmap_roi_offset(d, w) for d, w in pairs]
    # numDisp / 2 whenever that is even -- all a reachable configuration produces with numDisp % 16 == 0
    assert all(api.dmap_roi_offset(d, 752) == d // 2 for d in range(16, 257, 16))
