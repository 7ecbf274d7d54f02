"""the obstacle detectors and the PLY writer of include/mvsv_detection.hpp against an
independent numpy restatement of the reference logic (src/MeanDisparityDetection.cpp:71-266,
src/SamplePointDetection.cpp:29-178, src/utility.cpp:176-240, src/ply.cpp:36-95).  The CPU test feeds synthetic
means; the GPU test feeds the means / min-max the engine computed on a real SGBM output."""
import os
import subprocess

import numpy as np
import pytest

import cases
from mvstereovision3_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def build_demo(tmp_path):
    exe = str(tmp_path / "detect_demo")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "detect_demo.cpp"), "-o", exe])
    return exe


# ---- restatement (float32 arithmetic where the reference uses float) ------------------------------------------
def calc_dmap_values(c, Q):                      # src/utility.cpp:224-240
    num = f32(Q[2, 3]) - f32(c[2]) * f32(Q[3, 3])
    den = f32(c[2]) * f32(Q[3, 2])
    return f32(num / den) * f32(16)


def calc_coordinate(x, y, dvalue, Q):            # src/utility.cpp:176-200
    v = np.array([x, y, f32(dvalue) / f32(16), 1], f32)
    c = (Q.astype(np.float64) @ v.astype(np.float64)).astype(f32)
    c = (c.astype(np.float64) * (1.0 / np.float64(c[3]))).astype(f32)
    if np.isinf(c[2] / f32(1000)):
        c[2] = 0
    return c


def fmt_g9(v):
    return "%.9g" % float(v)


def expected_dump(cols, rows, xoff, mind, maxd, Q, means, mn, mx):
    out = []
    lo = calc_dmap_values([0, 0, f32(mind) * f32(1000)], Q)
    hi = calc_dmap_values([0, 0, f32(maxd) * f32(1000)], Q)
    out.append("range %s %s" % (fmt_g9(lo), fmt_g9(hi)))
    dx, dy = cols // 9, rows // 9
    subs = [((c * dx, r * dy), (c * dx + dx, r * dy + dy)) for r in range(9) for c in range(9)]
    nx, ny = cols // 8, rows // 8
    sps = [(c * (cols // nx), r * (rows // ny)) for c in range(1, nx) for r in range(1, ny)]
    out.append("nroi %d %d" % (len(subs), len(sps)))
    out.append("roi0 %d %d %d %d | sp0 %d %d %d %d" % (xoff, 0, dx, dy, xoff + sps[0][0] - 2, sps[0][1] - 2, 5, 5))
    m_means, s_means = means[:81], means[81:]
    centers = [(tl[0] + (br[0] - tl[0]) // 2, tl[1] + (br[1] - tl[1]) // 2) for tl, br in subs]
    out.append("mode 1 ndist 81")                                  # MEAN_DISTANCE falls through into MEAN_VALUE
    for i in range(0, 81, 9):
        c = calc_coordinate(centers[i][0], centers[i][1], m_means[i], Q)
        d = c[2] / f32(1000)
        out.append("dist %d %s" % (i, fmt_g9(0 if np.isinf(d) else d)))
    found = [i for i in range(81) if m_means[i] < lo and m_means[i] > hi]
    out.append("found_mean %d counter %d" % (len(found), 1 if found else 0))
    pts = []
    for i in found:
        c = calc_coordinate(centers[i][0], centers[i][1], m_means[i], Q)
        pts.append(c)
        out.append("M %d %d %s %s %s %s" % (centers[i][0], centers[i][1], fmt_g9(m_means[i]), fmt_g9(c[0]), fmt_g9(c[1]), fmt_g9(c[2])))
    sfound = [i for i in range(len(sps)) if s_means[i] < lo and s_means[i] > hi]
    out.append("found_sp %d counter %d" % (len(sfound), 1 if sfound else 0))
    for i in sfound:
        c = calc_coordinate(sps[i][0], sps[i][1], s_means[i], Q)
        out.append("S %d %d %s %s %s %s" % (sps[i][0], sps[i][1], fmt_g9(s_means[i]), fmt_g9(c[0]), fmt_g9(c[1]), fmt_g9(c[2])))
    out.append("PLY")
    out += ["ply", "format ascii 1.0", "comment author: Hagen Hiller", "comment object:obstacle pointcloud",
            "element vertex %d" % len(pts), "property float x", "property float y", "property float z",
            "property uchar red", "property uchar green", "property uchar blue", "end_header"]
    for c in pts:
        g = int(np.float64(f32(f32(c[2]) - f32(mn)) / f32(int(mx) - int(mn))) * 255.0)
        out.append("%s %s %s %d %d %d" % ("%g" % c[0], "%g" % c[1], "%g" % c[2], g, g, g))
    return out


def run_demo(exe, tmp_path, cols, rows, xoff, mind, maxd, Q, means, mn, mx):
    (tmp_path / "q.bin").write_bytes(np.ascontiguousarray(Q, f32).tobytes())
    (tmp_path / "m.bin").write_bytes(np.ascontiguousarray(means, f32).tobytes())
    txt = subprocess.check_output([exe, str(cols), str(rows), str(xoff), repr(mind), repr(maxd), str(tmp_path / "q.bin"),
                                   str(tmp_path / "m.bin"), str(int(mn)), str(int(mx))]).decode()
    return txt.strip("\n").split("\n")


def test_detectors_and_ply_cpu(tmp_path):
    exe = build_demo(tmp_path)
    Q = cases.Q_REFERENCE
    cols, rows, xoff = 688, 480, 64
    lo = calc_dmap_values([0, 0, f32(0.1) * f32(1000)], Q)
    hi = calc_dmap_values([0, 0, f32(1.5) * f32(1000)], Q)
    rng = np.random.default_rng(5)
    n_sp = (cols // 8 - 1) * (rows // 8 - 1)
    # truncated-division means like calcMeanDisparity returns, spread around the detection range, plus zeros
    means = np.floor(rng.uniform(float(hi) * 0.5, float(lo) * 1.3, size=81 + n_sp)).astype(f32)
    means[::7] = 0
    got = run_demo(exe, tmp_path, cols, rows, xoff, 0.1, 1.5, Q, means, 17, 1009)
    want = expected_dump(cols, rows, xoff, 0.1, 1.5, Q, means, 17, 1009)
    assert any(l.startswith("M ") for l in want) and any(l.startswith("S ") for l in want)
    assert got == want


@pytest.mark.gpu
def test_detectors_on_gpu_means(oracle, tmp_path):
    exe = build_demo(tmp_path)
    p = cases.sgbm_params(minDisp=1, numDisp=64, blockSize=13, speckleWindowSize=150, speckleRange=2)
    H, W = 240, 400
    l, r, _ = cases.synth.stereogram(H, W, 1, 64, seed=21)
    Q = cases.Q_REFERENCE
    off = api.dmap_roi_offset(64, W)
    cols = W - off
    rois = api.subimage_rois(cols, H, off) + api.samplepoint_rois(cols, H, off)
    with api.Engine(W, H) as e:
        gp = dict(p)
        gp["disparityMode"] = gp.pop("mode")
        e.set_sgbm_params(**gp)
        e.set_mean_rois(rois)
        e.compute(l, r, api.STAGE_SGBM | api.STAGE_MEANS)
        out = e.download(1, means=True)
        mm = e.download_minmax(1)[0]
    disp = oracle.sgbm(l, r, p)
    pos = disp[disp > 0]
    assert (int(mm[0]), int(mm[1])) == (int(pos.min()), int(pos.max()))       # Utility::calcMinMaxDisparity
    want_means = np.array([oracle.mean(disp, roi) for roi in rois], f32)
    np.testing.assert_array_equal(out["means"][0], want_means)
    # detection range chosen around the ramp's mean disparity so that some, not all, ROIs fire
    got = run_demo(exe, tmp_path, cols, H, off, 0.6, 3.0, Q, out["means"][0], mm[0], mm[1])
    want = expected_dump(cols, H, off, 0.6, 3.0, Q, want_means, mm[0], mm[1])
    assert got == want
    assert 0 < sum(l_.startswith("M ") for l_ in got) < 81
