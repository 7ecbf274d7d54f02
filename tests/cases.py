"""Shared test-case definitions: parameter sets, seeded inputs, warp maps, reference calibration loader."""
import os

import numpy as np

from mvstereovision3_b200 import synth

SGBM_DEFAULTS = dict(minDisp=0, numDisp=64, blockSize=5, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
                     speckleWindowSize=0, speckleRange=0, mode=0, P1=0, P2=0)


def sgbm_params(**kw):
    d = dict(SGBM_DEFAULTS)
    d.update(kw)
    return d


# (name, params, H, W)
SGBM_CASES = [
    ("sgbm_yml_d64", sgbm_params(minDisp=1, numDisp=64, blockSize=13, speckleWindowSize=150, speckleRange=2), 72, 200),
    ("hh_cfg4_style", sgbm_params(numDisp=32, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, uniquenessRatio=10,
                                  speckleWindowSize=150, speckleRange=2, mode=1), 64, 161),
    ("neg_mind_d16", sgbm_params(minDisp=-2, numDisp=16, blockSize=3, P1=10, P2=120, uniquenessRatio=5), 60, 150),
    ("d48_hh", sgbm_params(minDisp=3, numDisp=48, blockSize=7, P1=50, P2=51, uniquenessRatio=15, disp12MaxDiff=3, mode=1), 81, 183),
    ("bs1_cap63", sgbm_params(numDisp=16, blockSize=1, uniquenessRatio=30, preFilterCap=63), 67, 172),
    ("live_disparity", sgbm_params(numDisp=32, blockSize=9, P1=8 * 81, P2=32 * 81, preFilterCap=31, uniquenessRatio=10,
                                   speckleWindowSize=100, speckleRange=32), 95, 205),
    ("d24_even_bs", sgbm_params(numDisp=24, blockSize=4, uniquenessRatio=-1, disp12MaxDiff=-1), 58, 140),
    ("d128_shipped", sgbm_params(minDisp=1, numDisp=128, blockSize=13, speckleWindowSize=150, speckleRange=2), 40, 300),
    ("d8", sgbm_params(numDisp=8, blockSize=3, P1=7, P2=40, uniquenessRatio=10), 33, 77),
    ("d256_hh", sgbm_params(numDisp=256, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, uniquenessRatio=10,
                            speckleWindowSize=50, speckleRange=2, mode=1), 24, 330),
    ("saturated_s", sgbm_params(numDisp=16, blockSize=5, P1=3000, P2=6000, mode=1, preFilterCap=63), 48, 120),
    ("cap100_wide_cost", sgbm_params(minDisp=2, numDisp=40, blockSize=5, P1=30, P2=200, preFilterCap=100, uniquenessRatio=5), 50, 210),
]

# name -> params for the oracle (BmParams of oracle.loader) ; GPU params are the 5 bm.yml-style fields
BM_CASES = [
    ("bm_yml", dict(numDisp=80, blockSize=21, preFilterCap=2, uniquenessRatio=0, textureThreshold=30), 90, 260),
    ("bm_d16", dict(numDisp=16, blockSize=5, preFilterCap=31, uniquenessRatio=15, textureThreshold=10), 70, 200),
    ("bm_d32_oddH", dict(numDisp=32, blockSize=9, preFilterCap=63, uniquenessRatio=5, textureThreshold=0), 71, 180),
    ("bm_d48_tex", dict(numDisp=48, blockSize=15, preFilterCap=1, uniquenessRatio=0, textureThreshold=200), 74, 203),
    # window ring with one octet per lane (blockSize 23..~100), no ring (above), idle lanes beside a warp's rows (D = 96)
    ("bm_bs31_d32", dict(numDisp=32, blockSize=31, preFilterCap=5, uniquenessRatio=10, textureThreshold=5), 80, 190),
    ("bm_bs111_d16", dict(numDisp=16, blockSize=111, preFilterCap=2, uniquenessRatio=3, textureThreshold=24500), 150, 260),   # the texture sums are 24.4-24.6 K here: the threshold cuts through them (CTA-wide texture kernel)
    ("bm_d96_uniq", dict(numDisp=96, blockSize=7, preFilterCap=31, uniquenessRatio=12, textureThreshold=10), 50, 240),
    ("bm_bs111_cap1", dict(numDisp=16, blockSize=111, preFilterCap=1, uniquenessRatio=0, textureThreshold=0), 150, 260),
    ("bm_d80_bs25", dict(numDisp=80, blockSize=25, preFilterCap=2, uniquenessRatio=8, textureThreshold=30), 60, 230),
]


def sgbm_inputs(name, p, H, W, kind="ramp"):
    seed = sum(name.encode()) & 0xffff
    if kind == "ramp":
        l, r, _ = synth.stereogram(H, W, p["minDisp"], p["numDisp"], seed=seed)
    else:
        l, r = synth.random_pair(H, W, seed=seed + 1000)
    return l, r


def bm_inputs(name, p, H, W, kind="ramp"):
    seed = sum(name.encode()) & 0xffff
    if kind == "ramp":
        l, r, _ = synth.stereogram(H, W, 0, p["numDisp"], seed=seed)
    else:
        l, r = synth.random_pair(H, W, seed=seed + 1000)
    return l, r


def warp_maps(H, W, seed):
    """Smooth synthetic undistort/rectify-like float maps incl. out-of-frame samples."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    mx = xs + 3.0 * np.sin(ys / 17.0) + rng.uniform(-6, 6) + 0.013 * (xs - W / 2)
    my = ys + 2.0 * np.cos(xs / 23.0) + rng.uniform(-6, 6) - 0.011 * (ys - H / 2)
    return mx.astype(np.float32), my.astype(np.float32)


# A real Q of the reference rig (reference afterCalibrationParameters.yml:2-7: f = 303.5 px, baseline 118.7 mm,
# binned 376x240 rig), as the CV_32F copy the drivers hold (trgt/demo.cpp:179-180).
Q_REFERENCE = np.array([[1, 0, 0, -181.93], [0, 1, 0, -124.07], [0, 0, 0, 303.5], [0, 0, 1.0 / 118.7, 0]], np.float32)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
