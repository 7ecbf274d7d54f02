"""Pins the C oracle against cv2 4.13.0 -- the OpenCV build standing in for the un-vendored
OpenCV the reference calls (src/disparity.cpp:8,20; src/Stereosystem.cpp:252-253).  CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from mvstereovision3_b200 import synth


def cv_sgbm(left, right, p):
    m = cv2.StereoSGBM_create(minDisparity=p["minDisp"], numDisparities=p["numDisp"], blockSize=p["blockSize"],
                              P1=p["P1"], P2=p["P2"], disp12MaxDiff=p["disp12MaxDiff"],
                              preFilterCap=p["preFilterCap"], uniquenessRatio=p["uniquenessRatio"],
                              speckleWindowSize=p["speckleWindowSize"], speckleRange=p["speckleRange"],
                              mode=cv2.STEREO_SGBM_MODE_HH if p["mode"] == 1 else cv2.STEREO_SGBM_MODE_SGBM)
    return m.compute(left, right)


SGBM_CASES = [
    dict(minDisp=1, numDisp=64, blockSize=13, speckleWindowSize=150, speckleRange=2),            # sgbm.yml wiring, cfg 2
    dict(minDisp=0, numDisp=32, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, uniquenessRatio=10,
         speckleWindowSize=150, speckleRange=2, mode=1),                                           # cfg 4 style (HH)
    dict(minDisp=-2, numDisp=16, blockSize=3, P1=10, P2=120, uniquenessRatio=5),
    dict(minDisp=3, numDisp=48, blockSize=7, P1=50, P2=51, uniquenessRatio=15, disp12MaxDiff=3, mode=1),
    dict(minDisp=0, numDisp=16, blockSize=1, uniquenessRatio=30, preFilterCap=63),
    dict(minDisp=0, numDisp=32, blockSize=9, P1=8 * 81, P2=32 * 81, preFilterCap=31, uniquenessRatio=10,
         speckleWindowSize=100, speckleRange=32),                                                  # trgt/liveDisparity.cpp:19-20,61
    dict(minDisp=0, numDisp=24, blockSize=4, uniquenessRatio=-1, disp12MaxDiff=-1),
]


@pytest.mark.parametrize("i", range(len(SGBM_CASES)))
def test_sgbm_oracle_equals_cv2(oracle, i):
    p = oracle.make_sgbm_params(**SGBM_CASES[i])
    H, W = 60 + 7 * i, 150 + 11 * i
    for seed, kind in ((i, "ramp"), (100 + i, "noise")):
        if kind == "ramp":
            left, right, _ = synth.stereogram(H, W, p["minDisp"], p["numDisp"], seed=seed)
        else:
            left, right = synth.random_pair(H, W, seed=seed)
        ref = cv_sgbm(left, right, p)
        got = oracle.sgbm(left, right, p)
        assert ref.dtype == np.int16
        np.testing.assert_array_equal(got, ref)


def test_sgbm_oracle_saturated_s(oracle):
    # many paths x large P2: S saturates at 32767 (inside the contract)
    p = oracle.make_sgbm_params(minDisp=0, numDisp=16, blockSize=5, P1=3000, P2=6000, mode=1, preFilterCap=63)
    left, right = synth.random_pair(48, 120, seed=5)
    np.testing.assert_array_equal(oracle.sgbm(left, right, p), cv_sgbm(left, right, p))


def test_median_and_speckle(oracle):
    rng = np.random.default_rng(3)
    for k in range(6):
        img = (rng.integers(-2, 40, size=(37 + k, 53 + 2 * k)) * 16).astype(np.int16)
        np.testing.assert_array_equal(oracle.median3(img), cv2.medianBlur(img, 3))
        ref = img.copy()
        cv2.filterSpeckles(ref, -16, 5 + 3 * k, 16 * (k % 3))
        np.testing.assert_array_equal(oracle.speckle(img, -16, 5 + 3 * k, 16 * (k % 3)), ref)


def _warp_maps(H, W, seed):
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    mx = xs + 3.0 * np.sin(ys / 17.0) + rng.uniform(-6, 6) + 0.013 * (xs - W / 2)
    my = ys + 2.0 * np.cos(xs / 23.0) + rng.uniform(-6, 6) - 0.011 * (ys - H / 2)
    return mx.astype(np.float32), my.astype(np.float32)


def test_remap(oracle):
    H, W = 96, 140
    for seed in range(6):
        img, _ = synth.random_pair(H, W, seed=seed)
        mx, my = _warp_maps(H, W, seed)
        if seed == 1:   # integer / half / 1/64 tie offsets
            ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
            mx, my = xs + 1.0 + 1.0 / 64, ys - 0.5
        if seed == 2:   # identity
            ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
            mx, my = xs.copy(), ys.copy()
        ref = cv2.remap(img, mx, my, cv2.INTER_LINEAR)
        np.testing.assert_array_equal(oracle.remap(img, mx, my), ref)
        roi = (5, 3, W - 11, H - 9)
        np.testing.assert_array_equal(oracle.remap(img, mx, my, roi), ref[3:3 + H - 9, 5:5 + W - 11])


BM_CASES = [
    dict(numDisp=80, blockSize=21, preFilterCap=2, uniquenessRatio=0, textureThreshold=30),   # configs/bm.yml
    dict(numDisp=16, blockSize=5, preFilterCap=31, uniquenessRatio=15, textureThreshold=10),
    dict(numDisp=32, blockSize=9, preFilterCap=63, uniquenessRatio=5, textureThreshold=0),
    dict(numDisp=48, blockSize=15, preFilterCap=1, uniquenessRatio=0, textureThreshold=200),
    dict(numDisp=16, blockSize=7, preFilterCap=20, uniquenessRatio=10, textureThreshold=50),
]


def cv_bm(left, right, p):
    m = cv2.StereoBM_create(numDisparities=p["numDisp"], blockSize=p["blockSize"])
    m.setPreFilterCap(p["preFilterCap"])
    m.setUniquenessRatio(p["uniquenessRatio"])
    m.setTextureThreshold(p["textureThreshold"])
    m.setMinDisparity(p.get("minDisp", 0))
    return m.compute(left, right)


@pytest.mark.parametrize("i", range(len(BM_CASES)))
def test_bm_oracle_equals_cv2(oracle, i):
    p = BM_CASES[i]
    for H, W in ((70 + i, 200 + 3 * i), (71 + i, 180)):
        left, right, _ = synth.stereogram(H, W, 0, p["numDisp"], seed=i)
        np.testing.assert_array_equal(oracle.bm(left, right, p), cv_bm(left, right, p))
        left, right = synth.random_pair(H, W, seed=50 + i)
        np.testing.assert_array_equal(oracle.bm(left, right, p), cv_bm(left, right, p))


def test_bm_oracle_rejects_min_disp(oracle):
    left, right = synth.random_pair(40, 100, seed=1)
    with pytest.raises(ValueError):
        oracle.bm(left, right, dict(numDisp=16, blockSize=7, minDisp=4))


def test_reproject_and_mean(oracle):
    rng = np.random.default_rng(0)
    disp = (rng.integers(-1, 64 * 16, size=(40, 60))).astype(np.int16)
    Q = np.array([[1, 0, 0, -181.9], [0, 1, 0, -124.0], [0, 0, 0, 303.5], [0, 0, 1 / 118.7, 0]], np.float32)
    xyz, valid = oracle.reproject(disp, Q)
    ref = cv2.reprojectImageTo3D(disp.astype(np.float32) / 16, Q)
    v = valid.astype(bool)
    assert (v == (disp > 0)).all()
    np.testing.assert_allclose(xyz[v], ref[v], rtol=1e-5)
    # mean: int accumulate, truncating division (src/utility.cpp:265-285)
    roi = (7, 5, 20, 13)
    sub = disp[5:18, 7:27].astype(np.int64)
    sel = sub[sub > 1]
    want = float(int(sel.sum()) // len(sel)) if len(sel) and sel.sum() else 0.0
    assert oracle.mean(disp, roi) == want
