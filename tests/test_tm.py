"""Disparity::tm (reference src/disparity.cpp:25-58): per-pixel cv::matchTemplate
(TM_CCORR_NORMED) + cv::minMaxLoc.  The oracle and the device rank candidates by the exact rational N^2/B; OpenCV
evaluates N/sqrt(A*B) in floating point, so the two may pick different offsets only where scores tie to within
float rounding.  Tolerance for those pixels (written here as the task demands): the float64 score at our offset is
within 1e-5 relative of the score at cv2's offset; everywhere else the maps are identical.  Device vs oracle: bit-exact."""
import numpy as np
import pytest

import cases
from mvstereovision3_b200 import synth


def cv_tm(l, r, k):
    """the reference's loop, transcribed call for call onto cv2"""
    import cv2
    H, W = l.shape
    out = np.zeros((H, W), np.uint8)
    for i in range(max(H - k, 0)):
        for j in range(max(W - k, 0)):
            templ = l[i:i + k, j:j + k]
            strip = r[i:i + k, j:j + (W - j - 1)]
            res = cv2.matchTemplate(strip, templ, cv2.TM_CCORR_NORMED)
            out[i, j] = cv2.minMaxLoc(res)[3][0] & 255
    return out


def score(l, r, i, j, x, k):
    a = l[i:i + k, j:j + k].astype(np.float64)
    b = r[i:i + k, j + x:j + x + k].astype(np.float64)
    den = np.sqrt((a * a).sum() * (b * b).sum())
    return (a * b).sum() / den if den > 0 else 0.0


def assert_tm_equivalent(l, r, k, got, want):
    assert got.shape == want.shape
    diff = np.argwhere(got != want)
    for i, j in diff:
        s_got, s_want = score(l, r, i, j, int(got[i, j]), k), score(l, r, i, j, int(want[i, j]), k)
        assert abs(s_got - s_want) <= 1e-5 * max(s_want, 1e-12), (i, j, got[i, j], want[i, j], s_got, s_want)
    return len(diff)


TM_CASES = [("noise", 40, 70, 5), ("noise", 33, 61, 3), ("ramp", 40, 70, 8), ("ramp", 36, 90, 5), ("flat", 24, 40, 4)]


def tm_inputs(kind, H, W):
    if kind == "noise":
        return synth.random_pair(H, W, seed=H + W)
    if kind == "ramp":
        return synth.stereogram(H, W, 0, 16, seed=H)[:2]
    l, r = synth.random_pair(H, W, seed=3)           # coarse quantisation + constant areas: many exact ties
    l, r = (l // 64) * 64, (r // 64) * 64
    l[:, : W // 3] = 0
    r[: H // 2, W // 2:] = 128
    return l, r


@pytest.mark.parametrize("kind,H,W,k", TM_CASES)
def test_oracle_tm_matches_cv2(oracle, kind, H, W, k):
    pytest.importorskip("cv2")
    l, r = tm_inputs(kind, H, W)
    got, want = oracle.tm(l, r, k), cv_tm(l, r, k)
    ndiff = assert_tm_equivalent(l, r, k, got, want)
    if kind != "flat":
        assert ndiff == 0, ndiff          # textured inputs: no ties, identical maps
    assert not got[H - k:].any() and not got[:, W - k:].any()       # untouched border stays 0


def test_oracle_tm_degenerate(oracle):
    l, r = synth.random_pair(6, 9, seed=0)
    assert not oracle.tm(l, r, 6).any() and not oracle.tm(l, r, 9).any()        # k >= rows or cols: all zero
    z = np.zeros((8, 12), np.uint8)
    assert not oracle.tm(z, z, 3).any()                                          # zero energy: score 0, first offset


@pytest.mark.gpu
@pytest.mark.parametrize("kind,H,W,k", TM_CASES + [("noise", 61, 300, 5), ("ramp", 50, 333, 7), ("noise", 20, 64, 1),
                                                    ("noise", 40, 100, 31)])
def test_device_tm_matches_oracle(oracle, kind, H, W, k):
    from mvstereovision3_b200 import api
    l, r = tm_inputs(kind, H, W)
    with api.Engine(W, H) as e:
        got = api.tm(api.Stereopair(l, r), e, k)
    np.testing.assert_array_equal(got, oracle.tm(l, r, k))


@pytest.mark.gpu
def test_device_tm_batch_views_and_cv2(oracle):
    from mvstereovision3_b200 import api
    H, W, k, B = 30, 80, 5, 3
    big = np.zeros((B, H + 4, W + 16), np.uint8)
    big2 = np.zeros_like(big)
    for b in range(B):
        big[b, 2:2 + H, 8:8 + W], big2[b, 2:2 + H, 8:8 + W] = synth.random_pair(H, W, seed=40 + b)
    l, r = big[:, 2:2 + H, 8:8 + W], big2[:, 2:2 + H, 8:8 + W]            # cv::Mat ROI views: step > cols
    with api.Engine(W, H, max_batch=B) as e:
        got = e.tm(l, r, k)
        for b in range(B):
            np.testing.assert_array_equal(got[b], oracle.tm(l[b], r[b], k))
        try:
            import cv2  # noqa: F401
            assert assert_tm_equivalent(l[0], r[0], k, got[0], cv_tm(l[0], r[0], k)) == 0
        except ImportError:
            pass
        for bad in (0, 32):
            with pytest.raises(api.MvsvError):
                e.tm(l, r, bad)
        assert not e.tm(l, r, 30).any()                                   # k >= rows: the reference's loops do not run


@pytest.mark.gpu
def test_device_tm_full_frame_properties():
    """752x480, k = 5 (the value at the reference's call site, trgt/disparityTest.cpp:269): identical images must
    give offset 0 everywhere (score 1 at x = 0 is the first maximum); a right image shifted by s gives s."""
    from mvstereovision3_b200 import api
    l, _ = synth.random_pair(480, 752, seed=9)
    s = 7
    r = np.roll(l, s, axis=1)
    with api.Engine(752, 480) as e:
        same = e.tm(l, l, 5)[0]
        shifted = e.tm(l, r, 5)[0]
    assert not same.any()
    core = shifted[:475, :752 - 5 - s - 1]
    assert (core == s).mean() > 0.999


@pytest.mark.gpu
def test_device_tm_limits(oracle):
    """widest supported image (4096 columns) and the largest kernel size against the oracle on a thin strip"""
    from mvstereovision3_b200 import api
    H, W = 9, 4096
    l, r = synth.random_pair(H, W, seed=77)
    with api.Engine(W, H) as e:
        np.testing.assert_array_equal(e.tm(l, r, 5)[0], oracle.tm(l, r, 5))
    with api.Engine(4100, 8) as e:
        with pytest.raises(api.MvsvError) as ei:
            e.tm(np.zeros((8, 4100), np.uint8), np.zeros((8, 4100), np.uint8), 5)
        assert ei.value.code == -5
