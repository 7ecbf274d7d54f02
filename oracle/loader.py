"""ctypes wrapper around oracle/_build/libmvsv_oracle.so (see mvsv_oracle.c header).

Every wrapper cites the reference call site the C function follows.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmvsv_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mvsv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"] if force else ["make", "-s", "-C", _HERE])
    return _SO


class SgbmParams(C.Structure):
    # field order == Disparity::sgbmParameters (reference inc/disparity.h:17-27) + P1,P2
    _fields_ = [(n, C.c_int) for n in (
        "minDisp", "numDisp", "blockSize", "disp12MaxDiff", "preFilterCap", "uniquenessRatio",
        "speckleWindowSize", "speckleRange", "mode", "P1", "P2")]


class BmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisp", "numDisp", "blockSize", "preFilterCap", "textureThreshold", "uniquenessRatio",
        "speckleWindowSize", "speckleRange", "disp12MaxDiff")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_mean.restype = C.c_float
        _lib.orc_sgbm.restype = C.c_int
        _lib.orc_bm.restype = C.c_int
        _lib.orc_sgbm_cost.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def sgbm_dims(W, p):
    """(minX1, W1, D) of the evaluated cost domain."""
    maxD = p["minDisp"] + p["numDisp"]
    minX1 = max(maxD, 0)
    maxX1 = W + min(p["minDisp"], 0)
    return minX1, maxX1 - minX1, p["numDisp"]


def make_sgbm_params(**kw):
    d = dict(minDisp=0, numDisp=64, blockSize=5, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
             speckleWindowSize=0, speckleRange=0, mode=0, P1=0, P2=0)
    d.update(kw)
    return d


def remap(src, mapx, mapy, roi=None):
    """reference src/Stereosystem.cpp:252-256 (cv::remap INTER_LINEAR + crop)."""
    src = np.ascontiguousarray(src, np.uint8)
    mapx = np.ascontiguousarray(mapx, np.float32)
    mapy = np.ascontiguousarray(mapy, np.float32)
    H, W = src.shape
    if roi is None:
        roi = (0, 0, mapx.shape[1], mapx.shape[0])
    x, y, w, h = roi
    dst = np.empty((h, w), np.uint8)
    lib().orc_remap(_p(src, C.c_uint8), H, W, C.c_size_t(W), _p(mapx, C.c_float), _p(mapy, C.c_float),
                    C.c_size_t(mapx.shape[1]), x, y, w, h, _p(dst, C.c_uint8), C.c_size_t(w))
    return dst


def sgbm(left, right, params, want_volumes=False, want_raw=False):
    """reference src/disparity.cpp:6-10 (StereoSGBM::compute)."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    H, W = left.shape
    sp = SgbmParams(**params)
    disp = np.empty((H, W), np.int16)
    _, W1, D = sgbm_dims(W, params)
    Cv = Sv = raw = None
    if want_volumes and W1 > 0:
        Cv = np.empty((H, W1, D), np.int16)
        Sv = np.empty((H, W1, D), np.int16)
    if want_raw:
        raw = np.empty((H, W), np.int16)
    rc = lib().orc_sgbm(_p(left, C.c_uint8), _p(right, C.c_uint8), H, W, C.c_size_t(W), C.c_size_t(W),
                        C.byref(sp), _p(disp, C.c_int16), C.c_size_t(W),
                        _p(Cv, C.c_int16), _p(Sv, C.c_int16), _p(raw, C.c_int16))
    if rc != 0:
        raise ValueError("orc_sgbm: invalid parameters")
    if want_volumes or want_raw:
        return disp, Cv, Sv, raw
    return disp


def bm(left, right, params, want_prefilter=False):
    """reference src/disparity.cpp:18-22 (StereoBM::compute)."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    H, W = left.shape
    d = dict(minDisp=0, numDisp=64, blockSize=21, preFilterCap=31, textureThreshold=10, uniquenessRatio=15,
             speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1)
    d.update(params)
    bp = BmParams(**d)
    disp = np.empty((H, W), np.int16)
    pl = np.empty((H, W), np.uint8) if want_prefilter else None
    pr = np.empty((H, W), np.uint8) if want_prefilter else None
    rc = lib().orc_bm(_p(left, C.c_uint8), _p(right, C.c_uint8), H, W, C.c_size_t(W), C.c_size_t(W),
                      C.byref(bp), _p(disp, C.c_int16), C.c_size_t(W), _p(pl, C.c_uint8), _p(pr, C.c_uint8))
    if rc != 0:
        raise ValueError("orc_bm: invalid parameters")
    return (disp, pl, pr) if want_prefilter else disp


def median3(img):
    img = np.ascontiguousarray(img, np.int16)
    out = np.empty_like(img)
    lib().orc_median3(_p(img, C.c_int16), _p(out, C.c_int16), img.shape[0], img.shape[1])
    return out


def speckle(img, new_val, max_size, max_diff):
    out = np.ascontiguousarray(img, np.int16).copy()
    lib().orc_speckle(_p(out, C.c_int16), out.shape[0], out.shape[1], int(new_val), int(max_size), int(max_diff))
    return out


def mean(disp, roi):
    """reference src/utility.cpp:265-285 via inc/Subimage.h:31-35."""
    disp = np.ascontiguousarray(disp, np.int16)
    x, y, w, h = roi
    return float(lib().orc_mean(_p(disp, C.c_int16), C.c_size_t(disp.shape[1]), x, y, w, h))


def reproject(disp, Q):
    """reference src/utility.cpp:176-200,242-262 (calcCoordinate over dmap2pcl's loop)."""
    disp = np.ascontiguousarray(disp, np.int16)
    Q = np.ascontiguousarray(Q, np.float32)
    H, W = disp.shape
    xyz = np.empty((H, W, 3), np.float32)
    valid = np.empty((H, W), np.uint8)
    lib().orc_reproject(_p(disp, C.c_int16), C.c_size_t(W), H, W, _p(Q, C.c_float), _p(xyz, C.c_float),
                        _p(valid, C.c_uint8))
    return xyz, valid


def dmap_values(c, Q):
    """reference src/utility.cpp:224-240 (calcDMapValues) -> (dValue, image_x, image_y)."""
    c = np.ascontiguousarray(c, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    out = np.empty(3, np.float32)
    lib().orc_dmap_values(_p(c, C.c_float), _p(Q, C.c_float), _p(out, C.c_float))
    return out


def rectify_maps(K, dist5, R, P, size):
    """reference src/Stereosystem.cpp:214-217 (cv::initUndistortRectifyMap, CV_32FC1); size = (W, H)."""
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    d = np.zeros(5, np.float64)
    dd = np.asarray(dist5, np.float64).ravel()
    d[:min(5, dd.size)] = dd[:5]
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    P = np.ascontiguousarray(P, np.float64).reshape(12)
    W, H = size
    mx = np.empty((H, W), np.float32)
    my = np.empty((H, W), np.float32)
    lib().orc_rectify_maps(_p(K, C.c_double), _p(d, C.c_double), _p(R, C.c_double), _p(P, C.c_double), W, H,
                           _p(mx, C.c_float), _p(my, C.c_float))
    return mx, my


def resize(img, factor):
    """reference src/Stereosystem.cpp:294-295: cv::resize(img, dst, Size(0,0), factor, factor) for CV_8UC1."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    L = lib()
    L.orc_resize_dim.argtypes = [C.c_int, C.c_double]
    L.orc_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
    dw, dh = L.orc_resize_dim(W, float(factor)), L.orc_resize_dim(H, float(factor))
    out = np.empty((dh, dw), np.uint8)
    L.orc_resize(img.ctypes.data, W, H, float(factor), out.ctypes.data)
    return out


def tm(left, right, kernel_size):
    """reference src/disparity.cpp:25-58 (Disparity::tm): CV_8U map of the best-correlating offset."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    H, W = left.shape
    out = np.empty((H, W), np.uint8)
    L = lib()
    L.orc_tm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orc_tm(left.ctypes.data, right.ctypes.data, W, H, int(kernel_size), out.ctypes.data)
    return out
