"""Writes tests/golden/calibrations.json: the reference's three calibration fixtures (parameters/{baseline_small,
smallBL,foobar}/{intrinsic,extrinsic}.yml -- they exist only in the build container under /root/reference, which
is why the numbers are committed) together with what cv2 4.13.0 derives from them along
Stereosystem::initRectification (src/Stereosystem.cpp:193-242): stereoRectify's R0/R1/P0/P1/Q/valid ROIs, the display
ROI, and CRC-32s of the fixed-point maps rint(initUndistortRectifyMap * 32), full frame and 2x2-binned
(src/Stereosystem.cpp:203-208 halves the camera matrices).  Run from the repo root:
    python tests/golden/make_calibrations.py"""
import json
import os
import zlib

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "calibrations.json")
REF = "/root/reference/parameters"


def fixed_crc(mx, my):
    f = np.stack([np.rint(mx * np.float32(32)), np.rint(my * np.float32(32))], -1).astype(np.int32)
    return zlib.crc32(np.ascontiguousarray(f).tobytes())


def main():
    rigs = {}
    for name in ("baseline_small", "smallBL", "foobar"):
        fi = cv2.FileStorage(os.path.join(REF, name, "intrinsic.yml"), cv2.FILE_STORAGE_READ)
        fe = cv2.FileStorage(os.path.join(REF, name, "extrinsic.yml"), cv2.FILE_STORAGE_READ)
        g = lambda fs, k: fs.getNode(k).mat()
        KL, KR = g(fi, "cameraMatrixLeft"), g(fi, "cameraMatrixRight")
        DL, DR = g(fi, "distCoeffsLeft"), g(fi, "distCoeffsRight")
        R, T = g(fe, "R"), g(fe, "T")
        rig = {"KL": KL.tolist(), "KR": KR.tolist(), "DL": DL.ravel().tolist(), "DR": DR.ravel().tolist(),
               "R": R.tolist(), "T": T.ravel().tolist(), "modes": {}}
        for mode, scale, size in (("full", 1.0, (752, 480)), ("binned", 0.5, (376, 240))):
            kl, kr = KL * scale, KR * scale
            R0, R1, P0, P1, Q, roi0, roi1 = cv2.stereoRectify(kl, DL, kr, DR, size, R, T, flags=cv2.CALIB_ZERO_DISPARITY,
                                                              alpha=0, newImageSize=size)
            x0, y0 = max(roi0[0], roi1[0]), max(roi0[1], roi1[1])
            x1 = min(roi0[0] + roi0[2], roi1[0] + roi1[2])
            y1 = min(roi0[1] + roi0[3], roi1[1] + roi1[3])
            m = {"size": list(size), "R0": R0.tolist(), "R1": R1.tolist(), "P0": P0.tolist(), "P1": P1.tolist(),
                 "Q": Q.tolist(), "roi0": list(map(int, roi0)), "roi1": list(map(int, roi1)),
                 "display_roi": [int(x0), int(y0), int(x1 - x0), int(y1 - y0)]}
            m["crc_fixed_left"] = fixed_crc(*cv2.initUndistortRectifyMap(kl, DL, R0, P0, size, cv2.CV_32FC1))
            m["crc_fixed_right"] = fixed_crc(*cv2.initUndistortRectifyMap(kr, DR, R1, P1, size, cv2.CV_32FC1))
            rig["modes"][mode] = m
        rigs[name] = rig
    with open(OUT, "w") as f:
        json.dump({"cv2": cv2.__version__, "rigs": rigs}, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
